"""CPU study (numpy/torch): would SINGLE-PASS reduced-precision operands in the mapping network keep the PCA parity bar?

The tcgen05 layer kernel represents both operands of every product as fp16 hi + lo pairs (3 MMAs per product, ~22 bits) --
"fp32-grade".  One MMA per product with plain fp16 or bf16 operands would be 3x less tensor work and would let the eight layers
be fused on chip.  This script rounds the operands of every layer (activations and scaled weights) to the
given format, accumulates in fp32, runs the config-2-shaped decomposition (W space, b = 10000, c = 80) on N samples through the
oracle's chain, and compares with the fp32 run the way the parity tests compare .npz files.
usage: python tools/study_mapping_precision.py [N]"""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import ganspace_oracle as orc


def rounder(fmt):
    if fmt == "fp32":
        return lambda t: t
    if fmt == "fp16":
        return lambda t: t.half().float()
    if fmt == "bf16":
        return lambda t: t.bfloat16().float()
    if fmt == "fp16 hi+lo":
        def split(t):
            hi = t.half().float()
            return hi + (t - hi).half().float()
        return split
    raise ValueError(fmt)


def mapping(z, ws, bs, fmt):
    r = rounder(fmt)
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(z, np.float32))
        x = x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)
        for w, b in zip(ws, bs):
            wt = torch.from_numpy(w) * ((1 / np.sqrt(w.shape[1])) * orc.LR_MLP)
            # the device kernel scales W by a power of two into fp16's range before splitting; a power-of-two scale does not
            # change the rounding, so rounding the scaled weights directly is the same thing
            k = 2.0 ** np.round(np.log2(1.0 / float(wt.abs().max())))
            out = (r(x) @ (r(wt * k) / k).T) if fmt != "fp32" else x @ wt.T
            x = (2 ** 0.5) * torch.nn.functional.leaky_relu(out + torch.from_numpy(b) * orc.LR_MLP, negative_slope=0.2)
        return x.numpy()


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    ws, bs = orc.mapping_random_init(1234)
    normals = lambda s, B_: orc.standard_normal_f32(s, 512 * B_).reshape(B_, 512)
    runs = {}
    for fmt in ("fp32", "fp16 hi+lo", "fp16", "bf16"):
        sample = lambda s, B_, fmt=fmt: mapping(normals(s, B_), ws, bs, fmt)
        runs[fmt] = orc.compute_path(sample, None, 512, 512, N, 10_000, 80, True, use_w=True, ipca="gram")
        z = normals(123, 2000)
        err = np.abs(mapping(z, ws, bs, fmt) - mapping(z, ws, bs, "fp32")).max() / np.abs(mapping(z, ws, bs, "fp32")).max()
        runs[fmt]["_err"] = err
    print(f"N={N} b=10000 c=80, W space; reference = fp32 operands")
    print("| operand format | max activation error (rel. to max) | min signed cos of 80 components | components < 0.999 | max d var_ratio | act_stdev rel | lat_stdev rel |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for fmt in ("fp16 hi+lo", "fp16", "bf16"):
        cmp = orc.compare_npz(runs[fmt], runs["fp32"])
        a = runs[fmt]["act_comp"].reshape(80, -1).astype(np.float64)
        b = runs["fp32"]["act_comp"].reshape(80, -1).astype(np.float64)
        cos = np.sum(a * b, axis=1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
        print(f"| {fmt} | {runs[fmt]['_err']:.1e} | {cmp['min_signed_cos']:.8f} | {(cos < 0.999).sum()} | {cmp['max_abs_dvar_ratio']:.1e} | "
              f"{cmp['act_stdev_rel']:.1e} | {cmp['lat_stdev_rel']:.1e} |")

"""ORACLE -- round-2 golden fixtures (tests/golden/), produced by the UNMODIFIED reference on CPU (oracle/ref_harness.py).

  G9   c5_stylegan2_ffhq_convs4_z_n4000_b500_c4.npz     BASELINE config 5's layer (convs.4, d = 524288), Z space + regression,
                                                         pure random init; act_comp stored as float16 (4.2 MB instead of 8.4)
  G10  c5n_stylegan2_ffhq_convs1_z_noise_n4000_b500_c8.npz   convs.1 with NON-ZERO NoiseInjection weights and activation biases
                                                         (both are 0 at random init; perturbation = gen_golden.perturb_synthesis)

Usage:  python oracle/gen_golden_r2.py [g9] [g10]
"""
import sys
import tempfile
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import ref_harness as rh          # noqa: E402
from oracle.gen_golden import perturb_synthesis  # noqa: E402

OUT = REPO / "tests" / "golden"


def run(layer, n, b, c, perturb=None):
    ref = rh.import_reference()
    dev = torch.device("cpu")
    m = rh.rand_init_stylegan2(ref, dev, "ffhq", 1234)
    if perturb:
        perturb_synthesis(m.model, perturb)
    inst = ref.wrappers.get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=m, use_w=False)
    cfg = ref.Config(model="StyleGAN2", layer=layer, output_class="ffhq", estimator="ipca", use_w=False, n=n, batch_size=b, components=c)
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        path = ref.decomposition.get_or_compute(cfg, inst, force_recompute=True, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp))
        with np.load(path) as data:
            out = {k: data[k].copy() for k in data.files}
        name = path.name
    print(f"{layer}: {time.time() - t0:.0f} s", flush=True)
    inst.close()
    return out, name


def main():
    which = set(sys.argv[1:]) or {"g9", "g10"}
    if "g10" in which:
        names = ["conv1", "convs.0", "convs.1"]
        out, name = run("convs.1", 4_000, 500, 8, perturb=names)
        np.savez_compressed(OUT / "c5n_stylegan2_ffhq_convs1_z_noise_n4000_b500_c8.npz", dump_name=np.array(name),
                            perturbed=np.array(names), **out)
    if "g9" in which:
        out, name = run("convs.4", 4_000, 500, 4)
        out["act_comp_f16"] = out.pop("act_comp").astype(np.float16)
        np.savez_compressed(OUT / "c5_stylegan2_ffhq_convs4_z_n4000_b500_c4.npz", dump_name=np.array(name), **out)


if __name__ == "__main__":
    main()

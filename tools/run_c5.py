"""BASELINE config 5 family on one B200: get_or_compute on a StyledConv feature map, timed end to end.
usage: python tools/run_c5.py [--layer convs.4] [--n 200000] [--b 2000] [--c 80] [--save f.npz] [--compare f.npz]"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ganspace_b200 import _native                                 # noqa: E402
from ganspace_b200.config import Config                           # noqa: E402
from ganspace_b200.decomposition import get_or_compute            # noqa: E402
from ganspace_b200.models import StyleGAN2, get_instrumented_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", default="convs.4")
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--b", type=int, default=2000)
    ap.add_argument("--c", type=int, default=80)
    ap.add_argument("--save")
    ap.add_argument("--compare")
    ap.add_argument("--sections", action="store_true")
    a = ap.parse_args()
    import os
    rank, world = 0, 1
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:                     # torchrun: one process per GPU, NCCL
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
        rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    model = StyleGAN2(dev, "ffhq", random_init=1234)
    inst = get_instrumented_model("StyleGAN2", "ffhq", a.layer, dev, model=model, use_w=False)
    cfg = Config(model="StyleGAN2", layer=a.layer, output_class="ffhq", components=a.c, n=a.n, batch_size=a.b, use_w=False,
                 estimator="ipca")
    _native.instrument.reset()
    _native.instrument.timing = a.sections
    with tempfile.TemporaryDirectory() as tmp:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        path = get_or_compute(cfg, inst, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp), force_recompute=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank != 0:
            return
        with np.load(path) as data:
            out = {k: data[k] for k in data.files}
    res = {"n_gpus": world, "layer": a.layer, "n": a.n, "b": a.b, "c": a.c, "seconds": dt, "samples_per_s": a.n / dt,
           "launches": _native.instrument.launches, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
           "var_ratio_sum": float(out["var_ratio"].sum()), "act_stdev_head": out["act_stdev"][:4].tolist()}
    if a.sections:
        res["sections_ms"] = {k: round(v[0], 1) for k, v in _native.instrument.section_ms().items()}
    comp = out["act_comp"].reshape(out["act_comp"].shape[0], -1).astype(np.float64)
    g = comp @ comp.T
    res["orthonormality"] = float(np.abs(g - np.eye(g.shape[0])).max())
    if a.save:
        np.savez(a.save, **out)
    if a.compare:
        with np.load(a.compare) as data:
            ref = {k: data[k] for k in data.files}
        rc = ref["act_comp"].reshape(comp.shape[0], -1).astype(np.float64)
        cos = np.sum(comp * rc, axis=1)
        lc = np.sum(out["lat_comp"].reshape(comp.shape[0], -1).astype(np.float64) * ref["lat_comp"].reshape(comp.shape[0], -1), axis=1)
        res["vs_other_chain"] = {"min_signed_cos": float(cos.min()), "min_lat_signed_cos": float(lc.min()),
                                 "max_abs_dvar_ratio": float(np.abs(out["var_ratio"] - ref["var_ratio"]).max()),
                                 "act_stdev_rel": float(np.abs(out["act_stdev"] / ref["act_stdev"] - 1).max())}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

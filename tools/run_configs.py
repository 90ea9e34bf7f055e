"""Runs the BASELINE.json configs that are built (1-4) at their full sizes on one GPU through the public API
and prints one JSON line per config with wall-clock and basic invariants (committed under profiles/)."""
import json
import os
import sys
import tempfile
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ganspace_b200.config import Config  # noqa: E402
from ganspace_b200.decomposition import get_or_compute  # noqa: E402
from ganspace_b200.models import get_instrumented_model, get_model  # noqa: E402

dev = torch.device("cuda:0")
CASES = [
    ("config1", dict(model="StyleGAN2", output_class="ffhq", layer="style", use_w=True, n=10_000, batch_size=1_000, components=32), 1234),
    ("config2", dict(model="StyleGAN2", output_class="ffhq", layer="style", use_w=True, n=1_000_000, batch_size=10_000, components=80), 1234),
    ("config3", dict(model="StyleGAN2", output_class="car", layer="style", use_w=False, n=1_000_000, batch_size=10_000, components=80), 1234),
    ("config4", dict(model="BigGAN-512", output_class="husky", layer="generator.gen_z", use_w=False, n=1_000_000, batch_size=2_000, components=80), 4321),
]
for name, kw, seed in CASES:
    model = get_model(kw["model"], kw["output_class"], dev, random_init=seed)
    inst = get_instrumented_model(kw["model"], kw["output_class"], kw["layer"], dev, model=model, use_w=kw["use_w"])
    times = []
    for rep in range(3):
        cfg = Config(estimator="ipca", **kw)
        with tempfile.TemporaryDirectory() as tmp:
            old = sys.stdout
            sys.stdout = open(os.devnull, "w")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            path = get_or_compute(cfg, inst, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp), force_recompute=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sys.stdout = old
            with np.load(path) as d:
                out = {k: d[k] for k in d.files}
        times.append(dt)
    c = out["act_comp"].shape[0]
    comp = out["act_comp"].reshape(c, -1).astype(np.float64)
    lat = out["lat_comp"].reshape(c, -1).astype(np.float64)
    print(json.dumps({
        "config": name, **{k: v for k, v in kw.items()}, "seconds_best_of_3": min(times), "samples_per_s": kw["n"] / min(times),
        "act_comp_shape": list(out["act_comp"].shape), "lat_comp_shape": list(out["lat_comp"].shape),
        "orthonormality_err": float(np.max(np.abs(comp @ comp.T - np.eye(c)))),
        "lat_rows_unit_norm_err": float(np.max(np.abs(np.linalg.norm(lat, axis=1) - 1))),
        "var_ratio_sum": float(out["var_ratio"].sum()), "act_stdev_sorted": bool(np.all(np.diff(out["act_stdev"]) <= 0)),
    }), flush=True)
    inst.close()
    del model, inst
    torch.cuda.empty_cache()

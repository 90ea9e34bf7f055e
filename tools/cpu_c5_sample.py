"""Reference-CPU cost of BASELINE config 5 on a bounded sample ("C4/C5 at reduced N and extrapolated
linearly in the number of partial_fits, stating the extrapolation").  Runs the ORACLE (numpy/torch-CPU restatement of the
reference's compute(), sklearn-form IncrementalPCA) for layer convs.4 (d = 524288) with N = 4000, b = 2000 -> 2 partial_fits,
and times the phases.  usage: python tools/cpu_c5_sample.py [out.json]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import ganspace_oracle as orc      # noqa: E402

t_fit, t_act = [], []
_fit, _fwd = orc.ipca_partial_fit, orc.synthesis_forward


def fit(st, X):
    t = time.perf_counter()
    r = _fit(st, X)
    t_fit.append(time.perf_counter() - t)
    return r


def fwd(*a, **k):
    t = time.perf_counter()
    r = _fwd(*a, **k)
    t_act.append(time.perf_counter() - t)
    return r


orc.ipca_partial_fit, orc.synthesis_forward = fit, fwd
ws, bs = orc.mapping_random_init(1234)
params = orc.synthesis_random_init(1234, 1024, "convs.4")
t0 = time.perf_counter()
out = orc.compute_stylegan2_layer(ws, bs, params, "convs.4", 4000, 2000, 80, regress=False, form="reference")
total = time.perf_counter() - t0
import torch
res = {"layer": "convs.4", "d": 524288, "n": 4000, "b": 2000, "c": 80, "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
       "total_s": total, "partial_fit_s": t_fit, "synthesis_s_total": sum(t_act), "synthesis_samples": 4000,
       "extrapolation_200k_s": (sum(t_act) / 4000) * 400_000 + t_fit[0] + 99 * t_fit[-1],
       "note": "PCA half only (regress=False); the extrapolation adds the regression pass's 200k forward samples, 1 float32 first fit "
               "and 99 float64 stacked fits; reference-form StyledConv (per-sample weights, grouped conv) as the reference computes it",
       "var_ratio_head": out["var_ratio"][:4].tolist()}
print(json.dumps(res))
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(res) + "\n")

"""GPU microbenchmark of the mapping network (8 x mapping_layer_tc_kernel + pixelnorm_split) alone: 1,010,000 rows, CUDA-event
time per forward, error against an fp64 evaluation on a row subset.  GANSPACE_B200_MAPPING_CLUSTER selects the cluster size.
usage: python tools/bench_mapping.py [rows] [free_sms]"""
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ganspace_b200 import _native as nat
from oracle import ganspace_oracle as orc

nat.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_010_000
free_sms = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
ws, bs = orc.mapping_random_init(1234)
pm = nat.PackedMapping(torch.tensor(np.stack(ws)).to(dev), torch.tensor(np.stack(bs)).to(dev), 0.01)
z = torch.randn((rows, 512), device=dev, dtype=torch.float32)
out = torch.empty_like(z)
for _ in range(3):
    pm.forward(z, out=out, leave_free_sms=free_sms)
torch.cuda.synchronize()
times = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pm.forward(z, out=out, leave_free_sms=free_sms); e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
pm.check()
sub = slice(0, 4096)
zs = z[sub].double().cpu().numpy()
x = zs / np.sqrt((zs ** 2).mean(1, keepdims=True) + 1e-8)
for w, b in zip(ws, bs):
    x = x @ (w.astype(np.float64) * (0.01 / np.sqrt(512))).T + b.astype(np.float64) * 0.01
    x = np.where(x >= 0, x, 0.2 * x) * np.sqrt(2.0)
err = float(np.abs(out[sub].cpu().numpy() - x).max() / np.abs(x).max())
ms = min(times)
print(json.dumps({"rows": rows, "cluster": os.environ.get("GANSPACE_B200_MAPPING_CLUSTER", "4"), "free_sms": free_sms,
                  "ms_best": ms, "ms_all": [round(t, 3) for t in times], "tflops_algorithmic": rows * 8 * 2 * 512 * 512 / ms / 1e9,
                  "max_rel_err_vs_fp64": err}))

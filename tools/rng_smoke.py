import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ganspace_b200 import _native as nat
from oracle import ganspace_oracle as orc
for n in (2, 7, 512, 4993, 512000):
    t = time.time()
    out = nat.legacy_normal([1791095845, 2135392491, 7], n, "cuda").cpu().numpy()
    ok = all(np.array_equal(out[i], orc.standard_normal_f32(s, n)) for i, s in enumerate([1791095845, 2135392491, 7]))
    print("n", n, "ok", ok, round(time.time() - t, 3), flush=True)
out = nat.legacy_normal(list(range(100, 112)), 512000, "cuda"); torch.cuda.synchronize(); print("12 streams ok", flush=True)

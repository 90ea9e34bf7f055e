"""RNG kernel alone: S streams of 512*B normals (config 2: S=101, B=10000); time per launch.  usage: rng_probe.py [S] [B] [parts]"""
import sys, torch
sys.path.insert(0, ".")
from ganspace_b200 import _native as nat
nat.load()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 101
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
seeds = list(range(1000, 1000 + S))
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = torch.empty((S, 512 * B), dtype=torch.float32, device="cuda")
for _ in range(2):
    nat.legacy_normal(seeds, 512 * B, "cuda", out=out, parts=parts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); nat.legacy_normal(seeds, 512 * B, "cuda", out=out, parts=parts); e1.record(); torch.cuda.synchronize()
print(f"S={S} B={B} parts={parts}: {e0.elapsed_time(e1):.3f} ms")

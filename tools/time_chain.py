import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ganspace_b200 import _native as nat
rng = np.random.RandomState(0)
d, c, NB = 512, 80, 10000
basis = rng.standard_normal((d, d)) * (0.98 ** np.arange(d))[None, :]
Xs = [torch.tensor((rng.standard_normal((NB, d)) @ basis.T).astype(np.float32)).cuda() for _ in range(6)]
stats = [nat.batch_stats(X) for X in Xs]
chain = nat.IPCAChain(d, c, "cuda", side_stream=False)
torch.cuda.synchronize()
for k, (m, G) in enumerate(stats):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); chain.step(NB, m, G); e1.record(); torch.cuda.synchronize()
    print("step", k, "ms", round(e0.elapsed_time(e1), 3), flush=True)
out = chain.export()
print(os.environ.get("GANSPACE_B200_CHAIN", "default"), float(out["singular_values"][0]), float(out["components"][79, :5].abs().sum()))

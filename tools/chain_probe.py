"""GPU probe of the round-2 chain step: config-2-shaped chain (d=512, c=80, b=10000) on real W-space statistics,
compared with the oracle's exact Gram-form chain; prints iterations / residual per step and the time per step."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, ".")
from ganspace_b200 import _native as nat
from oracle import ganspace_oracle as orc

nat.load()
dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B = 10000
ws, bs = orc.mapping_random_init(1234)
stats = []
st = orc.IPCAState(80)
for k in range(K):
    z = orc.standard_normal_f32(1000 + k, 512 * B).reshape(B, 512)
    X = orc.mapping_forward(z, ws, bs)
    n_b, m, G = orc.batch_stats(X)
    orc.ipca_gram_step(st, n_b, m, G)
    stats.append((torch.tensor(m, device=dev), torch.tensor(G, device=dev)))
chain = nat.IPCAChain(512, 80, dev, side_stream=False)
hdrs = []
for k in range(K):
    chain.step(B, *stats[k])
    hdrs.append(chain.state[:192].view(torch.float64).cpu().numpy().copy())
out = {kk: v.cpu().numpy() for kk, v in chain.export().items()}
for k, h in enumerate(hdrs):
    print(f"step {k}: n_seen {h[0]:.0f} form {h[2]:.0f} buf {h[3]:.0f} iters {h[4]:.0f} rel {h[5]:.2e}")
cos = np.sum(out["components"] * st.components, axis=1)
print(f"min signed cos {cos.min():.10f}; max |d ratio| {np.max(np.abs(out['explained_variance_ratio'] - st.explained_variance_ratio)):.2e}; "
      f"sv rel {np.max(np.abs(out['singular_values'] / st.singular_values - 1)):.2e}; "
      f"orth {np.max(np.abs(out['components'] @ out['components'].T - np.eye(80))):.2e}")
# timing: replay the chain several times
torch.cuda.synchronize()
for rep in range(2):
    chain = nat.IPCAChain(512, 80, dev, side_stream=False)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    chain.step(B, *stats[0])
    e1.record()
    for k in range(1, K):
        chain.step(B, *stats[k])
    e2.record()
    torch.cuda.synchronize()
    hd = chain.state[:192].view(torch.float64).cpu().numpy()
    tot_it = hd[7]
    names = ["setup+T", "buildG", "gemm", "partial", "reduce", "resid", "chol", "publish", "commit"]
    print("   clocks per phase (CTA 0): " + ", ".join(f"{n} {hd[8 + i] / 1.965e3:.0f}us" for i, n in enumerate(names)) +
          f"; iterations {hd[17]:.0f}, steps {hd[18]:.0f}")
    print(f"first step {e0.elapsed_time(e1) * 1e3:.0f} us; {K - 1} subspace steps {e1.elapsed_time(e2) * 1e3:.0f} us "
          f"({e1.elapsed_time(e2) * 1e3 / (K - 1):.0f} us/step, {tot_it:.0f} iterations, "
          f"{e1.elapsed_time(e2) * 1e3 / max(tot_it + K - 1, 1):.1f} us per GEMM-equivalent)")

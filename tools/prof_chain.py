import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from ganspace_b200 import _native as nat
rng = np.random.RandomState(0)
d, c = 512, 80
B = rng.standard_normal((d, 3 * d)) * (0.97 ** np.arange(d))[:, None]
A = torch.tensor(B @ B.T).cuda()
for _ in range(3):
    ev, evec = nat.sym_eig_top(A, c)
torch.cuda.synchronize()
out = nat.legacy_normal([1, 2, 3, 4], 512 * 2000, "cuda")
torch.cuda.synchronize()
print(float(ev[0]), float(out.sum()))

"""GPU probe of the round-2 eigensolver: accuracy vs LAPACK and event-timed cost per call, several sizes."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ganspace_b200 import _native as nat

nat.load()
dev = torch.device("cuda:0")
for d, c in [(96, 12), (128, 128), (240, 80), (256, 40), (512, 80), (512, 512), (384, 80)]:
    rng = np.random.RandomState(d + c)
    B = rng.standard_normal((d, 3 * d)) * (0.97 ** np.arange(d))[:, None]
    A = B @ B.T
    lam, Q = np.linalg.eigh(A)
    lam, Q = lam[::-1][:c], Q[:, ::-1][:, :c].T
    At = torch.tensor(A, device=dev)
    try:
        ev, evec = nat.sym_eig_top(At, c)
    except Exception as ex:
        print(f"d={d} c={c}: FAILED {ex}")
        continue
    ev, evec = ev.cpu().numpy(), evec.cpu().numpy()
    R = A @ evec.T - evec.T * ev[None, :]
    print(f"d={d} c={c}: eval err {np.max(np.abs(ev - lam)) / lam[0]:.2e} resid {np.max(np.linalg.norm(R, axis=0)) / lam[0]:.2e} "
          f"orth {np.max(np.abs(evec @ evec.T - np.eye(c))):.2e} min|cos| {np.min(np.abs(np.sum(evec * Q, axis=1))):.10f}", flush=True)
    lib = nat.load()
    import ctypes as C
    ws = torch.empty(lib.gsb_ipca_workspace_bytes(d, c), dtype=torch.uint8, device=dev)
    evals = torch.empty(c, dtype=torch.float64, device=dev)
    evecs = torch.empty((c, d), dtype=torch.float64, device=dev)
    a = At.clone()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        lib.gsb_sym_eig_top(C.c_void_p(a.data_ptr()), d, c, C.c_void_p(evals.data_ptr()), C.c_void_p(evecs.data_ptr()),
                            C.c_void_p(ws.data_ptr()), ws.numel(), st)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    print(f"    {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per sym_eig_top call", flush=True)
    nat.check_eig_status("probe")

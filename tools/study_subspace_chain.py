"""CPU study (numpy) for DESIGN.md section 5b: the IPCA chain needs only the top-c INVARIANT SUBSPACE per step.

sklearn keeps (V, S) and re-forms  V^T S^2 V  in the next partial_fit; with any orthonormal basis Q of the same subspace and
H = Q^T G Q that term is  Q H Q^T  exactly, so the per-step eigen-decomposition can be replaced by a subspace iteration
(GEMMs + orthonormalisation) and ONE eigen-decomposition of the final H at export.  This script runs config-2-shaped chains
(real W-space activations of the random-init mapping network, b=10000, c=80) three ways and compares the exported components:
   exact  : eigh(G) every step (oracle.ipca_gram_step)
   subsp  : subspace iteration to a residual tolerance, no per-step eigen-decomposition
It prints the iterations each step needed and the final signed cosines / variance-ratio differences."""
import sys
import time
import numpy as np

sys.path.insert(0, ".")
from oracle import ganspace_oracle as orc


def batches(nsteps, B=10000, seed0=1):
    ws, bs = orc.mapping_random_init(1234)
    for k in range(nsteps):
        z = orc.standard_normal_f32(1000 + k, 512 * B).reshape(B, 512)
        X = orc.mapping_forward(z, ws, bs)
        yield orc.batch_stats(X)


def orth(Y):
    W = Y.T @ Y
    L = np.linalg.cholesky(W)
    return np.linalg.solve(L, Y.T).T


def run(nsteps, c=80, tol=1e-7, maxit=60, verbose=True):
    st = orc.IPCAState(c)
    Q = H = None
    mean = unnorm = None
    n_seen = 0
    its = []
    t0 = time.time()
    for k, (n_b, mean_b, gram_b) in enumerate(batches(nsteps)):
        orc.ipca_gram_step(st, n_b, mean_b, gram_b)
        n_tot = n_seen + n_b
        if n_seen == 0:
            lam, E = np.linalg.eigh(gram_b)
            Q = E[:, ::-1][:, :c].copy()
            H = np.diag(lam[::-1][:c])
            mean = mean_b.copy(); unnorm = np.diag(gram_b).copy()
            its.append(0)
        else:
            m = np.sqrt((n_seen / n_tot) * n_b) * (mean - mean_b)
            G = Q @ H @ Q.T + gram_b + np.outer(m, m)
            it = 0
            while True:
                Y = G @ Q
                Hn = Q.T @ Y
                R = Y - Q @ Hn
                # convergence: residual relative to the smallest Ritz-ish scale (min diagonal of H)
                rel = np.linalg.norm(R) / np.min(np.diag(Hn))
                if rel < tol or it >= maxit:
                    H = 0.5 * (Hn + Hn.T)
                    break
                Q = orth(Y)
                it += 1
            its.append(it)
            unnorm = unnorm + np.diag(gram_b) + (n_seen * n_b / n_tot) * (mean - mean_b) ** 2
            mean = (mean * n_seen + mean_b * n_b) / n_tot
        n_seen = n_tot
    lam, Z = np.linalg.eigh(H)
    lam = lam[::-1]; Z = Z[:, ::-1]
    V = (Q @ Z).T
    V, _ = orc.svd_flip_v(V)
    ratio = lam / np.sum(unnorm)
    cos = np.sum(V * st.components, axis=1)
    dr = np.max(np.abs(ratio - st.explained_variance_ratio))
    if verbose:
        print(f"steps={nsteps} tol={tol:g}: iterations per step {its}")
        print(f"   total iterations {sum(its)}; min signed cos {cos.min():.10f}; max |d ratio| {dr:.2e}; "
              f"offdiag(H)/diag {np.abs(H - np.diag(np.diag(H))).max() / np.diag(H).min():.2e}  [{time.time() - t0:.0f}s]")
    return its, cos.min(), dr


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    for tol in (1e-4, 1e-6, 1e-8):
        run(n, tol=tol)

"""CPU study (numpy): the round-2 chain step (orthogonal iteration on (Q, H) with the residual test) on
spectra that are NOT those of random-init W space -- the advisor's request after round 1: strongly decaying spectra, a cluster
of nearly equal eigenvalues straddling the cut at component c, a flat (isotropic) tail, and a trained-checkpoint-like power law.
For each: K groups of NB samples of x = (g * sigma) R^T + mu, the exact chain (oracle.ipca_gram_step = sklearn's Gram form)
against the subspace chain at the device defaults (tol 1e-4, iteration cap 60), compared the way the parity tests compare
(.npz fields): signed cosine per component, variance ratios.  For components inside a cluster the individual eigenvectors are
ill-conditioned for ANY method, so the subspace angle of the top-c span is reported as well.
usage: python tools/study_subspace_spectra.py [d] [c] [K] [NB]"""
import sys
import numpy as np

sys.path.insert(0, ".")
from oracle import ganspace_oracle as orc


def orth(Y):
    L = np.linalg.cholesky(Y.T @ Y)
    return np.linalg.solve(L, Y.T).T


def chain(stats, c, tol=1e-4, maxit=60):
    st = orc.IPCAState(c)
    Q = H = mean = unnorm = None
    n_seen, its, capped = 0, [], 0
    for n_b, mean_b, gram_b in stats:
        orc.ipca_gram_step(st, n_b, mean_b.copy(), gram_b.copy())
        n_tot = n_seen + n_b
        if n_seen == 0:
            lam, E = np.linalg.eigh(gram_b)
            Q, H = E[:, ::-1][:, :c].copy(), np.diag(lam[::-1][:c])
            mean, unnorm = mean_b.copy(), np.diag(gram_b).copy()
            its.append(0)
        else:
            m = np.sqrt((n_seen / n_tot) * n_b) * (mean - mean_b)
            G = Q @ H @ Q.T + gram_b + np.outer(m, m)
            it = 0
            while True:
                Y = G @ Q
                Hn = Q.T @ Y
                rel = np.linalg.norm(Y - Q @ Hn) / np.min(np.diag(Hn))
                if rel < tol or it >= maxit:
                    capped += int(rel >= tol)
                    H = 0.5 * (Hn + Hn.T)
                    break
                Q = orth(Y)
                it += 1
            its.append(it)
            unnorm = unnorm + np.diag(gram_b) + (n_seen * n_b / n_tot) * (mean - mean_b) ** 2
            mean = (mean * n_seen + mean_b * n_b) / n_tot
        n_seen = n_tot
    lam, Z = np.linalg.eigh(H)
    lam, Z = lam[::-1], Z[:, ::-1]
    V, _ = orc.svd_flip_v((Q @ Z).T)
    cos = np.sum(V * st.components, axis=1)
    dr = np.max(np.abs(lam / np.sum(unnorm) - st.explained_variance_ratio))
    # largest principal angle between the two top-c spans
    s = np.linalg.svd(V @ st.components.T, compute_uv=False)
    gap = st.explained_variance[:-1] / st.explained_variance[1:] - 1.0
    return dict(its=its, capped=capped, cos=cos, dr=dr, span=float(np.sqrt(max(0.0, 1.0 - s.min() ** 2))), min_gap=float(gap.min()),
                n_bad=int((cos < 0.999).sum()))


def spectra(d, c):
    i = np.arange(d, dtype=np.float64)
    out = {"decay 0.8^i": 0.8 ** i + 1e-6,
           "power law i^-1": 1.0 / (1.0 + i),
           "flat tail": np.where(i < c // 2, 4.0 * 0.9 ** i, 0.05),
           "random-init-like i^-0.25": (1.0 + i) ** -0.25}
    cl = 1.0 / (1.0 + 0.15 * i)
    cl[c - 3:c + 3] = cl[c - 3] * (1.0 - 1e-3 * np.arange(6))        # six eigenvalues within 0.5 %, three on each side of the cut
    out["cluster across the cut (0.1 % apart)"] = cl
    return out


if __name__ == "__main__":
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    c = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    NB = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
    rng = np.random.RandomState(0)
    R = np.linalg.qr(rng.standard_normal((d, d)))[0]
    mu = rng.standard_normal(d)
    print(f"d={d} c={c} K={K} NB={NB}  (tol 1e-4, cap 60)")
    print("| spectrum (sigma_i) | iterations per step (first 6 ... last) | total | cap hit | min signed cos | components < 0.999 | span sin(theta_max) | max d var_ratio | min rel. gap of the exact chain |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
    for name, sig in spectra(d, c).items():
        stats = []
        for k in range(K):
            X = ((rng.standard_normal((NB, d)) * sig) @ R.T + mu).astype(np.float32)
            stats.append(orc.batch_stats(X))
        r = chain(stats, c)
        print(f"| {name} | {r['its'][:6]} ... {r['its'][-1]} | {sum(r['its'])} | {r['capped']} | {r['cos'].min():.8f} | {r['n_bad']} | "
              f"{r['span']:.2e} | {r['dr']:.1e} | {r['min_gap']:.1e} |")

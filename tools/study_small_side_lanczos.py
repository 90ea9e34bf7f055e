"""CPU study (numpy, fp64): how well does the warm-started block Lanczos / Rayleigh-Ritz step track the exact small-side
IPCA chain on CONV feature maps?  (On config 5 the 3-block variant reached only cos 0.99994 vs the direct solve.)
Activations: oracle StyledConv chain to convs.1 (d = 32768), groups of NB = 2000, c = 80.
usage: python tools/study_small_side_lanczos.py [groups=8]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import ganspace_oracle as orc      # noqa: E402


def orth_rows(R):
    q, _ = np.linalg.qr(R.T)
    return q.T


def rr_top(T, basis, c):
    H = basis @ T @ basis.T
    H = 0.5 * (H + H.T)
    lam, U = np.linalg.eigh(H)
    return lam[::-1][:c], (U[:, ::-1][:, :c].T @ basis)


def krylov(T, Q0, blocks):
    bs = [Q0]
    for _ in range(blocks - 1):
        B = np.vstack(bs)
        W = bs[-1] @ T
        for _ in range(2):
            W = W - (W @ B.T) @ B
        bs.append(orth_rows(W))
    return np.vstack(bs)


def solve(T, c, variant):
    n = T.shape[0]
    if variant == "exact":
        lam, U = np.linalg.eigh(T)
        return lam[::-1][:c], U[:, ::-1][:, :c].T
    E = np.eye(n)[:c]
    if variant == "lanczos3":
        return rr_top(T, krylov(T, E, 3), c)
    if variant == "lanczos4":
        return rr_top(T, krylov(T, E, 4), c)
    if variant == "lanczos3x2":                 # two Rayleigh-Ritz rounds, the second restarted from the first's Ritz vectors
        lam, U = rr_top(T, krylov(T, E, 3), c)
        return rr_top(T, krylov(T, U, 3), c)
    if variant == "lanczos2x2":
        lam, U = rr_top(T, krylov(T, E, 2), c)
        return rr_top(T, krylov(T, U, 2), c)
    raise ValueError(variant)


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    NB, c, layer = 2000, 80, "convs.1"
    ws, bs = orc.mapping_random_init(1234)
    params = orc.synthesis_random_init(1234, 1024, layer)
    noises = orc.fixed_noise(0, 1024)
    t0 = time.time()
    groups = []
    for k in range(K):
        z = orc.standard_normal_f32(1000 + k, 512 * NB).reshape(NB, 512)
        acts = [orc.synthesis_forward(orc.mapping_forward(z[i:i + 250], ws, bs), params, noises, layer, form="shared")
                for i in range(0, NB, 250)]
        groups.append(np.concatenate(acts).reshape(NB, -1).astype(np.float64))
    d = groups[0].shape[1]
    print(f"activations: {K} x [{NB}, {d}] in {time.time() - t0:.0f} s", flush=True)
    variants = ["exact", "lanczos3", "lanczos4", "lanczos3x2", "lanczos2x2"]
    state = {v: dict(D=np.zeros((c, d)), mean=np.zeros(d), n=0, S=None) for v in variants}
    for k, X in enumerate(groups):
        mb = X.mean(0)
        Xc = X - mb
        line = [f"step {k}"]
        for v in variants:
            st = state[v]
            corr = np.sqrt(st["n"] / (st["n"] + NB) * NB) * (st["mean"] - mb) if st["n"] else np.zeros(d)
            M = np.vstack([st["D"], Xc, corr[None]])
            T = M @ M.T
            lam, U = solve(T, c, "exact" if st["n"] == 0 else v)
            Dn = U @ M
            idx = np.argmax(np.abs(Dn), axis=1)
            st["D"] = Dn * np.sign(Dn[np.arange(c), idx])[:, None]
            st["S"] = np.sqrt(np.maximum(lam, 0))
            st["mean"] = (st["mean"] * st["n"] + mb * NB) / (st["n"] + NB)
            st["n"] += NB
            if v != "exact":
                ce = state["exact"]["D"] / state["exact"]["S"][:, None]
                cv = st["D"] / st["S"][:, None]
                cos = np.sum(ce * cv, axis=1)
                line.append(f"{v}: min cos {cos.min():.9f} (comp {int(cos.argmin())}), sv rel {np.abs(st['S'] / state['exact']['S'] - 1).max():.1e}")
        print(" | ".join(line), flush=True)
    S = state["exact"]["S"]
    print("relative eigen-gaps (lambda_i - lambda_{i+1}) / lambda_1, components 70..79:", np.round((S[70:79] ** 2 - S[71:80] ** 2) / S[0] ** 2, 6))


if __name__ == "__main__":
    main()

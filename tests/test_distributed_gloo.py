"""world_size-2 gloo test (CPU) of the multi-GPU host logic: round-robin group ownership, the per-round
all-gather of the statistics, and the ordered merge -- with the oracle standing in for the device
kernels.  The sharded result must equal the single-process chain exactly (it is independent of the
world size by construction)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _groups(seed_base, K, NB, d):
    rng = np.random.RandomState(seed_base)
    basis = rng.standard_normal((d, d)) * (0.9 ** np.arange(d))[None, :]
    return [(np.random.RandomState(seed_base + 1 + k).standard_normal((NB, d)) @ basis.T + k * 0.01).astype(np.float32)
            for k in range(K)]


def _worker(rank, world, port, out_path):
    """The protocol of decomposition.compute_arrays' small-d branch: plan.rounds, k mod world ownership, one all-gather of the
    round's statistics (decomposition._StatsExchange), merges in group order -- the oracle standing in for the device chain."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ganspace_b200 import plan, decomposition
    from oracle import ganspace_oracle as orc
    d, c = 48, 6
    pl = plan.make_plan(n=10_000, B=1_000, components=c)      # N=10000, NB=2000, K=5
    data = _groups(5, pl.K, pl.NB, d)
    touched = plan.groups_to_process(pl, rank, world)
    assert pl.K - 1 in touched
    st = orc.IPCAState(c)

    class Est:                                                 # IPCAEstimator.fit_partial_stats
        def fit_partial_stats(self, nb, m, g):
            orc.ipca_gram_step(st, nb, m.numpy().copy(), g.numpy().copy())
            return True

    ex = decomposition._StatsExchange(d, world)
    pending, seen = None, []
    for rnd in plan.rounds(0, pl.K, world, 1, 2):             # rounds of 2, 4, ... groups (ragged last round)
        own = [k for k in rnd if plan.owner(k, world) == rank]
        g_max = -(-len(rnd) // world)
        means = torch.zeros((g_max, d), dtype=torch.float64)
        grams = torch.zeros((g_max, d, d), dtype=torch.float64)
        for i, k in enumerate(own):
            assert k in touched
            n_b, m, G = orc.batch_stats(data[k])
            means[i] = torch.from_numpy(m)
            grams[i] = torch.from_numpy(G)
        cur = ex.start(rnd, means, grams)
        if pending is not None:
            assert ex.finish(pending, Est(), pl.NB)
        pending = cur
        seen += list(rnd)
    assert ex.finish(pending, Est(), pl.NB) and seen == list(range(pl.K))
    if rank == 0:
        np.savez(out_path, comp=st.components, sv=st.singular_values, mean=st.mean, ratio=st.explained_variance_ratio)
    # every rank holds the same state
    ref = torch.from_numpy(st.components.copy())
    gathered = [torch.zeros_like(ref) for _ in range(world)]
    dist.all_gather(gathered, ref)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    dist.destroy_process_group()


def test_sharded_stats_exchange_equals_single_process(tmp_path, oracle):
    from ganspace_b200 import plan
    world = 2
    out_path = str(tmp_path / "sharded.npz")
    mp.spawn(_worker, args=(world, _free_port(), out_path), nprocs=world, join=True)
    got = np.load(out_path)
    d, c = 48, 6
    pl = plan.make_plan(10_000, 1_000, c)
    st = oracle.IPCAState(c)
    for X in _groups(5, pl.K, pl.NB, d):
        oracle.ipca_gram_step(st, *oracle.batch_stats(X))
    assert np.array_equal(got["comp"], st.components)
    assert np.array_equal(got["sv"], st.singular_values)
    assert np.array_equal(got["mean"], st.mean)
    # and the Gram chain is the sklearn chain
    st2 = oracle.IPCAState(c)
    for X in _groups(5, pl.K, pl.NB, d):
        oracle.ipca_partial_fit(st2, X)
    assert np.min(np.sum(st2.components * got["comp"], axis=1)) > 1 - 1e-8


# ---- feature-sharded large-d chain: row-parallel generation, all-to-all, Gram all-reduce ----------
def _big_batches(d, nb, k):
    rng = np.random.RandomState(99)
    basis = rng.standard_normal((d, 24)).astype(np.float32) * (0.8 ** np.arange(24, dtype=np.float32))[None, :]
    return [((rng.standard_normal((nb, 24)).astype(np.float32) @ basis.T) + 0.05 * rng.standard_normal((nb, d)).astype(np.float32)
             + 1.5).astype(np.float32) for _ in range(k)]


def _sharded_worker(rank, world, port, out_path):
    """The protocol of _native.BigIPCA(shard=...).step with numpy standing in for the three device phases."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ganspace_b200 import _native
    d, nb, c, k = 256, 60, 5, 3
    dl, q = d // world, nb // world
    D = np.zeros((c, dl), np.float32)
    mean = np.zeros(dl)
    n_seen = 0
    for X in _big_batches(d, nb, k):
        stage = torch.from_numpy(X[rank * q:(rank + 1) * q].copy())            # this rank's rows, all features
        rows = torch.empty((nb, dl), dtype=torch.float32)
        _native.exchange_rows(stage, rows, world)                               # all rows, this rank's feature block
        assert np.array_equal(rows.numpy(), X[:, rank * dl:(rank + 1) * dl])
        Xl = rows.numpy().astype(np.float64)
        mb = Xl.mean(0)
        corr = np.sqrt(n_seen / (n_seen + nb) * nb) * (mean - mb) if n_seen else np.zeros(dl)
        M = np.vstack([D, (Xl - mb).astype(np.float32), corr.astype(np.float32)[None]]).astype(np.float64)
        T = torch.from_numpy(M @ M.T)
        dist.all_reduce(T)                                                      # phase 1 -> sum over the feature shards
        lam, U = np.linalg.eigh(T.numpy())
        lam, U = lam[::-1][:c], U[:, ::-1][:, :c].T
        Dn = (U @ M).astype(np.float32)                                         # phase 2
        idx = np.argmax(np.abs(Dn), axis=1)
        rowmax = torch.from_numpy(np.stack([np.abs(Dn)[np.arange(c), idx], Dn[np.arange(c), idx]], axis=1).astype(np.float32))
        allmax = torch.empty((world * c, 2))
        dist.all_gather_into_tensor(allmax, rowmax)
        signs = _native.pick_global_signs(allmax.view(world, c, 2)).numpy()
        D = Dn * signs[:, None]                                                 # phase 3
        S = np.sqrt(lam)
        mean = (mean * n_seen + mb * nb) / (n_seen + nb)
        n_seen += nb
    comp_l = torch.from_numpy((D / S[:, None]).astype(np.float32))
    full = torch.empty((world * c, dl))
    dist.all_gather_into_tensor(full, comp_l)
    if rank == 0:
        np.savez(out_path, comp=full.view(world, c, dl).permute(1, 0, 2).reshape(c, d).numpy(), sv=S)
    dist.destroy_process_group()


def test_feature_sharded_chain_protocol(tmp_path, oracle):
    world = 2
    out_path = str(tmp_path / "fs.npz")
    mp.spawn(_sharded_worker, args=(world, _free_port(), out_path), nprocs=world, join=True)
    got = np.load(out_path)
    st = oracle.IPCAState(5)
    for X in _big_batches(256, 60, 3):
        oracle.ipca_partial_fit(st, X)
    cos = np.sum(got["comp"].astype(np.float64) * st.components, axis=1)
    assert cos.min() > 1 - 1e-5, cos          # signed: the agreed svd_flip signs match sklearn's
    assert np.allclose(got["sv"], st.singular_values, rtol=1e-5)


def test_pick_global_signs_first_max_rule():
    from ganspace_b200 import _native
    allmax = torch.tensor([[[2.0, -2.0], [1.0, 1.0], [3.0, 3.0]],
                           [[2.0, 2.0], [5.0, -5.0], [3.0, -3.0]]])       # [W=2, c=3, 2]
    assert _native.pick_global_signs(allmax).tolist() == [-1.0, -1.0, 1.0]   # ties -> lowest shard


# ---- the estimator's sharded host logic (batch_buffer / partial_fit_inplace / gathers) with a CPU stand-in engine ---------
def _estimator_worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "tests"))
    import fakes                                                       # CPU stand-in engine with BigIPCA's interface
    from ganspace_b200 import _native, estimators
    _native.BigIPCA = fakes.FakeBig
    estimators.DeviceIncrementalPCA.SMALL_D_MAX = 64                  # make d = 256 a "large-d" problem
    d, nb, c, k = 256, 60, 5, 3
    est = estimators.get_estimator("ipca", c, 1.0, device="cpu")
    tr = est.transformer
    tr.enable_feature_sharding(rank, world)
    q = nb // world
    for X in _big_batches(d, nb, k):
        stage = tr.batch_buffer(nb, d, torch.device("cpu"))           # this rank's rows, all features
        assert stage.shape == (q, d)
        stage.copy_(torch.from_numpy(X[rank * q:(rank + 1) * q]))
        assert est.fit_partial_inplace(nb)
    comp, stdev, ratio = est.get_components()
    rows = tr.last_batch_rows(nb)
    bmean = tr.last_batch_mean()
    assert comp.shape == (c, d) and tr.mean_.shape == (d,) and rows.shape == (nb, d) and bmean.shape == (d,)
    Xlast = _big_batches(d, nb, k)[-1].astype(np.float64)
    assert np.allclose(rows.numpy(), Xlast - Xlast.mean(0), atol=1e-5) and np.allclose(bmean.numpy(), Xlast.mean(0))
    if rank == 0:
        np.savez(out_path, comp=comp, stdev=stdev, ratio=ratio, mean=tr.mean_, n=int(tr.n_samples_seen_))
    dist.destroy_process_group()


def test_estimator_feature_sharding_host_logic(tmp_path, oracle):
    world = 2
    out_path = str(tmp_path / "est.npz")
    mp.spawn(_estimator_worker, args=(world, _free_port(), out_path), nprocs=world, join=True)
    got = np.load(out_path)
    st = oracle.IPCAState(5)
    for X in _big_batches(256, 60, 3):
        oracle.ipca_partial_fit(st, X)
    assert np.sum(got["comp"].astype(np.float64) * st.components, axis=1).min() > 1 - 1e-5
    assert np.allclose(got["stdev"], np.sqrt(st.explained_variance), rtol=1e-5)
    assert np.allclose(got["ratio"], st.explained_variance_ratio, rtol=1e-5)
    assert np.allclose(got["mean"], st.mean, atol=1e-6) and int(got["n"]) == 180

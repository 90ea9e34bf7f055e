"""world_size-2 gloo test (CPU) of the multi-GPU host logic: group ownership, the slot layout of the
single statistics all-reduce, and the ordered replay -- with the oracle standing in for the device
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
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ganspace_b200 import plan
    from oracle import ganspace_oracle as orc
    d, c = 48, 6
    pl = plan.make_plan(n=10_000, B=1_000, components=c)      # N=10000, NB=2000, K=5
    data = _groups(5, pl.K, pl.NB, d)
    slots = torch.zeros((pl.K, plan.slot_width(d)), dtype=torch.float64)
    touched = plan.groups_to_process(pl, rank, world)
    assert pl.K - 1 in touched
    for k in touched:
        if plan.owner(k, world, pl.K) != rank:
            continue
        n_b, m, G = orc.batch_stats(data[k])
        slots[k, :d * d] = torch.from_numpy(G.reshape(-1))
        slots[k, d * d:] = torch.from_numpy(m)
    dist.all_reduce(slots)
    st = orc.IPCAState(c)
    plan.replay(pl, slots.numpy(), d, lambda nb, m, g: orc.ipca_gram_step(st, nb, np.array(m), np.array(g)))
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

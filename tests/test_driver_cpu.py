"""Host logic of decomposition.compute_arrays for a conv-feature-map layer, end to end on the CPU: CPU stand-ins replace
the device layer (tests/fakes.py), the ORACLE's restatement of the reference's compute() is the expected result.
Single process and world_size-2 gloo (row-parallel generation, feature-sharded chain, sample-sharded regression)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
N, B, C_COMP = 4000, 500, 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


SMALL = dict(C=3, H=4, W=4)          # d = 48: the small-d engine, hook read-back, statistics exchange + replay


def _run_driver(small=False):
    sys.path.insert(0, str(ROOT / "tests"))
    sys.path.insert(0, str(ROOT))
    import fakes
    from ganspace_b200 import _native, decomposition, estimators
    from ganspace_b200.config import Config
    from ganspace_b200.netdissect.nethook import InstrumentedModel
    fakes.install(_native, estimators)
    model = fakes.FakeFeatureModel(**(SMALL if small else {}))
    inst = InstrumentedModel(model)
    inst.retain_layer("feat")
    cfg = Config(model="Fake", layer="feat", output_class="none", components=C_COMP, n=N, batch_size=B, use_w=False,
                 estimator="ipca")
    return decomposition.compute_arrays(cfg, inst), model


def _expected(oracle, model):
    sample = lambda seed, b: np.random.RandomState(seed).standard_normal(model.latent * b).reshape(b, model.latent).astype(np.float32)
    activate = lambda z: model.act_nchw_flat(torch.from_numpy(np.ascontiguousarray(z, np.float32))).numpy()
    d = model.C * model.H * model.W
    return oracle.compute_path(sample, activate, model.latent, d, N, B, C_COMP, False)


def _check(out, ref, model, oracle):
    assert out["act_comp"].shape == (C_COMP, 1, model.C, model.H, model.W) and out["act_mean"].shape == (1, model.C, model.H, model.W)
    assert out["lat_comp"].shape == (C_COMP, 1, model.latent) and out["lat_mean"].shape == (1, model.latent)
    cmp = oracle.compare_npz(out, ref)
    assert cmp["min_signed_cos"] > 1 - 1e-5 and cmp["max_abs_dvar_ratio"] < 1e-5 and cmp["min_lat_signed_cos"] > 1 - 1e-5, cmp
    assert cmp["act_mean_rel"] < 1e-5 and cmp["act_stdev_rel"] < 1e-4 and cmp["random_stdevs_rel"] < 1e-4, cmp
    assert cmp["lat_mean_rel"] < 1e-5 and np.array_equal(out["lat_stdev"], np.ones(C_COMP, np.float32))


_PATCHED = ("BigIPCA", "IPCAChain", "batch_stats", "batch_stats_multi", "LinregAccumulator", "project_std", "require_cuda")


def test_small_d_driver_single_process(oracle, monkeypatch):
    """d = 48: generic layer through the retain hook, Gram-form chain, hook-based regression."""
    from ganspace_b200 import _native
    for name in _PATCHED:
        monkeypatch.setattr(_native, name, getattr(_native, name))            # restored after the test
    out, model = _run_driver(small=True)
    _check(out, _expected(oracle, model), model, oracle)


def test_large_d_driver_single_process(oracle, monkeypatch):
    """d = 2048 > 1024 -> large-d path: producer rows in NHWC order, permutation back to NCHW, native regression."""
    from ganspace_b200 import _native
    for name in _PATCHED:
        monkeypatch.setattr(_native, name, getattr(_native, name))            # restored after the test
    out, model = _run_driver()
    _check(out, _expected(oracle, model), model, oracle)


def _worker(rank, world, port, out_path, small):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out, _ = _run_driver(small)
    if rank == 0:
        np.savez(out_path, **out)
    dist.barrier()
    dist.destroy_process_group()


def test_large_d_driver_two_ranks_gloo(oracle, tmp_path):
    sys.path.insert(0, str(ROOT / "tests"))
    import fakes
    out_path = str(tmp_path / "driver2.npz")
    mp.spawn(_worker, args=(2, _free_port(), out_path, False), nprocs=2, join=True)
    with np.load(out_path) as data:
        out = {k: data[k] for k in data.files}
    model = fakes.FakeFeatureModel()
    _check(out, _expected(oracle, model), model, oracle)


def test_small_d_driver_two_ranks_gloo(oracle, tmp_path):
    """Round-robin group ownership, per-round all-gather of the statistics, ordered merge, sharded regression -- at the driver level."""
    sys.path.insert(0, str(ROOT / "tests"))
    import fakes
    out_path = str(tmp_path / "driver2s.npz")
    mp.spawn(_worker, args=(2, _free_port(), out_path, True), nprocs=2, join=True)
    with np.load(out_path) as data:
        out = {k: data[k] for k in data.files}
    model = fakes.FakeFeatureModel(**SMALL)
    _check(out, _expected(oracle, model), model, oracle)


# ---- W space (samples are the latents): the bench configuration's control flow -----------------------------------------
def _run_style(use_w):
    sys.path.insert(0, str(ROOT / "tests"))
    sys.path.insert(0, str(ROOT))
    import fakes
    from ganspace_b200 import _native, decomposition, estimators
    from ganspace_b200.config import Config
    from ganspace_b200.netdissect.nethook import InstrumentedModel
    fakes.install(_native, estimators)
    model = fakes.FakeStyleModel()
    inst = InstrumentedModel(model)
    inst.retain_layer("style")
    cfg = Config(model="FakeStyleGAN", layer="style", output_class="none", components=C_COMP, n=10_000, batch_size=1_000,
                 use_w=use_w, estimator="ipca")
    return decomposition.compute_arrays(cfg, inst), model


def _expected_style(oracle, model, use_w):
    normals = lambda seed, b: np.random.RandomState(seed).standard_normal(model.latent * b).reshape(b, model.latent).astype(np.float32)
    mapping = lambda z: model.mapping(torch.from_numpy(np.ascontiguousarray(z, np.float32))).numpy()
    if use_w:
        return oracle.compute_path(lambda s, b: mapping(normals(s, b)), None, model.latent, model.latent, 10_000, 1_000, C_COMP, True,
                                   use_w=True)
    return oracle.compute_path(normals, mapping, model.latent, model.latent, 10_000, 1_000, C_COMP, False)


def _check_style(out, ref, oracle, use_w):
    cmp = oracle.compare_npz(out, ref)
    assert cmp["min_signed_cos"] > 1 - 1e-6 and cmp["max_abs_dvar_ratio"] < 1e-6 and cmp["min_lat_signed_cos"] > 1 - 1e-5, cmp
    assert cmp["act_mean_rel"] < 1e-5 and cmp["act_stdev_rel"] < 1e-5 and cmp["random_stdevs_rel"] < 1e-4, cmp
    assert cmp["lat_stdev_rel"] < 1e-4 and cmp["lat_mean_rel"] < 1e-5, cmp


def _style_worker(rank, world, port, out_path, use_w):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out, _ = _run_style(use_w)
    if rank == 0:
        np.savez(out_path, **out)
    dist.barrier()
    dist.destroy_process_group()


def test_style_layer_driver_w_and_z_space(oracle, monkeypatch, tmp_path):
    from ganspace_b200 import _native
    sys.path.insert(0, str(ROOT / "tests"))
    import fakes
    for name in _PATCHED:
        monkeypatch.setattr(_native, name, getattr(_native, name))
    for use_w in (True, False):
        out, model = _run_style(use_w)
        _check_style(out, _expected_style(oracle, model, use_w), oracle, use_w)
    # W space on two ranks: the bench configuration's statistics exchange + replay, lat_stdev epilogue
    out_path = str(tmp_path / "style2.npz")
    mp.spawn(_style_worker, args=(2, _free_port(), out_path, True), nprocs=2, join=True)
    with np.load(out_path) as data:
        out = {k: data[k] for k in data.files}
    _check_style(out, _expected_style(oracle, fakes.FakeStyleModel(), True), oracle, True)


def test_style_layer_driver_four_ranks_gloo(oracle, tmp_path):
    """W space on four ranks: one ragged round (16 owned slots for the plan's 10 groups), owners with fewer groups than others,
    the final group's buffer on a rank that does not own it -- the N > 2 control flow of bench.py's scaling run."""
    sys.path.insert(0, str(ROOT / "tests"))
    import fakes
    for use_w in (True, False):            # False: Z space, the regression pass sharded over the ranks as well
        out_path = str(tmp_path / f"style4_{int(use_w)}.npz")
        mp.spawn(_style_worker, args=(4, _free_port(), out_path, use_w), nprocs=4, join=True)
        with np.load(out_path) as data:
            out = {k: data[k] for k in data.files}
        _check_style(out, _expected_style(oracle, fakes.FakeStyleModel(), use_w), oracle, use_w)

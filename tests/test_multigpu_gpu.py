"""Two-GPU runs over NCCL against the reference fixtures: a conv-layer decomposition (feature-sharded large-d chain) and the
small-d path (round-robin group ownership, per-round all-gather of the statistics, replicated merge chain) for config 1
(W space) and the Z-space + regression configuration.  Skipped on single-GPU boxes; the protocols themselves are covered on
CPU by tests/test_distributed_gloo.py and tests/test_driver_cpu.py."""
import os
import socket
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from ganspace_b200.config import Config
    from ganspace_b200.decomposition import get_or_compute
    from ganspace_b200.models import StyleGAN2, get_instrumented_model
    dev = torch.device("cuda", rank)
    model = StyleGAN2(dev, "ffhq", random_init=1234)
    inst = get_instrumented_model("StyleGAN2", "ffhq", "convs.1", dev, model=model, use_w=False)
    cfg = Config(model="StyleGAN2", layer="convs.1", output_class="ffhq", components=8, n=4000, batch_size=500,
                 use_w=False, estimator="ipca")
    with tempfile.TemporaryDirectory() as tmp:
        path = get_or_compute(cfg, inst, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp), force_recompute=True)
        if rank == 0:
            with np.load(path) as data:
                np.savez(out_path, **{k: data[k] for k in data.files})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_conv_layer_two_gpus_vs_reference_golden(golden, oracle, tmp_path):
    import torch.multiprocessing as mp
    out_path = str(tmp_path / "two_gpu.npz")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    with np.load(out_path) as data:
        out = {k: data[k] for k in data.files}
    g = golden("c5s_stylegan2_ffhq_convs1_z_n4000_b500_c8.npz")
    cmp = oracle.compare_npz(out, g)
    assert cmp["min_signed_cos"] >= 0.999 and cmp["max_abs_dvar_ratio"] <= 1e-3 and cmp["min_lat_signed_cos"] >= 0.999, cmp
    assert cmp["act_mean_rel"] < 1e-3 and cmp["act_stdev_rel"] < 1e-3 and cmp["random_stdevs_rel"] < 1e-3, cmp


def _style_worker(rank, world, port, out_path, kw):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from ganspace_b200.config import Config
    from ganspace_b200.decomposition import get_or_compute
    from ganspace_b200.models import StyleGAN2, get_instrumented_model
    dev = torch.device("cuda", rank)
    model = StyleGAN2(dev, kw["output_class"], random_init=1234)
    inst = get_instrumented_model("StyleGAN2", kw["output_class"], "style", dev, model=model, use_w=kw["use_w"])
    cfg = Config(model="StyleGAN2", layer="style", estimator="ipca", **kw)
    with tempfile.TemporaryDirectory() as tmp:
        path = get_or_compute(cfg, inst, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp), force_recompute=True)
        if rank == 0:
            with np.load(path) as data:
                np.savez(out_path, **{k: data[k] for k in data.files})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("fixture,kw", [
    ("c1_stylegan2_ffhq_style_w_n10000_b1000_c32.npz",
     dict(output_class="ffhq", use_w=True, n=10_000, batch_size=1_000, components=32)),
    ("c3s_stylegan2_car_style_z_n4000_b1000_c16.npz",
     dict(output_class="car", use_w=False, n=4_000, batch_size=1_000, components=16)),
    ("w_ragged_n5000_b700_c20_seed7.npz",
     dict(output_class="ffhq", use_w=True, n=5_000, batch_size=700, components=20, seed=7)),
])
def test_small_d_two_gpus_vs_reference_golden(golden, oracle, tmp_path, fixture, kw):
    import torch.multiprocessing as mp
    out_path = str(tmp_path / "two_gpu_style.npz")
    mp.spawn(_style_worker, args=(2, _free_port(), out_path, kw), nprocs=2, join=True)
    with np.load(out_path) as data:
        out = {k: data[k] for k in data.files}
    cmp = oracle.compare_npz(out, golden(fixture))
    assert cmp["min_signed_cos"] >= 0.999 and cmp["max_abs_dvar_ratio"] <= 1e-3 and cmp["min_lat_signed_cos"] >= 0.999, cmp
    assert cmp["act_mean_rel"] < 1e-4 and cmp["act_stdev_rel"] < 1e-4 and cmp["random_stdevs_rel"] < 1e-4, cmp
    assert cmp["lat_stdev_rel"] < 1e-4, cmp

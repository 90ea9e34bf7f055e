"""BigGAN-512 generator.gen_z (BASELINE.json config 4): sampler, layer and end-to-end parity on the GPU."""
import tempfile
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from ganspace_b200.models.biggan import BigGAN
    return BigGAN(torch.device("cuda:0"), 512, "husky", random_init=4321)


def test_random_init_and_sampler_match_reference(model, golden):
    g = golden("biggan_known_answers.npz")
    gz = model.model.generator.gen_z
    assert np.array_equal(gz.weight_orig[:4, :6].detach().cpu().numpy(), g["weight_orig_head"])
    assert np.isclose(float(gz.weight_orig.detach().double().sum()), float(g["weight_orig_sum"]))
    assert np.array_equal(gz.weight_u[:8].cpu().numpy(), g["u_head"])
    assert np.array_equal(model.model.embeddings.weight[:4, :6].detach().cpu().numpy(), g["emb_head"])
    z = model.sample_latent(8, seed=11).cpu().numpy()
    assert z.shape == (8, 128) and np.max(np.abs(z - g["z"])) < 1e-6
    assert np.max(np.abs(z)) <= 2.0
    assert np.max(np.abs(model.sample_latent(4, seed=5).cpu().numpy() - g["trunc_seed5"])) < 1e-6


def test_partial_forward_gen_z_matches_reference(model, golden):
    from ganspace_b200.models import get_instrumented_model
    g = golden("biggan_known_answers.npz")
    inst = get_instrumented_model("BigGAN-512", "husky", "generator.gen_z", torch.device("cuda:0"), model=model)
    assert tuple(inst.feature_shape["generator.gen_z"]) == (1, 32768)
    model.partial_forward(torch.tensor(g["z"]).cuda(), "generator.gen_z")
    act = inst.retained_features()["generator.gen_z"].cpu().numpy()
    scale = np.max(np.abs(g["act"]))            # random-init gen_z activations are O(500)
    assert act.shape == (8, 32768) and np.max(np.abs(act - g["act"])) < 1e-5 * scale
    # the thin factorisation reproduces the layer:  act = (z R^T) Q^T + offset
    aff = model.affine_layer("generator.gen_z")
    y = aff.coords(torch.tensor(g["z"]).cuda()).double()
    rec = (y @ aff.Q.T + aff.offset).cpu().numpy()
    assert np.max(np.abs(rec - g["act"])) < 1e-5 * scale
    inst.close()


def test_config4_small_vs_reference_golden(model, golden, oracle):
    from ganspace_b200.config import Config
    from ganspace_b200.decomposition import get_or_compute
    from ganspace_b200.models import get_instrumented_model
    g = golden("c4s_biggan512_husky_genz_n4000_b1000_c16.npz")
    inst = get_instrumented_model("BigGAN-512", "husky", "generator.gen_z", torch.device("cuda:0"), model=model)
    cfg = Config(model="BigGAN-512", layer="generator.gen_z", output_class="husky", components=16, n=4_000,
                 batch_size=1_000, estimator="ipca")
    with tempfile.TemporaryDirectory() as tmp:
        path = get_or_compute(cfg, inst, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp), force_recompute=True)
        assert path.name == str(g["dump_name"])
        with np.load(path) as data:
            out = {k: data[k] for k in data.files}
    for k in ("act_comp", "act_mean", "act_stdev", "lat_comp", "lat_mean", "lat_stdev", "var_ratio", "random_stdevs"):
        assert out[k].shape == g[k].shape and out[k].dtype == np.float32, k
    cmp = oracle.compare_npz(out, g)
    assert cmp["min_signed_cos"] >= 0.999 and cmp["min_lat_signed_cos"] >= 0.999, cmp     # BASELINE.json tolerance
    assert cmp["max_abs_dvar_ratio"] <= 1e-3, cmp
    for k in ("act_mean_rel", "act_stdev_rel", "random_stdevs_rel"):
        assert cmp[k] < 1e-4, (k, cmp)
    assert np.array_equal(out["lat_stdev"], np.ones(16, np.float32))
    inst.close()


def test_gen_z_tensor_core_path_matches_fp64(model):
    """Batches of >= 128 rows take the tcgen05 kernel (fp16 hi/lo split operands, bias epilogue, TMA-stored fp32 rows);
    against an fp64 evaluation of the same affine map and against the fp32 FMA kernel (the < 128-row path)."""
    from ganspace_b200 import _native as nat
    g = model.model.generator.gen_z
    z = model.sample_latent(300, seed=3)
    cond = torch.cat((z, model._embed().unsqueeze(0).expand(300, -1)), dim=1).contiguous()
    w, b = g.effective_weight(), g.bias.detach()
    ref = (cond.double() @ w.double().T + b.double()).cpu().numpy()
    tc = nat.linear(cond, w, b, bounded=True).cpu().numpy()
    fma = nat.linear(cond, w, b, bounded=False).cpu().numpy()
    scale = np.max(np.abs(ref))
    assert tc.shape == (300, 32768)
    assert np.max(np.abs(tc - ref)) < 2e-5 * scale, np.max(np.abs(tc - ref)) / scale      # tensor-core truncation: ~1e-5 relative
    assert np.max(np.abs(fma - ref)) < 1e-5 * scale
    # a ragged row count (not a multiple of the 128-row tile) and the module's own forward
    out = g(cond[:257]).cpu().numpy()
    assert np.max(np.abs(out - ref[:257])) < 2e-5 * scale


def test_config4_general_large_d_engine_cross_check(golden, oracle, monkeypatch):
    """GANSPACE_B200_BIGGAN_AFFINE=0: no low-rank shortcut -- the [N, 32768] activations are materialised (tensor-core gen_z)
    and run through the general large-d engine (csrc/bigd.cu); same reference fixture as the shortcut."""
    monkeypatch.setenv("GANSPACE_B200_BIGGAN_AFFINE", "0")
    from ganspace_b200.config import Config
    from ganspace_b200.decomposition import get_or_compute
    from ganspace_b200.models import get_instrumented_model
    from ganspace_b200.models.biggan import BigGAN
    m = BigGAN(torch.device("cuda:0"), 512, "husky", random_init=4321)
    assert m.affine_layer("generator.gen_z") is None
    g = golden("c4s_biggan512_husky_genz_n4000_b1000_c16.npz")
    inst = get_instrumented_model("BigGAN-512", "husky", "generator.gen_z", torch.device("cuda:0"), model=m)
    cfg = Config(model="BigGAN-512", layer="generator.gen_z", output_class="husky", components=16, n=4_000,
                 batch_size=1_000, estimator="ipca")
    with tempfile.TemporaryDirectory() as tmp:
        path = get_or_compute(cfg, inst, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp), force_recompute=True)
        with np.load(path) as data:
            out = {k: data[k] for k in data.files}
    cmp = oracle.compare_npz(out, g)
    assert cmp["min_signed_cos"] >= 0.999 and cmp["min_lat_signed_cos"] >= 0.999 and cmp["max_abs_dvar_ratio"] <= 1e-3, cmp
    for k in ("act_mean_rel", "act_stdev_rel", "random_stdevs_rel"):
        assert cmp[k] < 1e-3, (k, cmp)
    inst.close()


def test_config4_n100k_vs_oracle(model, oracle):
    """Config 4 at N = 100k (50 partial_fit groups of 2000) against the oracle's restatement of the reference."""
    from ganspace_b200.config import Config
    from ganspace_b200.decomposition import compute_arrays
    from ganspace_b200.models import get_instrumented_model
    params = oracle.biggan_genz_random_init(4321)
    ref = oracle.compute_biggan_genz(params, 100_000, 2_000, 80, ipca="small")
    inst = get_instrumented_model("BigGAN-512", "husky", "generator.gen_z", torch.device("cuda:0"), model=model)
    cfg = Config(model="BigGAN-512", layer="generator.gen_z", output_class="husky", components=80, n=100_000,
                 batch_size=2_000, estimator="ipca")
    out = compute_arrays(cfg, inst)
    cmp = oracle.compare_npz(out, ref)
    assert cmp["min_signed_cos"] >= 0.999 and cmp["min_lat_signed_cos"] >= 0.999 and cmp["max_abs_dvar_ratio"] <= 1e-3, cmp
    inst.close()

"""Pins the oracle (numpy restatement) to the golden fixtures written by the unmodified reference
(oracle/gen_golden.py).  Tolerances: fp32 summation-order noise only."""
import numpy as np
import pytest


def test_mapping_init_and_forward(oracle, golden, mapping_weights):
    g = golden("mapping_known_answers.npz")
    ws, bs = mapping_weights
    assert np.allclose([float(w.astype(np.float64).sum()) for w in ws], g["style_weight_sums"])
    assert np.array_equal(ws[0][:4, :8], g["style1_weight_head"])
    out = oracle.mapping_forward(g["z"], ws, bs)
    assert np.max(np.abs(out - g["w"])) < 2e-5


@pytest.mark.parametrize("form", ["svd", "gram"])
def test_ipca_chain_restatement_vs_sklearn(oracle, golden, form):
    g = golden("ipca_chain_d96_c12.npz")
    st = oracle.IPCAState(12)
    for k, X in enumerate(g["X"]):
        if form == "svd":
            oracle.ipca_partial_fit(st, X)
        else:
            oracle.ipca_gram_step(st, *oracle.batch_stats(X))
        cos = np.sum(st.components * g[f"comp_{k}"], axis=1)
        assert np.min(cos) > 1 - 1e-6
        assert np.allclose(st.singular_values, g[f"sv_{k}"], rtol=1e-5)
        assert np.allclose(np.sqrt(st.explained_variance), g[f"stdev_{k}"], rtol=1e-5)
        assert np.allclose(st.explained_variance_ratio, g[f"ratio_{k}"], atol=1e-7)
        assert np.allclose(st.mean, g[f"mean_{k}"], rtol=1e-10, atol=1e-10)
        assert np.allclose(st.var, g[f"var_{k}"], rtol=1e-6)


CASES = [
    ("c1_stylegan2_ffhq_style_w_n10000_b1000_c32.npz", dict(n=10_000, B=1_000, c=32, use_w=True)),
    ("w_ragged_n5000_b700_c20_seed7.npz", dict(n=5_000, B=700, c=20, use_w=True, seed=7)),
    ("c3s_stylegan2_car_style_z_n4000_b1000_c16.npz", dict(n=4_000, B=1_000, c=16, use_w=False)),
]


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("form", ["svd", "gram"])
def test_compute_restatement_vs_reference(oracle, golden, mapping_weights, name, kw, form):
    g = golden(name)
    ws, bs = mapping_weights
    out = oracle.compute_stylegan2_style(ws, bs, kw["n"], kw["B"], kw["c"], kw["use_w"], seed=kw.get("seed"),
                                         ipca=form)
    cmp = oracle.compare_npz(out, g)
    assert cmp["min_signed_cos"] > 1 - 1e-6 and cmp["min_lat_signed_cos"] > 1 - 1e-6, cmp
    assert cmp["max_abs_dvar_ratio"] < 1e-6, cmp
    for k in ("act_mean_rel", "act_stdev_rel", "lat_mean_rel", "lat_stdev_rel", "random_stdevs_rel"):
        assert cmp[k] < 1e-5, (k, cmp)
    for k in out:
        assert out[k].shape == g[k].shape and out[k].dtype == g[k].dtype, k


def test_biggan_genz_known_answers(oracle, golden):
    g = golden("biggan_known_answers.npz")
    p = oracle.biggan_genz_random_init(4321)
    assert np.isclose(float(p["weight_orig"].astype(np.float64).sum()), float(g["weight_orig_sum"]))
    assert np.array_equal(p["weight_orig"][:4, :6], g["weight_orig_head"])
    assert np.array_equal(p["bias"][:8], g["bias_head"]) and np.array_equal(p["emb"][:4, :6], g["emb_head"])
    assert np.array_equal(p["u"][:8], g["u_head"]) and np.array_equal(p["v"][:8], g["v_head"])
    assert np.max(np.abs(oracle.truncated_noise_sample(5, 4) - g["trunc_seed5"])) < 1e-6
    assert np.max(np.abs(oracle.truncated_noise_sample(11, 8) - g["z"])) < 1e-6
    act = oracle.genz_forward(g["z"], p)
    assert np.max(np.abs(act - g["act"])) < 1e-5


def test_biggan_genz_compute_restatement_vs_reference(oracle, golden):
    g = golden("c4s_biggan512_husky_genz_n4000_b1000_c16.npz")
    out = oracle.compute_biggan_genz(oracle.biggan_genz_random_init(4321), 4_000, 1_000, 16)
    cmp = oracle.compare_npz(out, g)
    assert cmp["min_signed_cos"] > 1 - 1e-5 and cmp["min_lat_signed_cos"] > 1 - 1e-4, cmp
    assert cmp["max_abs_dvar_ratio"] < 1e-5, cmp
    for k in ("act_mean_rel", "act_stdev_rel", "lat_stdev_rel", "random_stdevs_rel"):
        assert cmp[k] < 1e-4, (k, cmp)
    for k in out:
        assert out[k].shape == g[k].shape and out[k].dtype == g[k].dtype, k


# ---- StyleGAN2 synthesis -------------------------------------------------------------
def _perturbed_params(oracle, upto):
    p = oracle.synthesis_random_init(1234, 1024, upto)
    for i, name in enumerate(oracle.synthesis_layer_names(upto)):
        p["layers"][name]["noise_weight"] = np.float32(0.1 * (i + 1))
        p["layers"][name]["act_bias"] = (0.1 * np.sin(np.arange(512, dtype=np.float32) + i)).astype(np.float32)
    return p


def test_synthesis_init_and_known_answers(oracle, golden, mapping_weights):
    """Oracle restatement of the StyledConv chain vs the unmodified reference's partial_forward outputs."""
    g = golden("synthesis_known_answers.npz")
    p = _perturbed_params(oracle, "convs.4")
    assert np.isclose(float(p["const"].astype(np.float64).sum()), float(g["const_sum"]), rtol=1e-12)
    for name in oracle.synthesis_layer_names("convs.4"):
        L = p["layers"][name]
        sums = [float(L["weight"].astype(np.float64).sum()), float(L["mod_weight"].astype(np.float64).sum())]
        assert np.allclose(sums, g[f"wsum_{name.replace('.', '_')}"], rtol=1e-12), name
    noises = oracle.fixed_noise(0, 1024)
    assert np.array_equal(np.stack([n.reshape(-1)[:4] for n in noises[:6]]), g["noise_heads"])
    w = oracle.mapping_forward(g["z"], *mapping_weights)
    assert np.abs(w - g["w"]).max() < 1e-6
    for layer, keep in (("conv1", 4), ("convs.1", 4), ("convs.2", 2)):
        ref = g[f"act_{layer.replace('.', '_')}"]
        act = oracle.synthesis_forward(g["w"][:keep], p, noises, layer)
        assert act.shape == ref.shape
        assert np.abs(act - ref).max() <= 2e-5 * np.abs(ref).max(), layer
    act = oracle.synthesis_forward(g["w"][:1], p, noises, "convs.4")
    assert np.abs(act[:, ::4, ::2, ::2] - g["act_convs_4_sub"]).max() <= 2e-5 * np.abs(act).max()
    sums = np.array([act.astype(np.float64).sum(), (act.astype(np.float64) ** 2).sum()])
    assert np.allclose(sums, g["sum_convs_4"], rtol=1e-5)


def test_synthesis_tap_form_equals_reference_form(oracle):
    """The contraction-per-tap form the CUDA kernels implement == the reference's per-sample grouped conv."""
    p = _perturbed_params(oracle, "convs.2")
    noises = oracle.fixed_noise(0, 1024)
    w = np.random.RandomState(3).standard_normal((3, 512)).astype(np.float32)
    a = oracle.synthesis_forward(w, p, noises, "convs.2")
    b = oracle.synthesis_forward(w, p, noises, "convs.2", form="taps")
    assert np.abs(a - b).max() < 5e-5 * np.abs(a).max()
    # and the shared-weight torch form used for bulk oracle runs (compute_stylegan2_layer)
    c = oracle.synthesis_forward(w, p, noises, "convs.2", form="shared")
    assert np.abs(a - c).max() < 5e-5 * np.abs(a).max()


def test_conv_layer_pca_restatement_vs_reference(oracle, golden, mapping_weights):
    """PCA half of compute() on layer convs.1 (d = 32768): oracle vs the unmodified reference's .npz."""
    g = golden("c5s_stylegan2_ffhq_convs1_z_n4000_b500_c8.npz")
    p = oracle.synthesis_random_init(1234, 1024, "convs.1")
    out = oracle.compute_stylegan2_layer(*mapping_weights, p, "convs.1", 4000, 500, 8, regress=False)
    a = out["act_comp"].reshape(8, -1).astype(np.float64)
    b = g["act_comp"].reshape(8, -1).astype(np.float64)
    assert np.sum(a * b, axis=1).min() > 0.99999
    assert np.abs(out["var_ratio"] - g["var_ratio"]).max() < 1e-6
    assert np.allclose(out["act_stdev"], g["act_stdev"], rtol=1e-4)
    assert np.allclose(out["act_mean"].reshape(-1), g["act_mean"].reshape(-1), atol=1e-5)
    assert np.allclose(out["random_stdevs"], g["random_stdevs"], rtol=1e-4)


def test_small_side_restatement_equals_svd_form(oracle):
    """ipca_partial_fit_small_side (used for d >> rows, config 4 at N = 100k) is the same factorisation as the gesdd form."""
    rng = np.random.RandomState(3)
    basis = rng.standard_normal((600, 24)) * (0.8 ** np.arange(24))[None, :]
    a, b = oracle.IPCAState(6), oracle.IPCAState(6)
    for k in range(3):
        X = (rng.standard_normal((50, 24)) @ basis.T + 0.05 * rng.standard_normal((50, 600)) + 1.5).astype(np.float32)
        oracle.ipca_partial_fit(a, X)
        oracle.ipca_partial_fit_small_side(b, X)
    assert np.sum(a.components * b.components, axis=1).min() > 1 - 1e-6
    assert np.allclose(a.singular_values, b.singular_values, rtol=1e-5)
    assert np.allclose(a.explained_variance_ratio, b.explained_variance_ratio, rtol=1e-5) and np.allclose(a.mean, b.mean)


def test_render_restatement_vs_reference(oracle, golden, mapping_weights):
    """Row f2: the oracle's Generator.forward (all 17 StyledConv layers, ToRGB / skip up-sampling, per-layer latents) against
    the unmodified reference's known answers (oracle/gen_golden_r2.py G11 / G11b): activations of convs.5 .. convs.15, the
    running skip image after every ToRGB, the final image and a style-mixed image."""
    from conftest import GOLDEN
    if not (GOLDEN / "synthesis_deep_known_answers.npz").exists():
        pytest.skip("fixture synthesis_deep_known_answers.npz not generated")
    g = golden("synthesis_deep_known_answers.npz")
    p = oracle.synthesis_random_init(1234, 1024, "convs.15")
    for i, name in enumerate(str(x) for x in g["conv_names"]):
        co = p["layers"][name]["weight"].shape[0]
        p["layers"][name]["noise_weight"] = np.float32(0.1 * (i + 1))
        p["layers"][name]["act_bias"] = (0.1 * np.sin(np.arange(co, dtype=np.float32) + i)).astype(np.float32)
    assert len(p["to_rgbs"]) == len(g["rgb_names"]) == 9
    for i, R in enumerate(p["to_rgbs"]):
        R["bias"] = (0.05 * np.array([1.0, -2.0, 3.0], np.float32) * (i + 1)).astype(np.float32)
    noises = oracle.fixed_noise(0, 1024)
    w = oracle.mapping_forward(g["z"], *mapping_weights)                     # [2, 512]
    n_lat = 18
    layers = [f"convs.{i}" for i in range(5, 16)] + [str(x) for x in g["rgb_names"]]
    img, kept = oracle.render_forward(np.repeat(w[:1, None, :], n_lat, axis=1), p, noises, keep=layers)
    for layer in layers:
        act, key = kept[layer], layer.replace(".", "_")
        assert tuple(act.shape) == tuple(g[f"shape_{key}"]), layer
        step = max(1, act.shape[-1] // 32)
        sub = act[:, ::max(1, act.shape[1] // 16), ::step, ::step]
        ref = g[f"act_{key}_sub"]
        assert np.abs(sub - ref).max() <= 5e-5 * np.abs(ref).max(), (layer, np.abs(sub - ref).max() / np.abs(ref).max())
        s2 = (act.astype(np.float64) ** 2).sum()
        assert abs(s2 - g[f"sum_{key}"][1]) < 1e-4 * g[f"sum_{key}"][1], layer
    ref_img = g["img_sub"][:1]
    assert np.abs(0.5 * (img + 1)[:, :, ::4, ::4] - ref_img).max() <= 5e-5 * np.abs(ref_img - 0.5).max()
    # style mixing: latents 0..7 from z0, 8..17 from z1 (the fixture's forward([z0] * 8 + [z1] * 10))
    wl = np.concatenate([np.repeat(w[:1, None, :], 8, axis=1), np.repeat(w[1:2, None, :], n_lat - 8, axis=1)], axis=1)
    mixed, _ = oracle.render_forward(wl, p, noises)
    assert np.abs(0.5 * (mixed + 1)[:, :, ::4, ::4] - g["mixed_sub"]).max() <= 5e-5 * np.abs(g["mixed_sub"] - 0.5).max()

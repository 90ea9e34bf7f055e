"""BASELINE.json sizes: size-independent properties at N=1e6 (config 2) and oracle parity at N=300k."""
import tempfile
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(n, b, c, mapping=None, monkeypatch=None):
    from ganspace_b200.config import Config
    from ganspace_b200.decomposition import compute_arrays
    from ganspace_b200.models import get_instrumented_model, StyleGAN2
    dev = torch.device("cuda:0")
    model = StyleGAN2(dev, "ffhq", random_init=1234)
    inst = get_instrumented_model("StyleGAN2", "ffhq", "style", dev, model=model, use_w=True)
    cfg = Config(model="StyleGAN2", layer="style", output_class="ffhq", components=c, n=n, batch_size=b,
                 use_w=True, estimator="ipca")
    out = compute_arrays(cfg, inst)
    inst.close()
    return out


def test_config2_full_size_properties(monkeypatch):
    out = _run(1_000_000, 10_000, 80)
    comp = out["act_comp"].reshape(80, 512).astype(np.float64)
    assert np.max(np.abs(comp @ comp.T - np.eye(80))) < 1e-5              # orthonormal basis
    assert np.all(np.diff(out["act_stdev"]) <= 0) and np.all(out["act_stdev"] > 0)   # sorted spectrum
    assert 0.5 < float(out["var_ratio"].sum()) < 1.0 and np.all(np.diff(out["var_ratio"]) <= 0)
    idx = np.argmax(np.abs(comp), axis=1)
    assert np.all(comp[np.arange(80), idx] > 0)                           # svd_flip sign rule
    assert np.array_equal(out["act_comp"], out["lat_comp"]) and np.array_equal(out["act_mean"], out["lat_mean"])
    assert np.all(out["random_stdevs"] < out["act_stdev"][0])             # random directions explain less
    # the run is deterministic, and both mapping paths (tcgen05 fp16x3 / fp32 FMA) give the same directions
    out2 = _run(1_000_000, 10_000, 80)
    for k in out:      # repeatable up to the order of the fp64 atomics in the Gram accumulation
        assert np.allclose(out[k], out2[k], rtol=1e-6, atol=1e-7), k
    monkeypatch.setenv("GANSPACE_B200_MAPPING", "simt")
    out3 = _run(1_000_000, 10_000, 80)
    cos = np.sum(comp * out3["act_comp"].reshape(80, 512), axis=1)
    assert np.min(cos) >= 0.999, np.min(cos)
    assert np.max(np.abs(out["var_ratio"] - out3["var_ratio"])) <= 1e-3


@pytest.mark.parametrize("mapping", ["tc", "simt"])
def test_c80_n300k_vs_oracle(oracle, mapping_weights, monkeypatch, mapping):
    monkeypatch.setenv("GANSPACE_B200_MAPPING", mapping)
    ws, bs = mapping_weights
    ref = oracle.compute_stylegan2_style(ws, bs, 300_000, 10_000, 80, True)
    out = _run(300_000, 10_000, 80)
    cmp = oracle.compare_npz(out, ref)
    assert cmp["min_signed_cos"] >= 0.999 and cmp["max_abs_dvar_ratio"] <= 1e-3, cmp
    assert cmp["act_mean_rel"] < 1e-4 and cmp["lat_stdev_rel"] < 1e-4 and cmp["random_stdevs_rel"] < 1e-4, cmp


def test_config3_n100k_vs_oracle(oracle, mapping_weights):
    """Config 3 (StyleGAN2-car, Z space, layer=style + latent regression) at N = 100k against the oracle."""
    import tempfile
    from types import SimpleNamespace
    from ganspace_b200.config import Config
    from ganspace_b200.decomposition import compute_arrays
    from ganspace_b200.models import get_instrumented_model, StyleGAN2
    ws, bs = mapping_weights
    ref = oracle.compute_stylegan2_style(ws, bs, 100_000, 10_000, 80, False)
    dev = torch.device("cuda:0")
    model = StyleGAN2(dev, "car", random_init=1234)
    inst = get_instrumented_model("StyleGAN2", "car", "style", dev, model=model, use_w=False)
    cfg = Config(model="StyleGAN2", layer="style", output_class="car", components=80, n=100_000, batch_size=10_000,
                 use_w=False, estimator="ipca")
    out = compute_arrays(cfg, inst)
    cmp = oracle.compare_npz(out, ref)
    assert cmp["min_signed_cos"] >= 0.999 and cmp["min_lat_signed_cos"] >= 0.999 and cmp["max_abs_dvar_ratio"] <= 1e-3, cmp
    assert cmp["act_mean_rel"] < 1e-4 and cmp["act_stdev_rel"] < 1e-4 and cmp["random_stdevs_rel"] < 1e-4, cmp
    inst.close()

"""Full synthesis to RGB: every StyledConv of the 1024^2 generator (512 ... 32 channels), the
ToRGB / skip chain, StyleGAN2.forward with one latent, per-layer latents and style mixing -- against known answers written by
the unmodified reference (oracle/gen_golden_r2.py G11), plus the reference's own test invariants (tests/partial_forward_test.py:
partial == full at the hooked layer; tests/layerwise_z_test.py: forward(z) == forward(n_latents * [z])).
convs.10 .. convs.15 are checked against the reference's full forward: its partial_forward never reaches them (substring match on
the layer name, wrappers.py:241-246 -- INTEGRATION.md section 3)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ACT_TOL = 5e-4         # max |diff| / max |ref| after up to 17 fused layers (fp16 hi/lo tensor-core products, ~1e-5 per layer)


def _perturb(model, conv_names, rgb_names):
    mods = dict(model.named_modules())
    for i, name in enumerate(conv_names):
        with torch.no_grad():
            mods[name].noise.weight.fill_(0.1 * (i + 1))
            b = mods[name].activate.bias
            b.copy_((0.1 * torch.sin(torch.arange(b.shape[0], dtype=torch.float32) + i)).to(b.device))
    for i, name in enumerate(rgb_names):
        with torch.no_grad():
            mods[name].bias.copy_((0.05 * torch.tensor([1.0, -2.0, 3.0]).view(1, 3, 1, 1) * (i + 1)).to(mods[name].bias.device))


@pytest.fixture(scope="module")
def deep(golden):
    from conftest import GOLDEN
    if not (GOLDEN / "synthesis_deep_known_answers.npz").exists():
        pytest.skip("fixture synthesis_deep_known_answers.npz not generated")
    return golden("synthesis_deep_known_answers.npz")


@pytest.fixture(scope="module")
def model(deep):
    from ganspace_b200.models import StyleGAN2
    m = StyleGAN2(torch.device("cuda:0"), "ffhq", random_init=1234)
    _perturb(m.model, [str(x) for x in deep["conv_names"]], [str(x) for x in deep["rgb_names"]])
    return m


def _sub(act):
    step = max(1, act.shape[-1] // 32)
    return act[:, ::max(1, act.shape[1] // 16), ::step, ::step]


def test_deep_layers_and_to_rgb_known_answers(deep, model):
    from ganspace_b200.models import get_instrumented_model
    dev = torch.device("cuda:0")
    model.use_z()
    z = torch.tensor(deep["z"]).to(dev)
    layers = [f"convs.{i}" for i in range(5, 16)] + [str(x) for x in deep["rgb_names"]]
    for layer in layers:
        inst = get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=model, use_w=False)
        model.partial_forward(z[:1], layer)
        act = inst.retained_features()[layer].float().cpu().numpy()
        key = layer.replace(".", "_")
        assert tuple(act.shape) == tuple(deep[f"shape_{key}"]), layer
        ref = deep[f"act_{key}_sub"]
        scale = np.abs(ref).max()
        assert np.abs(_sub(act) - ref).max() < ACT_TOL * scale, (layer, np.abs(_sub(act) - ref).max() / scale)
        s1, s2 = act.astype(np.float64).sum(), (act.astype(np.float64) ** 2).sum()
        assert abs(s2 - deep[f"sum_{key}"][1]) < 2e-3 * deep[f"sum_{key}"][1], layer
        assert abs(s1 - deep[f"sum_{key}"][0]) < 2e-3 * np.sqrt(deep[f"sum_{key}"][1] * act.size), layer
        inst.close()
    model.check_numerics()


def test_forward_images_vs_reference(deep, model):
    dev = torch.device("cuda:0")
    model.use_z()
    z = torch.tensor(deep["z"]).to(dev)
    img = model.forward(z).float().cpu().numpy()
    assert img.shape == (2, 3, 1024, 1024)
    ref = deep["img_sub"]
    scale = np.abs(ref - 0.5).max()
    assert np.abs(img[:, :, ::4, ::4] - ref).max() < 1e-3 * scale, np.abs(img[:, :, ::4, ::4] - ref).max() / scale
    assert abs((img.astype(np.float64) ** 2).sum() - deep["img_sum"][1]) < 2e-3 * deep["img_sum"][1]
    # one latent per layer (tests/layerwise_z_test.py:59-69) and style mixing with a per-layer list
    n_lat = model.get_max_latents()
    same = model.forward([z[:1]] * n_lat).float().cpu().numpy()
    assert np.abs(same - img[:1]).max() < 1e-5 * max(1.0, np.abs(img).max())
    mixed = model.forward([z[:1]] * 8 + [z[1:2]] * (n_lat - 8)).float().cpu().numpy()
    refm = deep["mixed_sub"]
    assert np.abs(mixed[:, :, ::4, ::4] - refm).max() < 1e-3 * np.abs(refm - 0.5).max()
    model.check_numerics()


def test_partial_forward_equals_forward_at_hooked_layer(model):
    """The reference's tests/partial_forward_test.py:112-121 invariant on random-init weights, plus latent lists in partial_forward
    (tests/layerwise_z_test.py:51-56) and the batch-size probe without a layer (decomposition.get_max_batch_size)."""
    from ganspace_b200.decomposition import get_max_batch_size
    from ganspace_b200.models import get_instrumented_model
    dev = torch.device("cuda:0")
    model.use_z()
    for layer in ("convs.0", "convs.7", "to_rgbs.2"):
        inst = get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=model, use_w=False)
        z = model.sample_latent(3, seed=5)
        model.partial_forward(z, layer)
        a = inst.retained_features()[layer].clone()
        model.forward(z)
        b = inst.retained_features()[layer].clone()
        assert torch.equal(a, b), layer
        model.partial_forward(model.get_max_latents() * [z], layer)
        c = inst.retained_features()[layer]
        assert torch.allclose(a, c, rtol=0, atol=1e-6 * float(a.abs().max())), layer
        # negative control (partial_forward_test.py:93-98): different latents differ
        model.partial_forward(model.sample_latent(3, seed=6), layer)
        assert not torch.equal(a, inst.retained_features()[layer])
        inst.close()
    inst = get_instrumented_model("StyleGAN2", "ffhq", "convs.0", dev, model=model, use_w=False)
    assert get_max_batch_size(inst, dev, None) >= 2
    # an activation edit cannot be re-fed into the fused chain: loud, not silent
    inst.edit_layer("convs.0", offset=torch.ones(1, 512, 8, 8, device=dev))
    with pytest.raises(NotImplementedError):
        model.forward(model.sample_latent(1, seed=1))
    inst.remove_edits()
    inst.close()

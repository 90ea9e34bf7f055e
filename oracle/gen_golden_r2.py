"""ORACLE -- round-2 golden fixtures (tests/golden/), produced by the UNMODIFIED reference on CPU (oracle/ref_harness.py).

  G9   c5_stylegan2_ffhq_convs4_z_n4000_b500_c4.npz     BASELINE config 5's layer (convs.4, d = 524288), Z space, pure random init;
                                                         act_comp stored as float16 (4.2 MB instead of 8.4).  PCA stage only: the
                                                         reference's regression stage needs > 62 GB at this d (see run())
  G10  c5n_stylegan2_ffhq_convs1_z_noise_n4000_b500_c8.npz   convs.1 with NON-ZERO NoiseInjection weights and activation biases
                                                         (both are 0 at random init; perturbation = gen_golden.perturb_synthesis)

Usage:  python oracle/gen_golden_r2.py [g9] [g10]
"""
import sys
import tempfile
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import ref_harness as rh          # noqa: E402
from oracle.gen_golden import perturb_synthesis  # noqa: E402

OUT = REPO / "tests" / "golden"


def run(layer, n, b, c, perturb=None, stub_regression=False):
    ref = rh.import_reference()
    if stub_regression:
        # the reference's regression stage (decomposition.py:77-139) ran out of memory at d = 524288 on the 62 GB container that
        # generates the fixtures (killed by the kernel at 65 GB resident); the PCA stage fits.  Only that one function is replaced
        # -- at run time, nothing under /root/reference is edited -- and the lat_* arrays of such a fixture are placeholders.
        ref.decomposition.linreg_lstsq = lambda comp, mean, stdev, inst, config: (
            np.zeros((comp.shape[0], inst.model.get_latent_dims()), np.float32), np.zeros((1, inst.model.get_latent_dims()), np.float32))
    dev = torch.device("cpu")
    m = rh.rand_init_stylegan2(ref, dev, "ffhq", 1234)
    if perturb:
        perturb_synthesis(m.model, perturb)
    inst = ref.wrappers.get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=m, use_w=False)
    cfg = ref.Config(model="StyleGAN2", layer=layer, output_class="ffhq", estimator="ipca", use_w=False, n=n, batch_size=b, components=c)
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        path = ref.decomposition.get_or_compute(cfg, inst, force_recompute=True, submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp))
        with np.load(path) as data:
            out = {k: data[k].copy() for k in data.files}
        name = path.name
    print(f"{layer}: {time.time() - t0:.0f} s", flush=True)
    inst.close()
    return out, name


def main():
    which = set(sys.argv[1:]) or {"g9", "g10"}
    if "g10" in which:
        names = ["conv1", "convs.0", "convs.1"]
        out, name = run("convs.1", 4_000, 500, 8, perturb=names)
        np.savez_compressed(OUT / "c5n_stylegan2_ffhq_convs1_z_noise_n4000_b500_c8.npz", dump_name=np.array(name),
                            perturbed=np.array(names), **out)
    if "g9" in which:
        out, name = run("convs.4", 4_000, 500, 4, stub_regression=True)
        out["act_comp_f16"] = out.pop("act_comp").astype(np.float16)
        np.savez_compressed(OUT / "c5_stylegan2_ffhq_convs4_z_n4000_b500_c4.npz", dump_name=np.array(name),
                            lat_placeholder=np.array(1), **out)


def deep_synthesis_golden():
    """G11 (`python oracle/gen_golden_r2.py g11`): the layers after convs.4 and the full forward, from the unmodified reference.
    Perturbed noise weights / activation biases on EVERY StyledConv and non-zero ToRGB biases (all zero at random init).
      * partial_forward known answers at convs.5 .. convs.15 (strided subsample + sum / sum of squares), one latent
      * to_rgb1 / to_rgbs.k outputs (the running skip image) at every resolution, same latent
      * StyleGAN2.forward(z) for two latents (1024^2, stored ::4) and forward(list of 18 per-layer latents) (style mixing)"""
    ref = rh.import_reference()
    dev = torch.device("cpu")
    m = rh.rand_init_stylegan2(ref, dev, "ffhq", 1234)
    conv_names = ["conv1"] + [f"convs.{i}" for i in range(16)]
    perturb_synthesis(m.model, conv_names)
    rgb_names = ["to_rgb1"] + [f"to_rgbs.{i}" for i in range(8)]
    mods = dict(m.model.named_modules())
    with torch.no_grad():
        for i, nme in enumerate(rgb_names):
            mods[nme].bias.copy_(0.05 * torch.tensor([1.0, -2.0, 3.0]).view(1, 3, 1, 1) * (i + 1))
    m.use_z()
    z = m.sample_latent(2, seed=21)
    ka = dict(z=z.numpy(), rgb_names=np.array(rgb_names), conv_names=np.array(conv_names))
    for layer in [f"convs.{i}" for i in range(5, 16)] + rgb_names:
        inst = ref.wrappers.get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=m, use_w=False)
        with torch.no_grad():
            m.partial_forward(z[:1], layer)
        act = inst.retained_features()[layer].numpy()
        key = layer.replace(".", "_")
        step = max(1, act.shape[-1] // 32)
        ka[f"act_{key}_sub"] = act[:, ::max(1, act.shape[1] // 16), ::step, ::step].copy()
        ka[f"sum_{key}"] = np.array([act.astype(np.float64).sum(), (act.astype(np.float64) ** 2).sum()])
        ka[f"shape_{key}"] = np.array(act.shape)
        inst.close()
    with torch.no_grad():
        img = m.forward(z).numpy()                                   # [2, 3, 1024, 1024] in [0, 1]-ish (0.5 * (out + 1))
        n_lat = m.get_max_latents()
        mixed = m.forward([z[:1]] * 8 + [z[1:2]] * (n_lat - 8)).numpy()
        same = m.forward([z[:1]] * n_lat).numpy()
    assert np.abs(same - img[:1]).max() < 1e-4
    ka["img_sub"] = img[:, :, ::4, ::4].copy()
    ka["img_sum"] = np.array([img.astype(np.float64).sum(), (img.astype(np.float64) ** 2).sum()])
    ka["mixed_sub"] = mixed[:, :, ::4, ::4].copy()
    ka["mixed_sum"] = np.array([mixed.astype(np.float64).sum(), (mixed.astype(np.float64) ** 2).sum()])
    np.savez_compressed(OUT / "synthesis_deep_known_answers.npz", **ka)
    print("wrote synthesis_deep_known_answers.npz", {k: v.shape for k, v in ka.items() if k.endswith("_sub")})


def deep_synthesis_two_digit_layers():
    """G11b (`python oracle/gen_golden_r2.py g11b`): convs.10 .. convs.15 again, from StyleGAN2.forward instead of partial_forward.
    The reference's partial_forward matches layer names by substring (wrappers.py:241-246: `f'convs.{i}' in layer_name`), so for
    'convs.10' .. 'convs.15' it returns right after convs.1 and the hooked layer never runs: what G11 stored for them was the
    activation left over from get_instrumented_model's shape-annotation pass (a zero latent), not that of z.  The full forward
    visits every layer; its retained activations are the known answers a truthful partial_forward must reproduce."""
    ref = rh.import_reference()
    dev = torch.device("cpu")
    path = OUT / "synthesis_deep_known_answers.npz"
    with np.load(path) as data:
        ka = {k: data[k] for k in data.files}
    m = rh.rand_init_stylegan2(ref, dev, "ffhq", 1234)
    perturb_synthesis(m.model, [str(x) for x in ka["conv_names"]])
    mods = dict(m.model.named_modules())
    with torch.no_grad():
        for i, nme in enumerate(str(x) for x in ka["rgb_names"]):
            mods[nme].bias.copy_(0.05 * torch.tensor([1.0, -2.0, 3.0]).view(1, 3, 1, 1) * (i + 1))
    m.use_z()
    z = torch.from_numpy(ka["z"])
    layers = [f"convs.{i}" for i in range(10, 16)] + ["convs.9"]
    inst = ref.wrappers.get_instrumented_model("StyleGAN2", "ffhq", layers, dev, model=m, use_w=False)
    with torch.no_grad():
        m.forward(z[:1])
    feats = inst.retained_features()
    # control: convs.9 from the forward pass equals what partial_forward gave (G11)
    a9 = feats["convs.9"].numpy()
    assert np.abs(a9[:, ::max(1, a9.shape[1] // 16), ::4, ::4] - ka["act_convs_9_sub"]).max() < 1e-4 * np.abs(a9).max()
    for layer in layers[:-1]:
        act = feats[layer].numpy()
        key = layer.replace(".", "_")
        step = max(1, act.shape[-1] // 32)
        ka[f"act_{key}_sub"] = act[:, ::max(1, act.shape[1] // 16), ::step, ::step].copy()
        ka[f"sum_{key}"] = np.array([act.astype(np.float64).sum(), (act.astype(np.float64) ** 2).sum()])
        ka[f"shape_{key}"] = np.array(act.shape)
    inst.close()
    ka["two_digit_layers_from_forward"] = np.array(1)
    np.savez_compressed(path, **ka)
    print("rewrote convs.10..15 in synthesis_deep_known_answers.npz")


if __name__ == "__main__":
    if "g11" in sys.argv[1:]:
        deep_synthesis_golden()
    if "g11" in sys.argv[1:] or "g11b" in sys.argv[1:]:
        deep_synthesis_two_digit_layers()
    if not sys.argv[1:] or {"g9", "g10"} & set(sys.argv[1:]):
        main()

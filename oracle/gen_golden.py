"""ORACLE -- generator of the committed golden fixtures (tests/golden/*.npz).

Runs the UNMODIFIED reference (/root/reference: decomposition.get_or_compute, models.wrappers.StyleGAN2,
the vendored stylegan2-pytorch Generator, estimators.IPCAEstimator -> scikit-learn IncrementalPCA) on
CPU in the build container and stores its outputs.  /root/reference does not exist on the GPU box, so
nothing at test/bench time imports it -- only these small fixtures travel.

Recipe = SURVEY.md Appendix A: four sys.modules stubs for absent non-hot-path imports (fbpca, skimage,
boto3, botocore) and a random-init ``load_model`` override (no network for checkpoints; BASELINE.json
configs say "random-init weights").  Nothing else about the reference is modified.

Usage:  python oracle/gen_golden.py            (writes tests/golden/, prints oracle-vs-reference deltas)
"""
import os
import sys
import types
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = os.environ.get("GANSPACE_REFERENCE", "/root/reference")
OUT = REPO / "tests" / "golden"


def _import_reference():
    sys.path.insert(0, REF)

    def stub(name, **a):
        m = types.ModuleType(name)
        m.__dict__.update(a)
        sys.modules[name] = m
        return m

    stub("fbpca")
    sk = stub("skimage")
    sk.morphology = stub("skimage.morphology")
    stub("boto3")
    bc = stub("botocore")
    bc.exceptions = stub("botocore.exceptions", ClientError=Exception)
    cwd = os.getcwd()
    from models import wrappers, stylegan2          # noqa: chdir side effect (models/stylegan2/__init__.py:8-16)
    from config import Config
    import decomposition
    import estimators
    os.chdir(cwd)
    return wrappers, stylegan2, Config, decomposition, estimators


def perturb_synthesis(model, names):
    """Deterministic non-zero NoiseInjection weights / FusedLeakyReLU biases (both are 0 at random init, which would
    leave the noise and bias paths of the kernels untested)."""
    for i, name in enumerate(names):
        mod = dict(model.named_modules())[name]
        with torch.no_grad():
            mod.noise.weight.fill_(0.1 * (i + 1))
            mod.activate.bias.copy_(0.1 * torch.sin(torch.arange(mod.activate.bias.shape[0], dtype=torch.float32) + i))


def synthesis_golden(wrappers, stylegan2, Config, decomposition, orc, dev):
    """G7/G8: StyleGAN2 synthesis up to a hooked StyledConv (BASELINE config 5 family)."""
    class RandInitStyleGAN2(wrappers.StyleGAN2):
        def load_model(self):
            torch.manual_seed(1234)
            self.model = stylegan2.Generator(self.resolution, 512, 8).to(self.device)
            self.latent_avg = torch.zeros(512, device=self.device)

    names = ["conv1"] + [f"convs.{i}" for i in range(5)]
    # ---- G7: known answers of partial_forward, with perturbed noise weights / biases ---------------------------
    m = RandInitStyleGAN2(dev, "ffhq")
    sd = {k: v.clone() for k, v in m.model.state_dict().items()}
    perturb_synthesis(m.model, names)
    m.use_z()
    z = m.sample_latent(4, seed=21)
    ka = dict(z=z.numpy())
    with torch.no_grad():
        ka["w"] = m.model.style(z).numpy()
    for layer, keep in (("conv1", 4), ("convs.0", 4), ("convs.1", 4), ("convs.2", 2), ("convs.3", 1), ("convs.4", 1)):
        inst = wrappers.get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=m, use_w=False)
        with torch.no_grad():
            m.partial_forward(z[:keep], layer)
        act = inst.retained_features()[layer].numpy()
        key = layer.replace(".", "_")
        if act[0].size <= 131072:
            ka[f"act_{key}"] = act
        else:                                                # 2 MB per sample: keep a strided subsample + moments
            ka[f"act_{key}_sub"] = act[:, ::4, ::2, ::2].copy()
        ka[f"sum_{key}"] = np.array([act.astype(np.float64).sum(), (act.astype(np.float64) ** 2).sum()])
        inst.close()
    ka["const_sum"] = np.array(float(sd["input.input"].double().sum()))
    for nme in names:
        ka[f"wsum_{nme.replace('.', '_')}"] = np.array([float(sd[f"{nme}.conv.weight"].double().sum()),
                                                        float(sd[f"{nme}.conv.modulation.weight"].double().sum())])
    ka["noise_heads"] = np.stack([n.numpy().reshape(-1)[:4] for n in m.noise[:6]])
    np.savez_compressed(OUT / "synthesis_known_answers.npz", **ka)

    # ---- G8: layer=convs.1 (d = 32768) and convs.2 (d = 131072), Z space, pure random init -----------------------
    for layer, n, b, c in (("convs.1", 4_000, 500, 8), ("convs.2", 4_000, 250, 6)):
        m2 = RandInitStyleGAN2(dev, "ffhq")
        inst = wrappers.get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=m2, use_w=False)
        cfg = Config(model="StyleGAN2", layer=layer, output_class="ffhq", estimator="ipca", use_w=False,
                     n=n, batch_size=b, components=c)
        with tempfile.TemporaryDirectory() as tmp:
            path = decomposition.get_or_compute(cfg, inst, force_recompute=True,
                                                submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp))
            with np.load(path) as data:
                out = {k: data[k].copy() for k in data.files}
            name = path.name
        np.savez_compressed(OUT / f"c5s_stylegan2_ffhq_{layer.replace('.', '')}_z_n{n}_b{b}_c{c}.npz",
                            dump_name=np.array(name), **out)
        inst.close()

    # oracle vs reference
    p = orc.synthesis_random_init(1234, 1024, "convs.4")
    assert np.array_equal(p["const"], sd["input.input"].numpy()[0])
    for nme in names:
        assert np.array_equal(p["layers"][nme]["weight"], sd[f"{nme}.conv.weight"].numpy()[0]), nme
        assert np.array_equal(p["layers"][nme]["mod_weight"], sd[f"{nme}.conv.modulation.weight"].numpy()), nme
    print("synthesis init: oracle == reference")


def main():
    wrappers, stylegan2, Config, decomposition, estimators = _import_reference()
    sys.path.insert(0, str(REPO))
    from oracle import ganspace_oracle as orc
    OUT.mkdir(parents=True, exist_ok=True)
    dev = torch.device("cpu")
    if "--only-synthesis" in sys.argv:
        synthesis_golden(wrappers, stylegan2, Config, decomposition, orc, dev)
        return

    class RandInitStyleGAN2(wrappers.StyleGAN2):
        def load_model(self):                      # replaces checkpoint download (wrappers.py:153-165)
            torch.manual_seed(1234)
            self.model = stylegan2.Generator(self.resolution, 512, 8).to(self.device)
            self.latent_avg = torch.zeros(512, device=self.device)

    def run(cfg_kwargs, outclass, use_w):
        m = RandInitStyleGAN2(dev, outclass)
        inst = wrappers.get_instrumented_model("StyleGAN2", outclass, "style", dev, model=m, use_w=use_w)
        cfg = Config(model="StyleGAN2", layer="style", output_class=outclass, estimator="ipca",
                     use_w=use_w, **cfg_kwargs)
        with tempfile.TemporaryDirectory() as tmp:
            path = decomposition.get_or_compute(cfg, inst, force_recompute=True,
                                                submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp))
            with np.load(path) as data:
                out = {k: data[k].copy() for k in data.files}
            name = path.name
        return out, name, m

    # ---- G1: BASELINE config 1 (StyleGAN2-ffhq style --use_w, N=10k b=1k c=32) --------------------
    out1, name1, m = run(dict(n=10_000, batch_size=1_000, components=32), "ffhq", True)
    np.savez_compressed(OUT / "c1_stylegan2_ffhq_style_w_n10000_b1000_c32.npz", dump_name=np.array(name1), **out1)

    # ---- G2: Z-space layer=style with the regression pass (config-3 shape, small N) ---------------
    out2, name2, _ = run(dict(n=4_000, batch_size=1_000, components=16), "car", False)
    np.savez_compressed(OUT / "c3s_stylegan2_car_style_z_n4000_b1000_c16.npz", dump_name=np.array(name2), **out2)

    # ---- G3: W-space with a ragged plan (B does not divide NB; explicit seed) ----------------------
    out3, name3, _ = run(dict(n=5_000, batch_size=700, components=20, seed=7), "ffhq", True)
    np.savez_compressed(OUT / "w_ragged_n5000_b700_c20_seed7.npz", dump_name=np.array(name3), **out3)

    # ---- G4: mapping-network known answers + sample_latent stream heads ---------------------------
    np.random.seed(1)
    m.use_z()
    z = m.sample_latent(16)                         # consumes the first global seed after seed(1)
    with torch.no_grad():
        w = m.model.style(z)
    heads = []
    for s in (1791095845, 2135392491, 5, 0, 2147483646):
        heads.append(np.random.RandomState(s).standard_normal(64).astype(np.float32))
    tail = np.random.RandomState(1791095845).standard_normal(512 * 1000).astype(np.float32)[-64:]
    np.random.seed(1)
    seeds_1 = np.array([np.random.randint(np.iinfo(np.int32).max) for _ in range(16)], np.int64)
    np.random.seed(3)
    seeds_3 = np.array([np.random.randint(np.iinfo(np.int32).max) for _ in range(16)], np.int64)
    sd = m.model.state_dict()
    wsum = np.array([float(sd[f"style.{i + 1}.weight"].double().sum()) for i in range(8)])
    np.savez_compressed(OUT / "mapping_known_answers.npz", z=z.numpy(), w=w.numpy(),
                        head_seeds=np.array([1791095845, 2135392491, 5, 0, 2147483646], np.int64),
                        heads=np.stack(heads), tail_1791095845_512000=tail,
                        seeds_after_seed1=seeds_1, seeds_after_seed3=seeds_3, style_weight_sums=wsum,
                        style1_weight_head=sd["style.1.weight"][:4, :8].numpy())

    # ---- G5: IPCAEstimator chain on synthetic data (d=96, c=12, 5 batches of 300) ------------------
    rng = np.random.RandomState(11)
    basis = rng.standard_normal((96, 96)) * (0.9 ** np.arange(96))[None, :]
    Xs = [(rng.standard_normal((300, 96)) @ basis.T + 3.0 * rng.standard_normal(96)).astype(np.float32)
          for _ in range(5)]
    est = estimators.get_estimator("ipca", 12, 1.0)
    chain = {}
    for k, X in enumerate(Xs):
        assert est.fit_partial(X.copy())
        comp, stdev, ratio = est.get_components()
        chain[f"comp_{k}"] = np.asarray(comp, np.float64)
        chain[f"stdev_{k}"] = np.asarray(stdev, np.float64)
        chain[f"ratio_{k}"] = np.asarray(ratio, np.float64)
        chain[f"mean_{k}"] = np.asarray(est.transformer.mean_, np.float64)
        chain[f"var_{k}"] = np.asarray(est.transformer.var_, np.float64)
        chain[f"sv_{k}"] = np.asarray(est.transformer.singular_values_, np.float64)
    np.savez_compressed(OUT / "ipca_chain_d96_c12.npz", X=np.stack(Xs), param_str=np.array(est.get_param_str()),
                        **chain)

    # ---- G6: BigGAN-512 husky, layer=generator.gen_z (config-4 shape, small N), random-init -------------
    from models import biggan
    import models.wrappers as mw
    one_hot = lambda names: biggan.one_hot_from_int([248])          # 'husky' = ImageNet class 248 (no WordNet here)
    biggan.one_hot_from_names = one_hot
    mw.biggan.one_hot_from_names = one_hot
    deep512 = [(False, 16, 16), (True, 16, 16), (False, 16, 16), (True, 16, 8), (False, 8, 8), (True, 8, 8),
               (False, 8, 8), (True, 8, 4), (False, 4, 4), (True, 4, 2), (False, 2, 2), (True, 2, 1),
               (False, 1, 1), (True, 1, 1)]

    class RandInitBigGAN(wrappers.BigGAN):
        def load_model(self, name):
            torch.manual_seed(4321)
            cfg = biggan.BigGANConfig(output_dim=512, layers=deep512, attention_layer_position=8)
            self.model = biggan.BigGAN(cfg).to(self.device)

    bm = RandInitBigGAN(dev, 512, "husky")
    binst = wrappers.get_instrumented_model("BigGAN-512", "husky", "generator.gen_z", dev, model=bm)
    bcfg = Config(model="BigGAN-512", layer="generator.gen_z", output_class="husky", estimator="ipca",
                  n=4_000, batch_size=1_000, components=16)
    with tempfile.TemporaryDirectory() as tmp:
        path = decomposition.get_or_compute(bcfg, binst, force_recompute=True,
                                            submit_config=SimpleNamespace(run_dir=tmp, run_dir_root=tmp))
        with np.load(path) as data:
            out6 = {k: data[k].copy() for k in data.files}
        name6 = path.name
    np.savez_compressed(OUT / "c4s_biggan512_husky_genz_n4000_b1000_c16.npz", dump_name=np.array(name6), **out6)
    gz = bm.model.generator.gen_z
    zs = bm.sample_latent(8, seed=11)
    with torch.no_grad():
        bm.partial_forward(zs, "generator.gen_z")
    act = binst.retained_features()["generator.gen_z"]
    np.savez_compressed(OUT / "biggan_known_answers.npz", z=zs.numpy(), act=act.numpy(),
                        weight_orig_sum=np.array(float(gz.weight_orig.double().sum())),
                        weight_orig_head=gz.weight_orig[:4, :6].detach().numpy(),
                        bias_head=gz.bias[:8].detach().numpy(), u_head=gz.weight_u[:8].numpy(), v_head=gz.weight_v[:8].numpy(),
                        emb_head=bm.model.embeddings.weight[:4, :6].detach().numpy(),
                        trunc_seed5=biggan.truncated_noise_sample(batch_size=4, seed=5))

    synthesis_golden(wrappers, stylegan2, Config, decomposition, orc, dev)

    # ---- report oracle-vs-reference (also asserted by tests/test_oracle_golden.py) -----------------
    ws, bs = orc.mapping_random_init(1234)
    for i in range(8):
        assert np.array_equal(ws[i], sd[f"style.{i + 1}.weight"].numpy()), "mapping init mismatch"
    print("mapping w max abs diff:", np.abs(orc.mapping_forward(z.numpy(), ws, bs) - w.numpy()).max())
    for form in ("svd", "gram"):
        o1 = orc.compute_stylegan2_style(ws, bs, 10_000, 1_000, 32, True, ipca=form)
        print("G1", form, orc.compare_npz(o1, out1))
        o2 = orc.compute_stylegan2_style(ws, bs, 4_000, 1_000, 16, False, ipca=form)
        print("G2", form, orc.compare_npz(o2, out2))
        o3 = orc.compute_stylegan2_style(ws, bs, 5_000, 700, 20, True, seed=7, ipca=form)
        print("G3", form, orc.compare_npz(o3, out3))
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()

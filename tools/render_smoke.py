import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from ganspace_b200.models import StyleGAN2, get_instrumented_model
dev = torch.device("cuda:0")
m = StyleGAN2(dev, "ffhq", random_init=1234)
m.use_z()
z = m.sample_latent(2, seed=21)
t = time.time(); img = m.forward(z); torch.cuda.synchronize(); print("forward", img.shape, float(img.min()), float(img.max()), bool(torch.isfinite(img).all()), time.time() - t)
t = time.time(); img2 = m.forward(z); torch.cuda.synchronize(); print("forward again", time.time() - t, bool(torch.equal(img, img2)))
n = m.get_max_latents()
same = m.forward([z[:1]] * n); print("layerwise max diff", float((same - img[:1]).abs().max()))
for layer in ("convs.9", "convs.11", "convs.13", "convs.15", "to_rgbs.7"):
    inst = get_instrumented_model("StyleGAN2", "ffhq", layer, dev, model=m, use_w=False)
    m.partial_forward(z[:1], layer)
    a = inst.retained_features()[layer]
    print(layer, tuple(a.shape), float(a.abs().max()), bool(torch.isfinite(a).all()))
    inst.close()
m.check_numerics()
print("ok")

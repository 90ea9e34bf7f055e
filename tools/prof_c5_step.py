"""One config-5 group on one B200 (for ncu): 2000 latents -> mapping -> synthesis to convs.4 -> one large-d IPCA step,
twice (the second step has a previous state).  usage: python tools/prof_c5_step.py [nb=2000]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ganspace_b200 import _native                     # noqa: E402
from ganspace_b200.models import StyleGAN2            # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
m = StyleGAN2(dev, "ffhq", random_init=1234)
m.use_z()
big = _native.BigIPCA(512 * 32 * 32, 80, nb, dev)
for k in range(2):
    z = m.sample_latent(nb, seed=100 + k)
    m.activations_into(z, "convs.4", big.batch_rows(nb))
    big.step(nb)
torch.cuda.synchronize()
print("singular values head", big.export()["singular_values"][:3].cpu().numpy())

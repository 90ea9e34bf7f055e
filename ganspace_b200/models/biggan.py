"""BigGAN wrapper of the hot path: latent sampler + ``generator.gen_z`` (BASELINE.json config 4).

Mirror of /root/reference/models/wrappers.py:525-648 (``BigGAN(BaseModel)``) restricted to what
``decomposition.compute`` reaches for ``layer='generator.gen_z'`` (``partial_forward`` stops after gen_z
with n_layers = 0, wrappers.py:611-648):

    z      = truncated_noise_sample(seed)                  biggan utils.py:21-33   -> gsb_legacy_truncnorm_f32
    embed  = embeddings(one_hot(class))                    biggan model.py:291     (constant per class)
    act    = gen_z(cat(z, embed))                          biggan model.py:211-212,232
           = spectral_norm(Linear(256, 4*4*16*ch)): weight_orig / sigma, sigma = u^T W v with the stored u, v
             (eval mode: no power iteration)               -> folded once into W_eff, gsb_linear_forward

Low-rank activation path.  gen_z is affine in z:  act = z A^T + const,  A = W_eff[:, :128]  (d = 32768, r = 128).
With the thin QR  A = Q R  every centred activation is  (z - zbar) R^T Q^T, so sklearn's IncrementalPCA on the
32768-dim activations equals IncrementalPCA on y = z R^T (128-dim) followed by  components = components_y Q^T
(the SVD of the stacked matrix is equivariant under the isometry Q).  ``affine_layer()`` exposes that
factorisation; the decomposition driver then never materialises the [N, 32768] activations (131 GB at N=1e6)
and reuses the small-d engine.  ``partial_forward`` still produces the full activation for API users.

The synthesis blocks after gen_z (GenBlock / SelfAttn / BigGANBatchNorm) are outside the hot path
 and raise.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch
from torch import nn

from .. import _native
from .wrappers import BaseModel, _global_seed

# ImageNet class ids for the names GANSpace's configs use (the reference resolves names through
# nltk/WordNet, biggan utils.py:174-216, which is not available offline)
_CLASS_IDS = {"husky": 248, "siberian_husky": 250, "golden_retriever": 207, "lion": 291, "tiger": 292,
              "mushroom": 947, "barn": 425, "church": 497, "castle": 483, "volcano": 980, "lakeside": 975}

# (up-sample, in, out) of biggan-deep-{128,256,512}: only len(layers) matters here (n_latents)
_N_LAYERS = {128: 10, 256: 12, 512: 14}
_CHANNEL_WIDTH = 128


class _Config:
    def __init__(self, resolution):
        self.output_dim = resolution
        self.z_dim = 128
        self.class_embed_dim = 128
        self.channel_width = _CHANNEL_WIDTH
        self.num_classes = 1000
        self.layers = [None] * _N_LAYERS[resolution]
        self.eps = 1e-4


class SNLinear(nn.Module):
    """``spectral_norm(nn.Linear)`` in eval mode: y = x (W_orig / sigma)^T + b, sigma = u^T W_orig v."""

    def __init__(self, lin_sn):
        super().__init__()
        self.weight_orig = nn.Parameter(lin_sn.weight_orig.detach().clone())
        self.bias = nn.Parameter(lin_sn.bias.detach().clone())
        self.register_buffer("weight_u", lin_sn.weight_u.detach().clone())
        self.register_buffer("weight_v", lin_sn.weight_v.detach().clone())
        self._eff = None
        self._key = None

    def effective_weight(self) -> torch.Tensor:
        key = (self.weight_orig._version, self.weight_orig.data_ptr(), self.weight_u._version, self.weight_v._version)
        if self._eff is None or self._key != key:
            w = self.weight_orig.detach()
            sigma = torch.dot(self.weight_u, torch.mv(w, self.weight_v))      # torch spectral_norm, eval mode
            self._eff = (w / sigma).contiguous()
            self._key = key
        return self._eff

    def forward(self, x):
        # latents are truncated normals in [-2, 2] * truncation and the class embedding is a fixed vector: inside fp16's range,
        # so batches of >= 128 rows run on the tensor cores (fp32-grade hi/lo split)
        return _native.linear(x, self.effective_weight(), self.bias.detach(), bounded=bool(x.abs().max() < 6e4) if x.shape[0] >= 128 else False)


class _Generator(nn.Module):
    def __init__(self, gen_z, config):
        super().__init__()
        self.config = config
        self.gen_z = gen_z
        self.layers = nn.ModuleList()        # GenBlock / SelfAttn: outside the hot path


class _BigGANNet(nn.Module):
    """Parameter layout of pytorch_pretrained_biggan.BigGAN restricted to embeddings + generator.gen_z."""

    def __init__(self, resolution):
        super().__init__()
        self.config = _Config(resolution)
        # creation order == the reference's (embeddings, then Generator.gen_z first), so that
        # torch.manual_seed(s) reproduces the reference's random init of these tensors bit-for-bit
        self.embeddings = nn.Linear(self.config.num_classes, self.config.z_dim, bias=False)
        lin = nn.utils.spectral_norm(nn.Linear(2 * self.config.z_dim, 4 * 4 * 16 * self.config.channel_width),
                                     eps=self.config.eps)
        self.generator = _Generator(SNLinear(lin), self.config)
        self.n_latents = len(self.config.layers) + 1


class AffineLayer:
    """act = y Q^T + offset,  y = z R^T  (see module docstring).  All device tensors."""

    def __init__(self, Q64, R64, offset64):
        self.Q = Q64                                   # [d, r] fp64, orthonormal columns
        self.Q32 = Q64.float().contiguous()
        self.R32 = R64.float().contiguous()           # [r, r]
        self.offset = offset64                         # [d] fp64: b + W_eff[:, r:] @ embed
        self.rank = Q64.shape[1]
        self.dim = Q64.shape[0]

    def coords(self, z: torch.Tensor) -> torch.Tensor:
        """y = z R^T through the fp32 GEMM kernel."""
        return _native.linear(z, self.R32)

    def lift_rows(self, rows64: torch.Tensor) -> torch.Tensor:
        """rows [k, r] in y-space -> [k, d] in activation space (directions: no offset); in-tree GEMM kernel, fp32 (the
        results are stored as float32)."""
        return _native.linear(rows64.float().contiguous(), self.Q32).double()


class BigGAN(BaseModel):
    def __init__(self, device, resolution, class_name, truncation=1.0, random_init=None):
        super().__init__(f"BigGAN-{resolution}", class_name)
        self.device = _native.require_cuda(device)
        self.truncation = truncation
        self._random_init = random_init
        self.resolution = int(resolution)
        self.load_model(f"biggan-deep-{resolution}")
        self.set_output_class(class_name or "husky")
        self.name = f"BigGAN-{resolution}-{self.outclass}-t{self.truncation}"
        self.has_latent_residual = True
        self._affine = {}

    def load_model(self, name):
        if self.resolution not in _N_LAYERS:
            raise RuntimeError("Unknown BigGAN model name", name)
        root = os.environ.get("GANCONTROL_CHECKPOINT_DIR", Path(__file__).parent / "checkpoints")
        weights = Path(root) / name / "pytorch_model.bin"
        seed = self._random_init
        if seed is None and os.environ.get("GANSPACE_B200_RANDOM_INIT"):
            seed = int(os.environ["GANSPACE_B200_RANDOM_INIT_BIGGAN"]) if os.environ.get("GANSPACE_B200_RANDOM_INIT_BIGGAN") \
                else int(os.environ["GANSPACE_B200_RANDOM_INIT"])
        if weights.is_file() and seed is None:
            net = _BigGANNet(self.resolution)
            sd = torch.load(weights, map_location="cpu")
            g = net.generator.gen_z
            net.embeddings.weight.data.copy_(sd["embeddings.weight"])
            g.weight_orig.data.copy_(sd["generator.gen_z.weight_orig"])
            g.bias.data.copy_(sd["generator.gen_z.bias"])
            g.weight_u.copy_(sd["generator.gen_z.weight_u"])
            g.weight_v.copy_(sd["generator.gen_z.weight_v"])
        elif seed is not None:
            torch.manual_seed(int(seed))
            net = _BigGANNet(self.resolution)
        else:
            raise RuntimeError(f"BigGAN weights {weights} not found and no network access; pass random_init=<seed>")
        self.model = net.to(self.device)

    # ---- latents ----------------------------------------------------------------------------------
    def sample_latent(self, n_samples=1, truncation=None, seed=None):
        if seed is None:
            seed = _global_seed()
        t = truncation or self.truncation
        return _native.legacy_truncnorm([seed], 128 * n_samples, -2.0, 2.0, float(t), self.device).view(n_samples, 128)

    def sample_latents_multi(self, n_samples, seeds, out=None):
        z = _native.legacy_truncnorm(list(seeds), 128 * n_samples, -2.0, 2.0, float(self.truncation), self.device)
        return z.view(len(seeds) * n_samples, 128)

    def get_max_latents(self):
        return len(self.model.config.layers) + 1

    def get_conditional_state(self, z):
        return self.v_class

    def set_conditional_state(self, z, c):
        self.v_class = c

    def is_valid_class(self, class_id):
        if isinstance(class_id, int):
            return class_id < 1000
        if isinstance(class_id, str):
            return class_id.replace(" ", "_").lower() in _CLASS_IDS
        raise RuntimeError(f"Unknown class identifier {class_id}")

    def set_output_class(self, class_id):
        if isinstance(class_id, int):
            idx = class_id
            self.outclass = f"class{class_id}"
        elif isinstance(class_id, str):
            key = class_id.replace(" ", "_").lower()
            if key not in _CLASS_IDS:
                raise RuntimeError(f"Unknown class identifier {class_id} (WordNet lookup is not available offline; "
                                   f"known names: {sorted(_CLASS_IDS)}; or pass the ImageNet class index)")
            idx = _CLASS_IDS[key]
            self.outclass = class_id.replace(" ", "_")
        else:
            raise RuntimeError(f"Unknown class identifier {class_id}")
        one_hot = torch.zeros(1, 1000, dtype=torch.float32)
        one_hot[0, idx] = 1.0
        self.v_class = one_hot.to(self.device)
        self._class_idx = idx
        self._affine = {}
        if hasattr(self, "model"):
            self.affine_layer("generator.gen_z")       # thin QR of the folded weight: model/class set-up, not part of a run

    def _embed(self) -> torch.Tensor:
        """embeddings(one_hot) = column `idx` of the embedding matrix  -> [128] fp32."""
        return self.model.embeddings.weight.detach()[:, self._class_idx].contiguous()

    def forward(self, x):
        raise NotImplementedError("BigGAN image synthesis (GenBlock / SelfAttn) is outside the B200 hot path "
                                  "(SURVEY.md section 2 row 8)")

    def partial_forward(self, x, layer_name):
        if layer_name not in ("embeddings", "generator.gen_z"):
            raise NotImplementedError(f"BigGAN.partial_forward to '{layer_name}': only generator.gen_z is on the hot path")
        z = x[0] if isinstance(x, list) else x
        cond = torch.cat((z, self._embed().unsqueeze(0).expand(z.shape[0], -1)), dim=1).contiguous()
        self.model.generator.gen_z(cond)             # hook retains [B, 4*4*16*ch]
        return None

    # ---- low-rank structure of gen_z ----------------------------------------------------------------
    def affine_layer(self, layer_name):
        # GANSPACE_B200_BIGGAN_AFFINE=0: no low-rank shortcut -- the activations are materialised and go through the general
        # large-d engine (csrc/bigd.cu), a cross-check of the shortcut
        if layer_name != "generator.gen_z" or os.environ.get("GANSPACE_B200_BIGGAN_AFFINE", "1") == "0":
            return None
        if layer_name not in self._affine:
            g = self.model.generator.gen_z
            w = g.effective_weight().double()                                # [d, 256]
            r = self.model.config.z_dim
            Q, R = torch.linalg.qr(w[:, :r].contiguous(), mode="reduced")   # one-time setup (like weight packing)
            offset = g.bias.detach().double() + w[:, r:] @ self._embed().double()
            self._affine[layer_name] = AffineLayer(Q.contiguous(), R.contiguous(), offset.contiguous())
        return self._affine[layer_name]

"""CPU stand-ins for the device layer, so that the HOST logic of the decomposition driver (plan, seeding protocol, large-d
path, feature sharding, layout permutation, regression, post-processing) can be exercised without a GPU -- single process
and under a world_size-2 gloo job.  Test infrastructure only: numpy / torch-CPU arithmetic, same interfaces as
ganspace_b200._native.BigIPCA / LinregAccumulator / project_std and as a models.wrappers.BaseModel with a producer
(activations_into / feature_layout)."""
import numpy as np
import torch
import torch.distributed as dist

from ganspace_b200 import _native as _real_native

_REAL_GATHERED = _real_native.BigIPCA.gathered          # captured before install() swaps the class


class FakeBig:
    """Interface of _native.BigIPCA; the three device phases in numpy (fp64), collectives over the default group."""

    def __init__(self, d, c, nb_max, device, shard=None, gram=None):
        from ganspace_b200 import _native
        self._pick = _native.pick_global_signs
        self.shard = shard if (shard is not None and shard[1] > 1) else None
        self.d_full = int(d)
        W = self.shard[1] if self.shard else 1
        self.d, self.c, self.nb_max, self.dev = int(d) // W, int(c), int(nb_max), torch.device("cpu")
        self.flags = 0
        self.rows = c + nb_max + 1
        self.M = torch.zeros((self.rows, self.d), dtype=torch.float32)
        self.mean = np.zeros(self.d)
        self.unnorm = np.zeros(self.d)
        self.S = np.zeros(c)
        self.batch_mean = torch.zeros(self.d, dtype=torch.float64)
        self.n_seen = self.last_nb = 0

    def batch_rows(self, nb):
        return self.M[self.c:self.c + nb]

    def step(self, nb):
        c = self.c
        X = self.M[c:c + nb].numpy().astype(np.float64)
        mb = X.mean(0)
        ss = ((X - mb) ** 2).sum(0)
        corr = np.sqrt(self.n_seen / (self.n_seen + nb) * nb) * (self.mean - mb) if self.n_seen else np.zeros(self.d)
        self.M[c:c + nb] = torch.from_numpy((X - mb).astype(np.float32))
        self.M[c + nb] = torch.from_numpy(corr.astype(np.float32))
        Mm = self.M[:c + nb + 1].numpy().astype(np.float64)
        T = torch.from_numpy(Mm @ Mm.T)
        if self.shard:
            dist.all_reduce(T)
        lam, U = np.linalg.eigh(T.numpy())
        lam, U = lam[::-1][:c], U[:, ::-1][:, :c].T
        Dn = (U @ Mm).astype(np.float32)
        idx = np.argmax(np.abs(Dn), axis=1)
        rowmax = torch.from_numpy(np.stack([np.abs(Dn)[np.arange(c), idx], Dn[np.arange(c), idx]], 1).astype(np.float32))
        if self.shard:
            allmax = torch.empty((self.shard[1] * c, 2))
            dist.all_gather_into_tensor(allmax, rowmax)
            signs = self._pick(allmax.view(self.shard[1], c, 2)).numpy()
        else:
            signs = np.where(rowmax[:, 1].numpy() < 0, -1.0, 1.0)
        self.M[:c] = torch.from_numpy(Dn * signs[:, None].astype(np.float32))
        self.S = np.sqrt(np.maximum(lam, 0))
        if self.n_seen:
            r = self.n_seen / nb
            tq = (self.mean * self.n_seen) / r - mb * nb
            self.unnorm = self.unnorm + ss + r / (self.n_seen + nb) * tq * tq
        else:
            self.unnorm = ss
        self.mean = (self.mean * self.n_seen + mb * nb) / (self.n_seen + nb)
        self.batch_mean = torch.from_numpy(mb)
        self.n_seen += nb
        self.last_nb = nb

    def gathered(self, local):
        return _REAL_GATHERED(self, local)                   # the real gather logic (pure torch.distributed)

    def export(self):
        out = {"components": (self.M[:self.c] / torch.from_numpy(self.S)[:, None]).float(),
               "singular_values": torch.from_numpy(self.S.copy()), "mean": torch.from_numpy(self.mean.copy()),
               "var": torch.from_numpy(self.unnorm / self.n_seen),
               "explained_variance": torch.from_numpy(self.S ** 2 / (self.n_seen - 1)),
               "explained_variance_ratio": torch.from_numpy(self.S ** 2 / self.unnorm.sum())}
        if self.shard:
            out["components"] = self.gathered(out["components"])
            out["mean"] = self.gathered(out["mean"].unsqueeze(0)).reshape(-1)
            out["var"] = self.gathered(out["var"].unsqueeze(0)).reshape(-1)
            out["explained_variance_ratio"] = out["singular_values"] ** 2 / (out["var"].sum() * self.n_seen)
        return out


class FakeLinreg:
    """Interface of _native.LinregAccumulator (decomposition.py:77-139 normal equations), torch-CPU fp64."""

    def __init__(self, c, latent_dim, device):
        self.c, self.L = int(c), int(latent_dim)
        self.state = torch.zeros(self.c * self.c + self.c * self.L + self.L, dtype=torch.float64)
        self.n_total = 0

    def accumulate(self, act, comp32, mean32, stdev32, z):
        A = ((act - mean32[None, :]) @ comp32.T / stdev32[None, :]).double()
        Z = z.double()
        c, L = self.c, self.L
        self.state[:c * c] += (A.T @ A).reshape(-1)
        self.state[c * c:c * c + c * L] += (A.T @ Z).reshape(-1)
        self.state[c * c + c * L:] += Z.sum(0)
        self.n_total += act.shape[0]

    def solve(self):
        c, L = self.c, self.L
        AtA = self.state[:c * c].view(c, c)
        AtZ = self.state[c * c:c * c + c * L].view(c, L)
        return torch.linalg.solve(AtA, AtZ), self.state[c * c + c * L:] / self.n_total


def fake_project_std(x, dirs, sub=None):
    xx = x.double()
    if sub is not None:
        xx = (xx - sub.double()[None, :]).float().double()      # numpy's float32 -= float64, then float32 products
    p = xx.float() @ dirs.float().T
    return p.double().std(dim=0, unbiased=False).float()


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


class FakeFeatureModel(torch.nn.Module):
    """A BaseModel-shaped generator on the CPU: latents -> a [C,H,W] feature map (tanh of an affine map), hookable as
    layer 'feat', with the producer interface (NHWC rows) the large-d path uses."""

    def __init__(self, C=8, H=16, W=16, latent=32, seed=5):
        super().__init__()
        self.device = torch.device("cpu")
        self.model_name, self.outclass, self.name = "Fake", "none", "Fake-none"
        self.model = torch.nn.Module()
        self.model.feat = _Identity()
        self.C, self.H, self.W, self.latent = C, H, W, latent
        g = torch.Generator().manual_seed(seed)
        d = C * H * W
        basis = torch.randn(d, latent, generator=g) * (0.8 ** torch.arange(latent, dtype=torch.float32))[None, :]
        self.A = basis * 0.5
        self.b = 0.3 * torch.randn(d, generator=g)

    def named_modules(self, *a, **k):
        return self.model.named_modules(*a, **k)

    def sample_latent(self, n_samples=1, seed=None, truncation=None):
        if seed is None:
            seed = np.random.randint(np.iinfo(np.int32).max)
        z = np.random.RandomState(seed).standard_normal(self.latent * n_samples).reshape(n_samples, self.latent)
        return torch.from_numpy(z).float()

    def act_nchw_flat(self, z):
        return torch.tanh(z.reshape(-1, self.latent) @ self.A.T + self.b)          # [n, C*H*W] in the reference's NCHW order

    def partial_forward(self, x, layer_name):
        assert layer_name == "feat"
        self.model.feat(self.act_nchw_flat(x).view(-1, self.C, self.H, self.W))

    def feature_layout(self, layer_name):
        return ("nhwc", (self.H, self.W, self.C))

    def activations_into(self, x, layer_name, out):
        a = self.act_nchw_flat(x).view(-1, self.C, self.H, self.W).permute(0, 2, 3, 1).reshape(-1, self.C * self.H * self.W)
        out.copy_(a)
        return out

    def latent_space_name(self):
        return "Z"

    def get_latent_shape(self):
        return tuple(self.sample_latent(1).shape)

    def get_latent_dims(self):
        return np.prod(self.get_latent_shape())

    def set_output_class(self, c):
        pass

    def get_max_latents(self):
        return 1


class FakeStyleModel(FakeFeatureModel):
    """StyleGAN-shaped: a mapping network behind layer 'style', W space via use_w() (wrappers.py:167-179,194-222)."""

    def __init__(self, latent=32, seed=9):
        super().__init__(C=1, H=1, W=latent, latent=latent, seed=seed)
        self.model.style = _Identity()
        g = torch.Generator().manual_seed(seed + 1)
        self.Mw = torch.randn(latent, latent, generator=g) * (0.7 ** torch.arange(latent, dtype=torch.float32))[:, None]
        self.w_primary = False

    def use_w(self):
        self.w_primary = True

    def use_z(self):
        self.w_primary = False

    def mapping(self, z):
        return torch.tanh(z @ self.Mw.T) + 0.1 * z

    def latent_space_name(self):
        return "W" if self.w_primary else "Z"

    def sample_latent(self, n_samples=1, seed=None, truncation=None):
        z = super().sample_latent(n_samples, seed)
        return self.mapping(z) if self.w_primary else z

    def partial_forward(self, x, layer_name):
        assert layer_name == "style"
        self.model.style(x if self.w_primary else self.mapping(x))

    def feature_layout(self, layer_name):
        return None


class FakeChain:
    """Interface of _native.IPCAChain (small-d Gram-form engine); the step is the oracle's Gram-form restatement."""

    def __init__(self, d, c, device, side_stream=True):
        from oracle import ganspace_oracle as orc
        self._orc = orc
        self.d, self.c, self.dev = int(d), int(c), torch.device("cpu")
        self.st = orc.IPCAState(self.c)
        self.n_seen = 0

    def step(self, n_batch, mean_b, gram_b):
        self._orc.ipca_gram_step(self.st, int(n_batch), mean_b.numpy().copy(), gram_b.numpy().copy())
        self.n_seen += int(n_batch)

    def join(self):
        pass

    def export(self):
        st = self.st
        t = lambda a: torch.from_numpy(np.array(a, dtype=np.float64))
        return {"components": t(st.components), "singular_values": t(st.singular_values), "mean": t(st.mean), "var": t(st.var),
                "explained_variance": t(st.explained_variance), "explained_variance_ratio": t(st.explained_variance_ratio)}


def fake_batch_stats(x, mean_out=None, gram_out=None):
    x64 = x.double()
    mean = x64.mean(0)
    xc = x64 - mean
    return mean, xc.T @ xc


def fake_batch_stats_multi(x, n_groups, rows_per_group, mean_out=None, gram_out=None):
    d = x.shape[1]
    mean = mean_out if mean_out is not None else torch.empty((n_groups, d), dtype=torch.float64)
    gram = gram_out if gram_out is not None else torch.empty((n_groups, d, d), dtype=torch.float64)
    for g in range(n_groups):
        m, G = fake_batch_stats(x[g * rows_per_group:(g + 1) * rows_per_group])
        mean[g] = m
        gram[g] = G
    return mean, gram


def install(native, estimators):
    """Swap the device layer of the product for the CPU stand-ins (call inside the process under test)."""
    native.BigIPCA = FakeBig
    native.IPCAChain = FakeChain
    native.batch_stats = fake_batch_stats
    native.batch_stats_multi = fake_batch_stats_multi
    native.LinregAccumulator = FakeLinreg
    native.project_std = fake_project_std
    native.require_cuda = lambda device=None: torch.device("cpu")

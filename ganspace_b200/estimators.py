"""Estimators of the hot path: ``IPCAEstimator`` / ``get_estimator``.

Mirror of /root/reference/estimators.py:55-81,206-218.  ``IPCAEstimator`` keeps the reference's
interface (``batch_support``, ``get_param_str``, ``fit``, ``fit_partial``, ``get_components`` and a
``.transformer`` object exposing scikit-learn's fitted attributes ``mean_``, ``var_``, ``components_``,
``singular_values_``, ``explained_variance_``, ``explained_variance_ratio_``, ``n_samples_seen_``) but the
arithmetic of ``IncrementalPCA.partial_fit`` runs on the device: per-batch statistics (csrc/stats.cu)
feed the fp64 Gram-form merge chain (csrc/ipca.cu).  Batches may be host ndarrays (copied in, as the
reference API allows) or CUDA tensors (no copy: the samples never leave HBM).  For d > 1024 (conv feature
maps) the d x d Gram is out of reach and the small-side engine (csrc/bigd.cu) takes over behind the same
interface; its stacked matrix lives in HBM and producers can write batches in place (``batch_buffer``).

The other estimators of the reference (pca / fbpca / ica / spca, estimators.py:18-52,84-204) are not
batched and appear in no BASELINE config; asking for them raises (SURVEY.md section 2 marks them out of
scope).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native


class DeviceIncrementalPCA:
    """The ``.transformer`` of IPCAEstimator: sklearn's fitted-attribute names, device-resident state."""

    def __init__(self, n_components, whiten=False, batch_size=None, device=None):
        if whiten:
            raise NotImplementedError("whiten=True is not used by GANSpace (estimators.py:58)")
        self.n_components = n_components
        self.whiten = whiten
        self.batch_size = batch_size
        self._device = device
        self._chain = None
        self._host = None            # cached export (host view)
        self._dev = None             # ... and its device tensors
        self._shard = None           # (rank, world): feature-sharded large-d engine
        self._stage = None
        self.n_samples_seen_ = np.int64(0)

    def enable_feature_sharding(self, rank, world):
        """Large-d engine over a torch.distributed job: every rank keeps d/world features of the stacked matrix and
        produces nb/world rows of each batch; ``partial_fit_inplace`` exchanges them with one all-to-all."""
        if self._chain is not None:
            raise RuntimeError("enable_feature_sharding must precede the first partial_fit")
        self._shard = (int(rank), int(world)) if world > 1 else None

    # -- device side ---------------------------------------------------------------------------------
    SMALL_D_MAX = 1024          # above this the d x d Gram engine (csrc/ipca.cu) gives way to the small-side engine (bigd.cu)

    def _ensure(self, d, device, nb=None):
        if self._chain is None:
            if self.n_components > d:
                raise ValueError(
                    "n_components=%r invalid for n_features=%d, need more rows than columns for "
                    "IncrementalPCA processing" % (self.n_components, d))
            if d > self.SMALL_D_MAX:
                self._chain = _native.BigIPCA(d, self.n_components, nb, device, shard=self._shard)
            else:
                self._chain = _native.IPCAChain(d, self.n_components, device)
        elif getattr(self._chain, "d_full", self._chain.d) != d:     # (a feature-sharded engine's .d is its local width)
            raise ValueError("Number of input features has changed from %i to %i between calls to partial_fit!"
                             % (getattr(self._chain, "d_full", self._chain.d), d))
        return self._chain

    @property
    def is_large_d(self):
        return isinstance(self._chain, _native.BigIPCA)

    def batch_buffer(self, nb, d, device):
        """Large-d engine only: the [nb, d] rows of the device-resident stacked matrix that the NEXT partial_fit will
        consume.  Producers (e.g. the synthesis kernels) write the raw batch there; then ``partial_fit_inplace(nb)``."""
        if d <= self.SMALL_D_MAX:
            raise ValueError("batch_buffer is the large-d engine's interface (d > %d)" % self.SMALL_D_MAX)
        chain = self._ensure(int(d), device, nb=int(nb))
        if self._shard is not None:
            # sharded: the caller fills a staging buffer with ITS rows [rank*q, (rank+1)*q) of the batch, all d features
            world = self._shard[1]
            if nb % world != 0 or nb > chain.nb_max:
                raise ValueError(f"feature-sharded IPCA needs equal batches divisible by the world size (nb={nb}, world={world})")
            q = nb // world
            if self._stage is None or self._stage.shape != (q, int(d)):
                self._stage = torch.empty((q, int(d)), dtype=torch.float32, device=chain.dev)
            return self._stage
        if nb > chain.nb_max:                                   # a later batch larger than the first: grow, keep the state
            big = _native.BigIPCA(chain.d, chain.c, int(nb), chain.dev, gram="tc" if chain.flags & 1 else "simt")
            big.M[:chain.c].copy_(chain.M[:chain.c])
            big.state.copy_(chain.state)
            big.n_seen = chain.n_seen
            self._chain = chain = big
        return chain.batch_rows(int(nb))

    def partial_fit_inplace(self, nb):
        chain = self._chain
        if int(self.n_samples_seen_) == 0 and self.n_components > nb:
            raise ValueError(f"n_components={self.n_components} must be less or equal to the batch number of "
                             f"samples {nb} for the first partial_fit call.")
        if self._shard is not None:
            _native.exchange_rows(self._stage, chain.batch_rows(int(nb)), self._shard[1])
        chain.step(int(nb))          # centres the batch rows in place (chain.batch_mean holds the batch mean)
        self.n_samples_seen_ = np.int64(chain.n_seen)
        self._host = None
        return self

    def batch_stats(self, X):
        """(n, mean[d], centred Gram[d,d]) of a batch; host arrays are copied to the device first."""
        if isinstance(X, np.ndarray):
            dev = _native.require_cuda(self._device)
            X = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(dev)
        if X.dim() != 2:
            raise ValueError("Expected 2D array, got %dD" % X.dim())
        if X.dtype != torch.float32:
            X = X.float()
        mean, gram = _native.batch_stats(X)
        return int(X.shape[0]), mean, gram

    def begin_run(self, n_groups, n_batch, d, device):
        """Small-d engine: announce a run of ``n_groups`` partial_fit calls of ``n_batch`` rows each.  From the second group on
        the merge chain then runs in ONE resident cluster kernel that takes the groups' statistics from a device queue
        (csrc/subspace.cu) instead of one launch per group."""
        if d > self.SMALL_D_MAX or self._chain is not None or self.n_components > min(d, n_batch):
            return False
        chain = self._ensure(int(d), device)
        return chain.begin_run(int(n_groups), int(n_batch)) if hasattr(chain, "begin_run") else False

    def end_run(self):
        if self._chain is not None and hasattr(self._chain, "end_run"):
            self._chain.end_run()

    def merge(self, n_batch, mean_b, gram_b):
        """One partial_fit step from precomputed batch statistics (must follow the reference's batch order)."""
        chain = self._ensure(int(mean_b.shape[0]), mean_b.device)
        if int(self.n_samples_seen_) == 0 and self.n_components > n_batch:
            raise ValueError(f"n_components={self.n_components} must be less or equal to the batch number of "
                             f"samples {n_batch} for the first partial_fit call.")
        run = getattr(chain, "_run", None)
        if run is not None and not run["closed"]:
            chain.run_step(n_batch, mean_b, gram_b)
        else:
            chain.step(n_batch, mean_b, gram_b)
        self.n_samples_seen_ = np.int64(chain.n_seen)
        self._host = None

    def partial_fit(self, X, y=None):
        d = int(X.shape[1]) if getattr(X, "ndim", 2) == 2 else 0
        if d > self.SMALL_D_MAX:
            if isinstance(X, np.ndarray):
                X = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
            dev = X.device if X.is_cuda else _native.require_cuda(self._device)
            self.batch_buffer(X.shape[0], d, dev).copy_(X)      # host arrays are copied in, CUDA tensors stay in HBM
            return self.partial_fit_inplace(X.shape[0])
        self.merge(*self.batch_stats(X))
        return self

    def last_batch_mean(self):
        """Large-d engine: fp64 [d] mean of the batch the last partial_fit centred in place."""
        return self._chain.gathered(self._chain.batch_mean.unsqueeze(0)).reshape(-1)

    def last_batch_rows(self, n):
        """Large-d engine: the first n rows of the last batch as centred by partial_fit, full feature width."""
        return self._chain.gathered(self._chain.batch_rows(self._chain.last_nb)[:n])

    # -- sklearn attribute names (host copies, fetched lazily) ---------------------------------------
    def device_attributes(self):
        """sklearn's fitted attributes as device tensors; one export per fit state (the host view below reuses it)."""
        if self._chain is None:
            raise AttributeError("IncrementalPCA is not fitted yet")
        if self._host is None or self._dev is None:
            self._dev = self._chain.export()
        return self._dev

    def _export(self):
        if self._chain is None:
            raise AttributeError("IncrementalPCA is not fitted yet")
        if self._host is None:
            dev = self._chain.export()
            # ONE device->host transfer: the six arrays packed into one fp64 buffer (each .cpu() is a synchronising copy)
            keys = list(dev)
            if sum(dev[k].numel() for k in keys) <= (1 << 22):
                flat = torch.cat([dev[k].reshape(-1).double() for k in keys]).cpu().numpy()
                host, off = {}, 0
                for k in keys:
                    n = dev[k].numel()
                    host[k] = flat[off:off + n].reshape(tuple(dev[k].shape)).astype(
                        np.float32 if dev[k].dtype == torch.float32 else np.float64)
                    off += n
            else:                                        # conv feature maps: 168 MB of components, copied as they are
                host = {k: v.cpu().numpy() for k, v in dev.items()}
            self._dev = dev
            self._host = host
        return self._host

    components_ = property(lambda self: self._export()["components"])
    singular_values_ = property(lambda self: self._export()["singular_values"])
    mean_ = property(lambda self: self._export()["mean"])
    var_ = property(lambda self: self._export()["var"])
    explained_variance_ = property(lambda self: self._export()["explained_variance"])
    explained_variance_ratio_ = property(lambda self: self._export()["explained_variance_ratio"])


class IPCAEstimator:
    def __init__(self, n_components, device=None):
        self.n_components = n_components
        self.whiten = False
        self.transformer = DeviceIncrementalPCA(n_components, whiten=self.whiten,
                                                batch_size=max(100, 2 * n_components), device=device)
        self.batch_support = True

    def get_param_str(self):
        return "ipca_c{}{}".format(self.n_components, "_w" if self.whiten else "")

    def fit(self, X):
        # sklearn IncrementalPCA.fit: partial_fit over gen_batches(n, batch_size, min_batch_size=n_components)
        n = X.shape[0]
        bs = self.transformer.batch_size
        start = 0
        while start < n:
            end = start + bs
            if end + self.n_components > n:
                end = n
            self.transformer.partial_fit(X[start:end])
            start = end

    def fit_partial(self, X):
        try:
            self.transformer.partial_fit(X)
            self.transformer.n_samples_seen_ = np.int64(self.transformer.n_samples_seen_)   # estimators.py:71-72
            return True
        except ValueError as e:
            print("\nIPCA error:", e)
            return False

    def fit_partial_stats(self, n_batch, mean_b, gram_b):
        """fit_partial from the batch's sufficient statistics (n, mean [d], centred Gram [d, d]; fp64 device tensors)
        instead of its rows -- small-d engine; must be called in the reference's batch order."""
        try:
            self.transformer.merge(int(n_batch), mean_b, gram_b)
            self.transformer.n_samples_seen_ = np.int64(self.transformer.n_samples_seen_)
            return True
        except ValueError as e:
            print("\nIPCA error:", e)
            return False

    def fit_partial_inplace(self, nb):
        """fit_partial on the batch already written into ``transformer.batch_buffer(nb, d, device)`` (large-d engine)."""
        try:
            self.transformer.partial_fit_inplace(nb)
            return True
        except ValueError as e:
            print("\nIPCA error:", e)
            return False

    def get_components(self):
        stdev = np.sqrt(self.transformer.explained_variance_)
        var_ratio = self.transformer.explained_variance_ratio_
        return self.transformer.components_, stdev, var_ratio


def get_estimator(name, n_components, alpha, device=None):
    if name == "ipca":
        return IPCAEstimator(n_components, device=device)
    if name in ("pca", "fbpca", "ica", "spca"):
        raise RuntimeError(f"estimator '{name}' is not batched and outside the B200 hot path "
                           "(SURVEY.md section 2); use 'ipca'")
    raise RuntimeError("Unknown estimator")

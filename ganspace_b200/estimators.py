"""Estimators of the hot path: ``IPCAEstimator`` / ``get_estimator``.

Mirror of /root/reference/estimators.py:55-81,206-218.  ``IPCAEstimator`` keeps the reference's
interface (``batch_support``, ``get_param_str``, ``fit``, ``fit_partial``, ``get_components`` and a
``.transformer`` object exposing scikit-learn's fitted attributes ``mean_``, ``var_``, ``components_``,
``singular_values_``, ``explained_variance_``, ``explained_variance_ratio_``, ``n_samples_seen_``) but the
arithmetic of ``IncrementalPCA.partial_fit`` runs on the device: per-batch statistics (csrc/stats.cu)
feed the fp64 Gram-form merge chain (csrc/ipca.cu).  Batches may be host ndarrays (copied in, as the
reference API allows) or CUDA tensors (no copy: the samples never leave HBM).

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
        self._host = None            # cached export
        self.n_samples_seen_ = np.int64(0)

    # -- device side ---------------------------------------------------------------------------------
    def _ensure(self, d, device):
        if self._chain is None:
            if self.n_components > d:
                raise ValueError(
                    "n_components=%r invalid for n_features=%d, need more rows than columns for "
                    "IncrementalPCA processing" % (self.n_components, d))
            self._chain = _native.IPCAChain(d, self.n_components, device)
        elif self._chain.d != d:
            raise ValueError("Number of input features has changed from %i to %i between calls to partial_fit!"
                             % (self._chain.d, d))
        return self._chain

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

    def merge(self, n_batch, mean_b, gram_b):
        """One partial_fit step from precomputed batch statistics (must follow the reference's batch order)."""
        chain = self._ensure(int(mean_b.shape[0]), mean_b.device)
        if int(self.n_samples_seen_) == 0 and self.n_components > n_batch:
            raise ValueError(f"n_components={self.n_components} must be less or equal to the batch number of "
                             f"samples {n_batch} for the first partial_fit call.")
        chain.step(n_batch, mean_b, gram_b)
        self.n_samples_seen_ = np.int64(chain.n_seen)
        self._host = None

    def partial_fit(self, X, y=None):
        self.merge(*self.batch_stats(X))
        return self

    # -- sklearn attribute names (host copies, fetched lazily) ---------------------------------------
    def device_attributes(self):
        return self._chain.export()

    def _export(self):
        if self._chain is None:
            raise AttributeError("IncrementalPCA is not fitted yet")
        if self._host is None:
            self._host = {k: v.cpu().numpy() for k, v in self._chain.export().items()}
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

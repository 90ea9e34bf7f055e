"""``InstrumentedModel`` -- the activation read-back hook of the hot path.

Mirror of the surface of /root/reference/netdissect/nethook.py:15-240 that GANSpace uses
(``decomposition.py:182-185,253-258``, ``visualize.py``, ``interactive.py``): retain a named layer's
output on every forward, optionally edit it (ablation / replacement / offset), undo everything on
``close()``.  Implemented with PyTorch forward hooks (the reference swaps ``layer.forward``); retained
tensors stay on the device -- nothing is copied to the host here.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch


def _as_matching(store: dict, key: str, like: torch.Tensor):
    """Fetch ``store[key]`` as a tensor broadcastable against ``like`` (cached after conversion)."""
    value = store.get(key)
    if value is None:
        return None
    if not torch.is_tensor(value):
        value = torch.from_numpy(np.array(value))
    if value.device != like.device or value.dtype != like.dtype:
        value = value.to(device=like.device, dtype=like.dtype)
    if value.dim() < like.dim():
        # leading batch axis, trailing singleton axes (reference nethook.py:258-263)
        value = value.view((1,) + tuple(value.shape) + (1,) * (like.dim() - value.dim() - 1))
    store[key] = value
    return value


class InstrumentedModel(torch.nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model
        self._retained = OrderedDict()
        self._ablation, self._replacement, self._offset = {}, {}, {}
        self._hooked_layer = {}      # aka -> layer name
        self._handles = {}           # aka -> hook handle

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self.close()

    def forward(self, *inputs, **kwargs):
        return self.model(*inputs, **kwargs)

    # ---- retaining ---------------------------------------------------------------------------
    def retain_layer(self, layername):
        self.retain_layers([layername])

    def retain_layers(self, layernames):
        self.add_hooks(layernames)
        for entry in layernames:
            aka = entry if isinstance(entry, str) else entry[1]
            self._retained.setdefault(aka, None)

    def retained_features(self):
        return OrderedDict(self._retained)

    def retained_layer(self, aka=None, clear=False):
        if aka is None:
            aka = next(iter(self._retained))
        value = self._retained[aka]
        if clear:
            self._retained[aka] = None
        return value

    # ---- editing -----------------------------------------------------------------------------
    def edit_layer(self, layername, ablation=None, replacement=None, offset=None):
        name, aka = (layername, layername) if isinstance(layername, str) else layername
        if ablation is None and replacement is not None:
            ablation = 1.0
        self.add_hooks([(name, aka)])
        if ablation is not None:
            self._ablation[aka] = ablation
        if replacement is not None:
            self._replacement[aka] = replacement
        if offset is not None:
            self._offset[aka] = offset

    def remove_edits(self, layername=None, remove_offset=True, remove_replacement=True):
        if layername is None:
            if remove_replacement:
                self._ablation.clear()
                self._replacement.clear()
            if remove_offset:
                self._offset.clear()
            return
        aka = layername if isinstance(layername, str) else layername[1]
        if remove_replacement:
            self._ablation.pop(aka, None)
            self._replacement.pop(aka, None)
        if remove_offset:
            self._offset.pop(aka, None)

    # ---- hooks -------------------------------------------------------------------------------
    def add_hooks(self, layernames):
        wanted = {}
        for entry in layernames:
            name, aka = (entry, entry) if isinstance(entry, str) else entry
            if self._hooked_layer.get(aka) != name:
                wanted[name] = aka
        if not wanted:
            return
        for name, layer in self.model.named_modules():
            aka = wanted.pop(name, None)
            if aka is not None:
                self._hook_layer(layer, name, aka)
        for name in wanted:
            raise ValueError("Layer %s not found in model" % name)

    def _hook_layer(self, layer, layername, aka):
        if aka in self._hooked_layer or layername in self._hooked_layer.values():
            raise ValueError("Layer %s already hooked" % aka)
        self._hooked_layer[aka] = layername

        def hook(_module, _inputs, output, _aka=aka):
            return self._postprocess_forward(output, _aka)

        self._handles[aka] = layer.register_forward_hook(hook)

    def _postprocess_forward(self, x, aka):
        if aka in self._retained:                   # retained before edits, detached, still on device
            self._retained[aka] = x.detach()
        a = _as_matching(self._ablation, aka, x)
        if a is not None:
            x = x * (1 - a)
            v = _as_matching(self._replacement, aka, x)
            if v is not None:
                x = x + v * a
        b = _as_matching(self._offset, aka, x)
        if b is not None:
            x = x + b
        return x

    def _unhook_layer(self, aka):
        handle = self._handles.pop(aka, None)
        if handle is None:
            return
        handle.remove()
        del self._hooked_layer[aka]
        for store in (self._ablation, self._replacement, self._offset, self._retained):
            store.pop(aka, None)

    def close(self):
        for aka in list(self._handles):
            self._unhook_layer(aka)
        assert not self._handles

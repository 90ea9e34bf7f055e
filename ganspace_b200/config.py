"""``Config`` -- the boundary type of the hot path (mirror of /root/reference/config.py:16-72).

Same flags, destinations and defaults as the reference's argparse definition (config.py:56-69), and
the same two construction modes: ``Config(**overrides)`` (defaults from the parser, then keyword
overrides) and ``Config().from_args(argv)``.
"""
from __future__ import annotations

import argparse
import copy
import json
import sys

# (flag, dest, type, default, action)  -- reference config.py:56-69
_FLAGS = [
    ("--model", "model", str, "StyleGAN", None),
    ("--layer", "layer", str, "g_mapping", None),
    ("--class", "output_class", str, None, None),
    ("--est", "estimator", str, "ipca", None),
    ("--sparsity", "sparsity", float, 1.0, None),
    ("--video", "make_video", None, False, "store_true"),
    ("--batch", "batch_mode", None, False, "store_true"),
    ("-b", "batch_size", int, None, None),
    ("-c", "components", int, 80, None),
    ("-n", "n", int, 300_000, None),
    ("--use_w", "use_w", None, False, "store_true"),
    ("--sigma", "sigma", float, 2.0, None),
    ("--inputs", "inputs", str, None, None),
    ("--seed", "seed", int, None, None),
]


def _parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="GAN component analysis config")
    for flag, dest, typ, default, action in _FLAGS:
        if action:
            p.add_argument(flag, dest=dest, action=action)
        else:
            p.add_argument(flag, dest=dest, type=typ, default=default)
    return p


class Config:
    def __init__(self, **kwargs):
        self.from_args([])
        self.default_args = copy.deepcopy(self.__dict__)
        self.from_dict(kwargs)

    def from_dict(self, dictionary):
        for key, value in dictionary.items():
            setattr(self, key, value)
        return self

    def from_args(self, args=None):
        ns = _parser().parse_args(sys.argv[1:] if args is None else args)
        return self.from_dict(vars(ns))

    def __str__(self):
        defaults = getattr(self, "default_args", {})
        custom, default = {}, {}
        for key, value in self.__dict__.items():
            if key == "default_args":
                continue
            (default if key in defaults and defaults[key] == value else custom)[key] = value
        return json.dumps({"custom": custom, "default": default}, indent=4)

    __repr__ = __str__

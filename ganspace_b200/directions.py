"""Named-direction ``.pkl`` files: the exchange format between a decomposition (.npz) and the reference's GUI.

Mirror of /root/reference/interactive.py:526-578 (``export_direction``: the dict it pickles and the file name it builds) and
:88-127 (``load_named_components``: how such files are selected and turned into edit directions).  The reference does both inside
its Tk application; here they are plain functions so that a decomposition computed on the B200 can be exported to -- and
directions saved by the GUI can be read back from -- ``out/directions/*.pkl`` without the GUI.  Wire format = the reference's:
a pickled dict with the keys of ``KEYS`` (numpy arrays for the two component rows, Python scalars otherwise).
"""
from __future__ import annotations

import glob
import pickle
import string
from pathlib import Path
from types import SimpleNamespace

import numpy as np

KEYS = ("name", "sigma_range", "component_index", "act_comp", "lat_comp", "latent_space", "act_stdev", "lat_stdev", "model_name",
        "output_class", "decomposition", "edit_type", "truncation", "edit_start", "edit_end", "example_seed")


def prettify_name(name: str) -> str:
    """utils.py:19-21: everything outside [-_A-Za-z0-9] becomes '_'."""
    valid = "-_%s%s" % (string.ascii_letters, string.digits)
    return "".join(c if c in valid else "_" for c in name)


def direction_file_ident(params: dict, estimator: str, layer: str, component_class: str) -> str:
    """File stem of interactive.py:557-567."""
    mode = params["edit_type"]
    if mode == "latent":
        mode = params["latent_space"].lower()
    cls = component_class if component_class == params["output_class"] else f"{component_class}_onto_{params['output_class']}"
    return "{model}-{name}-{cls}-{est}-{mode}-{layer}-comp{idx}-range{start}-{end}".format(
        model=params["model_name"], name=prettify_name(params["name"]), cls=cls, est=estimator, mode=mode, layer=layer,
        idx=params["component_index"], start=params["edit_start"], end=params["edit_end"])


def export_direction(npz_path, out_dir, config, component_index: int, name: str, latent_space: str, edit_start: int, edit_end: int,
                     sigma_range: float = 2.0, edit_type: str = "latent", truncation: float = 1.0, example_seed: int = 0,
                     output_class: str = None) -> Path:
    """Write component ``component_index`` of a decomposition ``.npz`` as a named direction (interactive.py:526-571).
    ``edit_end`` is exclusive, as saved by the reference.  ``config``: the Config the decomposition was computed with."""
    with np.load(npz_path, allow_pickle=False) as data:
        act_comp, lat_comp = data["act_comp"][component_index], data["lat_comp"][component_index]
        act_stdev, lat_stdev = float(data["act_stdev"][component_index]), float(data["lat_stdev"][component_index])
    params = {
        "name": name, "sigma_range": float(sigma_range), "component_index": int(component_index),
        "act_comp": np.asarray(act_comp), "lat_comp": np.asarray(lat_comp), "latent_space": latent_space,
        "act_stdev": act_stdev, "lat_stdev": lat_stdev, "model_name": config.model,
        "output_class": output_class or config.output_class,
        "decomposition": {"name": config.estimator, "components": config.components, "samples": config.n, "layer": config.layer,
                          "class_name": config.output_class},
        "edit_type": edit_type, "truncation": float(truncation), "edit_start": int(edit_start), "edit_end": int(edit_end),
        "example_seed": int(example_seed),
    }
    ident = direction_file_ident(params, config.estimator, config.layer, config.output_class)
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    path = out_dir / f"{ident}.pkl"
    with open(path, "wb") as f:
        pickle.dump(params, f)
    return path


def get_edit_name(idx, s, e, name=None):
    """interactive.py's slider title (inclusive layer range)."""
    return "{}{} ({}-{})".format("" if name is None else name + ": ", idx, s, e) if name else f"{idx} ({s}-{e})"


def load_named_components(path, model_name: str, class_name: str, latent_space: str, device=None) -> SimpleNamespace:
    """interactive.py:88-127: every ``*.pkl`` under ``path`` whose model / class / latent space match, as the namespace the
    reference's GUI edits with (X_comp, Z_comp: lists of tensors on ``device`` -- numpy arrays when ``device`` is None)."""
    selected = []
    for dump_path in sorted(glob.glob(f"{path}/*.pkl")):
        with open(dump_path, "rb") as f:
            data = pickle.load(f)
        if data["model_name"] != model_name or data["output_class"] != class_name:
            continue
        if data["latent_space"] != latent_space:
            print("Skipping", dump_path, "(wrong latent space)")
            continue
        selected.append(data)
        print("Using", dump_path)
    if len(selected) == 0:
        raise RuntimeError("No valid components in given path.")
    comp = SimpleNamespace(X_comp=[], Z_comp=[], X_stdev=[], Z_stdev=[], names=[], types=[], layer_names=[], ranges=[], latent_types=[])

    def place(a):
        if device is None:
            return a
        import torch
        return torch.from_numpy(a).to(device)

    for d in selected:
        s, e = d["edit_start"], d["edit_end"]
        comp.X_comp.append(place(d["act_comp"]))
        comp.Z_comp.append(place(d["lat_comp"]))
        comp.X_stdev.append(d["act_stdev"])
        comp.Z_stdev.append(d["lat_stdev"])
        comp.names.append(get_edit_name(d["component_index"], s, e - 1, d["name"]))      # shown inclusive
        comp.types.append(d["edit_type"])
        comp.layer_names.append(d["decomposition"]["layer"])
        comp.ranges.append((s, e))
        comp.latent_types.append(d["latent_space"])
    return comp

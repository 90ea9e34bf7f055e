"""Full-size parity record for BASELINE.json config 2 (N=1e6, b=10k, c=80): runs the CUDA path and the
oracle's restatement of the reference on the same seeds on this box and prints the comparison as JSON
(committed under profiles/).  Uses the oracle, so it lives outside the product package."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ganspace_oracle as orc  # noqa: E402
from ganspace_b200.config import Config  # noqa: E402
from ganspace_b200.decomposition import compute_arrays  # noqa: E402
from ganspace_b200.models import get_instrumented_model, StyleGAN2  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
res = {"n": n, "b": 10_000, "c": 80, "cores": os.cpu_count()}
ws, bs = orc.mapping_random_init(1234)
t0 = time.time()
ref = orc.compute_stylegan2_style(ws, bs, n, 10_000, 80, True)
res["oracle_seconds"] = time.time() - t0
for mapping in ("tc", "simt"):
    os.environ["GANSPACE_B200_MAPPING"] = mapping
    dev = torch.device("cuda:0")
    model = StyleGAN2(dev, "ffhq", random_init=1234)
    inst = get_instrumented_model("StyleGAN2", "ffhq", "style", dev, model=model, use_w=True)
    cfg = Config(model="StyleGAN2", layer="style", output_class="ffhq", components=80, n=n, batch_size=10_000,
                 use_w=True, estimator="ipca")
    sys.stdout = open(os.devnull, "w")
    out = compute_arrays(cfg, inst)
    torch.cuda.synchronize()
    t0 = time.time()
    out = compute_arrays(cfg, inst)
    torch.cuda.synchronize()
    dt = time.time() - t0
    sys.stdout = sys.__stdout__
    a = out["act_comp"].reshape(80, 512).astype(np.float64)
    b = ref["act_comp"].reshape(80, 512).astype(np.float64)
    cos = np.sum(a * b, axis=1)
    res[mapping] = {"seconds": dt, **orc.compare_npz(out, ref), "cos_components_72_79": cos[72:].tolist()}
print(json.dumps(res, indent=1))

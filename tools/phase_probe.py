"""Config-2 run with GANSPACE_B200_TIMING=1 (synchronising phase timers) -- where the wall-clock of one step goes."""
import os, sys, time
os.environ.setdefault("GANSPACE_B200_TIMING", "1")
sys.path.insert(0, ".")
import torch
from ganspace_b200 import _native, decomposition
from ganspace_b200.config import Config
from ganspace_b200.models import get_instrumented_model, StyleGAN2
dev = torch.device("cuda:0")
model = StyleGAN2(dev, "ffhq", random_init=1234)
inst = get_instrumented_model("StyleGAN2", "ffhq", "style", dev, model=model, use_w=True)
cfg = lambda: Config(model="StyleGAN2", layer="style", output_class="ffhq", components=80, n=1_000_000, batch_size=10_000, use_w=True, estimator="ipca")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if _native.instrument.timeline is not None:
        _native.instrument.timeline = []
        _native.instrument.mark("start")
    decomposition.compute_arrays(cfg(), inst)
    torch.cuda.synchronize(); print(f"== rep {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
if _native.instrument.timeline:
    for name, ms in _native.instrument.timeline_ms():
        print(f"   {ms:8.3f} ms  {name}")

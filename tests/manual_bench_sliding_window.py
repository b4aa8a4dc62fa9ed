"""Manual GPU measurement (not collected by pytest): whole-volume sliding-window prediction, product vs the reference's
schedule.  Usage on the GPU box:  python tests/manual_bench_sliding_window.py [--model dinounet_l] [--slices 4] [--hw 1024]
  product   : dinounet_b200.SlidingWindowPredictor (resident volume, batched tiles x mirror variants, fused accumulate)
  reference : the oracle port run by torch eager on the same GPU in the reference's precision regime, driven by the
              restated reference loop (batch 1, one forward per mirror variant) — predict_from_raw_data.py:572-621.
Prints one JSON line -> gpurun_out/sliding_window_bench.json."""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DINOUNET_B200_ALLOW_RANDOM_BACKBONE", "1")
import dinounet_b200  # noqa: E402
from dinounet_b200 import config  # noqa: E402
from oracle import dinounet_oracle as O  # noqa: E402
from oracle import sliding_window_oracle as SWO  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="dinounet_l")
ap.add_argument("--slices", type=int, default=4)
ap.add_argument("--hw", type=int, default=1024)
ap.add_argument("--tile-batch", type=int, default=8)
ap.add_argument("--ref-slices", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sd = O.make_state_dict(a.model, 2, seed=0)
net = dinounet_b200.DinoUNet.from_config({"architecture": dict(config.DEFAULT_ARCHITECTURE)}, 3, 2, None, a.model)
net.load_state_dict(sd, strict=True)
net = net.to(dev).eval()
x = torch.randn(3, a.slices, a.hw, a.hw, generator=torch.Generator().manual_seed(0)).pin_memory()
p = dinounet_b200.SlidingWindowPredictor(tile_step_size=0.5, use_gaussian=True, use_mirroring=True, device=dev,
                                         tile_batch=a.tile_batch)
p.manual_initialization(net, None, SimpleNamespace(patch_size=[512, 512]), None, {}, "DinoUNetTrainer", (0, 1))
n_tiles = len(p._internal_get_sliding_window_slicers((a.slices, a.hw, a.hw)))
p.predict_sliding_window_return_logits(x[:, :1])          # plan + graph build
torch.cuda.synchronize()
times = []
for _ in range(3):
    t0 = time.perf_counter()
    y = p.predict_sliding_window_return_logits(x)
    y_host = y.cpu()                                        # the caller's D2H of the fp16 logits
    times.append(time.perf_counter() - t0)
prod_s = min(times)

sd_dev = {k: v.to(dev) for k, v in sd.items()}


def ref_net(t):
    return O.forward(sd_dev, a.model, t, autocast_like_reference=True).half()


xr = x[:, :a.ref_slices]
SWO.predict_sliding_window_return_logits(ref_net, xr[:, :, :512, :512], (512, 512), 2, 0.5, True, (0, 1), results_device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
yr = SWO.predict_sliding_window_return_logits(ref_net, xr, (512, 512), 2, 0.5, True, (0, 1), results_device=dev).cpu()
ref_s = time.perf_counter() - t0
ref_tiles = n_tiles * a.ref_slices // a.slices
agree = float((y_host[:, :a.ref_slices].argmax(0) == yr.argmax(0)).float().mean()) if torch.isfinite(yr.float()).all() else None
out = {"model": a.model, "volume": [3, a.slices, a.hw, a.hw], "tiles": n_tiles, "forwards": n_tiles * 4,
       "product_s": prod_s, "product_forwards_per_s": n_tiles * 4 / prod_s, "product_tiles_per_s": n_tiles / prod_s,
       "reference_schedule_eager_s": ref_s, "reference_tiles": ref_tiles,
       "reference_forwards_per_s": ref_tiles * 4 / ref_s, "speedup_per_tile": (ref_s / ref_tiles) / (prod_s / n_tiles),
       "argmax_agreement_on_compared_slices": agree, "includes": "H2D of the volume, D2H of fp16 logits"}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sliding_window_bench.json", "w"), indent=1)
print(json.dumps(out))

"""One-off full-size dinounet_7b check on a GPU box: (1) parity of the kernel path against the CPU fp32 oracle at 256^2,
(2) forward throughput at 512^2, batch 16 (the per-GPU shard of BASELINE config 5).  Needs ~60 GB host RAM."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DINOUNET_B200_ALLOW_RANDOM_BACKBONE", "1")
import dinounet_b200  # noqa: E402
from dinounet_b200 import config, lib  # noqa: E402
from oracle import dinounet_oracle as O  # noqa: E402

model = "dinounet_7b"
t0 = time.time()
sd = O.make_state_dict(model, 2, seed=0)
print(f"state dict: {sum(v.numel() for k, v in sd.items() if not k.startswith('decoder.encoder.') and 'all_modules' not in k) / 1e9:.2f} B params, {time.time() - t0:.0f} s", flush=True)
net = dinounet_b200.DinoUNet.from_config({"architecture": dict(config.DEFAULT_ARCHITECTURE)}, 3, 2, None, model)
net.load_state_dict(sd, strict=True)
net = net.to("cuda").eval()
print(f"model on GPU, {time.time() - t0:.0f} s, {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
res = {}
if "--no-parity" not in sys.argv:
    x = O.make_input(1, 256, 0)
    with torch.no_grad():
        y = net(x.cuda()).float().cpu()
    torch.set_num_threads(min(32, os.cpu_count()))
    t1 = time.time()
    ref = O.forward(sd, model, x)
    res["oracle_cpu_s"] = time.time() - t1
    err = ((y - ref).abs().max() / ref.abs().max()).item()
    flips = int((y.argmax(1) != ref.argmax(1)).sum())
    res.update(rel_err=err, flips=flips, pixels=ref[:, 0].numel(), finite=bool(torch.isfinite(y).all()))
    print("parity", res, flush=True)
del sd
B, S = 16, 512
eng = net._get_engine(torch.device("cuda", 0))
xs = O.make_input(B, S, 1).cuda()
with torch.no_grad():
    for _ in range(3):
        eng.forward(xs, use_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        eng.forward(xs, use_graph=True)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
res.update(batch=B, ms_per_step=ms, patches_per_s=B / ms * 1e3, model_tflops=B / ms * 1e3 * O.algorithmic_flops_per_patch(model, S) / 1e12,
           mem_gb=torch.cuda.max_memory_allocated() / 1e9, kernels=len(eng.get_plan(B, S)[0]))
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/check_7b.json", "w"))

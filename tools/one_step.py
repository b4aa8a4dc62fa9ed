"""One eager (no CUDA graph) forward of the launch plan after one warm-up step: the target of the per-kernel ncu passes.
    ncu --metrics ... -k regex:b2u -s <kernels_per_step> -c <kernels_per_step> python tools/one_step.py dinounet_l 32 512
Prints the number of kernels per step on the first line when called with --count."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DINOUNET_B200_ALLOW_RANDOM_BACKBONE", "1")
import dinounet_b200  # noqa: E402
from dinounet_b200 import config  # noqa: E402
from oracle import dinounet_oracle as O  # noqa: E402  (synthetic weights / inputs only)

model = sys.argv[1] if len(sys.argv) > 1 else "dinounet_l"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 512
sd = O.make_state_dict(model, 2, seed=0)
net = dinounet_b200.DinoUNet.from_config({"architecture": dict(config.DEFAULT_ARCHITECTURE)}, 3, 2, None, model)
net.load_state_dict(sd, strict=True)
net = net.cuda().eval()
eng = net._get_engine(torch.device("cuda", 0))
plan, bufs = eng.get_plan(B, S)
if "--count" in sys.argv:
    print(len(plan.calls))
    names = [n for n, _, _ in plan.calls]
    import json
    json.dump(names, open("gpurun_out/plan_names.json", "w"))
    sys.exit(0)
x = O.make_input(B, S, 5).cuda()
with torch.no_grad():
    for _ in range(2):
        eng.forward(x, use_graph=False)
torch.cuda.synchronize()
print("done", len(plan.calls))

"""Two training steps (warm-up + one) for per-kernel ncu timing of the training path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DINOUNET_B200_ALLOW_RANDOM_BACKBONE", "1")
import dinounet_b200
from dinounet_b200 import config
from dinounet_b200.loss import DC_and_CE_loss
from dinounet_b200.train_path import FusedSGD, train_step
from oracle import dinounet_oracle as O
model = sys.argv[1] if len(sys.argv) > 1 else "dinounet_s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sd = O.make_state_dict(model, 2, seed=0)
net = dinounet_b200.DinoUNet.from_config({"architecture": dict(config.DEFAULT_ARCHITECTURE)}, 3, 2, None, model)
net.load_state_dict(sd, strict=True)
net = net.cuda().train()
crit = DC_and_CE_loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {}, weight_ce=1, weight_dice=1)
opt = FusedSGD(net.parameters(), lr=1e-2)
x = O.make_input(B, 512, 1).cuda()
t = torch.randint(0, 2, (B, 1, 512, 512)).float().cuda()
for _ in range(2):
    train_step(net, crit, opt, x, t)
torch.cuda.synchronize()
print("done")

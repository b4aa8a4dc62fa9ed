"""Bring-up probe used in round 1: halo conv (opt 0) vs the per-tap walk (opt 1).  Result recorded in gemm_tc2.cu:
shifted-window descriptors need no base_offset on B200 (the variant with base_offset=(addr>>7)&7 produced garbage)."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dinounet_b200 import lib as L
from tests.gpu_helpers import gemm, rel_err
from tests.test_gpu_kernels import _pack_conv, _rand
lib = L.load()
for (cin, cout, hw, B) in [(64, 32, 256, 1), (64, 64, 128, 2)]:
    td = torch.float16
    x = _rand(B, hw, hw, cin, dt=td); w = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=1); bias = _rand(cout, seed=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(td).float(), bias, stride=1, padding=1)
    for opt in (1, 0):
        lib.b2u_set_option(2, opt)
        out = torch.full((B * hw * hw, cout), float("nan"), device="cuda", dtype=td)
        gemm(x.view(-1, cin), _pack_conv(w, td), out, L.F16, M=0, K=9 * cin, lda=cin, bias=bias, conv=L.CONV3X3_S1, img=(B, hw, hw, cin))
        torch.cuda.synchronize()
        got = out.view(B, hw, hw, cout).permute(0, 3, 1, 2).float()
        e = (got - ref).abs()
        print((cin, cout, hw, B), "opt", opt, "rel err", rel_err(got, ref), "bad px frac", (e.amax(1) > 0.05).float().mean().item(),
              "err by x%8:", [round(e[..., i::8].max().item(), 3) for i in range(8)])
lib.b2u_set_option(2, 0)

"""Three launches of the tf32 GEMM for an `ncu --set full` capture: a 768-wide linear (row-mode staging, BN 128), a 3x3
convolution forward 64 -> 32 at 512^2 (window on the A side, BN 32) and its weight gradient (block-mode staging, split-K)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dinounet_b200 import lib as L          # noqa: E402
from dinounet_b200 import train_path as TP  # noqa: E402

dev = "cuda"
M, N, K = 2 * 8 * 5376, 768, 768
x, W, y = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev) * 0.03, torch.empty(M, N, device=dev)
B, H, Cin, Cout = 2, 512, 64, 32
npix = B * H * H
xi, Wp = torch.randn(npix, Cin, device=dev), torch.randn(Cout, 9 * Cin, device=dev) * 0.05
yo, dWp = torch.empty(npix, Cout, device=dev), torch.zeros(Cout, 9 * Cin, device=dev)
for _ in range(2):   # launches 0-2 warm, 3-5 the ones to read
    TP._gemm(x, W, y, M, N, K, tier="tf32")
    TP._gemm(xi, Wp, yo, npix, Cout, 9 * Cin, conv=L.CONV3X3_S1, img=(H, H, Cin), cpad=Cin, tier="tf32")
    TP._gemm(yo, xi, dWp, Cout, 9 * Cin, npix, a_trans=1, lda=Cout, w_mode=3, conv=L.CONV3X3_S1, img=(H, H, Cin), cpad=Cin,
             ksplit=max(1, min(256, npix // 4096)), tier="tf32")
torch.cuda.synchronize()
print("done")

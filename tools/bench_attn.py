"""Timing of the tcgen05 attention kernel (and its single-pass softmax variant, option 4 = 4) at the bench shape (dinounet_l, B=32: 512 (batch, head) pairs x 1029 tokens)
and at the cfg-5 sweep points (head_dim 64 / 128, N 261 / 1029, several batches): CUDA events on the launch stream,
inputs rotated so that no launch starts on a warm L2 copy of its own K/V.
    python tools/bench_attn.py [out.json]
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dinounet_b200 import lib as L  # noqa: E402
from tests.gpu_helpers import P, stream  # noqa: E402

lib = L.load()
dev = "cuda"
bf = torch.bfloat16
peak = 1416.7
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("bf16_tflops_sustained", peak)


def make(B, H, N, hd, seed):
    g = torch.Generator().manual_seed(seed)
    q, k, v = ((torch.randn(B, H, N, hd, generator=g)).to(dev).to(bf) for _ in range(3))
    npad = (N + 7) // 8 * 8
    vt = torch.zeros(B, H, hd, npad, device=dev, dtype=bf)
    vt[..., :N] = v.transpose(2, 3)
    return q, k, v, vt, npad


def run(q, k, vt, o, B, H, N, npad, hd):
    if hd == 64:
        L.check(lib.b2u_attention_tc(P(q), P(k), P(vt), P(o), B, H, N, npad, 0, hd ** -0.5, L.BF16, stream()), "attn")
    else:
        L.check(lib.b2u_attention_tc_hd(P(q), P(k), P(vt), P(o), B, H, N, npad, hd, hd ** -0.5, L.BF16, stream()), "attn")


def time_case(B, H, N, hd, opt, reps=12):
    lib.b2u_set_option(4, opt)
    sets = [make(B, H, N, hd, s) for s in range(3)]
    o = torch.empty(B, N, H * hd, device=dev, dtype=bf)
    for i in range(3):
        q, k, v, vt, npad = sets[i % 3]
        run(q, k, vt, o, B, H, N, npad, hd)
    torch.cuda.synchronize()
    q, k, v, vt, npad = sets[0]
    run(q, k, vt, o, B, H, N, npad, hd)
    nb = min(B, 2)
    ref = F.scaled_dot_product_attention(q[:nb].float(), k[:nb].float(), v[:nb].float()).transpose(1, 2).reshape(nb, N, H * hd)
    err = ((o[:nb].float() - ref).abs().max() / ref.abs().max()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        q, k, v, vt, npad = sets[i % 3]
        run(q, k, vt, o, B, H, N, npad, hd)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    tf = 4.0 * N * N * hd * B * H / (us * 1e-6) / 1e12
    lib.b2u_set_option(4, 0)
    return {"B": B, "heads": H, "N": N, "head_dim": hd, "kernel": {0: "gen3 (default: every 3rd exp2 on the FMA pipe)", 4: "gen3-onepass", 6: "gen3 (all MUFU)", 7: "gen3 (every 4th)"}[opt], "us": round(us, 1),
            "tflops": round(tf, 1), "frac_of_sustained_bf16_peak": round(tf / peak, 3), "rel_err_vs_sdpa": err}


rows = []
for opt in (6, 7, 0):
    rows.append(time_case(32, 16, 1029, 64, opt))
    print(rows[-1], flush=True)
if "--sweep" in sys.argv:
    for hd, H in ((64, 16), (128, 32)):
        for N in (261, 1029):
            for B in (4, 16, 32):
                if hd == 128 and B == 32:
                    continue
                rows.append(time_case(B, H, N, hd, 0, reps=6))
                print(rows[-1], flush=True)
out = [a for a in sys.argv[1:] if a.endswith(".json")]
if out:
    json.dump(rows, open(out[0], "w"), indent=1)

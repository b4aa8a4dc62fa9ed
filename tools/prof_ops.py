"""Standalone launches of the hot kernels at dinounet_l / B=32 / 512^2 shapes, for ncu captures:
   ncu --set full --import-source on -k regex:<kernel> ... python tools/prof_ops.py <op> [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dinounet_b200 import lib as L  # noqa: E402
from tests.gpu_helpers import P, gemm, stream  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "fc1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = L.load()
dev = "cuda"
T, D, Hd = 32 * 1029, 1024, 4096
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
rnd = lambda *s, dt=bf, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev).to(dt)

if op in ("fc1", "fc1_noact", "fc2", "proj"):
    if op == "fc1_noact":
        A, W, bias = rnd(T, D), rnd(Hd, D, sc=D ** -0.5), rnd(Hd, dt=torch.float32)
        out = torch.empty(T, Hd, device=dev, dtype=bf)
        run = lambda: gemm(A, W, out, L.BF16, bias=bias)
    elif op == "fc1":
        A, W, bias = rnd(T, D), rnd(Hd, D, sc=D ** -0.5), rnd(Hd, dt=torch.float32)
        out = torch.empty(T, Hd, device=dev, dtype=bf)
        run = lambda: gemm(A, W, out, L.BF16, bias=bias, act1=L.ACT_GELU)
    elif op == "fc2":
        A, W, bias = rnd(T, Hd), rnd(D, Hd, sc=Hd ** -0.5), rnd(D, dt=torch.float32)
        X, ls = rnd(T, D, dt=torch.float32), rnd(D, dt=torch.float32)
        run = lambda: gemm(A, W, X, L.BF16, out_fp32=True, bias=bias, scale=ls, residual=X, ldres=D)
    else:
        A, W, bias = rnd(T, D), rnd(D, D, sc=D ** -0.5), rnd(D, dt=torch.float32)
        X, ls = rnd(T, D, dt=torch.float32), rnd(D, dt=torch.float32)
        run = lambda: gemm(A, W, X, L.BF16, out_fp32=True, bias=bias, scale=ls, residual=X, ldres=D)
elif op in ("outproj", "ffn2"):   # extractor GEMMs into the fp32 query stream (dinounet_l, B=32: 172032 query rows)
    Mq, Kq = 32 * 5376, (512 if op == "outproj" else 256)
    A, W, bias = rnd(Mq, Kq, dt=torch.float16), rnd(D, Kq, dt=torch.float16, sc=Kq ** -0.5), rnd(D, dt=torch.float32)
    X = rnd(Mq, D, dt=torch.float32)
    run = lambda: gemm(A, W, X, L.F16, out_fp32=True, bias=bias, residual=X, ldres=D)
elif op == "qkv":
    import ctypes as C
    from oracle import dinounet_oracle as O
    A = rnd(T, D)
    W = rnd(3 * D, D, sc=D ** -0.5)
    bias = rnd(3 * D, dt=torch.float32)
    periods = 100.0 ** (2 * torch.arange(16, dtype=torch.float32) / 32)
    sin, cos = [t.to(dev).contiguous() for t in O.rope_sincos(periods, 32, 32)]
    q, k = (torch.empty(32, 16, 1029, 64, device=dev, dtype=bf) for _ in range(2))
    vt = torch.zeros(32, 16, 64, 1032, device=dev, dtype=bf)
    p = L.QkvParams()
    p.B, p.ntok, p.D, p.heads, p.prefix = 32, 1029, D, 16, 5
    p.A, p.lda, p.Wp, p.ldw, p.bias = P(A), D, P(W), D, P(bias)
    p.rope_sin, p.rope_cos, p.q, p.k, p.v, p.dtype = P(sin), P(cos), P(q), P(k), P(vt), L.BF16
    p.v_transposed, p.npad, p.rope_w = 1, 1032, 32
    run = lambda: L.check(lib.b2u_qkv_rope(C.byref(p), stream()), "qkv")
elif op == "attn_tc":
    q, k = (rnd(32, 16, 1029, 64) for _ in range(2))
    vt = torch.zeros(32, 16, 64, 1032, device=dev, dtype=bf)
    vt[..., :1029] = rnd(32, 16, 64, 1029)
    o = torch.empty(32, 1029, 1024, device=dev, dtype=bf)
    run = lambda: L.check(lib.b2u_attention_tc(P(q), P(k), P(vt), P(o), 32, 16, 1029, 1032, 0, 0.125, L.BF16, stream()), "attn_tc")
elif op == "conv":   # decoder 512^2 conv 64 -> 32
    x = rnd(32 * 512 * 512, 64, dt=torch.float16)
    w = rnd(32, 9 * 64, dt=torch.float16, sc=0.05)
    bias = rnd(32, dt=torch.float32)
    out = torch.empty(32 * 512 * 512, 32, device=dev, dtype=torch.float16)
    run = lambda: gemm(x, w, out, L.F16, M=0, K=9 * 64, lda=64, bias=bias, conv=L.CONV3X3_S1, img=(32, 512, 512, 64))
elif op == "msda":
    Bm, Hv, heads, dh = 32, 32, 16, 32
    Lq = Hv * Hv * 21 // 4
    value = rnd(Bm, Hv * Hv, heads, dh, dt=torch.float16)
    offaw = torch.cat([rnd(Bm * Lq, heads * 8, dt=torch.float32, sc=2.0), rnd(Bm * Lq, heads * 4, dt=torch.float32)], 1).contiguous()
    out = torch.empty(Bm * Lq, heads * dh, device=dev, dtype=torch.float16)
    run = lambda: L.check(lib.b2u_msda_forward(P(value), P(offaw), P(out), Bm, Hv, Hv, heads, dh, 4, L.F16, stream()), "msda")
else:
    raise SystemExit("unknown op")

if len(sys.argv) > 3:
    lib.b2u_set_option(int(sys.argv[3]), int(sys.argv[4]))
for _ in range(reps):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{op}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us / launch")

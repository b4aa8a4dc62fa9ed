"""CPU: the reference formulas of tests/test_gpu_tf32_gemm.py, checked without a GPU.

The GPU file compares `b2u_tf32_gemm` with fp64 torch references (conv2d / conv2d_input / conv2d_weight / conv_transpose2d
on TF32-rounded operands).  Here the same test bodies run against a slow torch EMULATION of the b2u_f32_gemm_params
addressing (include/dinounet_b200.h: a_trans / w_mode 1-3 / conv windows / row remaps / pixel shuffle / split-K /
accumulate), written from the header's description, so a wrong permute or stride in a reference formula shows up here
instead of costing a GPU run.  It also pins the bit-level TF32 rounding emulation."""
import numpy as np
import torch
import torch.nn.functional as F

import tests.test_gpu_tf32_gemm as G
from dinounet_b200 import lib as L


def _act(v, act):
    if act == L.ACT_GELU:
        return F.gelu(v)
    if act == L.ACT_RELU:
        return F.relu(v)
    if act == L.ACT_LRELU:
        return F.leaky_relu(v, 0.01)
    return v


def _window(img, Hin, Win, Cc, stride, pix, tap, c):
    """element (tap, c) of the 3x3 / pad 1 window of output pixel `pix` (flat over batch, Ho, Wo); broadcasting index tensors"""
    Ho, Wo = Hin // stride, Win // stride
    cb, r = pix // (Ho * Wo), pix % (Ho * Wo)
    cy, cx = r // Wo, r % Wo
    iy, ix = cy * stride + tap // 3 - 1, cx * stride + tap % 3 - 1
    ok = (c < Cc) & (iy >= 0) & (iy < Hin) & (ix >= 0) & (ix < Win)
    idx = ((cb * Hin + iy) * Win + ix) * Cc + c
    return torch.where(ok, img.reshape(-1)[idx.clamp(0, img.numel() - 1)], torch.zeros((), dtype=img.dtype))


def emulate(fn="b2u_tf32_gemm", **f):
    g = lambda k, d=0: f.get(k, d)
    A, W, out = f["A"], f["W"], f["out"]
    M, N, K = f["M"], f["N"], f["K"]
    conv, w_mode = g("conv"), g("w_mode")
    stride = 2 if conv == L.CONV3X3_S2 else 1
    m = torch.arange(M).view(M, 1)
    n = torch.arange(N).view(N, 1)
    k = torch.arange(K).view(1, K)
    Af, Wf = A.reshape(-1), W.reshape(-1)
    if g("a_trans"):
        Ap = Af[k * f["lda"] + m]
    elif not conv or w_mode == 3:
        arow = m
        if g("a_rows_in") > 0:
            arow = (m // f["a_rows_in"]) * f["a_rows_out"] + f["a_row_off"] + m % f["a_rows_in"]
        Ap = Af[arow * f["lda"] + k]
    else:
        Ap = _window(A, f["Hin"], f["Win"], f["C"], stride, m, k // f["Cpad"], k % f["Cpad"])
    if w_mode == 0:
        Wp = Wf[n * f["ldw"] + k]
    elif w_mode == 1:
        Wp = Wf[k * f["ldw"] + n]
    elif w_mode == 2:
        tap, nn = k // f["Cpad"], k % f["Cpad"]
        idx = nn * f["ldw"] + (8 - tap) * f["w_cpad"] + n
        Wp = torch.where(nn < f["C"], Wf[idx.clamp(0, Wf.numel() - 1)], torch.zeros(()))
    else:
        Wp = _window(W, f["Hin"], f["Win"], f["C"], stride, k, n // f["Cpad"], n % f["Cpad"])
    v = G.tf32(Ap) @ G.tf32(Wp).t()
    if g("bias", None) is not None:
        v = v + f["bias"].double()
    v = _act(v, g("act1"))
    if g("scale", None) is not None:
        v = v * f["scale"].double()
    if g("shift", None) is not None:
        v = v + f["shift"].double()
    v = _act(v, g("act2"))
    orow, oc = m.expand(M, N), n.view(1, N).expand(M, N)
    if g("ps_cout") > 0:
        ph, pw, pc = f["ps_h"], f["ps_w"], f["ps_cout"]
        pb, rem = m // (ph * pw), m % (ph * pw)
        base = (pb * (2 * ph) + 2 * (rem // pw)) * (2 * pw) + 2 * (rem % pw)
        q = n.view(1, N) // pc
        orow, oc = base + (q >> 1) * (2 * pw) + (q & 1), n.view(1, N) % pc + torch.zeros_like(base)
    elif g("rows_in") > 0:
        orow = ((m // f["rows_in"]) * f["rows_out"] + f["row_off"] + m % f["rows_in"]).expand(M, N)
    oc = oc + g("col_off")
    if g("residual", None) is not None:
        v = v + f["residual"].reshape(-1)[orow * f["ldres"] + oc].double()
    flat = out.view(-1)
    idx = (orow * f["ldc"] + oc).reshape(-1)
    if g("ksplit") > 1 or g("accumulate"):
        flat.index_add_(0, idx, v.reshape(-1).float())
    else:
        flat[idx] = v.reshape(-1).float()


def _cases(fn):
    for mk in getattr(fn, "pytestmark", []):
        if mk.name == "parametrize":
            names = [s.strip() for s in mk.args[0].split(",")]
            return [dict(zip(names, vals)) for vals in mk.args[1]]
    return [{}]


def test_reference_formulas_of_the_gpu_test_agree_with_the_parameter_block_semantics(monkeypatch):
    monkeypatch.setattr(G, "DEV", "cpu")
    monkeypatch.setattr(G, "raw_gemm", emulate)
    monkeypatch.setattr(G, "TOL", 1e-5)
    ran = 0
    for fn in (G.test_plain_rows_bias_activation_residual, G.test_unaligned_leading_dimensions_take_the_scalar_paths,
               G.test_transposed_operands_and_split_k, G.test_conv3x3_forward_data_gradient_weight_gradient, G.test_conv3x3_channel_padding_rows_are_zero,
               G.test_pixel_shuffle_and_row_remap_epilogues):
        for kw in _cases(fn):
            if kw.get("M", 0) * kw.get("N", 0) * kw.get("K", 0) > 2e9:      # keep the CPU suite short
                continue
            fn(**kw)
            ran += 1
    assert ran >= 20


def test_tf32_rounding_emulation_is_round_to_nearest_ties_away_on_10_mantissa_bits():
    x = torch.tensor([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 - 2.0 ** -23, 1.0 + 2.0 ** -10, -1.0 - 2.0 ** -11, 3.14159265, 0.0, -0.0, 1e-30])
    r = G.tf32(x)
    assert r[0] == 1.0 and r[1] == 1.0 + 2.0 ** -10 and r[2] == 1.0 and r[3] == 1.0 + 2.0 ** -10 and r[4] == -1.0 - 2.0 ** -10
    bits = r.float().view(torch.int32).numpy()
    assert (bits & 0x1FFF == 0).all()
    assert abs(r[5].item() - 3.14159265) <= 3.14159265 * 2.0 ** -11
    y = torch.randn(10000)
    assert ((G.tf32(y) - y.double()).abs() <= y.abs().double() * 2.0 ** -11 + 1e-45).all()
    assert np.signbit(r[7].item()) and r[6] == 0

"""CPU: the matrix-product tier plumbing of dinounet_b200/train_path.py with a recording stand-in for the native library.
Each autograd Function must call the tier that was active at ITS forward in its backward too (the global tier is only the
default of new Functions), `tier=` overrides win, and `trainable_forward` validates / restores the tier."""
import contextlib

import pytest
import torch

from dinounet_b200 import lib as L
from dinounet_b200 import train_path as TP


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*a):
            self.calls.append(name)
            return 0
        return fn


@pytest.fixture
def rec(monkeypatch):
    r = _Recorder()
    monkeypatch.setattr(L, "load", lambda: r)
    monkeypatch.setattr(TP, "_s", lambda t: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: type("S", (), {"cuda_stream": 0})())
    return r


def _gemms(r):
    return [c for c in r.calls if c.endswith("_gemm")]


@pytest.mark.parametrize("tier,fn", [("tf32", "b2u_tf32_gemm"), ("fp32", "b2u_f32_gemm")])
def test_functions_remember_the_tier_of_their_forward(rec, monkeypatch, tier, fn):
    other = "fp32" if tier == "tf32" else "tf32"
    monkeypatch.setattr(TP, "_GEMM_TIER", tier)
    x = torch.randn(64, 8, requires_grad=True)
    W = torch.randn(4, 8, requires_grad=True)
    b = torch.randn(4, requires_grad=True)
    y = TP.LinearF.apply(x, W, b, None)
    xc = torch.randn(2 * 8 * 8, 4, requires_grad=True)
    Wc = torch.randn(6, 4, 3, 3, requires_grad=True)
    yc = TP.Conv3x3F.apply(xc, Wc, None, 2, 8, 8, 1)
    xt = torch.randn(2 * 4 * 4, 4, requires_grad=True)
    Wt = torch.randn(4, 3, 2, 2, requires_grad=True)
    bt = torch.randn(3, requires_grad=True)
    yt = TP.ConvT2x2F.apply(xt, Wt, bt, 2, 4, 4)
    assert _gemms(rec) == [fn] * 3
    monkeypatch.setattr(TP, "_GEMM_TIER", other)           # the default of NEW Functions changes ...
    rec.calls.clear()
    (y.sum() + yc.sum() + yt.sum()).backward()
    g = _gemms(rec)
    assert len(g) == 6 and set(g) == {fn}                  # ... the backward of the old ones does not: dx + dW for each
    assert "b2u_f32_colsum" in rec.calls and "b2u_f32_unshuffle" in rec.calls


def test_explicit_tier_overrides_the_default(rec, monkeypatch):
    monkeypatch.setattr(TP, "_GEMM_TIER", "tf32")
    A, W, out = torch.zeros(4, 4), torch.zeros(4, 4), torch.zeros(4, 4)
    TP._gemm(A, W, out, 4, 4, 4, tier="fp32")
    TP._gemm(A, W, out, 4, 4, 4)
    assert _gemms(rec) == ["b2u_f32_gemm", "b2u_tf32_gemm"]


def test_trainable_forward_validates_and_restores_the_tier(monkeypatch):
    monkeypatch.setattr(TP, "_GEMM_TIER", "fp32")
    seen = []
    monkeypatch.setattr(TP, "_trainable_forward", lambda *a: seen.append(TP._GEMM_TIER) or "logits")
    assert TP.trainable_forward({}, "dinounet_s", None, [], 2, gemm="tf32") == "logits"
    assert seen == ["tf32"] and TP._GEMM_TIER == "fp32"
    with pytest.raises(ValueError):
        TP.trainable_forward({}, "dinounet_s", None, [], 2, gemm="bf16")
    monkeypatch.setattr(TP, "_trainable_forward", lambda *a: (_ for _ in ()).throw(RuntimeError("boom")))
    with pytest.raises(RuntimeError):
        TP.trainable_forward({}, "dinounet_s", None, [], 2, gemm="tf32")
    assert TP._GEMM_TIER == "fp32"                         # restored on the error path too

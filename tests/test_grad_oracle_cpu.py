"""Gradient oracle (oracle/grad_oracle.py = autograd through the oracle forward + loss oracle), the round-2 backward
target: pinned against autograd through the REAL reference module + REAL reference loss (build container only), and
against committed golden gradient norms / samples (any container)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dinounet_oracle as O
from oracle import grad_oracle as G
from oracle.ref_loader import reference_available


def _case(model, B, S, ncls, seed):
    sd = O.make_state_dict(model, ncls, seed=seed)
    x = O.make_input(B, S, seed)
    target = torch.randint(0, ncls, (B, 1, S, S), generator=torch.Generator().manual_seed(seed + 7)).float()
    return sd, x, target


def test_grad_oracle_matches_golden():
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "grads_*.npz")))
    assert files
    for f in files:
        model, b, s, c, w = os.path.basename(f)[len("grads_"):-4].rsplit("_", 4)
        g = np.load(f)
        sd, x, target = _case(model, int(b[1:]), int(s[1:]), int(c[1:]), int(w[1:]))
        loss, grads = G.loss_and_grads(sd, model, x, target)
        assert abs(loss.item() - float(g["loss"])) < 1e-6
        names = [str(n) for n in g["names"]]
        assert names == sorted(grads)
        norms = np.array([grads[k].double().norm().item() for k in names])
        assert np.allclose(norms, g["norms"], rtol=1e-4, atol=1e-9)
        for i, k in enumerate(names):
            fl = grads[k].reshape(-1)
            samp = fl[:: max(1, fl.numel() // 16)][:16].numpy()
            assert np.allclose(samp, g[f"s{i}"], rtol=1e-3, atol=1e-7), k


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_grad_oracle_equals_reference_autograd():
    from oracle.ref_loader import build_reference_model, load_reference_module
    from oracle import loss_oracle as LO
    import sys
    model, ncls = "dinounet_s", 2
    sd, x, target = _case(model, 1, 128, ncls, 3)
    net = build_reference_model(model, ncls, sd)          # eval mode: BN running stats, DropPath off
    load_reference_module()
    msda_mod = sys.modules["dinounet.dinov3.eval.segmentation.models.utils.ms_deform_attn"]

    class _Differentiable:                                 # the extension-backed backward cannot run on CPU (see module doc)
        @staticmethod
        def apply(value, shapes, lsi, loc, aw, step):
            return msda_mod.ms_deform_attn_core_pytorch(value, shapes, loc, aw)

    orig = msda_mod.MSDeformAttnFunction
    msda_mod.MSDeformAttnFunction = _Differentiable
    try:
        DC_and_CE_loss, MemDice, _ = LO.load_reference_loss()
        crit = DC_and_CE_loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {}, weight_ce=1,
                              weight_dice=1, ignore_label=None, dice_class=MemDice)
        loss_ref = crit(net(x), target)
        loss_ref.backward()
    finally:
        msda_mod.MSDeformAttnFunction = orig
    loss, grads = G.loss_and_grads(sd, model, x, target)
    assert abs(loss.item() - loss_ref.item()) < 1e-6
    ref = {n: p.grad for n, p in net.named_parameters() if p.requires_grad}
    # the reference's named_parameters() lists each shared Parameter once; every trainable oracle key must be among them
    assert set(ref) == set(grads), (sorted(set(ref) ^ set(grads))[:10])
    worst = 0.0
    for k, gr in ref.items():
        go = grads[k]
        if gr is None:
            assert float(go.abs().max()) == 0.0, k
            continue
        denom = float(gr.abs().max()) + 1e-12
        worst = max(worst, float((go - gr).abs().max()) / denom)
    assert worst < 1e-4, worst

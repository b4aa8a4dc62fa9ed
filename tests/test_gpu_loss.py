"""-m gpu: fused Dice+CE / validation-statistics kernels (csrc/loss.cu through dinounet_b200.loss) against the loss
oracle (bit-identical to the reference classes, tests/test_loss_cpu.py).  Floating point: loss within 2e-6 relative
(fp64 partial sums here vs torch's fp32 tree), gradient within 1e-5 of max|grad|; integer counts exact."""
import numpy as np
import pytest
import torch

from dinounet_b200.loss import DC_and_CE_loss, validation_statistics
from oracle import loss_oracle as LO

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _case(B, C, H, W, seed, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, C, H, W, generator=g) * scale).to(DEV), torch.randint(0, C, (B, 1, H, W), generator=g).to(DEV)


@pytest.mark.parametrize("B,C,H,W", [(2, 2, 512, 512), (3, 4, 96, 200), (1, 3, 33, 17), (4, 16, 64, 64), (2, 5, 128, 128), (2, 7, 64, 96),
                                     (1, 21, 128, 128), (2, 32, 64, 64)])
@pytest.mark.parametrize("batch_dice", [True, False])
def test_loss_and_gradient_match_oracle(B, C, H, W, batch_dice):
    z, t = _case(B, C, H, W, B * 100 + C)
    mod = DC_and_CE_loss({"batch_dice": batch_dice, "smooth": 1e-5, "do_bg": False, "ddp": False}, {}, 1, 1)
    for tt in (t.float(), t, t.to(torch.uint8), t.int()):                 # nnU-Net hands float label maps
        z1 = z.clone().requires_grad_(True)
        got = mod(z1, tt)
        z2 = z.double().requires_grad_(True)
        want, ce, dc = LO.dc_and_ce_loss(z2, tt, batch_dice=batch_dice)
        assert abs(float(got) - float(want)) <= 2e-6 * max(1.0, abs(float(want))), (float(got), float(want))
    (got * 3.0).backward()
    (want * 3.0).backward()
    err = float((z1.grad.double() - z2.grad).abs().max() / z2.grad.abs().max())
    assert err < 1e-5, err
    # fp32 torch (what the reference computes on the GPU) lands in the same place
    w32, _, _ = LO.dc_and_ce_loss(z, t, batch_dice=batch_dice)
    assert abs(float(got) - float(w32)) < 1e-5 * max(1.0, abs(float(w32)))


def test_weights_do_bg_and_determinism():
    z, t = _case(2, 3, 100, 100, 7)
    mod = DC_and_CE_loss({"batch_dice": True, "smooth": 1.0, "do_bg": True, "ddp": False}, {}, weight_ce=0.5, weight_dice=2)
    want, _, _ = LO.dc_and_ce_loss(z.double(), t, batch_dice=True, do_bg=True, smooth=1.0, weight_ce=0.5, weight_dice=2)
    a, b = mod(z, t), mod(z, t)
    assert abs(float(a) - float(want)) < 2e-6 and torch.equal(a, b)
    z1 = z.clone().requires_grad_(True)
    mod(z1, t).backward()
    z2 = z.double().requires_grad_(True)
    LO.dc_and_ce_loss(z2, t, batch_dice=True, do_bg=True, smooth=1.0, weight_ce=0.5, weight_dice=2)[0].backward()
    assert float((z1.grad.double() - z2.grad).abs().max() / z2.grad.abs().max()) < 1e-5


def test_validation_statistics_match_reference_step():
    """nnUNetTrainer.py:961-1005 on one batch: loss + hard tp/fp/fn without background, float arrays."""
    z, t = _case(4, 3, 256, 256, 11, scale=1.0)
    z[0, :, :10, :10] = 0.0                                          # ties: argmax takes the first maximum
    mod = DC_and_CE_loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {}, 1, 1)
    out = validation_statistics(mod, z, t.float())
    tp, fp, fn = LO.validation_hard_counts(z, t)
    want, _, _ = LO.dc_and_ce_loss(z.double(), t)
    assert out["tp_hard"].dtype == np.float32 and out["tp_hard"].shape == (2,)
    assert np.array_equal(out["tp_hard"], tp.cpu().numpy()[1:]) and np.array_equal(out["fp_hard"], fp.cpu().numpy()[1:])
    assert np.array_equal(out["fn_hard"], fn.cpu().numpy()[1:])
    assert abs(float(out["loss"]) - float(want)) < 2e-6 * max(1.0, abs(float(want)))
    bad = t.clone()
    bad[0, 0, 0, 0] = 3
    with pytest.raises(RuntimeError):
        validation_statistics(mod, z, bad)


def test_trainer_validation_step_on_the_b200_forward():
    """`DinoUNetTrainer.validation_step` (forward + fused statistics) against the oracle applied to the same logits."""
    import os
    from types import SimpleNamespace
    import dinounet_b200
    from dinounet_b200 import config
    from oracle import dinounet_oracle as O
    os.environ["DINOUNET_B200_ALLOW_RANDOM_BACKBONE"] = "1"
    net = dinounet_b200.DinoUNet.from_config({"architecture": dict(config.DEFAULT_ARCHITECTURE)}, 3, 2, None, "dinounet_s")
    net.load_state_dict(O.make_state_dict("dinounet_s", 2, seed=0), strict=True)
    net = net.to(DEV).eval()
    tr = dinounet_b200.DinoUNetTrainer_s.__new__(dinounet_b200.DinoUNetTrainer_s)
    tr.device, tr.network, tr.is_ddp = DEV, net, False
    tr.label_manager = SimpleNamespace(has_regions=False, ignore_label=None, has_ignore_label=False)
    tr.configuration_manager = SimpleNamespace(batch_dice=True)
    tr.loss = tr._build_loss()
    g = torch.Generator().manual_seed(5)
    batch = {"data": torch.randn(2, 3, 256, 256, generator=g), "target": torch.randint(0, 2, (2, 1, 256, 256), generator=g).float()}
    out = tr.validation_step(batch)
    with torch.no_grad():
        logits = net(batch["data"].to(DEV))
    want, _, _ = LO.dc_and_ce_loss(logits.double(), batch["target"].to(DEV))
    tp, fp, fn = LO.validation_hard_counts(logits, batch["target"].to(DEV))
    assert abs(float(out["loss"]) - float(want)) < 2e-6 * max(1.0, abs(float(want)))
    assert np.array_equal(out["tp_hard"], tp.cpu().numpy()[1:]) and np.array_equal(out["fn_hard"], fn.cpu().numpy()[1:])
    assert np.array_equal(out["fp_hard"], fp.cpu().numpy()[1:])

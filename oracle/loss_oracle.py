"""TEST INFRASTRUCTURE ONLY — torch restatement of the reference's Dice+CE loss and online-validation statistics.

Follows: training/loss/compound_losses.py:31-56 (`DC_and_CE_loss.forward`, no ignore label),
training/loss/dice.py:72-119 (`MemoryEfficientSoftDiceLoss.forward`, ddp off), robust_ce_loss.py:12-16,
dice.py:122-178 (`get_tp_fp_fn_tn`), nnUNetTrainer.py:363-365 (construction: batch_dice from the plans, smooth 1e-5,
do_bg False, weights 1:1) and :969-1003 (validation_step's hard tp/fp/fn).
Pinned against the REAL reference classes in tests/test_loss_cpu.py (`load_reference_loss`).
"""
import os
import sys

import torch
import torch.nn.functional as F


def dice_term(logits, target, batch_dice=True, do_bg=False, smooth=1e-5):
    x = torch.softmax(logits, 1)
    axes = tuple(range(2, x.ndim))
    with torch.no_grad():
        y = target if target.ndim == x.ndim else target.view(target.shape[0], 1, *target.shape[1:])
        onehot = torch.zeros(x.shape, device=x.device, dtype=torch.bool)
        onehot.scatter_(1, y.long(), 1)
        if not do_bg:
            onehot = onehot[:, 1:]
        sum_gt = onehot.sum(axes)
    if not do_bg:
        x = x[:, 1:]
    intersect = (x * onehot).sum(axes)
    sum_pred = x.sum(axes)
    if batch_dice:
        intersect, sum_pred, sum_gt = intersect.sum(0), sum_pred.sum(0), sum_gt.sum(0)
    dc = (2 * intersect + smooth) / torch.clip(sum_gt + sum_pred + smooth, 1e-8)
    return -dc.mean()


def dc_and_ce_loss(logits, target, batch_dice=True, do_bg=False, smooth=1e-5, weight_ce=1, weight_dice=1):
    """target [B,1,...] (any numeric dtype) -> (loss, ce, dice term)"""
    dc = dice_term(logits, target, batch_dice, do_bg, smooth)
    ce = F.cross_entropy(logits, target[:, 0].long())
    return weight_ce * ce + weight_dice * dc, ce, dc


def validation_hard_counts(logits, target):
    """nnUNetTrainer.py:969-991 + get_tp_fp_fn_tn with axes (0, 2, ...): per-class tp, fp, fn of argmax vs target."""
    axes = [0] + list(range(2, logits.ndim))
    seg = logits.argmax(1)[:, None]
    pred = torch.zeros(logits.shape, device=logits.device, dtype=torch.float32)
    pred.scatter_(1, seg, 1)
    onehot = torch.zeros(logits.shape, device=logits.device)
    onehot.scatter_(1, target.long(), 1)
    tp = (pred * onehot).sum(dim=axes)
    fp = (pred * (1 - onehot)).sum(dim=axes)
    fn = ((1 - pred) * onehot).sum(dim=axes)
    return tp, fp, fn


def load_reference_loss():
    """The REAL `DC_and_CE_loss`, `MemoryEfficientSoftDiceLoss`, `get_tp_fp_fn_tn` (build container only)."""
    from . import ref_loader
    ref_loader._install_shims()
    tr = sys.modules["dinounet.training"]
    real = os.path.join(ref_loader.REF_ROOT, "dinounet", "training")
    if real not in tr.__path__:
        tr.__path__.append(real)
    from dinounet.training.loss.compound_losses import DC_and_CE_loss
    from dinounet.training.loss.dice import MemoryEfficientSoftDiceLoss, get_tp_fp_fn_tn
    return DC_and_CE_loss, MemoryEfficientSoftDiceLoss, get_tp_fp_fn_tn

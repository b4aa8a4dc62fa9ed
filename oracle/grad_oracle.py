"""TEST INFRASTRUCTURE ONLY — gradients of the reference's training objective through the oracle forward.

Target for the round-2 backward kernels (BASELINE.json config 3: forward + backward of the trainable parameters with the
Dice+CE loss).  The DINOv3 backbone is frozen and evaluated under no_grad in the reference
(dinov3_adapter.py:326, :422-426), so the trainable set is everything outside `encoder.dinov3_adapter.backbone.` that
is a Parameter (BatchNorm running statistics and `num_batches_tracked` are buffers).

Semantics pinned here: EVAL-mode modules (BatchNorm uses running statistics, DropPath is the identity) - the
deterministic part of the reference's backward.  The reference's own backward cannot run on CPU
(`MSDeformAttnFunction.backward` needs the CUDA extension, ms_deform_attn.py:47-68); the pinning test
(tests/test_grad_oracle_cpu.py) therefore swaps that Function for the differentiable
`ms_deform_attn_core_pytorch` (ms_deform_attn.py:71-92) - the same math - before comparing against autograd through the
REAL reference module.
"""
from typing import Dict, List, Tuple

import torch

from . import dinounet_oracle as O
from . import loss_oracle as LO

_BACKBONE = "encoder.dinov3_adapter.backbone."
_BUFFER_KINDS = {"rm", "rv", "nbt", "periods"}


def trainable_keys(model: str, num_classes: int = 2) -> List[str]:
    return [k for k, _, kind in O.param_spec(model, num_classes) if not k.startswith(_BACKBONE) and kind not in _BUFFER_KINDS]


def loss_and_grads(sd: Dict[str, torch.Tensor], model: str, x: torch.Tensor, target: torch.Tensor,
                   batch_dice: bool = True) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """sd: reference-keyed state dict (aliases allowed).  Returns (loss, {key: dLoss/dparam}) for the trainable keys."""
    num_classes = sd["decoder.seg_layers.2.weight"].shape[0]
    keys = trainable_keys(model, num_classes)
    P = {k: v.detach().clone() for k, v in sd.items()}
    leaves = {}
    for k in keys:
        leaves[k] = P[k].requires_grad_(True)
    # the duplicated state-dict entries are the same Parameter objects in the reference: point them at the same leaves
    for k, v in O.expand_aliases({k: leaves[k] for k in keys}).items():
        P[k] = v
    v = O.VARIANTS[model]                      # O.forward() itself runs under no_grad: call its two halves directly
    logits = O.decoder_forward(P, O.encoder_forward(P, v, x, False, None), None)
    loss, _, _ = LO.dc_and_ce_loss(logits, target, batch_dice=batch_dice)
    grads = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(keys, grads)}

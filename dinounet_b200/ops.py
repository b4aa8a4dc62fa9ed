"""Python face of the reference's one native extension, `MultiScaleDeformableAttention`
(reference: dinov3/eval/segmentation/models/utils/ops/src/vision.cpp:18-21, ms_deform_attn.h:26-67).

Same two entry points, same argument order and meaning, same "contiguous CUDA tensors or RuntimeError" contract; the
work is done by `b2u_msda_forward_f32` / `b2u_msda_backward_f32` of the C-ABI library.  `install_as_reference_extension()`
makes `import MultiScaleDeformableAttention` resolve here, which is what the reference's
`MSDeformAttnFunction.backward` (ms_deform_attn.py:58-66) needs in order to train without compiling its own CUDA ops.
`MSDeformAttnFunction` is the autograd function with the reference's call signature built on these two kernels.

fp32 only (the reference force-casts to fp32 on CUDA, ms_deform_attn.py:30); other dtypes raise.
"""
import sys

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import lib as _L


def _check(name, t, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} tensor has to be contiguous")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")


def _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    _check("value", value, torch.float32)
    _check("spatial_shapes", spatial_shapes, torch.int64)
    _check("level_start_index", level_start_index, torch.int64)
    _check("sampling_loc", sampling_loc, torch.float32)
    _check("attn_weight", attn_weight, torch.float32)
    B, S, M, D = value.shape
    _, Lq, M2, Lv, Pn, two = sampling_loc.shape
    if M2 != M or two != 2 or spatial_shapes.shape != (Lv, 2) or level_start_index.numel() != Lv \
            or attn_weight.shape != (B, Lq, M, Lv, Pn):
        raise RuntimeError("ms_deform_attn: inconsistent shapes")
    return B, S, M, D, Lq, Lv, Pn


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=64):
    """-> [B, Lq, heads*dh] fp32.  `im2col_step` (a launch-chunking knob of the reference kernel) is accepted, unused."""
    B, S, M, D, Lq, Lv, Pn = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    out = torch.empty(B, Lq, M * D, device=value.device, dtype=torch.float32)
    with torch.cuda.device(value.device):
        rc = _L.load().b2u_msda_forward_f32(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                            sampling_loc.data_ptr(), attn_weight.data_ptr(), out.data_ptr(),
                                            B, S, Lq, M, D, Lv, Pn, _stream(value))
    _L.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step=64):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight] (fresh fp32 tensors shaped like their primals)."""
    B, S, M, D, Lq, Lv, Pn = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    _check("grad_output", grad_output, torch.float32)
    if grad_output.shape != (B, Lq, M * D):
        raise RuntimeError("ms_deform_attn_backward: grad_output shape")
    gv = torch.empty_like(value)
    gl = torch.empty_like(sampling_loc)
    ga = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device):
        rc = _L.load().b2u_msda_backward_f32(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                             sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
                                             gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), B, S, Lq, M, D, Lv, Pn,
                                             _stream(value))
    _L.check(rc, "ms_deform_attn_backward")
    return [gv, gl, ga]


class MSDeformAttnFunction(Function):
    """Call-compatible with the reference's `MSDeformAttnFunction.apply(value, shapes, level_start_index, loc, w, step)`
    (ms_deform_attn.py:28-68); both directions run the library kernels."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        args = [value.float().contiguous(), value_spatial_shapes, value_level_start_index,
                sampling_locations.float().contiguous(), attention_weights.float().contiguous()]
        ctx.im2col_step = im2col_step
        ctx.save_for_backward(*args)
        return ms_deform_attn_forward(*args, im2col_step)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        gv, gl, ga = ms_deform_attn_backward(*ctx.saved_tensors, grad_output.float().contiguous(), ctx.im2col_step)
        return gv, None, None, gl, ga, None


def install_as_reference_extension():
    """`import MultiScaleDeformableAttention as MSDA` (ms_deform_attn.py:20) then finds these kernels."""
    sys.modules["MultiScaleDeformableAttention"] = sys.modules[__name__]

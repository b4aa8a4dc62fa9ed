"""Label-map export feed: what `export_prediction_from_logits` needs from the network when the
argmax already happened on the GPU.

The reference (`dinounet/inference/export_prediction.py:15-68`) takes fp32/fp16 LOGITS on the host, resamples them to the
pre-resampling shape, applies softmax + argmax on CPU, pastes the result into the pre-crop bounding box and reverts the
plans' axis transpose.  For the Dino U-Net 2D plans (`force_target_shape`, no spacing change between the configuration and
the case: `shape_after_cropping_and_before_resampling == logits.shape[1:]`) resampling is the identity, and softmax does
not change an argmax, so the uint8 label maps that `b2u_seg_head` writes (8x fewer bytes over NVLink / PCIe than fp32
2-class logits, `parallel.AsyncGatherer(dtype=torch.uint8)`) are all that has to leave the GPU:

    seg = labels_to_original_geometry(labels, properties_dict, plans_manager.transpose_backward)
    rw.write_seg(seg, output_file_truncated + file_ending, properties_dict)        # the reference's writer, unchanged

Cases that DO need resampling keep the reference route (logits -> `convert_predicted_logits_to_segmentation_with_correct_shape`).
Pure numpy on the host: this is I/O glue, not the hot path.
"""
from typing import Sequence

import numpy as np


def needs_resampling(label_shape: Sequence[int], properties_dict: dict) -> bool:
    """True when the reference's `resampling_fn_probabilities` would change the grid (export_prediction.py:26-33)."""
    return tuple(int(s) for s in label_shape) != tuple(int(s) for s in properties_dict["shape_after_cropping_and_before_resampling"])


def labels_to_original_geometry(labels, properties_dict: dict, transpose_backward: Sequence[int],
                                num_foreground_labels: int = 1) -> np.ndarray:
    """export_prediction.py:43-52: paste the predicted label volume into the pre-crop bounding box (background 0 outside)
    and revert the plans' transpose.  `labels`: integer array / tensor shaped like the cropped case [z, y, x] (for a 2D
    configuration the slice stack the sliding-window predictor returns)."""
    if hasattr(labels, "detach"):
        labels = labels.detach().cpu().numpy()
    labels = np.asarray(labels)
    if needs_resampling(labels.shape, properties_dict):
        raise ValueError("this case needs resampling to its original spacing: export it through the logits route "
                         "(convert_predicted_logits_to_segmentation_with_correct_shape)")
    dtype = np.uint8 if num_foreground_labels < 255 else np.uint16
    out = np.zeros(properties_dict["shape_before_cropping"], dtype=dtype)
    slicer = tuple(slice(int(lo), int(hi)) for lo, hi in properties_dict["bbox_used_for_cropping"])   # bounding_box_to_slice
    out[slicer] = labels.astype(dtype, copy=False)
    return out.transpose(list(transpose_backward))

"""TEST INFRASTRUCTURE — writes tests/golden/sliding_window_kat.npz from the REAL reference
(`nnUNetPredictor` + `sliding_window_prediction.py`, imported through oracle/ref_predictor_loader.py).
Run in the build container only:  python oracle/make_golden_sliding_window.py
Contents: step lists for a set of (image, tile, step) cases, the fp16 gaussian for two tile sizes (full 2D maps are
small), and the fp16 result of the reference sliding-window loop around a fixed toy network (seeded conv) for three
image shapes — with / without gaussian and mirroring.
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_predictor_loader import load_reference_predictor_class  # noqa: E402

STEP_CASES = [((110,), (64,), 0.5), ((512, 512), (512, 512), 0.5), ((700, 900), (512, 512), 0.5),
              ((1024, 1024), (512, 512), 0.5), ((513, 2000), (512, 512), 0.25), ((600, 512), (512, 512), 1.0)]
LOOP_CASES = [  # (c, d, H, W), patch, step, gaussian, mirror axes
    ((1, 2, 40, 56), (32, 32), 0.5, True, (0, 1)),
    ((3, 1, 20, 70), (32, 32), 0.5, True, (0, 1)),
    ((2, 3, 64, 33), (32, 32), 0.25, False, None),
    ((4, 1, 50, 50), (32, 32), 0.5, True, (1,)),
]


def toy_network(cin, heads, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(heads, cin, 3, 3, generator=g) * 0.5
    b = torch.randn(heads, generator=g)

    def net(x):
        return torch.nn.functional.conv2d(x.float(), w, b, padding=1).half()   # fp16 like autocast's conv output
    return net


def reference_loop(shape, patch, step, use_gaussian, mirror_axes, heads=2, seed=0):
    cls = load_reference_predictor_class()
    pred = cls(tile_step_size=step, use_gaussian=use_gaussian, use_mirroring=mirror_axes is not None,
               perform_everything_on_device=False, device=torch.device("cpu"), verbose=False, allow_tqdm=False)
    net = toy_network(shape[0], heads, seed)

    class Net(torch.nn.Module):
        def forward(self, x):
            return net(x)
    pred.network = Net()
    pred.configuration_manager = SimpleNamespace(patch_size=list(patch))
    pred.label_manager = SimpleNamespace(num_segmentation_heads=heads)
    pred.allowed_mirroring_axes = mirror_axes
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(seed + 1))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return x, pred.predict_sliding_window_return_logits(x)


def main():
    load_reference_predictor_class()          # installs the import shims
    from dinounet.inference.sliding_window_prediction import compute_gaussian, compute_steps_for_sliding_window
    out = {}
    out["steps_json"] = np.frombuffer(json.dumps(
        [[list(i), list(t), s, compute_steps_for_sliding_window(i, t, s)] for i, t, s in STEP_CASES]).encode(), dtype=np.uint8)
    for n in (32, 512):
        out[f"gaussian_{n}"] = compute_gaussian((n, n), sigma_scale=1. / 8, value_scaling_factor=10,
                                                device=torch.device("cpu")).numpy()
    out["gaussian_48x20_scale1"] = compute_gaussian((48, 20), device=torch.device("cpu")).numpy()
    for k, (shape, patch, step, ug, ma) in enumerate(LOOP_CASES):
        _, y = reference_loop(shape, patch, step, ug, ma)
        out[f"loop_{k}"] = y.numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "sliding_window_kat.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE — writes tests/golden/grads_*.npz (loss + per-parameter gradient norms and samples) from the
oracle's autograd (pinned to the real reference in tests/test_grad_oracle_cpu.py).  Any container:
    python oracle/make_golden_grads.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dinounet_oracle as O  # noqa: E402
from oracle import grad_oracle as G  # noqa: E402

CASES = [("dinounet_s", 2, 128, 2, 0)]   # model, batch, size, classes, seed


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    for model, B, S, ncls, seed in CASES:
        sd = O.make_state_dict(model, ncls, seed=seed)
        x = O.make_input(B, S, seed)
        target = torch.randint(0, ncls, (B, 1, S, S), generator=torch.Generator().manual_seed(seed + 7)).float()
        loss, grads = G.loss_and_grads(sd, model, x, target)
        arrays = {"loss": np.float64(loss.item())}
        names = sorted(grads)
        arrays["names"] = np.array(names)
        arrays["norms"] = np.array([grads[k].double().norm().item() for k in names])
        for i, k in enumerate(names):
            f = grads[k].reshape(-1)
            arrays[f"s{i}"] = f[:: max(1, f.numel() // 16)][:16].numpy().astype(np.float32)
        path = os.path.join(out_dir, f"grads_{model}_b{B}_s{S}_c{ncls}_w{seed}.npz")
        np.savez_compressed(path, **arrays)
        print(path, float(loss), len(names), "tensors; largest grad norm", arrays["norms"].max())


if __name__ == "__main__":
    main()

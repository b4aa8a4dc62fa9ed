"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the REAL reference forward.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
Weights/inputs are NOT stored: they are regenerated bit-identically from seeds by
`oracle.dinounet_oracle.make_state_dict / make_input` (CPU torch RNG), so the fixtures stay small.
Each fixture holds the reference logits (fp32) plus strided samples of intermediate activations captured with
forward hooks on the reference modules (adapter outputs f1..f4, encoder skips, decoder stages).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dinounet_oracle as O  # noqa: E402
from oracle.ref_loader import build_reference_model  # noqa: E402

CASES = [
    # (model, batch, size, weight seed, input seed)
    ("dinounet_s", 2, 256, 0, 0),
    ("dinounet_s", 1, 512, 0, 1),
    ("dinounet_b", 1, 256, 0, 0),
    ("dinounet_l", 1, 256, 0, 0),
    ("dinounet_7b_tiny", 1, 256, 0, 0),
]
# Benchmarked shapes (BASELINE.json configs 2/4 and the dinounet_b line): 512^2, batch >= 2 where affordable.  To keep
# the fixtures small these store the reference logits subsampled [:, :, ::SUB, ::SUB] (fp32) plus the FULL-resolution
# argmax mask bit-packed; the -m gpu test additionally compares every pixel with the oracle run live on the host (the
# oracle is pinned bit-identical to the reference by tests/test_oracle_vs_reference.py and by these samples).
BENCH_CASES = [
    ("dinounet_l", 2, 512, 0, 3),
    ("dinounet_b", 1, 512, 0, 4),
    ("dinounet_s", 4, 512, 0, 5),
]
SUB = 3
NSAMP = 2048


def sample(t: torch.Tensor) -> np.ndarray:
    f = t.detach().float().reshape(-1)
    stride = max(1, f.numel() // NSAMP)
    return f[::stride][:NSAMP].numpy().copy()


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only_bench = "--bench-only" in sys.argv
    for model, B, S, wseed, xseed in ([] if only_bench else CASES) + BENCH_CASES:
        bench_case = (model, B, S, wseed, xseed) in BENCH_CASES
        sd = O.make_state_dict(model, 2, seed=wseed)
        net = build_reference_model(model, 2, sd)
        x = O.make_input(B, S, xseed)
        cap = {}
        hooks = [
            net.encoder.dinov3_adapter.register_forward_hook(
                lambda m, i, o: cap.update({f"f{k}": o[k] for k in ("1", "2", "3", "4")})),
            net.encoder.register_forward_hook(lambda m, i, o: cap.update({f"skip{k}": t for k, t in enumerate(o)})),
        ]
        for s in range(3):
            hooks.append(net.decoder.stages[s].register_forward_hook(
                lambda m, i, o, s=s: cap.update({f"dec{s}": o})))
        with torch.no_grad():
            y = net(x)
        for h in hooks:
            h.remove()
        if bench_case:
            arrays = {"logits_sub": y[:, :, ::SUB, ::SUB].numpy().astype(np.float32).copy(), "sub": np.int64(SUB),
                      "argmax_bits": np.packbits(y.argmax(1).numpy().astype(np.uint8).reshape(-1)),
                      "absmax": np.float32(y.abs().max().item())}
            ref = O.forward(sd, model, x)
            assert torch.equal(ref, y), "oracle restatement must be bit-identical to the reference on this case"
        else:
            arrays = {"logits": y.numpy().astype(np.float32)}
        for k, t in cap.items():
            arrays["samp_" + k] = sample(t)
        name = f"{model}_b{B}_s{S}_w{wseed}_x{xseed}.npz"
        np.savez_compressed(os.path.join(out_dir, name), **arrays)
        print(name, y.shape, float(y.abs().max()), sorted(cap))


if __name__ == "__main__":
    main()

"""CPU: the oracle restatement reproduces the golden vectors made from the REAL reference
(oracle/make_golden.py).  Tolerance 2e-5 abs (fp32 reduction-order noise across host core counts is
~5e-6); argmax must agree wherever the golden class margin exceeds 1e-4."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dinounet_oracle as O


def _cases(golden_dir):
    out = []
    for f in sorted(glob.glob(os.path.join(golden_dir, "dinounet_*.npz"))):
        m, b, s, w, x = os.path.basename(f)[:-4].rsplit("_", 4)
        out.append((f, m, int(b[1:]), int(s[1:]), int(w[1:]), int(x[1:])))
    return out


def _sample(t, n=2048):
    f = t.detach().float().reshape(-1)
    stride = max(1, f.numel() // n)
    return f[::stride][:n].numpy()


@pytest.mark.parametrize("pick", ["dinounet_s_b2_s256", "dinounet_b_b1_s256"])
def test_oracle_matches_reference_golden(golden_dir, pick):
    case = [c for c in _cases(golden_dir) if pick in c[0]]
    assert case, "golden fixture missing"
    f, model, B, S, wseed, xseed = case[0]
    g = np.load(f)
    sd = O.make_state_dict(model, 2, seed=wseed)
    x = O.make_input(B, S, xseed)
    cap = {}
    y = O.forward(sd, model, x, collect=cap).numpy()
    ref = g["logits"]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-5
    margin = np.abs(ref[:, 0] - ref[:, 1])
    flips = (y.argmax(1) != ref.argmax(1)) & (margin > 1e-4)
    assert flips.sum() == 0
    for k in ("f1", "f4", "skip0", "skip3", "dec0", "dec2"):
        a, b = _sample(cap[k]), g["samp_" + k]
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), k


def test_state_dict_is_deterministic_and_aliased():
    a = O.make_state_dict("dinounet_s", 2, seed=0)
    b = O.make_state_dict("dinounet_s", 2, seed=0)
    assert len(a) == 1002  # 1002 keys for dinounet_s incl. duplicates
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["decoder.encoder.fapm.shared_basis.weight"] is a["encoder.fapm.shared_basis.weight"]
    assert a["decoder.stages.0.convs.0.all_modules.0.weight"] is a["decoder.stages.0.convs.0.conv.weight"]
    c = O.make_state_dict("dinounet_s", 2, seed=1)
    assert not torch.equal(a["encoder.fapm.shared_basis.weight"], c["encoder.fapm.shared_basis.weight"])


def test_parity_inputs_exercise_the_transformer():
    """default random-init hides the ViT (LayerScale 1e-5).  Our init must not."""
    sd = O.make_state_dict("dinounet_s", 2, seed=0)
    x = O.make_input(1, 128, 0)
    y0 = O.forward(sd, "dinounet_s", x)
    sd2 = dict(sd)
    for k in sd:
        if k.endswith("ls1.gamma") and k.startswith("encoder."):
            sd2[k] = torch.zeros_like(sd[k])
    y1 = O.forward(sd2, "dinounet_s", x)
    assert (y0 - y1).abs().max() > 1e-2
    sd3 = dict(sd)
    for k in sd:
        if "sampling_offsets.weight" in k and k.startswith("encoder."):
            sd3[k] = torch.zeros_like(sd[k])
    y2 = O.forward(sd3, "dinounet_s", x)
    assert (y0 - y2).abs().max() > 1e-3


def test_flop_model_matches_survey():
    assert abs(O.algorithmic_flops_per_patch("dinounet_s") / 1e9 - 161.9) < 0.5
    assert abs(O.algorithmic_flops_per_patch("dinounet_b") / 1e9 - 373.9) < 1.0
    assert abs(O.algorithmic_flops_per_patch("dinounet_l") / 1e9 - 943.5) < 2.0

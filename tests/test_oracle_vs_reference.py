"""CPU, build container only: the oracle restatement is bit-identical (fp32) to the REAL reference forward
imported from /root/reference through oracle/ref_loader.py.  Skipped where the reference is absent (GPU box)."""
import pytest
import torch

from oracle import dinounet_oracle as O
from oracle.ref_loader import build_reference_model, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("model,size", [("dinounet_s", 128), ("dinounet_b", 64), ("dinounet_7b_tiny", 64)])
def test_bit_identical_to_reference(model, size):
    sd = O.make_state_dict(model, 2, seed=3)
    net = build_reference_model(model, 2)
    ref_sd = net.state_dict()
    assert set(ref_sd) == set(sd)
    for k, t in ref_sd.items():
        assert tuple(t.shape) == tuple(sd[k].shape), k
    net.load_state_dict(sd, strict=True)
    x = O.make_input(2, size, 5)
    with torch.no_grad():
        yr = net(x)
    yo = O.forward(sd, model, x)
    assert torch.equal(yr, yo)


def test_single_channel_input_path():
    """dinounet_training.py:491-497 channel fix-up."""
    sd = O.make_state_dict("dinounet_s", 2, seed=3)
    net = build_reference_model("dinounet_s", 2, sd)
    x = O.make_input(1, 64, 7, channels=1)
    with torch.no_grad():
        yr = net(x)
    assert torch.equal(yr, O.forward(sd, "dinounet_s", x))

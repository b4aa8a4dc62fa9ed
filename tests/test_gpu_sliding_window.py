"""-m gpu: the sliding-window predictor (gather kernel -> batched engine forward -> accumulate/finalize kernels) against
the oracle restatement of the reference loop (oracle/sliding_window_oracle.py, pinned to the real `nnUNetPredictor` in
tests/test_sliding_window_cpu.py) wrapped around the SAME network, called the reference's way (batch 1, one forward per
mirror variant, torch flips, fp16 torch accumulation).  Integer-exact bar: the fp16 results must be bit-identical."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import dinounet_b200
from dinounet_b200 import config, lib as L
from dinounet_b200.sliding_window import SlidingWindowPredictor, compute_gaussian
from oracle import dinounet_oracle as O
from oracle import sliding_window_oracle as SWO

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _net(model="dinounet_s", in_ch=3, ncls=2):
    os.environ["DINOUNET_B200_ALLOW_RANDOM_BACKBONE"] = "1"
    net = dinounet_b200.DinoUNet.from_config({"architecture": dict(config.DEFAULT_ARCHITECTURE)}, in_ch, ncls, None, model)
    net.load_state_dict(O.make_state_dict(model, ncls, seed=0), strict=True)
    return net.to(DEV).eval()


def _predictor(net, patch, step, gaussian, mirror, tile_batch, on_device=True):
    p = SlidingWindowPredictor(tile_step_size=step, use_gaussian=gaussian, use_mirroring=mirror is not None,
                               perform_everything_on_device=on_device, device=DEV, tile_batch=tile_batch)
    p.manual_initialization(net, None, SimpleNamespace(patch_size=list(patch)), None, {}, "DinoUNetTrainer_s", mirror)
    return p


def _reference_way(net, x, patch, step, gaussian, mirror):
    def network(t):                      # autocast's fp16 conv output (predict_from_raw_data.py:695)
        with torch.no_grad():
            return net(t).half()
    return SWO.predict_sliding_window_return_logits(network, x.to(DEV), patch, net.num_classes, step, gaussian, mirror,
                                                    results_device=DEV)


def test_kernels_gather_and_accumulate_against_torch():
    """The three kernels alone, on synthetic logits: crops / mirrors / channel rule; fp16 accumulation sequence."""
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(0)
    for cin in (1, 2, 3, 5):
        vol = torch.randn(cin, 3, 70, 90, generator=g).to(DEV)
        desc = torch.tensor([[0, 0, 0, 0], [1, 5, 7, 1], [2, 38, 58, 2], [1, 17, 3, 3]], dtype=torch.int32, device=DEV)
        out = torch.empty(4, 3, 32, 32, device=DEV)
        L.check(lib.b2u_sw_gather_tiles(vol.data_ptr(), out.data_ptr(), desc.data_ptr(), 4, cin, 3, 70, 90, 32, 32, st))
        for n, (d, y0, x0, f) in enumerate(desc.tolist()):
            t = vol[:, d, y0:y0 + 32, x0:x0 + 32][None]
            t = t.repeat(1, 3, 1, 1)[:, :3] if cin < 3 else t[:, :3]       # dinounet_training.py:491-497
            axes = [a for a, b in ((2, 1), (3, 2)) if f & b]
            t = torch.flip(t, axes) if axes else t
            assert torch.equal(out[n], t[0]), (cin, n)
    C, D, H, W, th = 3, 2, 50, 60, 32
    gauss = compute_gaussian((th, th), 1. / 8, 10, device=DEV)
    acc = torch.zeros(C, D, H, W, dtype=torch.half, device=DEV)
    npred = torch.zeros(D, H, W, dtype=torch.half, device=DEV)
    racc, rn = acc.clone(), npred.clone()
    tiles = [(0, 0, 0), (0, 10, 20), (1, 18, 28), (0, 18, 0)]
    for nvar, bits in ((4, [0, 1, 2, 3]), (2, [0, 2]), (1, [0])):
        for (d, y0, x0) in tiles:
            logits = (torch.randn(nvar, C, th, th, generator=g) * 30).to(DEV)
            desc = torch.tensor([[d, y0, x0, b] for b in bits], dtype=torch.int32, device=DEV)
            L.check(lib.b2u_sw_accumulate(logits.data_ptr(), desc.data_ptr(), 0, nvar, gauss.data_ptr(), acc.data_ptr(),
                                          npred.data_ptr(), C, D, H, W, th, th, st))
            pred = logits[0].half()
            for v in range(1, nvar):
                axes = [a for a, b in ((1, 1), (2, 2)) if bits[v] & b]
                pred += torch.flip(logits[v].half(), axes)
            if nvar > 1:
                pred /= nvar
            racc[:, d, y0:y0 + th, x0:x0 + th] += pred * gauss
            rn[d, y0:y0 + th, x0:x0 + th] += gauss
    assert torch.equal(acc, racc) and torch.equal(npred, rn)
    # without gaussian: += prediction, += 1
    L.check(lib.b2u_sw_accumulate(logits.data_ptr(), desc.data_ptr(), 0, 1, None, acc.data_ptr(), npred.data_ptr(),
                                  C, D, H, W, th, th, st))
    d, y0, x0 = tiles[-1]
    racc[:, d, y0:y0 + th, x0:x0 + th] += logits[0].half()
    rn[d, y0:y0 + th, x0:x0 + th] += 1
    assert torch.equal(acc, racc) and torch.equal(npred, rn)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    npred.clamp_(min=0.5)
    rn.clamp_(min=0.5)
    a2, n2 = acc.clone(), npred.clone()
    L.check(lib.b2u_sw_finalize(acc.data_ptr(), npred.data_ptr(), C, D * H * W, flag.data_ptr(), st))
    racc /= rn
    assert torch.equal(acc, racc) and int(flag.item()) == 0 and not torch.isinf(racc).any()
    a2[1, 1, 7, 9], n2[1, 7, 9] = 60000.0, 0.5          # 120000 is not representable in fp16
    L.check(lib.b2u_sw_finalize(a2.data_ptr(), n2.data_ptr(), C, D * H * W, flag.data_ptr(), st))
    assert int(flag.item()) == 1 and torch.isinf(a2[1, 1, 7, 9])


@pytest.mark.parametrize("shape,patch,step,gaussian,mirror,tile_batch", [
    ((3, 2, 200, 300), (128, 128), 0.5, True, (0, 1), 4),      # 2 slices x 3 x 4 tiles, 4 mirror variants, ragged last batch
    ((1, 1, 100, 260), (128, 128), 0.5, True, (1,), 3),        # single channel, padded rows, 2 variants
    ((2, 3, 128, 128), (128, 128), 0.25, False, None, 2),      # exactly one tile per slice, no gaussian, no mirroring
])
def test_predictor_bit_identical_to_reference_loop(shape, patch, step, gaussian, mirror, tile_batch):
    net = _net("dinounet_s", in_ch=shape[0])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(3))
    want = _reference_way(net, x, patch, step, gaussian, mirror)
    got = _predictor(net, patch, step, gaussian, mirror, tile_batch).predict_sliding_window_return_logits(x)
    assert got.dtype == torch.half and got.shape == (2, *shape[1:]) and got.is_cuda
    neq = int((got != want).sum())
    assert neq == 0, f"{neq}/{got.numel()} fp16 values differ, max |d| {float((got.float() - want.float()).abs().max())}"
    cpu = _predictor(net, patch, step, gaussian, mirror, tile_batch, on_device=False).predict_sliding_window_return_logits(x)
    assert not cpu.is_cuda and torch.equal(cpu, want.cpu())


def test_predictor_512_tiles_and_inf_error():
    """The reference's real tile size (main_dinov3 forces 512x512): 700x900 slice -> 2x3 tiles x 4 mirror variants."""
    net = _net("dinounet_s")
    x = torch.randn(3, 1, 700, 900, generator=torch.Generator().manual_seed(4))
    want = _reference_way(net, x, (512, 512), 0.5, True, (0, 1))
    p = _predictor(net, (512, 512), 0.5, True, (0, 1), 6)
    got = p.predict_sliding_window_return_logits(x)
    assert torch.equal(got, want)
    with pytest.raises(AssertionError):
        p.predict_sliding_window_return_logits(x[0])
    # fp16 accumulator overflow -> the reference's RuntimeError (predict_from_raw_data.py:603-606)
    sd = O.make_state_dict("dinounet_s", 2, seed=0)
    sd["decoder.seg_layers.2.weight"] = sd["decoder.seg_layers.2.weight"] * 1e5
    net.load_state_dict(sd, strict=True)
    net.repack()
    with pytest.raises(RuntimeError, match="Encountered inf"):
        p.predict_sliding_window_return_logits(x[:, :, :512, :512])

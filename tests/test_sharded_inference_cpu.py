"""CPU, world_size 2, gloo: the N>1 path (batch sharding + one logits all-gather, dinounet_b200/parallel.py).
The per-rank forward is a stand-in (the kernels need a GPU); what is tested is the host logic the 8-GPU run uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dinounet_b200.parallel import AsyncGatherer, gather_logits, shard_bounds, sharded_forward, sharded_sliding_window


def test_shard_bounds_cover_batch_exactly():
    for B in (1, 2, 7, 32, 256):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _fake_forward(x):  # stand-in "network": per-sample deterministic function, [b,3,H,W] -> [b,2,H,W]
    return torch.stack([x.sum(1), x.mean(1) * 2.0], 1)


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(B, 3, 8, 8, generator=g)
        full = sharded_forward(_fake_forward, x)
        ok = torch.equal(full, _fake_forward(x))
        lo, hi = shard_bounds(B, world, rank)
        ok = ok and torch.equal(gather_logits(_fake_forward(x[lo:hi]), B), _fake_forward(x))
        ok = ok and torch.equal(gather_logits(_fake_forward(x[lo:hi]), B, dtype=torch.float16), _fake_forward(x).half())
        if B % world == 0:     # the side-stream gatherer (synchronous on CPU/gloo): fp16 logits and uint8 label maps
            ag, al = AsyncGatherer(B, x.device, torch.float16), AsyncGatherer(B, x.device, torch.uint8)
            for _ in range(3):
                t = ag.submit(_fake_forward(x[lo:hi]))
                tl = al.submit(_fake_forward(x[lo:hi]).argmax(1).to(torch.uint8))
            ok = ok and torch.equal(ag.result(t), _fake_forward(x).half())
            ok = ok and torch.equal(al.result(tl), _fake_forward(x).argmax(1).to(torch.uint8))
        q.put((rank, bool(ok), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_two_rank_gloo_gather_matches_single_process(B):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (B, 2, 8, 8) for _, _, shape in res)


def _fake_predict(vol):  # stand-in sliding-window predictor: [c, d, H, W] -> fp16 [2, d, H, W], slice-independent
    return torch.stack([vol.sum(0), vol.mean(0) - 1.0], 0).half()


def _sw_worker(rank, world, port, n_slices, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = torch.randn(3, n_slices, 6, 5, generator=torch.Generator().manual_seed(1))
        got = sharded_sliding_window(_fake_predict, vol)
        q.put((rank, bool(torch.equal(got, _fake_predict(vol))), tuple(got.shape), str(got.dtype)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_slices", [1, 4, 5])
def test_two_rank_gloo_slice_sharded_sliding_window(n_slices):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sw_worker, args=(r, 2, port, n_slices, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert all(shape == (2, n_slices, 6, 5) and dt == "torch.float16" for _, _, shape, dt in res)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dinounet_b200.train_path import all_reduce_gradients
        g = torch.Generator().manual_seed(5)
        ps = [torch.randn(7, 3, generator=g).requires_grad_(), torch.randn(11, generator=g).requires_grad_(),
              torch.randn(2, 2, generator=g).requires_grad_(), torch.randn(4, generator=g)]       # last one: frozen
        grads = [[torch.full_like(p, float(r + 1) * (i + 1)) for i, p in enumerate(ps)] for r in range(world)]
        ps[0].grad, ps[1].grad = grads[rank][0].clone(), grads[rank][1].clone()
        if rank == 0:
            ps[2].grad = grads[rank][2].clone()          # rank 1 has no gradient for this tensor: contributes zeros
        all_reduce_gradients(ps)
        ok = torch.allclose(ps[0].grad, sum(grads[r][0] for r in range(world)) / world)
        ok = ok and torch.allclose(ps[1].grad, sum(grads[r][1] for r in range(world)) / world)
        ok = ok and torch.allclose(ps[2].grad, grads[0][2] / world) and ps[3].grad is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_all_reduce():
    """Data-parallel training exchange (train_path.all_reduce_gradients): one flat all-reduce, mean over ranks, ranks
    without a gradient for some tensor contribute zeros, frozen parameters are left alone."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res

"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

  * the NCCL path of BASELINE config 4: every rank runs its batch shard, ONE all-gather (fp32 / fp16 logits, uint8 label
    maps; synchronous `gather_logits` and the side-stream `AsyncGatherer`) -> must equal the single-GPU forward of the
    whole batch bit for bit (patches are independent and every reduction has a fixed order);
  * the engine on a device that is not the process's current device (device guard).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
needs2 = pytest.mark.skipif(NGPU < 2, reason="needs two GPUs")


def _build(model, dev):
    os.environ["DINOUNET_B200_ALLOW_RANDOM_BACKBONE"] = "1"
    import dinounet_b200
    from dinounet_b200 import config
    from oracle import dinounet_oracle as O
    sd = O.make_state_dict(model, 2, seed=0)
    net = dinounet_b200.DinoUNet.from_config({"architecture": dict(config.DEFAULT_ARCHITECTURE)}, 3, 2, None, model)
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval()


def _worker(rank, world, port, B, S, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from dinounet_b200.parallel import AsyncGatherer, gather_logits, shard_bounds
        from oracle import dinounet_oracle as O
        net = _build("dinounet_s", dev)
        x = O.make_input(B, S, 7)
        lo, hi = shard_bounds(B, world, rank)
        with torch.no_grad():
            eng = net._get_engine(dev)
            logits, labels = eng.forward(x[lo:hi].to(dev), use_graph=True)
            full32 = gather_logits(logits, B)
            full16 = gather_logits(logits, B, dtype=torch.float16)
            lab = gather_logits(labels, B)
            ag = AsyncGatherer(B, dev, torch.float16)
            al = AsyncGatherer(B, dev, torch.uint8)
            tickets = []
            for it in range(3):      # overlapped with the next forwards, double-buffered
                lg, lb = eng.forward(x[lo:hi].to(dev), use_graph=True)
                tickets.append((ag.submit(lg), al.submit(lb)))
            a16 = ag.result(tickets[-1][0]).clone()
            a8 = al.result(tickets[-1][1]).clone()
            ok = True
            if rank == 0:
                ref, ref_lab = eng.forward(x.to(dev), use_graph=False)
                ok = (torch.equal(full32, ref) and torch.equal(full16, ref.half()) and torch.equal(lab, ref_lab)
                      and torch.equal(a16, ref.half()) and torch.equal(a8, ref_lab))
        torch.cuda.synchronize()
        q.put((rank, bool(ok), tuple(full32.shape)))
    finally:
        dist.destroy_process_group()


@needs2
def test_nccl_gathered_logits_equal_single_gpu_bitwise():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    B, S = 4, 128
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (B, 2, S, S) for _, _, shape in res)


@needs2
def test_engine_on_non_current_device():
    from oracle import dinounet_oracle as O
    torch.cuda.set_device(0)
    x = O.make_input(2, 128, 3)
    with torch.no_grad():
        y0 = _build("dinounet_s", torch.device("cuda", 0))(x.cuda(0)).cpu()
        net1 = _build("dinounet_s", torch.device("cuda", 1))
        assert torch.cuda.current_device() == 0
        y1 = net1(x.cuda(1)).cpu()
        yg, _ = net1._engine.forward(x.cuda(1), use_graph=True)
        yg = yg.cpu()
    assert torch.equal(y0, y1) and torch.equal(yg, y1)

"""CPU (gloo, world_size 2): the gradient-bucket reducer that replaces the reference's DataParallel reduce-add
(sync_batchnorm/replicate.py:50-67, train_generator.py:171-178). Checks: averaged gradients equal the mean of the per-rank
gradients; parameters without a gradient on every rank are skipped; a parameter missing a gradient on ONE rank is handled."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # spawned workers re-import this module
import hrv_loader  # noqa: E402

hrv_loader.load()
from hrviton_b200 import ddp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    dead = torch.nn.Linear(2, 2)  # never used: no gradient on any rank (the reference's conv2.* case)
    params = list(net.parameters()) + list(dead.parameters())
    x = torch.full((2, 6), float(rank + 1))
    loss = net(x).sum() if rank == 0 else net[1](net[0](x)).sum()  # rank 1 leaves net[2] without a gradient
    loss.backward()
    local = [p.grad.clone() if p.grad is not None else None for p in params]
    red = ddp.GradBucketReducer(params, bucket_bytes=64)  # tiny buckets: several all-reduces
    nb = red.reduce()
    out[rank] = (local, [p.grad.clone() if p.grad is not None else None for p in params], nb)
    dist.destroy_process_group()


def test_bucket_reducer_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (l0, r0, nb0), (l1, r1, nb1) = out[0], out[1]
    assert nb0 == nb1 and nb0 >= 2
    for i in range(len(l0)):
        if l0[i] is None and l1[i] is None:
            assert r0[i] is None and r1[i] is None  # skipped everywhere
            continue
        a = l0[i] if l0[i] is not None else torch.zeros_like(r0[i])
        b = l1[i] if l1[i] is not None else torch.zeros_like(r0[i])
        want = (a + b) / 2
        assert torch.allclose(r0[i], want, atol=1e-6) and torch.allclose(r1[i], want, atol=1e-6)


def test_wrapper_keeps_module_attribute():
    m = torch.nn.Linear(2, 2)
    w = ddp.DataParallelWithCallback(m, device_ids=[0])
    assert w.module is m
    assert torch.equal(w(torch.ones(1, 2)), m(torch.ones(1, 2)))
    assert w.reduce_gradients() == 0  # no process group: pass-through

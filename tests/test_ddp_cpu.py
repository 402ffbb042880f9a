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


def _worker_auto(rank, world, port, out):
    """The reference loop's pattern (train_generator.py:314-322): wrap, forward, loss.backward(), optimizer.step() — no explicit
    reduce call.  Replicas must start identical (broadcast at wrap time) and stay identical."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WORLD_SIZE"] = str(world)
    os.environ["RANK"] = str(rank)
    torch.manual_seed(100 + rank)  # different initial weights per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    dead = torch.nn.Linear(2, 2)
    net.add_module("dead", dead)
    net.forward = lambda x: net[2](net[1](net[0](x)))
    wrapped = ddp.DataParallelWithCallback(net, device_ids=[0])  # initialises gloo from the environment, broadcasts rank 0's state
    w0 = [p.detach().clone() for p in net.parameters()]
    b0 = [b.detach().clone() for b in net.buffers()]
    opt = torch.optim.SGD([p for p in net.parameters()], lr=0.1)
    xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(7 + 10 * it + rank)) for it in range(3)]
    locals_, launched = [], []
    for it in range(3):
        opt.zero_grad()  # set_to_none=True, like the reference's optimizer.zero_grad() on a modern torch
        loss = wrapped(xs[it]).pow(2).sum()
        loss.backward()
        launched.append(wrapped._reducer.launched)
        opt.step()
    out[rank] = (w0, b0, [p.detach().clone() for p in net.parameters()], [p.grad.clone() if p.grad is not None else None for p in net.parameters()],
                 launched)
    dist.destroy_process_group()


def test_wrapper_synchronises_without_explicit_reduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_auto, args=(world, _free_port(), out), nprocs=world, join=True)
    (w0a, b0a, wa, ga, la), (w0b, b0b, wb, gb, lb) = out[0], out[1]
    for x, y in zip(w0a + b0a, w0b + b0b):
        assert torch.equal(x, y)  # broadcast at wrap time
    for x, y in zip(wa, wb):
        assert torch.equal(x, y)  # identical after 3 optimiser steps: gradients were averaged automatically
    assert any(not torch.equal(x, y) for x, y in zip(w0a, wa))  # and training actually moved them
    for x, y in zip(ga, gb):
        assert (x is None) == (y is None) and (x is None or torch.equal(x, y))
    assert ga[-1] is None  # the dead branch got no gradient and was skipped
    assert la == lb and all(n >= 1 for n in la)


def test_wrapper_keeps_module_attribute():
    m = torch.nn.Linear(2, 2)
    w = ddp.DataParallelWithCallback(m, device_ids=[0])
    assert w.module is m
    assert torch.equal(w(torch.ones(1, 2)), m(torch.ones(1, 2)))
    assert w.reduce_gradients() == 0  # no process group: pass-through


def _buffers_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bn = torch.nn.BatchNorm2d(3)
    with torch.no_grad():
        bn.running_mean.fill_(float(rank + 1))       # per-replica running statistics (what nn.BatchNorm2d under data parallelism gives)
        bn.running_var.fill_(float(10 * (rank + 1)))
        bn.num_batches_tracked.fill_(7 + rank)       # integer buffer: rank 0's value wins
    ddp.average_module_buffers(bn)
    out[rank] = (bn.running_mean.clone(), bn.running_var.clone(), int(bn.num_batches_tracked))
    dist.destroy_process_group()


def test_average_module_buffers_two_ranks():
    """ddp.average_module_buffers: float buffers (BatchNorm running statistics) become the mean over ranks, integer buffers rank 0's —
    the opt-in step before save_checkpoint for runs that want checkpoint statistics of the global batch (the reference's own semantics,
    plain nn.BatchNorm2d under DataParallel, are per-replica statistics with replica 0's running buffers saved)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_buffers_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        m, v, n = out[r]
        assert torch.allclose(m, torch.full((3,), 1.5)) and torch.allclose(v, torch.full((3,), 15.0)) and n == 7

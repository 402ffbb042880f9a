"""Data-parallel plumbing: one process per GPU, gradient buckets all-reduced over NCCL/NVLink (gloo in CPU tests).

Replaces the reference's single-process nn.DataParallel + replicate callback (sync_batchnorm/replicate.py:50-67,
train_generator.py:171-178): no parameter re-broadcast per forward, no scatter/gather, only a gradient all-reduce.
The hot path shards by samples, so the only exchange step is the gradient average."""
import torch
import torch.distributed as dist
import torch.nn as nn


class GradBucketReducer:
    """Averages .grad of the given parameters across ranks using a few large flat buckets (default 64 MiB: on
    NVSwitch the cost is launch latency, not link count).  Parameters without a gradient (e.g. the reference's dead
    conv2.* branch) are skipped consistently on every rank by exchanging a presence mask first."""

    def __init__(self, params, bucket_bytes=64 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.group = group
        self._present = None

    def reduce(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return 0
        world = dist.get_world_size(self.group)
        if not self.params:
            return 0
        dev = self.params[0].device
        if self._present is None:  # decided once (the set of parameters that receive gradients is static): keeps later calls
            mask = torch.tensor([1 if p.grad is not None else 0 for p in self.params], dtype=torch.int32, device=dev)  # sync-free
            dist.all_reduce(mask, op=dist.ReduceOp.MAX, group=self.group)
            self._present = mask.tolist()
        present = self._present
        buckets, cur, cur_bytes = [], [], 0
        for p, has in zip(self.params, present):
            if not has:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            nbytes = p.grad.numel() * p.grad.element_size()
            if cur and (cur_bytes + nbytes > self.bucket_bytes or p.grad.dtype != cur[0].grad.dtype):
                buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            buckets.append(cur)
        for b in buckets:
            flat = torch.cat([p.grad.reshape(-1) for p in b])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(world)
            off = 0
            for p in b:
                n = p.grad.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        return len(buckets)


class DataParallelWithCallback(nn.Module):
    """Same name/constructor as sync_batchnorm.DataParallelWithCallback(module, device_ids=...).  forward() calls the
    wrapped module on this rank's shard; call .reduce_gradients() after backward() (the bundled training step does)."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super().__init__()
        self.module = module
        self.device_ids = device_ids
        self._reducer = None

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def reduce_gradients(self):
        if self._reducer is None:
            self._reducer = GradBucketReducer(list(self.module.parameters()))
        return self._reducer.reduce()

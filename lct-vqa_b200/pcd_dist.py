"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo on CPU tests).

The search batch is sharded across ranks, every rank holds a full replica, BatchNorm statistics stay
rank-local (standard DDP semantics; the reference is single-process) and the
gradients are averaged at the points SURVEY.md §8(e) lists.  `GradReducer` is what Architect and
SearchStep call at those points.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


class GradReducer:
    """Average a list of tensors across ranks in place, through flat buckets (few large collectives:
    NVSwitch makes the cost latency- not link-bound, so bucket count is kept small)."""

    def __init__(self, group=None, bucket_bytes=256 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_elems = bucket_bytes // 4
        self.calls = 0
        self.bytes = 0

    def __call__(self, tensors):
        if self.world == 1 or not tensors:
            return
        bucket, n = [], 0
        for t in tensors:
            bucket.append(t)
            n += t.numel()
            if n >= self.bucket_elems:
                self._reduce(bucket)
                bucket, n = [], 0
        if bucket:
            self._reduce(bucket)

    def _reduce(self, bucket):
        if len(bucket) == 1 and bucket[0].is_contiguous():
            flat = bucket[0].view(-1)
            dist.all_reduce(flat, group=self.group)
            flat.div_(self.world)
        else:
            flat = torch.cat([t.reshape(-1) for t in bucket])
            dist.all_reduce(flat, group=self.group)
            flat.div_(self.world)
            torch._foreach_copy_(bucket, [v.view(t.shape) for v, t in
                                          zip(flat.split_with_sizes([t.numel() for t in bucket]), bucket)])
        self.calls += 1
        self.bytes += flat.numel() * 4

    def report(self):
        return {"allreduce_calls": self.calls, "allreduce_bytes": self.bytes}

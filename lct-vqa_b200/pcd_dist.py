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


def _flat_runs(tensors):
    """Coalesce tensors that sit back to back in one storage (the per-parameter views of a cell's flat gradient arena)
    into single flat views: no copy, fewer and larger collectives."""
    runs, cur = [], None          # cur = [storage_ptr, first tensor, start ptr, end ptr]
    for t in tensors:
        ok = t.is_contiguous() and t.numel() > 0
        if ok and cur is not None and t.untyped_storage().data_ptr() == cur[0] and t.data_ptr() == cur[3] and t.dtype == cur[1].dtype:
            cur[3] += t.numel() * t.element_size()
            continue
        if cur is not None:
            runs.append(cur)
        cur = [t.untyped_storage().data_ptr(), t, t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()] if ok else None
        if not ok and t.numel() > 0:
            runs.append([None, t, 0, 0])
    if cur is not None:
        runs.append(cur)
    out = []
    for sp, t0, a, b in runs:
        if sp is None:
            out.append(t0)
            continue
        n = (b - a) // t0.element_size()
        out.append(t0 if n == t0.numel() else
                   torch.empty(0, dtype=t0.dtype, device=t0.device).set_(t0.untyped_storage(), t0.storage_offset(), (n,)))
    return out


class GradReducer:
    """Average tensors across ranks in place.

    `start(tensors)` launches the collectives on a side stream as soon as everything enqueued so far on the compute
    stream is done and returns at once; `finish()` makes the compute stream wait for them.  Callers start a bucket as soon
    as its gradients exist and keep computing (staged_grads below: the question-encoder / head gradients — 3/4 of the bytes
    — are reduced while the image encoder's backward runs).  Large tensors are reduced in place, adjacent views are
    coalesced without a copy, only the remaining small tensors are packed into one bucket.  NCCL: ReduceOp.AVG, no
    separate division.  Everything is capturable in a CUDA graph (event fork / join, no host synchronisation)."""

    SMALL = 1 << 18           # elements: below this a tensor goes into the packed bucket

    def __init__(self, group=None, bucket_bytes=256 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0
        self.bytes = 0
        self.collectives = 0
        self.overlapped_bytes = 0
        self.stream = None
        self._pending = []        # (packed bucket, [views], [tensors]) to copy back at finish(), plus keep-alives
        self._avg = None

    # ---- low level ------------------------------------------------------------------------------------------------
    def _all_reduce(self, t):
        if self._avg is None:
            self._avg = dist.get_backend(self.group) == "nccl"
        if self._avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, group=self.group)
            t.div_(self.world)
        self.collectives += 1
        self.bytes += t.numel() * t.element_size()

    def start(self, tensors, overlapped=False):
        if self.world == 1 or not tensors:
            return
        views = _flat_runs(list(tensors))
        big = [v for v in views if v.numel() >= self.SMALL]
        small = [v for v in views if v.numel() < self.SMALL]
        cuda = views[0].is_cuda
        if cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            self.stream.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(self.stream) if cuda else _Null()
        with ctx:
            for v in big:
                self._all_reduce(v)
            packed = None
            if small:
                packed = torch.cat([v.reshape(-1) for v in small]) if len(small) > 1 else small[0].reshape(-1)
                self._all_reduce(packed)
                if len(small) > 1:
                    torch._foreach_copy_(small, [p.view(v.shape) for p, v in
                                                 zip(packed.split_with_sizes([v.numel() for v in small]), small)])
        self._pending.append((views, packed, list(tensors)))
        self.calls += 1
        if overlapped:
            self.overlapped_bytes += sum(v.numel() * v.element_size() for v in views)

    def finish(self):
        if self.stream is not None and self._pending:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._pending.clear()

    def __call__(self, tensors):
        self.start(tensors)
        self.finish()

    def report(self):
        return {"allreduce_calls": self.calls, "collectives": self.collectives, "allreduce_bytes": self.bytes,
                "overlapped_bytes": self.overlapped_bytes}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def staged_grads(model, batch, params, reducer, qst_only=False, extra=()):
    """d loss / d params (and d loss / d extra, tensors that only feed the image encoder: the alphas / betas) with the
    data-parallel averaging overlapped with the backward pass: the loss depends on the image encoder only through the image
    embedding (VqaModel._loss_staged), so the backward is cut there — stage 1 yields the gradients of every other parameter
    (embedding, LSTM, vocabulary projection, heads: 74 of the 99 MB) and their all-reduce starts on the side stream; stage 2
    (the search network's backward, the long part) runs meanwhile, its gradients are reduced at the end.
    Returns (loss, grads in `params` order with None where a parameter took no part, grads of `extra`)."""
    loss, feat = model._loss_staged(*batch, qst_only)
    enc_ids = {id(p) for p in model.img_encoder.parameters()}
    rest = [p for p in params if id(p) not in enc_ids]
    enc = [p for p in params if id(p) in enc_ids]
    got = torch.autograd.grad(loss, rest + [feat], retain_graph=True, allow_unused=True)
    g_rest, g_feat = list(got[:-1]), got[-1]
    reducer.start([g for g in g_rest if g is not None], overlapped=True)
    got2 = torch.autograd.grad(feat, enc + list(extra), grad_outputs=g_feat, allow_unused=True)
    g_enc, g_extra = list(got2[:len(enc)]), list(got2[len(enc):])
    reducer.start([g for g in g_enc + g_extra if g is not None])
    reducer.finish()
    by_id = {id(p): g for p, g in zip(rest + enc, g_rest + g_enc)}
    assert len({g.data_ptr() for g in by_id.values() if g is not None and g.numel()}) == \
        sum(1 for g in by_id.values() if g is not None and g.numel()), "two parameters share one gradient buffer"
    return loss, [by_id[id(p)] for p in params], g_extra

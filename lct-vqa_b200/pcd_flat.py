"""Optimizer-side operations over flat runs (C ABI: pcd_flat_*).

The 732 parameter tensors of the model are, in memory, a few dozen contiguous runs: the search network's parameters are
views of one arena (pcd_ops.Arena) and its gradients come back as one flat buffer per cell.  `co_runs` finds the runs that
are contiguous in EVERY operand list at once; the kernels then do w' = w - eta g, w +- R v, |v|, clip_grad_norm_ and Adam
in one launch each instead of a dozen multi-tensor launches.
"""
import ctypes as C

import torch

import pcd_native as N


def co_runs(*lists):
    """Parallel lists of same-shaped tensors -> (sizes, [ptr list per operand]): position i+1 is merged into the run of
    position i when in every list tensor i+1 starts exactly where tensor i ends."""
    n = len(lists[0])
    k = len(lists)
    sizes, ptrs = [], [[] for _ in range(k)]
    ends = None
    for i in range(n):
        ts = [lst[i] for lst in lists]
        t0 = ts[0]
        cnt = t0.numel()
        if cnt == 0:
            continue
        for t in ts:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != cnt:
                raise ValueError("flat runs need contiguous fp32 tensors of matching sizes")
        starts = [t.data_ptr() for t in ts]
        if ends is not None and starts == ends:
            sizes[-1] += cnt
        else:
            sizes.append(cnt)
            for j in range(k):
                ptrs[j].append(starts[j])
        ends = [p + 4 * cnt for p in starts]
    return sizes, ptrs


def _tables(lib, sizes, ptrs):
    """ctypes arrays per chunk of at most pcd_flat_max_runs() runs."""
    cap = lib.pcd_flat_max_runs()
    for a in range(0, len(sizes), cap):
        b = min(len(sizes), a + cap)
        sz = (C.c_longlong * (b - a))(*sizes[a:b])
        tabs = [(C.c_void_p * (b - a))(*p[a:b]) for p in ptrs]
        yield b - a, sz, tabs


def axpy_(ys, xs, alpha=1.0, alpha_dev=None):
    """ys[i] += alpha * (alpha_dev or 1) * xs[i]   (alpha_dev: 0-dim / 1-element fp32 device tensor, stays on the device)."""
    if not ys:
        return
    lib = N.lib_for(ys[0])
    sizes, ptrs = co_runs(ys, xs)
    for n, sz, (ty, tx) in _tables(lib, sizes, ptrs):
        N.check(lib, lib.pcd_flat_axpy(n, sz, ty, tx, N.ptr(alpha_dev), float(alpha), N.stream_for(ys[0])), "pcd_flat_axpy")


def scale_(ys, scale_dev):
    if not ys:
        return
    lib = N.lib_for(ys[0])
    sizes, ptrs = co_runs(ys)
    for n, sz, (ty,) in _tables(lib, sizes, ptrs):
        N.check(lib, lib.pcd_flat_scale(n, sz, ty, N.ptr(scale_dev), N.stream_for(ys[0])), "pcd_flat_scale")


def norm(xs):
    """sqrt(sum of squares) over all tensors: 0-dim fp32 device tensor (accumulated in fp64, fixed order: bit-reproducible)."""
    lib = N.lib_for(xs[0])
    out = torch.zeros((), dtype=torch.float64, device=xs[0].device)
    sizes, ptrs = co_runs(xs)
    for n, sz, (tx,) in _tables(lib, sizes, ptrs):
        work = torch.empty(int(lib.pcd_flat_sumsq_work(sum(sz))), dtype=torch.float64, device=xs[0].device)
        N.check(lib, lib.pcd_flat_sumsq(n, sz, tx, N.ptr(out), N.ptr(work), N.stream_for(xs[0])), "pcd_flat_sumsq")
    return out.sqrt().to(torch.float32)


def clip_grad_norm_(params, max_norm):
    """nn.utils.clip_grad_norm_ (2-norm) over flat runs: returns the norm before clipping (device tensor), scales the
    gradients in place by min(1, max_norm / (norm + 1e-6))."""
    grads = [p.grad for p in params if p.grad is not None]
    total = norm(grads)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    scale_(grads, coef)
    return total


class FlatAdam:
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) with the update of all parameters in ONE launch over the flat
    runs (same formulas; the step counter lives on the device, so the step is CUDA-graph capturable).  The moment buffers
    are two flat tensors laid out in parameter order."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = list(params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self._m, self._v, off = [], [], 0
        for p in self.params:
            n = p.numel()
            self._m.append(self.exp_avg[off:off + n])
            self._v.append(self.exp_avg_sq[off:off + n])
            off += n
        self.param_groups = [{"params": self.params, "lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        self.state = {}

    def state_tensors(self):
        """Everything step() mutates besides the parameters (search._TrainingState snapshots / restores these in place)."""
        return [self.exp_avg, self.exp_avg_sq, self.step_t]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        idx = [i for i, p in enumerate(self.params) if p.grad is not None]
        if not idx:
            return
        ps = [self.params[i].data for i in idx]
        gs = [self.params[i].grad.reshape(self.params[i].shape) if self.params[i].grad.is_contiguous()
              else self.params[i].grad.contiguous() for i in idx]
        ms = [self._m[i] for i in idx]
        vs = [self._v[i] for i in idx]
        self.step_t += 1
        lib = N.lib_for(ps[0])
        sizes, ptrs = co_runs(ps, gs, ms, vs)
        g = self.param_groups[0]
        for n, sz, (tp, tg, tm, tv) in _tables(lib, sizes, ptrs):
            N.check(lib, lib.pcd_flat_adam(n, sz, tp, tg, tm, tv, float(g["lr"]), float(self.betas[0]), float(self.betas[1]),
                                           float(self.eps), float(self.weight_decay), N.ptr(self.step_t), N.stream_for(ps[0])),
                    "pcd_flat_adam")

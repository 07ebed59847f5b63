"""Second-order DARTS architect — B200 drop-in for darts_vqa/pcdarts/architect_vqa.py.

Same public surface: Architect(model, args).step(img_train, qst_train, label_train, img_valid,
qst_valid, label_valid, eta, network_optimizer=None, unrolled=True) with args.arch_learn_rate /
arch_wt_decay / qst_only; Adam(betas=(0.5, 0.999)) over model.arch_parameters(); the virtual step uses
momentum = weight decay = 0 and the finite-difference Hessian-vector product uses r = 1e-2
(architect_vqa.py:15-16,19-21,105).

Same arithmetic, different mechanics:
  * the unrolled model w' = w - eta * dL_train/dw is a PERSISTENT twin (built once with model.new())
    whose weights / BN buffers / alphas are overwritten in place with multi-tensor ops, instead of
    constructing and load_state_dict-ing a fresh 732-tensor model every step (:90-103);
  * w +- R v and ||v|| are multi-tensor (foreach) kernels over the parameter list, not 732 Python-level
    in-place ops (:106-118);
  * the two HVP backward passes only ask autograd for the alpha/beta grads, which switches the CUDA cell
    kernels to their activation-only mode (no weight-gradient work);
  * an optional `reducer` averages gradients across data-parallel ranks at the four points where the
    reference's single-process quantities become global sums.
"""
import torch

import pcd_flat
import pcd_ops


def _concat(xs):
    return torch.cat([x.reshape(-1) for x in xs])


class Architect(object):

    def __init__(self, model, args, reducer=None):
        self.network_momentum = 0
        self.network_weight_decay = 0
        self.model = model
        self.args = args
        self.reducer = reducer
        self.optimizer = torch.optim.Adam(self.model.arch_parameters(), lr=args.arch_learn_rate,
                                          betas=(0.5, 0.999), weight_decay=args.arch_wt_decay)
        self.exp_zero_grad = 6 if self.args.qst_only else 0
        self._twin = None
        self._cache = {}
        # True: keep R = r/||v|| on the device (no host sync) so the whole step can live in a CUDA graph
        self.device_scalars = False
        # True (CUDA): the two finite-difference passes of the Hessian-vector product, which are independent of each other,
        # run CONCURRENTLY — w + R v on the model (current stream), w - R v on the persistent twin (side stream)
        self.concurrent_hvp = False
        self._side = None
        self.last = {}            # quantities of the last unrolled step, for inspection / tests

    # ---- helpers -----------------------------------------------------------------------------------
    def _allreduce(self, tensors):
        if self.reducer is not None:
            self.reducer(tensors)

    def _lists(self, module):
        """(parameters, buffers) of a module, cached: walking 700 sub-modules costs milliseconds per call."""
        key = id(module)
        if key not in self._cache:
            self._cache[key] = (list(module.parameters()), list(module.buffers()))
        return self._cache[key]

    def _copy_state(self, twin, model):
        """twin <- model (weights, BN buffers, alphas): flat arena copies for the search network."""
        tp, tb = self._lists(twin)
        mp, mb = self._lists(model)
        nets = [(a, b) for a, b in zip(twin.modules(), model.modules()) if hasattr(a, '_arena') and hasattr(b, 'cells')] \
            if 'nets' not in self._cache else self._cache['nets']
        self._cache['nets'] = nets
        covered_p, covered_b = set(), set()
        for tn, mn in nets[:1]:
            ta, ma = tn._arena(), mn._arena()
            for which in ('params', 'running', 'nbt'):
                ta.flat(which).copy_(ma.flat(which))
            covered_p.update(id(p) for p in ta.params)
            covered_b.update(id(b) for b in ta.running + ta.nbt)
        rest = [(a, b) for a, b in zip(tp, mp) if id(a) not in covered_p] + \
               [(a, b) for a, b in zip(tb, mb) if id(a) not in covered_b]
        for a, b in rest:
            a.copy_(b)
        for x, y in zip(twin.arch_parameters(), model.arch_parameters()):
            x.copy_(y)

    def unrolled_model(self):
        """The persistent twin that holds w' (created on first use through model.new())."""
        if self._twin is None:
            self._twin = self.model.new()
            self._twin.train()
        return self._twin

    def _staged(self, model, batch, params, extra=()):
        """Data-parallel gradients with the all-reduce overlapped with the image encoder's backward (pcd_dist.staged_grads)."""
        from pcd_dist import staged_grads
        loss, grads, g_extra = staged_grads(model, batch, params, self.reducer, self.args.qst_only, extra=extra)
        missing = sum(g is None for g in grads)
        assert missing == self.exp_zero_grad, (missing, self.exp_zero_grad)
        grads = [torch.zeros_like(p) if g is None else g for g, p in zip(grads, params)]
        return grads, g_extra, loss

    def _calc_grad(self, loss, params, exp_zero_grad=0):
        grads = list(torch.autograd.grad(loss, params, allow_unused=True))
        missing = 0
        for i, p in enumerate(params):
            if grads[i] is None:
                grads[i] = torch.zeros_like(p)
                missing += 1
        assert missing == exp_zero_grad, (missing, exp_zero_grad)
        return grads

    # ---- public ------------------------------------------------------------------------------------
    def step(self, img_train, qst_train, label_train, img_valid, qst_valid, label_valid, eta,
             network_optimizer=None, unrolled=True):
        self.optimizer.zero_grad()
        if unrolled:
            self._backward_step_unrolled(img_train, qst_train, label_train, img_valid, qst_valid, label_valid,
                                         eta, network_optimizer)
        else:
            self._backward_step(img_valid, qst_valid, label_valid)
        pcd_ops.overlap_join(img_valid)        # no deferred weight-grad job of this step outlives it
        self.optimizer.step()

    def _backward_step(self, img_valid, qst_valid, label_valid):
        # first-order: d L_val / d alpha at the current weights (architect_vqa.py:53-55)
        arch = self.model.arch_parameters()
        with pcd_ops.weight_grads(False):      # only d/d(alpha, beta): activation-only cell backward, no weight-grad jobs
            loss = self.model._loss(img_valid, qst_valid, label_valid)
            grads = torch.autograd.grad(loss, arch)
        self._allreduce(list(grads))
        for a, g in zip(arch, grads):
            a.grad = g

    def _compute_unrolled_model(self, img, qst, label, eta, network_optimizer):
        model = self.model
        params, _ = self._lists(model)
        if self.reducer is not None and hasattr(model, "_loss_staged"):
            grads = self._staged(model, (img, qst, label), params)[0]
        else:
            loss = model._loss(img, qst, label, self.args.qst_only)
            grads = self._calc_grad(loss, params, self.exp_zero_grad)
            self._allreduce(grads)
        twin = self.unrolled_model()
        with torch.no_grad():
            self._copy_state(twin, model)                                # model_dict carries the live BN buffers
            if self.device_scalars:       # captured / benchmarked path: one launch over the flat runs
                pcd_flat.axpy_([p.data for p in self._lists(twin)[0]], grads, alpha=-eta)   # theta - eta * (0 + dtheta)
            else:                         # eager path: the multi-tensor op the goldens were pinned with (same rounding)
                torch._foreach_add_(self._lists(twin)[0], grads, alpha=-eta)
        return twin

    def _backward_step_unrolled(self, img_train, qst_train, label_train, img_valid, qst_valid, label_valid,
                                eta, network_optimizer):
        twin = self._compute_unrolled_model(img_train, qst_train, label_train, eta, network_optimizer)
        tparams = self._lists(twin)[0]
        tarch = twin.arch_parameters()
        if self.reducer is not None and hasattr(twin, "_loss_staged"):
            vector, extra, unrolled_loss = self._staged(twin, (img_valid, qst_valid, label_valid), tparams, extra=list(tarch))
            dalpha = [g.clone() for g in extra]
        else:
            unrolled_loss = twin._loss(img_valid, qst_valid, label_valid, self.args.qst_only)
            got = torch.autograd.grad(unrolled_loss, list(tarch) + tparams, allow_unused=True)
            dalpha = [g.clone() for g in got[:len(tarch)]]
            vector = [torch.zeros_like(p) if g is None else g for g, p in zip(got[len(tarch):], tparams)]
            self._allreduce(dalpha + vector)
        implicit = self._hessian_vector_product(vector, img_train, qst_train, label_train)
        with torch.no_grad():
            torch._foreach_add_(dalpha, implicit, alpha=-eta)
        for a, g in zip(self.model.arch_parameters(), dalpha):
            a.grad = g
        self.last.update(unrolled_loss=unrolled_loss.detach())

    def _hvp_passes_concurrent(self, pdata, vector, Rh, Rd, img, qst, label):
        """g+ = dL/d(alpha, beta) at w + R v on the model and g- at w - R v on the twin, as two independent branches (two
        streams; two branches of the CUDA graph when the step is captured).  The reference runs them one after the other on
        the same module (architect_vqa.py:106-118); the only state the second pass inherits from the first is the BatchNorm
        running statistics, r2 = 0.9 (0.9 r0 + 0.1 s+) + 0.1 s-, rebuilt below from the two branches' buffers."""
        model, twin = self.model, self._twin
        arch = model.arch_parameters()
        cuda = pdata[0].is_cuda              # (on the CPU emulation build the same code runs without streams: tests)
        if cuda:
            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
        nets = [m for m in model.modules() if hasattr(m, '_arena') and hasattr(m, 'cells')]
        tnets = [m for m in twin.modules() if hasattr(m, '_arena') and hasattr(m, 'cells')]
        with torch.no_grad():
            self._copy_state(twin, model)                        # twin <- (w, BN buffers r0, alphas)
            r0 = [n._arena().flat('running').clone() for n in nets]
            tdata = [p.data for p in self._lists(twin)[0]]
            pcd_flat.axpy_(pdata, vector, alpha=Rh, alpha_dev=Rd)            # model: w + R v
            pcd_flat.axpy_(tdata, vector, alpha=-Rh, alpha_dev=Rd)           # twin:  w - R v
        if cuda:
            self._side.wait_stream(cur)
        with (torch.cuda.stream(self._side) if cuda else pcd_ops.weight_grads(False)), pcd_ops.weight_grads(False):
            grads_n = list(torch.autograd.grad(twin._loss(img, qst, label, self.args.qst_only), twin.arch_parameters()))
        with pcd_ops.weight_grads(False):
            grads_p = list(torch.autograd.grad(model._loss(img, qst, label, self.args.qst_only), arch))
        if cuda:
            cur.wait_stream(self._side)
        with torch.no_grad():
            pcd_flat.axpy_(pdata, vector, alpha=-Rh, alpha_dev=Rd)           # back to w
            for n, t, r in zip(nets, tnets, r0):                 # BatchNorm buffers as after the two passes in sequence
                run = n._arena().flat('running')
                run.mul_(1.0 - pcd_ops.BN_MOMENTUM).add_(t._arena().flat('running')).sub_(r, alpha=1.0 - pcd_ops.BN_MOMENTUM)
                n._arena().flat('nbt').add_(1)
        return grads_p, grads_n

    def _hessian_vector_product(self, vector, img, qst, label, r=1e-2):
        model = self.model
        params = self._lists(model)[0]
        arch = model.arch_parameters()
        pdata = [p.data for p in params]
        with torch.no_grad():
            # |v| exactly as before (per-tensor norms, then the norm of those): one multi-tensor launch; R = r / |v| enters
            # w +- R v, where a different last bit already moves the toy-size goldens (ReLU ties)
            vnorm = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(vector)))
            if self.device_scalars:
                R = r / vnorm                                        # 0-dim device tensor: no host synchronisation
                Rd, Rh = R.reshape(1), 1.0
            else:
                R = (r / vnorm).item()
                Rd, Rh = None, R
        if self.concurrent_hvp and self._twin is not None:
            grads_p, grads_n = self._hvp_passes_concurrent(pdata, vector, Rh, Rd, img, qst, label)
            self._allreduce(grads_p + grads_n)
            self.last.update(g_pos=grads_p, g_neg=grads_n, R=R, vnorm=vnorm)
            return [(x - y).div_(2 * R) for x, y in zip(grads_p, grads_n)]

        def shift(k):       # w += k * R * v: flat kernel on the captured path, the goldens' multi-tensor op on the eager one
            if self.device_scalars:
                pcd_flat.axpy_(pdata, vector, alpha=k * Rh, alpha_dev=Rd)
            else:
                torch._foreach_add_(params, vector, alpha=k * R)
        with torch.no_grad():
            shift(1.0)
        with pcd_ops.weight_grads(False):      # only d/d(alpha, beta) is needed at w +- R v
            grads_p = list(torch.autograd.grad(model._loss(img, qst, label, self.args.qst_only), arch))
        with torch.no_grad():
            shift(-2.0)
        with pcd_ops.weight_grads(False):
            grads_n = list(torch.autograd.grad(model._loss(img, qst, label, self.args.qst_only), arch))
        with torch.no_grad():
            shift(1.0)
        self._allreduce(grads_p + grads_n)
        self.last.update(g_pos=grads_p, g_neg=grads_n, R=R, vnorm=vnorm)
        return [(x - y).div_(2 * R) for x, y in zip(grads_p, grads_n)]

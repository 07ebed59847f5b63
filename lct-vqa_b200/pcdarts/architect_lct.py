"""ArchitectLct — the 3-stage "Learning by Creating Tests" alpha-step (basic_vqa/pcdarts/architect_lct.py:13-236).

    unroll EF one SGD step on the training batch          (architect_lct.py:94-116)
    -> EF' generates pseudo questions / soft answers      (:53-55, temperature softmax)
    -> unroll W one step on real + pseudo QA              (:118-140)
    -> grad of W' validation loss w.r.t. W'               (:62-65)
    -> kappa = finite-difference HVP through W +- R v, w.r.t. EF' weights   (:211-235)
    -> gamma = finite-difference HVP through EF +- R kappa, w.r.t. alpha/beta (:181-209)
    -> alpha.grad = gamma * ef_lr * w_lr ; Adam step       (:84-88, :44)

Same public signatures and arithmetic as the reference.  The search-network passes (6 forward, 5 backward per step) run on
the sm_100a kernels through the EF model; the W model (VGG19) is stock PyTorch.
"""
import logging

import torch
import torch.nn.functional as F

import config
import pcd_ops


def _concat(xs):
    return torch.cat([x.reshape(-1) for x in xs])


_LAST_INSTANCE = None      # debugging aid: the most recently constructed architect


class ArchitectLct(object):
    def __init__(self, ef_model, w_model, ef_optimizer, w_optimizer, reducer=None):
        global _LAST_INSTANCE
        # data parallel (not in the reference, which is single-process): every gradient evaluation of the step is averaged
        # across ranks, so the finite-difference radii R = r / |.| and the final alpha-step are identical on all replicas
        self.reducer = reducer
        _LAST_INSTANCE = self
        self.ef_model = ef_model
        self.w_model = w_model
        self.ef_momentum = 0
        self.ef_weight_decay = 0
        self.optimizer = torch.optim.Adam(self.ef_model.arch_parameters(), lr=config.ARCH_LEARNING_RATE, betas=(0.5, 0.999),
                                          weight_decay=config.ARCH_WEIGHT_DECAY)
        self.ef_optimizer = ef_optimizer
        self.w_optimizer = w_optimizer
        self.last = {}          # intermediate quantities of the last step (tests, logging)
        self._twins = {}        # source model id -> persistent unrolled twin
        self.device_scalars = False     # True: no host reads inside step() (CUDA-graph capture)

    def step(self, img_train, qst_train, label_train, img_valid, qst_valid, label_valid, ef_lr, w_lr):
        self.ef_optimizer.zero_grad()
        self.w_optimizer.zero_grad()
        self.optimizer.zero_grad()
        self.last = {}
        # second-order (unrolled) unconditionally: a first-order approximation does not exist for this objective
        self._backward_step_unrolled(img_train, qst_train, label_train, img_valid, qst_valid, label_valid, ef_lr, w_lr)
        pcd_ops.overlap_join(img_train)       # no deferred weight-grad job of this step outlives it
        self.optimizer.step()

    def _backward_step_unrolled(self, img_train, qst_train, label_train, img_valid, qst_valid, label_valid, ef_lr, w_lr):
        unrolled_ef = self._compute_unrolled_model(img_train, qst_train, label_train, ef_lr, self.ef_optimizer, self.ef_model,
                                                   self.ef_model._loss)

        def pseudo_qa_fn():
            pseudo_qst, pseudo_ans = unrolled_ef.generate(img_train)
            return pseudo_qst, F.softmax(pseudo_ans / config.TEMPERATURE, dim=1)

        def qa_fn():
            return qst_train, label_train

        pseudo_qst, pseudo_ans = pseudo_qa_fn()
        unrolled_w = self._compute_unrolled_model_2(img_train, qst_train, label_train, pseudo_qst, pseudo_ans, w_lr,
                                                    self.w_optimizer, self.w_model, self.w_model._soft_loss, exp_zero_grad=36)
        unrolled_loss = unrolled_w._loss(img_valid, qst_valid, label_valid)
        grad_wprime = self._calc_grad(unrolled_loss, unrolled_w.parameters, exp_zero_grad=36)
        kappa = self._hessian_vector_product_2(grad_wprime, img_train, qa_fn, pseudo_qa_fn, self.w_model, self.w_model._soft_loss,
                                               unrolled_ef.parameters, exp_zero_grad=2)
        gamma = self._hessian_vector_product(kappa, img_train, qa_fn, self.ef_model, self.ef_model._loss,
                                             self.ef_model.arch_parameters, exp_zero_grad=0)
        for v, g in zip(self.ef_model.arch_parameters(), gamma):
            if v.grad is None:
                v.grad = (g.detach() * ef_lr * w_lr).clone()
            else:
                v.grad.data.copy_(g.detach() * ef_lr * w_lr)
        self.last.update(unrolled_loss=unrolled_loss.detach(), kappa_norm=_concat(kappa).norm().detach(),
                         grad_wprime_norm=_concat(grad_wprime).norm().detach())
        if not self.device_scalars:
            logging.info("| TRAIN SET | STAGE3 | W'-Val-Loss: {:.4f}".format(unrolled_loss.item()))

    def _compute_unrolled_model(self, img, qst, label, eta, optimizer, model, loss_fn, exp_zero_grad=0, weight_decay=0, momentum=0):
        loss = loss_fn(img, qst, label)
        return self._unroll(loss, eta, optimizer, model, exp_zero_grad, weight_decay)

    def _compute_unrolled_model_2(self, img, qst, label, pseudo_qst, pseudo_label, eta, optimizer, model, loss_fn, exp_zero_grad=0,
                                  weight_decay=0, momentum=0):
        loss = loss_fn(img, qst, label, pseudo_qst, pseudo_label)
        return self._unroll(loss, eta, optimizer, model, exp_zero_grad, weight_decay)

    def _unroll(self, loss, eta, optimizer, model, exp_zero_grad, weight_decay):
        theta = _concat(model.parameters()).data
        # the reference looks up SGD momentum buffers, finds none under Adam and falls back to zeros (:107-111)
        moment = torch.zeros_like(theta)
        grads = self._calc_grad(loss, model.parameters, exp_zero_grad)
        dtheta = _concat(grads).data + weight_decay * theta
        return self._construct_model_from_theta(theta.sub(moment + dtheta, alpha=eta), model)

    def _construct_model_from_theta(self, theta, model):
        """model.new() + load_state_dict(model's buffers, theta as parameters) (architect_lct.py:142-156).  The reference
        builds a fresh model (for W: a whole VGG19) on every call; here the twin is built once per source model and
        refilled in place — same parameters, same buffers, same dropout configuration."""
        twin = self._twins.get(id(model))
        if twin is None:
            twin = model.new().to(config.DEVICE)
            self._twins[id(model)] = twin
        with torch.no_grad():
            src_buf, dst_buf = dict(model.named_buffers()), dict(twin.named_buffers())
            if src_buf:
                keys = [k for k in dst_buf if k in src_buf]
                torch._foreach_copy_([dst_buf[k] for k in keys], [src_buf[k] for k in keys])
            views, offset = [], 0
            params = list(twin.parameters())
            for v in params:
                n = v.numel()
                views.append(theta[offset: offset + n].view(v.size()))
                offset += n
            assert offset == len(theta)
            torch._foreach_copy_([p.data for p in params], views)
            if hasattr(model, "arch_parameters"):       # Network.new() copies the current alphas / betas (model_search.py:138-143)
                torch._foreach_copy_([t.data for t in twin.arch_parameters()], [t.data for t in model.arch_parameters()])
        twin.train(model.training)
        return twin

    def _calc_grad(self, loss, param_fn, exp_zero_grad=0):
        grads = list(torch.autograd.grad(loss, list(param_fn()), allow_unused=True))
        num_zero_grad = 0
        for i, p in enumerate(param_fn()):
            if grads[i] is None:
                grads[i] = torch.zeros_like(p)
                num_zero_grad += 1
            else:
                assert grads[i].shape == p.shape
        assert num_zero_grad == exp_zero_grad, (num_zero_grad, exp_zero_grad)
        if self.reducer is not None:
            self.reducer(grads)
        self.last.setdefault("calls", []).append((loss.detach(), _concat(grads).norm().detach()))
        return grads

    def _perturb(self, params, vector, R):
        """Returns shift(k): params += k * R * vector.  With `device_scalars` R stays a 0-dim device tensor (R * v is formed
        once), so the step has no host synchronisation and can be captured in a CUDA graph (search.GraphedLctStep)."""
        data = [p.data for p in params]
        if self.device_scalars:
            step_v = torch._foreach_mul(list(vector), R)
            return lambda k: torch._foreach_add_(data, step_v, alpha=float(k))
        Rf = R.item()
        return lambda k: torch._foreach_add_(data, list(vector), alpha=k * Rf)

    def _hessian_vector_product(self, vector, img, qa_fn, model, loss_fn, param_fn, r=1e-2, exp_zero_grad=0):
        R = r / _concat(vector).norm()
        shift = self._perturb(list(model.parameters()), vector, R)
        # only d/d(alpha, beta) is wanted at EF +- R kappa: activation-only backward passes, no weight-grad jobs
        arch_ids = {id(t) for t in model.arch_parameters()} if hasattr(model, "arch_parameters") else set()
        wgrads = not all(id(t) in arch_ids for t in param_fn())
        shift(1)
        qst, ans = qa_fn()
        with pcd_ops.weight_grads(wgrads):
            grads_p = self._calc_grad(loss_fn(img, qst, ans), param_fn, exp_zero_grad)
        shift(-2)
        qst, ans = qa_fn()
        with pcd_ops.weight_grads(wgrads):
            grads_n = self._calc_grad(loss_fn(img, qst, ans), param_fn, exp_zero_grad)
        shift(1)
        self.last.update(gamma_p=grads_p, gamma_n=grads_n, gamma_R=R.detach())
        return [(x - y).div_(2 * R) for x, y in zip(grads_p, grads_n)]

    def _hessian_vector_product_2(self, vector, img, qa_fn, pseudo_qa_fn, model, loss_fn, param_fn, r=1e-2, exp_zero_grad=0):
        R = r / _concat(vector).norm()
        shift = self._perturb(list(model.parameters()), vector, R)
        shift(1)
        qst, ans = qa_fn()
        pseudo_qst, pseudo_ans = pseudo_qa_fn()
        grads_p = self._calc_grad(loss_fn(img, qst, ans, pseudo_qst, pseudo_ans), param_fn, exp_zero_grad)
        shift(-2)
        qst, ans = qa_fn()
        pseudo_qst, pseudo_ans = pseudo_qa_fn()
        grads_n = self._calc_grad(loss_fn(img, qst, ans, pseudo_qst, pseudo_ans), param_fn, exp_zero_grad)
        shift(1)
        self.last.update(kappa_p=grads_p, kappa_n=grads_n, kappa_R=R.detach())
        return [(x - y).div_(2 * R) for x, y in zip(grads_p, grads_n)]

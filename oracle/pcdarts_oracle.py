"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement ("port") of the PC-DARTS-VQA search hot path of aahamed/LCT-VQA,
written functionally over a flat ``{state_dict key: tensor}`` dictionary so that a
reference ``state_dict()`` can be fed to it unchanged.  All arithmetic goes
through stock ATen CPU ops (the same third-party kernels the reference reaches:
torch 2.11.0, see SURVEY.md §8c "Third-party arithmetic"); gradients come from
torch autograd exactly as in the reference.

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports the
unmodified reference from /root/reference (in the build container), runs it on
seeded inputs and commits the outputs; ``tests/test_oracle_golden.py`` replays
them through this file.

Citations are path:line under /root/reference/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# darts_vqa/pcdarts/genotypes.py:5-14 — order defines the alpha columns
PRIMITIVES = ("none", "max_pool_3x3", "avg_pool_3x3", "skip_connect",
              "sep_conv_3x3", "sep_conv_5x5", "dil_conv_3x3", "dil_conv_5x5")
BN_EPS = 1e-5       # nn.BatchNorm2d default
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default
K_PARTIAL = 4       # model_search.py:36


# --------------------------------------------------------------------------
# index maps (integer work, numpy; bit-exact contract)
# --------------------------------------------------------------------------
def shuffle_perm(channels: int, groups: int = K_PARTIAL) -> np.ndarray:
    """perm[o] = input channel that lands in output channel o.

    model_search.py:14-28: view (B, g, C/g, H, W) -> transpose(1, 2) -> flatten,
    i.e. out[:, j*g + q] = in[:, q*(C/g) + j].
    """
    per = channels // groups
    o = np.arange(channels)
    return (o % groups) * per + (o // groups)


def channel_shuffle_np(x: np.ndarray, groups: int = K_PARTIAL) -> np.ndarray:
    return x[:, shuffle_perm(x.shape[1], groups)]


def adaptive_windows(n_in: int, n_out: int) -> List[Tuple[int, int]]:
    """AdaptiveAvgPool2d window [start, end) per output index (model_search.py:129)."""
    return [((i * n_in) // n_out, -((-(i + 1) * n_in) // n_out)) for i in range(n_out)]


def edge_table(steps: int = 4) -> List[Tuple[int, int]]:
    """(node, source state index) for every edge, in `_ops` order (model_search.py:76-81)."""
    return [(i, j) for i in range(steps) for j in range(2 + i)]


def cell_plan(C: int, layers: int, multiplier: int = 4, stem_multiplier: int = 3):
    """Per cell: (C_prev_prev, C_prev, C_curr, reduction, reduction_prev) — model_search.py:115-127."""
    c_pp = c_p = stem_multiplier * C
    c_cur = C
    red_prev = False
    plan = []
    for i in range(layers):
        red = i in (layers // 3, 2 * layers // 3)
        if red:
            c_cur *= 2
        plan.append((c_pp, c_p, c_cur, red, red_prev))
        red_prev = red
        c_pp, c_p = c_p, multiplier * c_cur
    return plan


# --------------------------------------------------------------------------
# primitive ops (operations.py)
# --------------------------------------------------------------------------
class BNState:
    """Optional running-stat side effects: dict key -> tensor, mutated like nn.BatchNorm2d."""

    def __init__(self, state: Optional[Params] = None):
        self.state = state

    def get(self, prefix: str):
        if self.state is None:
            return None, None, None
        return (self.state.get(prefix + "running_mean"), self.state.get(prefix + "running_var"),
                self.state.get(prefix + "num_batches_tracked"))


def batch_norm(P: Params, bns: BNState, prefix: str, z: Tensor, training: bool = True) -> Tensor:
    """nn.BatchNorm2d forward; affine iff `<prefix>weight` exists (only the stem BN, model_search.py:112)."""
    rm, rv, nbt = bns.get(prefix)
    if training and nbt is not None:
        nbt += 1                                 # nn.BatchNorm2d.forward bumps the counter first
    use_batch = training or rm is None
    return F.batch_norm(z, rm, rv, P.get(prefix + "weight"), P.get(prefix + "bias"),
                        use_batch, BN_MOMENTUM, BN_EPS)


def relu_conv_bn(P, bns, pre, x, training=True):
    """operations.py:22-33 with kernel 1, stride 1, pad 0 (the only use: model_search.py:70-71)."""
    z = F.conv2d(F.relu(x), P[pre + "op.1.weight"])
    return batch_norm(P, bns, pre + "op.2.", z, training)


def factorized_reduce(P, bns, pre, x, training=True):
    """operations.py:90-104: two stride-2 1x1 convs, the second on the (1,1)-shifted grid."""
    r = F.relu(x)
    z = torch.cat([F.conv2d(r, P[pre + "conv_1.weight"], stride=2),
                   F.conv2d(r[:, :, 1:, 1:], P[pre + "conv_2.weight"], stride=2)], dim=1)
    return batch_norm(P, bns, pre + "bn.", z, training)


def sep_conv(P, bns, pre, x, k, stride, training=True):
    """operations.py:50-66 (padding = k // 2)."""
    c = x.shape[1]
    t = F.conv2d(F.relu(x), P[pre + "op.1.weight"], stride=stride, padding=k // 2, groups=c)
    z = F.conv2d(t, P[pre + "op.2.weight"])
    y = batch_norm(P, bns, pre + "op.3.", z, training)
    t = F.conv2d(F.relu(y), P[pre + "op.5.weight"], stride=1, padding=k // 2, groups=c)
    z = F.conv2d(t, P[pre + "op.6.weight"])
    return batch_norm(P, bns, pre + "op.7.", z, training)


def dil_conv(P, bns, pre, x, k, stride, training=True):
    """operations.py:35-47 with dilation 2, padding = k - 1 (OPS table operations.py:12-13)."""
    c = x.shape[1]
    t = F.conv2d(F.relu(x), P[pre + "op.1.weight"], stride=stride, padding=k - 1, dilation=2, groups=c)
    z = F.conv2d(t, P[pre + "op.2.weight"])
    return batch_norm(P, bns, pre + "op.3.", z, training)


def candidate_op(P, bns, pre, idx, x, stride, training=True):
    """One entry of MixedOp._ops (model_search.py:37-41): OPS[PRIMITIVES[idx]](c, stride, False)."""
    if idx == 0:      # Zero, operations.py:78-87
        return (x if stride == 1 else x[:, :, ::stride, ::stride]).mul(0.)
    if idx == 1:
        return batch_norm(P, bns, pre + "1.", F.max_pool2d(x, 3, stride, 1), training)
    if idx == 2:
        return batch_norm(P, bns, pre + "1.",
                          F.avg_pool2d(x, 3, stride, 1, count_include_pad=False), training)
    if idx == 3:
        return x if stride == 1 else factorized_reduce(P, bns, pre, x, training)
    if idx == 4:
        return sep_conv(P, bns, pre, x, 3, stride, training)
    if idx == 5:
        return sep_conv(P, bns, pre, x, 5, stride, training)
    if idx == 6:
        return dil_conv(P, bns, pre, x, 3, stride, training)
    return dil_conv(P, bns, pre, x, 5, stride, training)


def channel_shuffle(x: Tensor, groups: int = K_PARTIAL) -> Tensor:
    perm = torch.from_numpy(shuffle_perm(x.shape[1], groups))
    return x.index_select(1, perm)


def mixed_op(P, bns, pre, x, weights, stride, training=True):
    """model_search.py:44-58.  `pre` ends with '_ops.' of the MixedOp."""
    c = x.shape[1] // K_PARTIAL
    xs, rest = x[:, :c], x[:, c:]
    acc = 0
    for k in range(len(PRIMITIVES)):            # python sum(): ((0 + w0 o0) + w1 o1) + ...
        acc = acc + weights[k] * candidate_op(P, bns, f"{pre}{k}.", k, xs, stride, training)
    if stride != 1:
        rest = F.max_pool2d(rest, 2, 2)
    return channel_shuffle(torch.cat([acc, rest], dim=1))


def cell_forward(P, bns, pre, s0, s1, weights, weights2, reduction, reduction_prev,
                 steps=4, multiplier=4, training=True):
    """model_search.py:83-94."""
    if reduction_prev:
        s0 = factorized_reduce(P, bns, pre + "preprocess0.", s0, training)
    else:
        s0 = relu_conv_bn(P, bns, pre + "preprocess0.", s0, training)
    s1 = relu_conv_bn(P, bns, pre + "preprocess1.", s1, training)
    states = [s0, s1]
    e = 0
    for _ in range(steps):
        acc = 0
        for j, h in enumerate(states):
            stride = 2 if reduction and j < 2 else 1
            acc = acc + weights2[e + j] * mixed_op(P, bns, f"{pre}_ops.{e + j}._ops.", h,
                                                   weights[e + j], stride, training)
        e += len(states)
        states.append(acc)
    return torch.cat(states[-multiplier:], dim=1)


def arch_weights(alphas: Tensor, betas: Tensor, steps: int = 4):
    """model_search.py:153-174: row softmax of alphas; grouped softmax of betas over 2,3,4,5."""
    w = F.softmax(alphas, dim=-1)
    parts, start = [], 0
    for n in range(2, 2 + steps):
        parts.append(F.softmax(betas[start:start + n], dim=-1))
        start += n
    return w, torch.cat(parts, dim=0)


def network_forward(P, bns, arch: Sequence[Tensor], x, C=16, layers=4, pre="", training=True):
    """model_search.py:145-179.  arch = (alphas_normal, alphas_reduce, betas_normal, betas_reduce)."""
    n, _, h, w = x.shape
    x = x.expand(n, 3, h, w)
    z = F.conv2d(x, P[pre + "stem.0.weight"], padding=1)
    s0 = s1 = batch_norm(P, bns, pre + "stem.1.", z, training)
    for i, (_, _, _, red, red_prev) in enumerate(cell_plan(C, layers)):
        wts, wts2 = arch_weights(arch[1], arch[3]) if red else arch_weights(arch[0], arch[2])
        s0, s1 = s1, cell_forward(P, bns, f"{pre}cells.{i}.", s0, s1, wts, wts2, red, red_prev,
                                  training=training)
    return F.adaptive_avg_pool2d(s1, 7).flatten(1)


# --------------------------------------------------------------------------
# VQA model (darts_vqa/vqa_model.py)
# --------------------------------------------------------------------------
def lstm_forward(P, pre, x, h0, c0):
    """nn.LSTM, 1 layer, seq-first (vqa_model.py:91,182).  Gate order i,f,g,o."""
    w_ih, w_hh = P[pre + "weight_ih_l0"], P[pre + "weight_hh_l0"]
    b = P[pre + "bias_ih_l0"] + P[pre + "bias_hh_l0"]
    h, c = h0, c0
    xs = x @ w_ih.t() + b
    outs = []
    for t in range(x.shape[0]):
        i, f, g, o = (xs[t] + h @ w_hh.t()).chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs), h, c


def vqa_forward(P, bns, arch, img, qst, C=16, layers=4, training=True, dropout_p=0.5):
    """VqaModel.forward vqa_model.py:300-318 with DartsEncoder :58-66 and QstEncoder :170-196."""
    feat = network_forward(P, bns, arch, img, C, layers, "img_encoder.darts.", training)
    feat = F.linear(feat, P["img_encoder.fc.weight"], P["img_encoder.fc.bias"])
    feat = feat / feat.norm(p=2, dim=1, keepdim=True).detach()
    emb = torch.tanh(F.embedding(qst, P["qst_encoder.word2vec.weight"])).transpose(0, 1)
    out, h, c = lstm_forward(P, "qst_encoder.lstm.", emb, feat, feat)
    qf = torch.tanh(torch.cat([h, c], dim=1))
    qf = F.linear(qf, P["qst_encoder.fc2.weight"], P["qst_encoder.fc2.bias"])
    q_out = F.linear(torch.tanh(out.transpose(0, 1)), P["qst_encoder.fc1.weight"],
                     P["qst_encoder.fc1.bias"])
    z = F.dropout(torch.tanh(feat * qf), dropout_p, training)
    z = F.dropout(torch.tanh(F.linear(z, P["fc1.weight"], P["fc1.bias"])), dropout_p, training)
    return F.linear(z, P["fc2.weight"], P["fc2.bias"]), q_out


def vqa_loss(P, bns, arch, img, qst, label, qst_only=False, **kw):
    """VqaModel._loss vqa_model.py:351-364."""
    ans, q_out = vqa_forward(P, bns, arch, img, qst, **kw)
    q_loss = F.cross_entropy(q_out[:, :-1].flatten(end_dim=1), qst[:, 1:].flatten())
    return q_loss if qst_only else F.cross_entropy(ans, label) + q_loss


# --------------------------------------------------------------------------
# Architect (darts_vqa/pcdarts/architect_vqa.py)
# --------------------------------------------------------------------------
def adam_step(params, grads, state, lr, betas=(0.5, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam single step (L2-style weight decay), architect_vqa.py:19-21."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g + weight_decay * p
        m = state.setdefault(("m", i), torch.zeros_like(p))
        v = state.setdefault(("v", i), torch.zeros_like(p))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = (v.sqrt() / math.sqrt(1 - betas[1] ** t)).add_(eps)
        p.sub_((lr / (1 - betas[0] ** t)) * m / denom)


def _grads(loss, tensors):
    gs = torch.autograd.grad(loss, tensors, allow_unused=True)
    return [torch.zeros_like(t) if g is None else g for g, t in zip(gs, tensors)]


def arch_grad_first_order(P, bns, arch, valid, **kw):
    """Architect._backward_step architect_vqa.py:53-55."""
    return _grads(vqa_loss(P, bns, arch, *valid, **kw), list(arch))


def arch_grad_unrolled(P, bns, arch, train, valid, eta, param_keys, r=1e-2, qst_only=False,
                       debug=None, **kw):
    """Architect._backward_step_unrolled architect_vqa.py:57-88 (momentum = weight decay = 0, :15-16).

    `param_keys`: the reference's named_parameters() order.  BN buffers of the unrolled model are
    a copy of the live ones (load_state_dict of model_dict, :91-102); the live model's buffers are
    bumped by the three `_loss` calls on it (:25,:109,:114).
    """
    ws = [P[k] for k in param_keys]
    g_train = _grads(vqa_loss(P, bns, arch, *train, qst_only=qst_only, **kw), ws)
    P2 = dict(P)
    for k, w, g in zip(param_keys, ws, g_train):
        P2[k] = (w.detach() - eta * g).requires_grad_(True)
    bns2 = BNState(None if bns.state is None else {k: v.clone() for k, v in bns.state.items()})
    arch2 = [a.detach().clone().requires_grad_(True) for a in arch]
    loss2 = vqa_loss(P2, bns2, arch2, *valid, qst_only=qst_only, **kw)
    got = _grads(loss2, arch2 + [P2[k] for k in param_keys])
    dalpha, vector = got[:len(arch2)], got[len(arch2):]
    R = r / torch.cat([v.reshape(-1) for v in vector]).norm()
    with torch.no_grad():
        for w, v in zip(ws, vector):
            w.add_(v, alpha=R)
    g_pos = _grads(vqa_loss(P, bns, arch, *train, qst_only=qst_only, **kw), list(arch))
    with torch.no_grad():
        for w, v in zip(ws, vector):
            w.sub_(v, alpha=2 * R)
    g_neg = _grads(vqa_loss(P, bns, arch, *train, qst_only=qst_only, **kw), list(arch))
    with torch.no_grad():
        for w, v in zip(ws, vector):
            w.add_(v, alpha=R)
    if debug is not None:
        debug.update(g_pos=g_pos, g_neg=g_neg, R=R, dalpha=[d.clone() for d in dalpha], loss2=loss2,
                     w_prime={k: P2[k].detach().clone() for k in param_keys},
                     bn_prime=None if bns2.state is None else {k: v.clone() for k, v in bns2.state.items()})
    return [d - eta * (gp - gn) / (2 * R) for d, gp, gn in zip(dalpha, g_pos, g_neg)]


def architect_step(P, bns, arch, adam_state, train, valid, eta, param_keys, unrolled=True,
                   arch_lr=6e-4, arch_wd=1e-3, **kw):
    """Architect.step architect_vqa.py:40-51."""
    if unrolled:
        g = arch_grad_unrolled(P, bns, arch, train, valid, eta, param_keys, **kw)
    else:
        g = arch_grad_first_order(P, bns, arch, valid, **kw)
    with torch.no_grad():
        adam_step(list(arch), g, adam_state, arch_lr, weight_decay=arch_wd)
    return g


def w_step(P, bns, arch, batch, adam_state, param_keys, lr=1e-3, clip=5.0, debug=None, **kw):
    """The w-step of Experiment.train, darts_vqa/experiment.py:187-200."""
    ws = [P[k] for k in param_keys]
    loss = vqa_loss(P, bns, arch, *batch, **kw)
    if debug is not None:      # loss.backward() of the reference also accumulates into the alphas' / betas' .grad
        got = _grads(loss, ws + list(arch))
        gs, debug["arch_grads"] = got[:len(ws)], got[len(ws):]
    else:
        gs = _grads(loss, ws)
    total = torch.norm(torch.stack([g.norm(2) for g in gs]), 2)
    coef = torch.clamp(clip / (total + 1e-6), max=1.0)      # nn.utils.clip_grad_norm_
    if debug is not None:
        debug.update(grads=[g.clone() for g in gs], total_norm=total.clone(), clip_coef=coef.clone())
    with torch.no_grad():
        adam_step(ws, [g * coef for g in gs], adam_state, lr, betas=(0.9, 0.999))
    return loss.detach()


# --------------------------------------------------------------------------
# helpers shared by tests / bench (deterministic weights that do not depend on module init order)
# --------------------------------------------------------------------------
def split_state(sd: Params):
    """state_dict -> (params requiring grad, BN buffers)."""
    buf = {k: v for k, v in sd.items() if k.endswith(("running_mean", "running_var",
                                                       "num_batches_tracked"))}
    par = {k: v for k, v in sd.items() if k not in buf}
    return par, buf


def synth_fill_(sd: Params, seed: int = 1234) -> None:
    """Overwrite every float tensor of a state_dict with seeded values, key by key in sorted order.

    Used by the golden generator and by the tests so that reference, oracle and product hold
    identical weights without relying on module construction order.
    """
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        if not v.is_floating_point():
            v.zero_()
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif v.dim() <= 1:
            v.copy_(torch.randn(v.shape, generator=g) * 0.1 + (1.0 if k.endswith("stem.1.weight") else 0.0))
        else:
            fan_in = v[0].numel()
            v.copy_(torch.randn(v.shape, generator=g) * (1.0 / math.sqrt(fan_in)))


# --------------------------------------------------------------------------
# state_dict specs (key -> shape) in the reference's registration order
# --------------------------------------------------------------------------
def _bn_spec(pre, c, affine=False):
    d = {}
    if affine:
        d[pre + "weight"] = (c,)
        d[pre + "bias"] = (c,)
    d[pre + "running_mean"] = (c,)
    d[pre + "running_var"] = (c,)
    d[pre + "num_batches_tracked"] = ()
    return d


def mixed_op_spec(C: int, stride: int, pre: str = "_ops."):
    """Keys of MixedOp(C, stride).state_dict() (model_search.py:32-41, operations.py)."""
    c = C // K_PARTIAL
    d = {}
    d.update(_bn_spec(pre + "1.1.", c))
    d.update(_bn_spec(pre + "2.1.", c))
    if stride != 1:
        d[pre + "3.conv_1.weight"] = (c // 2, c, 1, 1)
        d[pre + "3.conv_2.weight"] = (c // 2, c, 1, 1)
        d.update(_bn_spec(pre + "3.bn.", c))
    for idx, k in ((4, 3), (5, 5)):
        p = f"{pre}{idx}.op."
        d[p + "1.weight"] = (c, 1, k, k)
        d[p + "2.weight"] = (c, c, 1, 1)
        d.update(_bn_spec(p + "3.", c))
        d[p + "5.weight"] = (c, 1, k, k)
        d[p + "6.weight"] = (c, c, 1, 1)
        d.update(_bn_spec(p + "7.", c))
    for idx, k in ((6, 3), (7, 5)):
        p = f"{pre}{idx}.op."
        d[p + "1.weight"] = (c, 1, k, k)
        d[p + "2.weight"] = (c, c, 1, 1)
        d.update(_bn_spec(p + "3.", c))
    return d


def cell_spec(cpp, cp, C, reduction, reduction_prev, pre="", steps=4):
    d = {}
    if reduction_prev:
        d[pre + "preprocess0.conv_1.weight"] = (C // 2, cpp, 1, 1)
        d[pre + "preprocess0.conv_2.weight"] = (C // 2, cpp, 1, 1)
        d.update(_bn_spec(pre + "preprocess0.bn.", C))
    else:
        d[pre + "preprocess0.op.1.weight"] = (C, cpp, 1, 1)
        d.update(_bn_spec(pre + "preprocess0.op.2.", C))
    d[pre + "preprocess1.op.1.weight"] = (C, cp, 1, 1)
    d.update(_bn_spec(pre + "preprocess1.op.2.", C))
    for e, (_, j) in enumerate(edge_table(steps)):
        stride = 2 if reduction and j < 2 else 1
        d.update(mixed_op_spec(C, stride, f"{pre}_ops.{e}._ops."))
    return d


def network_spec(C=16, layers=4, pre=""):
    d = {pre + "stem.0.weight": (3 * C, 3, 3, 3)}
    d.update(_bn_spec(pre + "stem.1.", 3 * C, affine=True))
    for i, (cpp, cp, cc, red, rp) in enumerate(cell_plan(C, layers)):
        d.update(cell_spec(cpp, cp, cc, red, rp, f"{pre}cells.{i}."))
    return d


def vqa_spec(embed_size, qst_vocab_size, ans_vocab_size, word_embed_size, num_layers, hidden_size,
             C=16, layers=4):
    assert num_layers == 1
    d = network_spec(C, layers, "img_encoder.darts.")
    d["img_encoder.fc.weight"] = (embed_size, 256 * 49)
    d["img_encoder.fc.bias"] = (embed_size,)
    d["qst_encoder.word2vec.weight"] = (qst_vocab_size, word_embed_size)
    d["qst_encoder.lstm.weight_ih_l0"] = (4 * hidden_size, word_embed_size)
    d["qst_encoder.lstm.weight_hh_l0"] = (4 * hidden_size, hidden_size)
    d["qst_encoder.lstm.bias_ih_l0"] = (4 * hidden_size,)
    d["qst_encoder.lstm.bias_hh_l0"] = (4 * hidden_size,)
    d["qst_encoder.fc1.weight"] = (qst_vocab_size, hidden_size)
    d["qst_encoder.fc1.bias"] = (qst_vocab_size,)
    d["qst_encoder.fc2.weight"] = (embed_size, 2 * num_layers * hidden_size)
    d["qst_encoder.fc2.bias"] = (embed_size,)
    d["fc1.weight"] = (ans_vocab_size, embed_size)
    d["fc1.bias"] = (ans_vocab_size,)
    d["fc2.weight"] = (ans_vocab_size, ans_vocab_size)
    d["fc2.bias"] = (ans_vocab_size,)
    return d


def alloc_state(spec, seed=None) -> Params:
    sd = {k: (torch.zeros(s, dtype=torch.long) if k.endswith("num_batches_tracked")
              else torch.zeros(s)) for k, s in spec.items()}
    if seed is not None:
        synth_fill_(sd, seed)
    return sd

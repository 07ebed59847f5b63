"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference (aahamed/LCT-VQA, darts_vqa/) is imported as-is with config.DEVICE='cpu', fed weights
from oracle.synth_fill_ (seeded, independent of module construction order) and seeded inputs.  The
outputs are committed so that the oracle — and through it the CUDA path — stay pinned on machines
where /root/reference does not exist (the GPU box).  torch.set_num_threads(1): the reference is
bit-deterministic for a fixed thread count.
"""
import argparse
import os
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("LCT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "darts_vqa"))

import config  # noqa: E402  (reference)
config.DEVICE = "cpu"
from pcdarts.model_search import MixedOp, Cell, Network, channel_shuffle  # noqa: E402
from pcdarts.architect_vqa import Architect  # noqa: E402
from vqa_model import VqaModel  # noqa: E402
from oracle.pcdarts_oracle import synth_fill_  # noqa: E402

MIXED_CASES = [  # name, C, stride, B, H
    ("t1", 16, 1, 2, 16), ("t2", 32, 2, 2, 16), ("t3", 32, 1, 2, 12),
    ("t4", 64, 2, 2, 8), ("t5", 64, 1, 3, 8), ("odd", 16, 1, 1, 9), ("rect2", 16, 2, 2, 20),
]
CELL_CASES = [  # name, Cpp, Cp, C, reduction, reduction_prev, B, H (of s1), s0 is 2H if reduction_prev
    ("normal", 48, 48, 16, False, False, 2, 16),
    ("reduce", 48, 64, 32, True, False, 2, 16),
    ("reduce_rp", 64, 128, 64, True, True, 2, 8),
    ("normal_rp", 128, 256, 64, False, True, 2, 8),
]
VQA_DIMS = dict(embed_size=16, qst_vocab_size=40, ans_vocab_size=12, word_embed_size=10,
                num_layers=1, hidden_size=16)


def gen(seed):
    return torch.Generator().manual_seed(seed)


def fill(module, seed):
    sd = module.state_dict()
    synth_fill_(sd, seed)
    module.load_state_dict(sd)


def npz(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def grads_by_name(module):
    return {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in module.named_parameters()}


def mixed_case(name, C, stride, B, H):
    m = MixedOp(C, stride)
    m.train()
    fill(m, 100 + C + stride)
    x = torch.randn(B, C, H, H, generator=gen(1)).requires_grad_(True)
    w = torch.softmax(torch.randn(8, generator=gen(2)), 0).requires_grad_(True)
    y = m(x, w)
    G = torch.randn(y.shape, generator=gen(3))
    (y * G).sum().backward()
    out = {"x": x, "w": w, "G": G, "y": y, "dx": x.grad, "dw": w.grad}
    out.update({"grad." + k: v for k, v in grads_by_name(m).items()})
    out.update({"buf." + k: v for k, v in m.state_dict().items() if "running" in k or "tracked" in k})
    return npz(out)


def cell_case(name, cpp, cp, C, red, red_prev, B, H):
    cell = Cell(4, 4, cpp, cp, C, red, red_prev)
    cell.train()
    fill(cell, 200 + C + red + 2 * red_prev)
    h0 = 2 * H if red_prev else H
    s0 = torch.randn(B, cpp, h0, h0, generator=gen(4)).requires_grad_(True)
    s1 = torch.randn(B, cp, H, H, generator=gen(5)).requires_grad_(True)
    w = torch.softmax(torch.randn(14, 8, generator=gen(6)), -1).requires_grad_(True)
    w2 = torch.rand(14, generator=gen(7)).requires_grad_(True)
    y = cell(s0, s1, w, w2)
    G = torch.randn(y.shape, generator=gen(8))
    (y * G).sum().backward()
    out = {"s0": s0, "s1": s1, "w": w, "w2": w2, "G": G, "y": y, "ds0": s0.grad, "ds1": s1.grad,
           "dw": w.grad, "dw2": w2.grad}
    out.update({"grad." + k: v for k, v in grads_by_name(cell).items()})
    out.update({"buf." + k: v for k, v in cell.state_dict().items() if "running" in k or "tracked" in k})
    return npz(out)


def network_case(H):
    net = Network(16, 10, 4)
    net.train()
    fill(net, 300)
    arch = net.arch_parameters()
    for i, a in enumerate(arch):
        a.data.copy_(0.5 * torch.randn(a.shape, generator=gen(20 + i)))
    x = torch.randn(2, 3, H, H, generator=gen(9)).requires_grad_(True)
    y = net(x)
    G = torch.randn(y.shape, generator=gen(10))
    (y * G).sum().backward()
    out = {"x": x, "G": G, "y": y, "dx": x.grad}
    for i, a in enumerate(arch):
        out[f"arch{i}"] = a.data
        out[f"darch{i}"] = a.grad
    g = grads_by_name(net)
    out["grad_keys"] = np.array(list(g.keys()))
    out["grad_l2"] = np.array([v.norm().item() for v in g.values()], dtype=np.float64)
    out["grad_sum"] = np.array([v.double().sum().item() for v in g.values()], dtype=np.float64)
    for k in list(g.keys())[:6] + list(g.keys())[200:204] + list(g.keys())[-6:]:
        out["grad." + k] = g[k]
    sd = net.state_dict()
    bk = [k for k in sd if "running" in k]
    out["buf_keys"] = np.array(bk)
    out["buf_sum"] = np.array([sd[k].double().sum().item() for k in bk], dtype=np.float64)
    return npz(out)


def make_vqa(seed=400):
    m = VqaModel(img_encoder_type="darts", **VQA_DIMS)
    m.train()
    fill(m, seed)
    for i, a in enumerate(m.arch_parameters()):
        a.data.copy_(0.5 * torch.randn(a.shape, generator=gen(30 + i)))
    return m


def batch(seed, B=2, H=32):
    g = gen(seed)
    img = torch.randn(B, 3, H, H, generator=g)
    qst = torch.randint(0, VQA_DIMS["qst_vocab_size"], (B, 30), generator=g)
    qst[:, 0] = 2
    lbl = torch.randint(0, VQA_DIMS["ans_vocab_size"], (B,), generator=g)
    return img, qst, lbl


def vqa_case():
    m = make_vqa()
    m.dropout.p = 0.0
    img, qst, lbl = batch(11)
    ans, qout = m(img, qst)
    loss = m._loss(img, qst, lbl)   # second forward: BN counters advance twice in total
    loss.backward()
    out = {"img": img, "qst": qst, "lbl": lbl, "ans": ans, "qout": qout, "loss": loss}
    for i, a in enumerate(m.arch_parameters()):
        out[f"arch{i}"] = a.data
        out[f"darch{i}"] = a.grad
    g = grads_by_name(m)
    out["grad_keys"] = np.array(list(g.keys()))
    out["grad_l2"] = np.array([v.norm().item() for v in g.values()], dtype=np.float64)
    for k in list(g.keys())[-15:]:
        out["grad." + k] = g[k]
    # dropout live (CPU generator): pins the oracle's RNG consumption order
    m2 = make_vqa()
    torch.manual_seed(77)
    out["loss_dropout"] = m2._loss(img, qst, lbl)
    return npz(out)


def _ulp_noise_(module, seed, frac=0.3):
    """Move a random 30% of every weight up by one ulp (what fp32 atomic-order noise does on a GPU)."""
    g = gen(seed)
    with torch.no_grad():
        for p in module.parameters():
            up = torch.nextafter(p, torch.full_like(p, float("inf")))
            p.copy_(torch.where(torch.rand(p.shape, generator=g) < frac, up, p))


def _architect_run(unrolled, seeds, ulp_seed=None):
    m = make_vqa()
    m.dropout.p = 0.0
    if ulp_seed is not None:
        _ulp_noise_(m, ulp_seed)
    # model.new() (vqa_model.py:342-349) builds a fresh VqaModel whose Dropout(0.5) is live and whose
    # masks depend on how much RNG the throw-away init consumed; the harness switches it off on the
    # unrolled copy as well — no reference code is modified.
    orig_new = m.new

    def new_without_dropout():
        fresh = orig_new()
        fresh.dropout.p = 0.0
        return fresh
    m.new = new_without_dropout
    args = Namespace(arch_learn_rate=6e-4, arch_wt_decay=1e-3, qst_only=False)
    arch = Architect(m, args)
    cap = {}
    orig = arch._hessian_vector_product

    def spy(vector, *a, **k):
        cap["vnorm"] = torch.cat([v.reshape(-1) for v in vector]).norm()
        res = orig(vector, *a, **k)
        cap["hvp"] = [r.clone() for r in res]
        return res
    arch._hessian_vector_product = spy
    tr, va = batch(seeds[0]), batch(seeds[1])
    torch.manual_seed(5)     # model.new() draws from the global RNG for its throw-away init
    arch.step(*tr, *va, 1e-3, None, unrolled=unrolled)
    return m, cap


def architect_case(unrolled):
    """Architect.step on a WELL-CONDITIONED point.  The finite-difference HVP evaluates gradients at
    w +- R v; if an activation there sits on a ReLU / max-pool tie, a one-ulp change of the weights flips the
    gradient by ~0.3% in the reference itself (observed for batch seeds 12/13).  GPU weight gradients carry
    one-ulp ordering noise, so the fixture is only accepted when the reference is stable under that noise."""
    seeds = (12, 13)
    for attempt in range(20):
        m, cap = _architect_run(unrolled, seeds)
        stable = True
        for ulp_seed in (101, 102, 103, 104):
            m2, _ = _architect_run(unrolled, seeds, ulp_seed)
            for a, b in zip(m.arch_parameters(), m2.arch_parameters()):
                rel = ((a.grad - b.grad).abs().max() / a.grad.abs().max()).item()
                stable = stable and rel < 2e-5
        if stable:
            break
        seeds = (seeds[0] + 10, seeds[1] + 10)
    else:
        raise RuntimeError("no well-conditioned test point found")
    out = {"seed_train": seeds[0], "seed_valid": seeds[1]}
    for i, a in enumerate(m.arch_parameters()):
        out[f"arch_after{i}"] = a.data
        out[f"darch{i}"] = a.grad
    if unrolled:
        out["vnorm"] = cap["vnorm"]
        for i, h in enumerate(cap["hvp"]):
            out[f"hvp{i}"] = h
    sd = m.state_dict()
    bk = [k for k in sd if k.endswith("num_batches_tracked")]
    out["nbt0"] = sd[bk[0]]
    rm = [k for k in sd if k.endswith("running_mean")]
    out["rm_keys"] = np.array(rm[:8])
    for k in rm[:8]:
        out["buf." + k] = sd[k]
    return npz(out)


def wstep_case():
    """darts_vqa/experiment.py:187-200 reproduced by hand (the Experiment class needs the dataset)."""
    m = make_vqa()
    m.dropout.p = 0.0
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()
    img, qst, lbl = batch(14)
    losses = []
    for _ in range(2):
        opt.zero_grad()
        ans, qout = m(img, qst)
        loss = crit(ans, lbl) + crit(qout[:, :-1].flatten(end_dim=1), qst[:, 1:].flatten())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5)
        opt.step()
        losses.append(loss.item())
    sd = dict(m.named_parameters())
    keys = list(sd.keys())
    out = {"losses": np.array(losses), "keys": np.array(keys),
           "param_l2": np.array([sd[k].norm().item() for k in keys], dtype=np.float64)}
    for k in keys[:4] + keys[-4:]:
        out["param." + k] = sd[k]
    return npz(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    torch.set_num_threads(1)
    x = torch.arange(2 * 16 * 2 * 3, dtype=torch.float32).reshape(2, 16, 2, 3)
    np.savez_compressed(os.path.join(a.out, "shuffle.npz"),
                        **{f"c{c}": channel_shuffle(torch.arange(c * 2.).reshape(1, c, 1, 2), 4).numpy()
                           for c in (16, 32, 64, 8)}, x=x.numpy(), y=channel_shuffle(x, 4).numpy())
    for case in MIXED_CASES:
        np.savez_compressed(os.path.join(a.out, f"mixed_{case[0]}.npz"), **mixed_case(*case))
    for case in CELL_CASES:
        np.savez_compressed(os.path.join(a.out, f"cell_{case[0]}.npz"), **cell_case(*case))
    np.savez_compressed(os.path.join(a.out, "network32.npz"), **network_case(32))
    np.savez_compressed(os.path.join(a.out, "vqa.npz"), **vqa_case())
    np.savez_compressed(os.path.join(a.out, "architect_first.npz"), **architect_case(False))
    np.savez_compressed(os.path.join(a.out, "architect_unrolled.npz"), **architect_case(True))
    np.savez_compressed(os.path.join(a.out, "wstep.npz"), **wstep_case())
    print("golden vectors written to", a.out)


if __name__ == "__main__":
    main()

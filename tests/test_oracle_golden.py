"""Pin the CPU oracle against vectors produced by the unmodified reference (tests/golden/make_golden.py).

CPU only.  Tolerance 1e-5 relative: both sides run the same ATen CPU kernels, differences come only
from op ordering (e.g. explicit LSTM loop vs nn.LSTM).
"""
import numpy as np
import pytest
import torch

from helpers import assert_close, load_golden
from oracle import pcdarts_oracle as O

TOL = 2e-5
torch.set_num_threads(max(1, min(4, torch.get_num_threads())))


def gen(seed):
    return torch.Generator().manual_seed(seed)


def test_shuffle_index_exact():
    g = load_golden("shuffle")
    for c in (16, 32, 64, 8):
        src = np.arange(c * 2.).reshape(1, c, 1, 2).astype(np.float32)
        assert np.array_equal(O.channel_shuffle_np(src), g[f"c{c}"].numpy())
    assert torch.equal(O.channel_shuffle(g["x"]), g["y"])
    assert list(O.shuffle_perm(16)) == [0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15]


def test_adaptive_windows():
    x = torch.randn(1, 1, 16, 16, generator=gen(0))
    ref = torch.nn.functional.adaptive_avg_pool2d(x, 7)
    win = O.adaptive_windows(16, 7)
    mine = torch.stack([torch.stack([x[0, 0, a:b, c:d].mean() for (c, d) in win]) for (a, b) in win])
    assert_close(mine, ref[0, 0], 1e-6)


MIXED = [("t1", 16, 1), ("t2", 32, 2), ("t3", 32, 1), ("t4", 64, 2), ("t5", 64, 1), ("odd", 16, 1),
         ("rect2", 16, 2)]


@pytest.mark.parametrize("name,C,stride", MIXED)
def test_mixed_op(name, C, stride):
    g = load_golden("mixed_" + name)
    sd = O.alloc_state(O.mixed_op_spec(C, stride), seed=100 + C + stride)
    par, buf = O.split_state(sd)
    for v in par.values():
        v.requires_grad_(True)
    x = g["x"].clone().requires_grad_(True)
    w = g["w"].clone().requires_grad_(True)
    y = O.mixed_op(par, O.BNState(buf), "_ops.", x, w, stride)
    assert_close(y, g["y"], TOL, "y")
    # bit-exact pass-through channels: out[:, 4j+q] (q=1..3) is a copy (or 2x2 max) of x[:, q*c+j]
    c = C // 4
    if stride == 1:
        for q in range(1, 4):
            assert torch.equal(y[:, q::4], x[:, q * c:(q + 1) * c])
    (y * g["G"]).sum().backward()
    assert_close(x.grad, g["dx"], TOL, "dx")
    assert_close(w.grad, g["dw"], TOL, "dw")
    for k, v in par.items():
        assert_close(v.grad, g["grad." + k], TOL, k)
    for k, v in buf.items():
        assert_close(v.float(), g["buf." + k].float(), TOL, k)


CELLS = [("normal", 48, 48, 16, False, False), ("reduce", 48, 64, 32, True, False),
         ("reduce_rp", 64, 128, 64, True, True), ("normal_rp", 128, 256, 64, False, True)]


@pytest.mark.parametrize("name,cpp,cp,C,red,rp", CELLS)
def test_cell(name, cpp, cp, C, red, rp):
    g = load_golden("cell_" + name)
    sd = O.alloc_state(O.cell_spec(cpp, cp, C, red, rp), seed=200 + C + red + 2 * rp)
    par, buf = O.split_state(sd)
    for v in par.values():
        v.requires_grad_(True)
    ins = [g[k].clone().requires_grad_(True) for k in ("s0", "s1", "w", "w2")]
    y = O.cell_forward(par, O.BNState(buf), "", *ins, red, rp)
    assert_close(y, g["y"], TOL, "y")
    (y * g["G"]).sum().backward()
    for t, k in zip(ins, ("ds0", "ds1", "dw", "dw2")):
        assert_close(t.grad, g[k], TOL, k)
    for k, v in par.items():
        assert_close(v.grad, g["grad." + k], TOL, k)
    for k, v in buf.items():
        assert_close(v.float(), g["buf." + k].float(), TOL, k)


def test_network():
    g = load_golden("network32")
    sd = O.alloc_state(O.network_spec(), seed=300)
    par, buf = O.split_state(sd)
    for v in par.values():
        v.requires_grad_(True)
    arch = [g[f"arch{i}"].clone().requires_grad_(True) for i in range(4)]
    x = g["x"].clone().requires_grad_(True)
    y = O.network_forward(par, O.BNState(buf), arch, x)
    assert_close(y, g["y"], TOL, "y")
    (y * g["G"]).sum().backward()
    assert_close(x.grad, g["dx"], TOL, "dx")
    for i in range(4):
        assert_close(arch[i].grad, g[f"darch{i}"], TOL, f"darch{i}")
    keys = [str(k) for k in g["grad_keys"]]
    assert keys == list(par.keys()), "parameter registration order differs from the reference"
    l2 = torch.tensor([par[k].grad.norm().item() for k in keys], dtype=torch.float64)
    assert_close(l2, torch.from_numpy(g["grad_l2"]) if isinstance(g["grad_l2"], np.ndarray) else g["grad_l2"], TOL)
    for k in g:
        if k.startswith("grad."):
            assert_close(par[k[5:]].grad, g[k], TOL, k)
    bsum = torch.tensor([buf[str(k)].double().sum().item() for k in g["buf_keys"]], dtype=torch.float64)
    assert_close(bsum, g["buf_sum"], TOL, "running stats")


VQA_DIMS = dict(embed_size=16, qst_vocab_size=40, ans_vocab_size=12, word_embed_size=10,
                num_layers=1, hidden_size=16)


def vqa_state(g=None, seed=400, arch_seed=30):
    sd = O.alloc_state(O.vqa_spec(**VQA_DIMS), seed=seed)
    par, buf = O.split_state(sd)
    for v in par.values():
        v.requires_grad_(True)
    arch = [(0.5 * torch.randn(s, generator=gen(arch_seed + i))).requires_grad_(True)
            for i, s in enumerate(((14, 8), (14, 8), (14,), (14,)))]
    return par, buf, arch


def batch(seed, B=2, H=32):
    g = gen(seed)
    img = torch.randn(B, 3, H, H, generator=g)
    qst = torch.randint(0, VQA_DIMS["qst_vocab_size"], (B, 30), generator=g)
    qst[:, 0] = 2
    lbl = torch.randint(0, VQA_DIMS["ans_vocab_size"], (B,), generator=g)
    return img, qst, lbl


def test_vqa_model():
    g = load_golden("vqa")
    par, buf, arch = vqa_state()
    for i in range(4):
        assert torch.equal(arch[i].detach(), g[f"arch{i}"])
    img, qst, lbl = batch(11)
    assert torch.equal(img, g["img"]) and torch.equal(qst, g["qst"]) and torch.equal(lbl, g["lbl"])
    bns = O.BNState(buf)
    ans, qout = O.vqa_forward(par, bns, arch, img, qst, dropout_p=0.0)
    assert_close(ans, g["ans"], TOL, "ans")
    assert_close(qout, g["qout"], TOL, "qout")
    loss = O.vqa_loss(par, bns, arch, img, qst, lbl, dropout_p=0.0)
    assert_close(loss, g["loss"], TOL, "loss")
    gs = torch.autograd.grad(loss, arch + list(par.values()))
    for i in range(4):
        assert_close(gs[i], g[f"darch{i}"], TOL, f"darch{i}")
    keys = [str(k) for k in g["grad_keys"]]
    assert keys == list(par.keys())
    l2 = torch.tensor([t.norm().item() for t in gs[4:]], dtype=torch.float64)
    assert_close(l2, g["grad_l2"], TOL, "grad norms")
    gmap = dict(zip(keys, gs[4:]))
    for k in g:
        if k.startswith("grad."):
            assert_close(gmap[k[5:]], g[k], TOL, k)
    # dropout live: same CPU generator consumption as the reference (vqa_model.py:296,310,314)
    par, buf, arch = vqa_state()
    torch.manual_seed(77)
    loss_d = O.vqa_loss(par, O.BNState(buf), arch, img, qst, lbl)
    assert_close(loss_d, g["loss_dropout"], TOL, "loss with dropout")


@pytest.mark.parametrize("unrolled", [False, True])
def test_architect_step(unrolled):
    g = load_golden("architect_unrolled" if unrolled else "architect_first")
    par, buf, arch = vqa_state()
    dbg = {}
    grads = O.architect_step(par, O.BNState(buf), arch, {}, batch(int(g["seed_train"])), batch(int(g["seed_valid"])),
                             1e-3, list(par.keys()),
                             unrolled=unrolled, dropout_p=0.0, **({"debug": dbg} if unrolled else {}))
    for i in range(4):
        assert_close(grads[i], g[f"darch{i}"], 1e-4, f"darch{i}")
        assert_close(arch[i].detach(), g[f"arch_after{i}"], 1e-5, f"arch_after{i}")
    if unrolled:
        assert_close(0.01 / dbg["R"], g["vnorm"], TOL, "|vector|")
        # raw finite-difference HVP: catastrophic cancellation, only loosely reproducible
        scale = max(float(dbg["g_pos"][i].abs().max()) for i in range(4)) / (2 * float(dbg["R"]))
        for i in range(4):
            hv = (dbg["g_pos"][i] - dbg["g_neg"][i]) / (2 * dbg["R"])
            assert (hv - g[f"hvp{i}"]).abs().max().item() <= 1e-4 * 2 * scale
        assert int(buf["img_encoder.darts.stem.1.num_batches_tracked"]) == int(g["nbt0"]) == 3
    for k in g["rm_keys"]:
        assert_close(buf[str(k)], g["buf." + str(k)], TOL, str(k))


def test_w_step():
    g = load_golden("wstep")
    par, buf, arch = vqa_state()
    keys = list(par.keys())
    st = {}
    b = batch(14)
    losses = []
    for _ in range(2):
        losses.append(float(O.w_step(par, O.BNState(buf), arch, b, st, keys, dropout_p=0.0)))
    assert_close(torch.tensor(losses), torch.as_tensor(g["losses"]), TOL, "losses")
    l2 = torch.tensor([par[k].norm().item() for k in keys], dtype=torch.float64)
    assert_close(l2, g["param_l2"], TOL, "param norms after 2 steps")
    for k in g:
        if k.startswith("param."):
            assert_close(par[k[6:]].detach(), g[k], 1e-4, k)

"""Golden vectors for the 3-stage LCT alpha-step from the UNMODIFIED reference (basic_vqa/), build container only.

    python tests/golden/make_golden_lct.py        # needs /root/reference; writes tests/golden/architect_lct.npz

basic_vqa shares top-level module names with darts_vqa (config, pcdarts, ...), hence a separate script/process.
Two things of the environment are neutralised, not the algorithm: torchvision's VGG19 is built without the ImageNet
download (models.py:23 hard-codes pretrained=True; there is no network here), and Dropout is constructed with p = 0
(the unrolled models are created inside ArchitectLct by model.new(); SURVEY.md App. C: parity needs matching masks).
Weights come from oracle.synth_fill_ (seeded by state_dict key), inputs from seeded generators.
Every _calc_grad result is recorded in call order so the test can apply the cancellation-aware HVP bounds.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("LCT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "basic_vqa"))

import torchvision.models as tvm  # noqa: E402
_vgg19 = tvm.vgg19
tvm.vgg19 = lambda pretrained=False, **kw: _vgg19(weights=None)
_drop_init = torch.nn.Dropout.__init__
torch.nn.Dropout.__init__ = lambda self, p=0.5, inplace=False: _drop_init(self, 0.0, inplace)

import config  # noqa: E402  (reference basic_vqa/config.py)
config.DEVICE = torch.device("cpu")
config.ARCH_TYPE = "darts"
from models_lct import VqaModel as EfModel  # noqa: E402
from models import VqaModel as WModel  # noqa: E402
from pcdarts.architect_lct import ArchitectLct  # noqa: E402
from oracle.pcdarts_oracle import synth_fill_  # noqa: E402

DIMS = dict(embed_size=16, qst_vocab_size=40, ans_vocab_size=12, word_embed_size=8, num_layers=1, hidden_size=16)


def batch(seed, B=2, H=32):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, H, H, generator=g)
    qst = torch.randint(0, DIMS["qst_vocab_size"], (B, 30), generator=g)
    qst[:, 0] = 2
    lbl = torch.randint(0, DIMS["ans_vocab_size"], (B,), generator=g)
    return img, qst, lbl


def main():
    torch.set_num_threads(1)
    torch.manual_seed(10)
    ef = EfModel(**DIMS)
    w = WModel(**DIMS)
    for m, seed in ((ef, 500), (w, 600)):
        sd = m.state_dict()
        synth_fill_(sd, seed)
        m.load_state_dict(sd)
        m.train()
    g = torch.Generator().manual_seed(77)
    for a in ef.arch_parameters():
        a.data.copy_(1e-1 * torch.randn(a.shape, generator=g))
    ef_opt = torch.optim.Adam(ef.parameters(), lr=1e-3)
    w_opt = torch.optim.Adam(w.parameters(), lr=1e-3)
    arch = ArchitectLct(ef, w, ef_opt, w_opt)
    calls = []
    orig = arch._calc_grad

    def rec(loss, param_fn, exp_zero_grad=0):
        out = orig(loss, param_fn, exp_zero_grad)
        calls.append((float(loss.detach()), [t.detach().clone() for t in out]))
        return out
    arch._calc_grad = rec
    gens = []
    gen_orig = EfModel.generate

    def gen_rec(self, img):
        q, ans = gen_orig(self, img)
        gens.append((q.detach().clone(), ans.detach().clone()))
        return q, ans
    EfModel.generate = gen_rec
    train, valid = batch(21), batch(22)
    arch.step(*train, *valid, 1e-3, 1e-3)
    EfModel.generate = gen_orig
    names = ["unroll_ef", "unroll_w", "grad_wprime", "kappa_p", "kappa_n", "gamma_p", "gamma_n"]
    assert len(calls) == len(names), len(calls)
    out = {}
    for n, (loss, grads) in zip(names, calls):
        out[f"loss.{n}"] = np.float64(loss)
        flat = torch.cat([t.reshape(-1) for t in grads])
        out[f"norm.{n}"] = np.float64(flat.norm().item())
        if n.startswith("gamma"):
            for i, t in enumerate(grads):
                out[f"{n}{i}"] = t.numpy()
    for j, (q, ans) in enumerate(gens):          # 3 generate() calls: unroll W, kappa +, kappa -
        out[f"pseudo_qst{j}"] = q.numpy()
        out[f"pseudo_ans{j}"] = ans.numpy()
    for i, a in enumerate(ef.arch_parameters()):
        out[f"darch{i}"] = a.grad.numpy()
        out[f"arch_after{i}"] = a.data.numpy()
    np.savez_compressed(os.path.join(HERE, "architect_lct.npz"), **out)
    print({k: (v if np.ndim(v) == 0 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()

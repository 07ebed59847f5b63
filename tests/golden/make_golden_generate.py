"""Golden vectors for QstEncoder.generate (greedy decode) from the UNMODIFIED reference (darts_vqa/vqa_model.py:103-136),
build container only.

    python tests/golden/make_golden_generate.py        # needs /root/reference; writes tests/golden/generate.npz

Two seeded encoders (hidden 32 / 64), the reference's own default initialisation under torch.manual_seed, image embeddings
from a seeded generator.  Stored: every parameter, the embeddings, the generated words, and per step the margin between the
best and the second-best logit (so that a consumer can tell a real disagreement from a rounding-level tie).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LCT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "darts_vqa"))

import config  # noqa: E402  (reference darts_vqa/config.py)
config.DEVICE = torch.device("cpu")
from vqa_model import QstEncoder  # noqa: E402

CASES = [dict(name="a", V=300, E=12, H=32, B=5, T=30, seed=11), dict(name="b", V=700, E=20, H=64, B=9, T=12, seed=12)]


def main():
    torch.set_num_threads(1)
    out = {}
    for c in CASES:
        torch.manual_seed(c["seed"])
        enc = QstEncoder(c["V"], c["E"], c["H"], 1, c["H"], max_length=c["T"])
        with torch.no_grad():
            enc.word2vec.weight.mul_(1.5)          # spread the logits a little: fewer near-ties
            enc.fc1.bias.copy_(0.05 * torch.randn(c["V"]))
        img = 0.5 * torch.randn(c["B"], c["H"], generator=torch.Generator().manual_seed(c["seed"] + 100))
        with torch.no_grad():
            words = enc.generate(img)
            # margins, teacher-forced with the reference's own words
            state = (img.view(1, -1, c["H"]), img.view(1, -1, c["H"]))
            cur = torch.tanh(enc.word2vec(torch.full((c["B"], 1), 2, dtype=torch.long))).transpose(0, 1)
            margins = []
            for t in range(c["T"]):
                o, state = enc.lstm(cur, state)
                logits = enc.fc1(torch.tanh(o.transpose(0, 1)))[:, 0]
                top2 = logits.topk(2, dim=1).values
                margins.append((top2[:, 0] - top2[:, 1]).numpy())
                assert torch.equal(logits.argmax(1), words[:, t])
                cur = enc.word2vec(words[:, t:t + 1]).transpose(0, 1)
        n = c["name"]
        for k, v in enc.state_dict().items():
            out[f"{n}.{k}"] = v.numpy()
        out[f"{n}.img"] = img.numpy()
        out[f"{n}.words"] = words.numpy()
        out[f"{n}.margin"] = np.stack(margins, 1)
        out[f"{n}.dims"] = np.array([c["V"], c["E"], c["H"], c["B"], c["T"]])
        print(n, "min margin", float(np.min(out[f"{n}.margin"])), "words[0,:8]", words[0, :8].tolist())
    np.savez_compressed(os.path.join(HERE, "generate.npz"), **out)


if __name__ == "__main__":
    main()

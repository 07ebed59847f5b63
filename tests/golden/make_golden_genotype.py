"""Golden genotypes from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden_genotype.py      # needs /root/reference; writes tests/golden/genotype.npz

Network.genotype() (darts_vqa/pcdarts/model_search.py:218-263) of the reference for several seeded alpha / beta settings,
including rows where 'none' is the strongest op (it must be skipped) and exact ties between edges (the lower index wins).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LCT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "darts_vqa"))

import config  # noqa: E402  (reference)
config.DEVICE = "cpu"
from pcdarts.model_search import Network  # noqa: E402


def main():
    out = {}
    net = Network(16, 10, 4)
    for case in range(6):
        g = torch.Generator().manual_seed(100 + case)
        scale = (1e-3, 1.0, 3.0, 1.0, 1.0, 0.5)[case]
        an, ar = scale * torch.randn(14, 8, generator=g), scale * torch.randn(14, 8, generator=g)
        bn, br = scale * torch.randn(14, generator=g), scale * torch.randn(14, generator=g)
        if case == 3:          # 'none' strongest everywhere
            an[:, 0] += 5.0
            ar[:, 0] += 5.0
        if case == 4:          # exact ties between the edges of a node and between ops of an edge
            an[:] = an[0]
            bn[:] = 0.0
            ar[:, 1:] = ar[:, 1:2]
        with torch.no_grad():
            for t, v in zip(net.arch_parameters(), (an, ar, bn, br)):
                t.copy_(v)
        geno = net.genotype()
        out[f"c{case}_alphas_normal"], out[f"c{case}_alphas_reduce"] = an.numpy(), ar.numpy()
        out[f"c{case}_betas_normal"], out[f"c{case}_betas_reduce"] = bn.numpy(), br.numpy()
        out[f"c{case}_normal"] = np.array([f"{n}:{j}" for n, j in geno.normal])
        out[f"c{case}_reduce"] = np.array([f"{n}:{j}" for n, j in geno.reduce])
        out[f"c{case}_concat"] = np.array(list(geno.normal_concat) + list(geno.reduce_concat))
    np.savez_compressed(os.path.join(HERE, "genotype.npz"), **out)
    print("wrote genotype.npz:", [str(x) for x in out["c1_normal"]])


if __name__ == "__main__":
    main()

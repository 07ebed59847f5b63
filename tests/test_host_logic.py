"""Host-side logic that needs neither kernels nor the reference tree: genotype derivation against the reference's golden."""
import os

import torch


def test_genotype_matches_reference_golden():
    """Network.genotype() against the unmodified reference (tests/golden/make_golden_genotype.py): 'none' never chosen, the two
    strongest edges per node, ties resolved like the reference's sorted() / argmax."""
    import numpy as np
    import config
    config.DEVICE = torch.device("cpu")
    from pcdarts.model_search import Network
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "genotype.npz"))
    net = Network(16, 10, 4)
    for case in range(6):
        with torch.no_grad():
            for t, k in zip(net.arch_parameters(), ("alphas_normal", "alphas_reduce", "betas_normal", "betas_reduce")):
                t.copy_(torch.from_numpy(z[f"c{case}_{k}"]))
        g = net.genotype()
        assert [f"{n}:{j}" for n, j in g.normal] == [str(s) for s in z[f"c{case}_normal"]], case
        assert [f"{n}:{j}" for n, j in g.reduce] == [str(s) for s in z[f"c{case}_reduce"]], case
        assert list(g.normal_concat) + list(g.reduce_concat) == [int(v) for v in z[f"c{case}_concat"]]

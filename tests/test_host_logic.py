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


def test_auto_split_fills_whole_rounds():
    """pcd_ops._auto_split: K splits for the persistent GEMM's round-robin schedule (one CTA per SM, 148 SMs)."""
    import pcd_ops
    f = pcd_ops._auto_split
    assert f(1920, 17858, 512) == 1                    # 1050 tiles: more than enough work items
    assert f(17858, 512, 1920) == 1
    s = f(1920, 512, 17860)                            # vocabulary dX: 30 tiles
    assert 2 <= s <= 32 and (30 * s) % 148 > 110 or (30 * s) <= 148      # the last round is nearly full
    assert f(1920, 300, 2048) == 4                     # 30 tiles x 4 = 120 items: one round (5 would be two)
    assert f(64, 512, 12544) == 32                     # 2 tiles: as many splits as the cap allows
    for m, n, k in ((64, 64, 64), (128, 256, 100), (5, 1000, 129), (300, 300, 300)):
        s = f(m, n, k)
        bk = 16 if n >= 256 else 32
        assert 1 <= s <= max(1, min(32, -(-k // bk) // 8))

"""Stand-alone candidate operations — the op kernels' SOURCES compiled with -DPCD_EMU, against the same
layers in stock torch at float64, plus the pin of that stock path to the reference's own modules."""
import os
import sys

import pytest
import torch

import parity_cases as P


@pytest.fixture(scope="module", autouse=True)
def emulation():
    import pcd_build
    import pcd_native
    pcd_native.enable_emulation(pcd_build.build_emu())
    yield
    pcd_native._emu_lib = None


@pytest.mark.parametrize("name,C,stride,affine,B,H", P.OPS_CASES)
def test_op_vs_stock_emulated(name, C, stride, affine, B, H):
    P.op_vs_stock(name, C, stride, affine, B, H, "cpu")


@pytest.mark.parametrize("name,stride", [("max_pool_3x3", 1), ("max_pool_3x3", 2), ("sep_conv_3x3", 1), ("dil_conv_5x5", 2)])
def test_op_exact_ties_emulated(name, stride):
    """Pool windows full of equal values (first maximum in scan order takes the gradient, like ATen), ReLU inputs exactly 0."""
    P.op_vs_stock(name, 8, stride, True, 2, 12, "cpu", quantized=True)


@pytest.mark.parametrize("cpb", ["2", "3"])
def test_pointwise_backward_walks_several_chunks(cpb, monkeypatch):
    """pw_bwd blocks that walk several 128-pixel chunks (what large batches use): 20 x 20 planes = 3 full chunks + a ragged one."""
    monkeypatch.setenv("PCD_PW_CPB", cpb)
    P.op_vs_stock("dil_conv_3x3", 8, 1, True, 2, 20, "cpu")
    P.op_vs_stock("sep_conv_3x3", 16, 2, True, 1, 32, "cpu")


def test_unsupported_ops_raise():
    from pcdarts.operations import OPS
    op = OPS["conv_7x1_1x7"](8, 1, True)
    with pytest.raises(RuntimeError, match="PCD_ERR_UNSUPPORTED"):
        op(torch.randn(1, 8, 8, 8))
    with pytest.raises(RuntimeError, match="PCD_ERR_UNSUPPORTED"):
        OPS["sep_conv_3x3"](6, 1, True)(torch.randn(1, 6, 8, 8))           # channels not a multiple of 4
    with pytest.raises(RuntimeError, match="PCD_ERR_UNSUPPORTED"):
        OPS["max_pool_3x3"](4, 1, True)(torch.randn(1, 4, 80, 80))         # planes above 64 x 64
    with pytest.raises(NotImplementedError):
        OPS["dil_conv_3x3"](8, 1, True).eval()(torch.randn(1, 8, 8, 8))    # eval-mode BatchNorm is out of scope


REF = "/root/reference/darts_vqa"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the development container")
@pytest.mark.parametrize("name,C,stride", [("sep_conv_3x3", 8, 2), ("sep_conv_5x5", 8, 1), ("sep_conv_7x7", 4, 1), ("dil_conv_3x3", 8, 1),
                                           ("dil_conv_5x5", 8, 2), ("max_pool_3x3", 4, 2), ("avg_pool_3x3", 4, 1),
                                           ("skip_connect", 8, 2), ("skip_connect", 8, 1), ("none", 4, 2)])
def test_stock_forward_is_the_reference_module(name, C, stride):
    """`stock_forward` (the yardstick of the op parity tests) against the reference's OPS[name] with the same state_dict:
    bit-equal outputs and gradients (both are the same ATen calls)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_operations", os.path.join(REF, "pcdarts", "operations.py"))
    ref_ops = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ops)
    from pcdarts.operations import OPS
    torch.manual_seed(3)
    mine, ref = OPS[name](C, stride, True).train(), ref_ops.OPS[name](C, stride, True).train()
    ref.load_state_dict(mine.state_dict())
    x1 = torch.randn(2, C, 12, 12, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    y1 = mine.stock_forward(x1) if hasattr(mine, "stock_forward") else mine(x1)
    y2 = ref(x2)
    assert torch.equal(y1, y2)
    if y1.requires_grad:
        gy = torch.randn_like(y1)
        g1 = torch.autograd.grad(y1, [x1] + list(mine.parameters()), gy)
        g2 = torch.autograd.grad(y2, [x2] + list(ref.parameters()), gy)
        for a, b in zip(g1, g2):
            assert torch.equal(a, b)


def test_derived_network_vs_stock_emulated():
    """NetworkDerived (pcdarts/model.py) end to end: stem, 4 derived cells (2 reductions), pooling — forward, all weight grads."""
    P.derived_vs_stock("cpu")


def test_derive_from_search_network():
    import config
    config.DEVICE = torch.device("cpu")
    from pcdarts.model import derive
    from pcdarts.model_search import Network
    torch.manual_seed(0)
    search = Network(16, 10, 4)
    net = derive(search)
    g = search.genotype()
    assert net.genotype() == g and len(net.cells) == 4 and net.output_ch == search.output_ch
    assert [type(c.preprocess0).__name__ for c in net.cells] == [type(c.preprocess0).__name__ for c in search.cells]
    names = [n for n, _ in g.normal]
    assert all(n in ("max_pool_3x3", "avg_pool_3x3", "skip_connect", "sep_conv_3x3", "sep_conv_5x5", "dil_conv_3x3", "dil_conv_5x5") for n in names)

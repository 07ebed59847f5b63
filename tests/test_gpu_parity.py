"""GPU parity: the sm_100a kernels, called through the C ABI, against the reference goldens and the
CPU oracle.  rel 1e-4 on outputs / weight grads / alpha-beta grads, bit-exact channel indexing."""
import pytest
import torch

import parity_cases as P

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name,C,stride", P.MIXED)
def test_mixed_op_golden(name, C, stride):
    P.mixed_case(name, C, stride, DEV)


@pytest.mark.parametrize("name,cpp,cp,C,red,rp", P.CELLS)
def test_cell_golden(name, cpp, cp, C, red, rp):
    P.cell_case(name, cpp, cp, C, red, rp, DEV)


def test_network_golden():
    P.network_case(DEV)


def test_shuffle_bit_exact():
    P.shuffle_case(DEV)


# the five production edge shapes of SURVEY.md §8(a) at a batch the oracle finishes in seconds
@pytest.mark.parametrize("C,stride,B,H", [(16, 1, 4, 64), (32, 2, 4, 64), (32, 1, 4, 32), (64, 2, 4, 32), (64, 1, 8, 16)])
def test_mixed_op_production_shapes_vs_oracle(C, stride, B, H):
    P.mixed_vs_oracle(C, stride, B, H, DEV)


def test_vqa_model_golden():
    P.vqa_case(DEV)


@pytest.mark.parametrize("unrolled", [False, True])
def test_architect_step_golden(unrolled):
    P.architect_case(DEV, unrolled)


def test_w_step_golden():
    P.wstep_case(DEV)


def test_native_library_is_the_one_running():
    import pcd_native
    lib = pcd_native.load_cuda()
    assert lib.pcd_is_cuda_build() == 1
    maps = open("/proc/self/maps").read()
    assert "libpcdarts_sm100.so" in maps

"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for f in os.listdir(os.path.join(ROOT, "include")):
        if f.endswith(".h"):
            src = open(os.path.join(ROOT, "include", f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"^\s*(?:const\s+)?(?:int|long long|size_t|char\s*\*|const char\s*\*)\s*(pcd_\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("pcd_cell_forward", "pcd_cell_backward", "pcd_mixedop_forward", "pcd_mixedop_backward",
                 "pcd_stem_forward", "pcd_stem_backward", "pcd_channel_shuffle", "pcd_adaptive_avgpool_forward",
                 "pcd_preprocess_forward", "pcd_version", "pcd_strerror"):
        assert must in names


def test_cuda_library_exports_every_declared_symbol():
    import pcd_build
    lib_path = pcd_build.build_cuda()          # nvcc cross-compiles for sm_100a without a GPU
    try:
        lib = ctypes.CDLL(lib_path)
    except OSError as e:                       # e.g. libcudart not loadable on an exotic host
        pytest.skip(f"cannot dlopen the CUDA build here: {e}")
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.pcd_version() == 100
    assert lib.pcd_is_cuda_build() == 1
    lib.pcd_strerror.restype = ctypes.c_char_p
    assert lib.pcd_strerror(-2) == b"shape not supported by the compiled kernels"


def test_python_binding_matches_header():
    import pcd_native
    assert sorted(pcd_native.EXPORTS) == declared_functions()


def test_sizes_query_runs_without_gpu():
    import pcd_build
    import pcd_native
    lib = pcd_native._declare(ctypes.CDLL(pcd_build.build_emu()))
    sh = pcd_native.CellShape(64, 48, 48, 16, 64, 64, 0, 0, 4, 1e-5, 0.1)
    sz = pcd_native.CellSizes()
    assert lib.pcd_cell_sizes_of(ctypes.byref(sh), ctypes.byref(sz)) == 0
    # 2 preprocess 1x1 (16x48) + 14 stride-1 edges of c=4: 102c + 6c^2 = 504 floats each
    assert sz.param_floats == 2 * 16 * 48 + 14 * 504
    assert (sz.out_height, sz.out_width) == (64, 64)
    bad = pcd_native.CellShape(64, 48, 48, 20, 64, 64, 0, 0, 4, 1e-5, 0.1)
    assert lib.pcd_cell_sizes_of(ctypes.byref(bad), ctypes.byref(sz)) == -2

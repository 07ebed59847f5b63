"""Build recipes for the native library (in-tree, so the .so travels with the repo snapshot).

    build_cuda()  -> lct-vqa_b200/libpcdarts_sm100.so   nvcc, sm_100a only (cross-compiles without a GPU)
    build_emu()   -> tests/emu/libpcd_emu.so            g++ -DPCD_EMU: CPU emulation of the same kernel
                                                         sources, TEST INFRASTRUCTURE ONLY
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = [os.path.join(CSRC, "pcd_api.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("pcd_common.cuh", "pcd_fwd.cuh", "pcd_bwd.cuh")] + \
          [os.path.join(ROOT, "include", "pcdarts_sm100.h")]
CUDA_LIB = os.path.join(HERE, "libpcdarts_sm100.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libpcd_emu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_cuda(force=False, verbose=False):
    if not force and not _stale(CUDA_LIB, SOURCES + HEADERS):
        return CUDA_LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", CUDA_LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return CUDA_LIB


def build_emu(force=False):
    if not force and not _stale(EMU_LIB, SOURCES + HEADERS):
        return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-DPCD_EMU", "-x", "c++", "-shared", "-fPIC", "-o", EMU_LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ (emulation build) failed:\n" + r.stdout + r.stderr)
    return EMU_LIB

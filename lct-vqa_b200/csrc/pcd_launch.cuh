// pcd_launch.cuh — the one __global__ entry template, its host-side launcher (with the launch counter and
// the optional per-launch CUDA-event profiler), and the CPU-emulation twin used by tests/emu.
#pragma once
#include "../../include/pcdarts_sm100.h"
#include "pcd_common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace pcd {

constexpr int kMaxKernels = 96, kMaxRecords = 1 << 16;
struct LaunchState {
    long long launches;
    const char* names[kMaxKernels];
    int num_kernels, prof_on, prof_n;
    int prof_kid[kMaxRecords];
    char last_err[256];
#if PCD_CUDA
    cudaEvent_t* ev;      // 2 events per record, created on first enable
#endif
};
LaunchState& launch_state();           // defined in pcd_api.cu
// forward passes launch from the Python main thread, backward passes from autograd's per-device worker thread: the one
// counter both bump is atomic; the event profiler (pcd_profile_enable) is a single-threaded measurement mode
inline void count_launch(LaunchState& L) { __atomic_fetch_add(&L.launches, 1LL, __ATOMIC_RELAXED); }
int register_kernel(const char* name);

#define PCD_TRY(x) do { int rc_ = (x); if (rc_ != PCD_OK) return rc_; } while (0)

#if PCD_CUDA
#define PCD_D __device__ __forceinline__

// Programmatic dependent launch (opt-in, PCD_PDL=1): every kernel (1) lets its successor in the stream start launching as
// soon as all of its own blocks are resident and (2) waits, before touching memory, until its predecessor has completed and
// flushed.  Without the launch attribute both instructions are no-ops, so correctness never depends on them.  Measured on
// the captured search step (B200, round 2, profiles/r02_bench_ab_pdl.json): 38.29 ms with, 36.84 ms without — the early
// successor blocks hold registers / shared memory of the last wave's SMs while they spin, which costs more than the
// drain-then-launch gap it hides inside a CUDA graph — so it stays off by default.
template <class Body, class Args>
__global__ void __launch_bounds__(kThreads, Body::kMinBlocks) pcd_kernel(const Args a) {
    extern __shared__ F4 pcd_smem4[];
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Body::run(a, blockIdx.x, blockIdx.y, blockIdx.z, reinterpret_cast<float*>(pcd_smem4));
}

inline bool pdl_enabled() {
    static const bool on = getenv("PCD_PDL") != nullptr && getenv("PCD_PDL")[0] == '1';
    return on;
}

template <class Body, class Args>
static int launch(const Args& a, int gx, int gy, int gz, size_t smem_floats, void* stream) {
    LaunchState& L = launch_state();
    const size_t bytes = smem_floats * sizeof(float);
    if (gx <= 0 || gy <= 0 || gz <= 0) return PCD_OK;
    if (bytes > 227 * 1024 || gy > 65535 || gz > 65535) return PCD_ERR_UNSUPPORTED;
    if (bytes > 48 * 1024) {
        static bool configured = false;   // per (Body,Args) instantiation; idempotent
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(pcd_kernel<Body, Args>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 227 * 1024);
            if (e != cudaSuccess) {
                snprintf(L.last_err, sizeof L.last_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return PCD_ERR_CUDA;
            }
            configured = true;
        }
    }
    static const int kid = register_kernel(Body::name());
    const int rec = (L.prof_on && L.prof_n < kMaxRecords) ? L.prof_n++ : -1;
    if (rec >= 0) { L.prof_kid[rec] = kid; cudaEventRecord(L.ev[2 * rec], (cudaStream_t)stream); }
    if (pdl_enabled() && rec < 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gx, gy, gz);
        cfg.blockDim = dim3(kThreads, 1, 1);
        cfg.dynamicSmemBytes = bytes;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, pcd_kernel<Body, Args>, a);
    } else {
        pcd_kernel<Body, Args><<<dim3(gx, gy, gz), kThreads, bytes, (cudaStream_t)stream>>>(a);
    }
    if (rec >= 0) cudaEventRecord(L.ev[2 * rec + 1], (cudaStream_t)stream);
    count_launch(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(L.last_err, sizeof L.last_err, "launch %s: %s", Body::name(), cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}
#else
#define PCD_D inline

template <class Body, class Args>
static int launch(const Args& a, int gx, int gy, int gz, size_t smem_floats, void*) {
    LaunchState& L = launch_state();
    if (gx <= 0 || gy <= 0 || gz <= 0) return PCD_OK;
    if (smem_floats * sizeof(float) > 227 * 1024) return PCD_ERR_UNSUPPORTED;
    static const int kid = register_kernel(Body::name());
    (void)kid;
    count_launch(L);
    float* smem = (float*)aligned_alloc(64, (smem_floats * sizeof(float) + 63) / 64 * 64 + 64);
    for (int z = 0; z < gz; ++z)
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) {
                for (size_t i = 0; i < smem_floats; ++i) smem[i] = NAN;   // catch reads of unwritten smem
                Body::run(a, x, y, z, smem);
            }
    free(smem);
    return PCD_OK;
}
#endif

}  // namespace pcd

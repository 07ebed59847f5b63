// pcd_tc.cuh — tcgen05 / TMEM / TMA / mbarrier primitives shared by the tensor-core kernels (sm_100a only).
#pragma once
#include "pcd_launch.cuh"

#if PCD_CUDA
#include <cuda.h>

namespace pcd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// K-major shared-memory matrix descriptor: rows of ROWB bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B), 8-row swizzle atoms
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);      // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused with swizzle, canonical value 1)
    d |= (uint64_t)((8 * ROWB) >> 4) << 32;        // stride byte offset: next 8-row group   bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;    // SWIZZLE_128B / SWIZZLE_64B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, K-major both, N >> 3, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// x = hi + lo with hi = x truncated to TF32 (kind::tf32 ignores the low 13 mantissa bits of the raw word, so the raw word
// itself is fed as the hi operand) and lo = x - hi, exact in fp32
__device__ __forceinline__ uint32_t tf32_lo(uint32_t v) {
    return __float_as_uint(__uint_as_float(v) - __uint_as_float(v & 0xFFFFE000u));
}

// 16 fp32 values of one operand row (one k-block of 16) into a K-major SWIZZLE_64B tile: 64-byte rows, 16-byte chunk c of row r
// lives at chunk position c ^ ((r >> 1) & 3) (Swizzle<2,4,3>); the tile base is 1024-byte aligned.  hi and lo tiles share it.
__device__ __forceinline__ void store_row_sw64(uint8_t* tile, int r, const uint32_t (&v)[16]) {
    uint8_t* row = tile + r * 64;
    const int x = (r >> 1) & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(row + ((c ^ x) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// sum over the 32 lanes of a warp of 16 per-lane values: lane L (even) ends up with the total of value index
// ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1); 16 shuffles
__device__ __forceinline__ float warp_sum16(const float (&v)[16], int& index) {
    const unsigned lane = threadIdx.x & 31u;
    float a8[8], a4[4], a2[2], a1;
    {
        const bool hi = (lane & 16u) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) a8[i] = (hi ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, hi ? v[i] : v[i + 8], 16);
    }
    {
        const bool hi = (lane & 8u) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) a4[i] = (hi ? a8[i + 4] : a8[i]) + __shfl_xor_sync(0xffffffffu, hi ? a8[i] : a8[i + 4], 8);
    }
    {
        const bool hi = (lane & 4u) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) a2[i] = (hi ? a4[i + 2] : a4[i]) + __shfl_xor_sync(0xffffffffu, hi ? a4[i] : a4[i + 2], 4);
    }
    {
        const bool hi = (lane & 2u) != 0;
        a1 = (hi ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, hi ? a2[0] : a2[1], 2);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    index = (int)(((lane >> 4) & 1u) * 8 + ((lane >> 3) & 1u) * 4 + ((lane >> 2) & 1u) * 2 + ((lane >> 1) & 1u));
    return a1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// fp32 tensor of `rank` dims (dims[0] innermost, contiguous), byte strides of dims 1.. in `strides`; zero fill out of bounds
inline int make_map_nd(CUtensorMap* m, const float* p, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                       const cuuint32_t* estr, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return PCD_ERR_CUDA;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(p), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(launch_state().last_err, sizeof launch_state().last_err, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}

}  // namespace tc
}  // namespace pcd
#endif

// pcd_gemm_sm100.cu — fp32-accurate GEMM on the Blackwell tensor cores (tcgen05 / TMEM / TMA), used for the dense
// contractions of the search step that sit above the roofline ridge: the question decoder's vocabulary projection
// (vqa_model.py:192-194: fc1 over B*30 rows, 512 -> V) and its two backward products.
//
//     C[M][N] (+)= A[M][K] * B[N][K]^T (+ bias[N])          all fp32, row-major, K contiguous in A and B
//
// tcgen05 has no fp32 MMA kind, so every fp32 operand is split into two TF32-representable parts
//     x = hi + lo,   hi = x with the low 13 mantissa bits cleared,  lo = x - hi   (exact in fp32)
// and the product is accumulated in fp32 in tensor memory as  hi*hi + hi*lo + lo*hi  (3xTF32; the dropped lo*lo term
// and the truncation of lo are ~2^-22 relative).  Algorithmic flops are counted once.
//
// Tile: 128 x 256 x 16 (64-byte SWIZZLE_64B rows) with 4 stages by default; 128 x 128 x 32 (SWIZZLE_128B, 3 stages) for
// narrow outputs.  Warp roles (192 threads, one 128 x BN tile per CTA, one CTA per SM):
//   warp 0    : TMA producer  — cp.async.bulk.tensor (swizzled, BK fp32 per row) into the hi tiles
//   warps 2-5 : splitters     — lo = x - trunc13(x) of each landed stage into a second tile (same swizzled positions;
//                               the raw tile itself serves as hi: the MMA ignores the low 13 mantissa bits), then the
//                               epilogue: tcgen05.ld of the accumulator, bias, store / atomic add (split-K)
//   warp 1    : MMA issuer    — one elected thread issues 3 tcgen05.mma.kind::tf32 per 8-wide k-step (hi*hi, hi*lo, lo*hi),
//                               tcgen05.commit releases the stage / signals the epilogue
// mbarriers: full[s] (TMA bytes landed) -> split[s] (128 splitter arrivals) -> MMA -> empty[s] (commit) -> TMA.
#include "../../include/pcdarts_sm100.h"
#include "pcd_launch.cuh"

#if PCD_CUDA
#include <cuda.h>

namespace pcd {
namespace gemm {

constexpr int BM = 128, UMMA_K = 8, kGemmThreads = 192, kSplitThreads = 128;

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

template <int BN, int BK, int STAGES>
struct Cfg {
    static constexpr uint32_t ROWB = BK * 4, A_BYTES = BM * ROWB, B_BYTES = BN * ROWB, STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both, N >> 3, M >> 4
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <int BN, int BK, int STAGES, bool REWRITE_HI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_3xtf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                      long long ldc, int M, int N, int K, const float* __restrict__ bias, int kb_per_split, int atomic) {
    using G = Cfg<BN, BK, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * G::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* split = bars + STAGES;
    uint64_t* empty = bars + 2 * STAGES;
    uint64_t* accum = bars + 3 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int total_kb = (K + BK - 1) / BK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int kb1 = (kb0 + kb_per_split < total_kb) ? kb0 + kb_per_split : total_kb;
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], kSplitThreads);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    auto a_hi = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES; };
    auto b_hi = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES + G::A_BYTES; };
    auto a_lo = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES + G::A_BYTES + G::B_BYTES; };
    auto b_lo = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES + 2 * G::A_BYTES + G::B_BYTES; };

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], G::A_BYTES + G::B_BYTES);
                tma_load_2d(a_hi(s), &tmA, &full[s], (kb0 + i) * BK, m0);
                tma_load_2d(b_hi(s), &tmB, &full[s], (kb0 + i) * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&split[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t dah = umma_desc<G::ROWB>(smem_u32(a_hi(s))), dbh = umma_desc<G::ROWB>(smem_u32(b_hi(s)));
                const uint64_t dal = umma_desc<G::ROWB>(smem_u32(a_lo(s))), dbl = umma_desc<G::ROWB>(smem_u32(b_lo(s)));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);      // bytes >> 4 along the 128-byte row
                    umma_tf32(tmem_base, dah + adv, dbh + adv, G::IDESC, (i > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, dah + adv, dbl + adv, G::IDESC, 1u);
                    umma_tf32(tmem_base, dal + adv, dbh + adv, G::IDESC, 1u);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(accum);
        }
    } else {
        // ---- splitters ---------------------------------------------------------------------------------------
        const int t = threadIdx.x - 64;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            uint4* hi = reinterpret_cast<uint4*>(a_hi(s));      // A then B are contiguous; so are their lo tiles
            uint4* lo = reinterpret_cast<uint4*>(a_lo(s));
            constexpr int NV = (G::A_BYTES + G::B_BYTES) / 16 / kSplitThreads;
#pragma unroll 8
            for (int j = 0; j < NV; ++j) {
                const uint4 v = hi[t + kSplitThreads * j];
                uint4 h, l;
                h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
                l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
                l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
                l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
                l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
                if (REWRITE_HI) hi[t + kSplitThreads * j] = h;
                lo[t + kSplitThreads * j] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&split[s]);
        }
        // ---- epilogue: TMEM lane quadrant of this warp = rows m0 + 32*(warp % 4) .. + 31 -----------------------------
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        if (nkb > 0) {
            mbar_wait(accum, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        float* crow = C + (long long)row * ldc;
        const bool vec_ok = !atomic && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        const bool add_bias = bias != nullptr && blockIdx.z == 0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            if (nkb > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (row < M) {
                const int col0 = n0 + c * 32;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const int col = col0 + 4 * j4;
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j] = __uint_as_float(r[4 * j4 + j]);
                        if (add_bias && col + j < N) v[j] += bias[col + j];
                    }
                    if (vec_ok && col + 3 < N) {
                        *reinterpret_cast<float4*>(crow + col) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col + j < N) {
                                if (atomic) atomicAdd(crow + col + j, v[j]);
                                else crow[col + j] = v[j];
                            }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ---- persistent variant (r2) ------------------------------------------------------------------------------------------------
// One CTA per SM walks the work list (output tile x K split); the accumulator is double-buffered in tensor memory (2 x BN
// columns), so the epilogue of tile i (tcgen05.ld -> global stores, done by four DEDICATED warps) overlaps the TMA / split /
// MMA pipeline of tile i + 1, which never drains between tiles.  Work order: m-tile fastest, so the CTAs running at one time
// share a few B (weight) tiles through L2.
//   warp 0 TMA | warp 1 MMA | warps 2-5 splitters | warps 6-9 epilogue (TMEM lane quadrant = warp % 4)
//   full[s] -> split[s] -> MMA -> empty[s];   MMA -> accf[a] -> epilogue -> acce[a] -> MMA
constexpr int kPersistThreads = 320;

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(kPersistThreads, 1)
gemm_tn_3xtf32_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                              long long ldc, int M, int N, int K, const float* __restrict__ bias, int kb_per_split, int nsplit,
                              int m_tiles, int n_tiles) {
    using G = Cfg<BN, BK, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * G::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* split = bars + STAGES;
    uint64_t* empty = bars + 2 * STAGES;
    uint64_t* accf = bars + 3 * STAGES;          // [2]
    uint64_t* acce = bars + 3 * STAGES + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_kb = (K + BK - 1) / BK;
    const int nwork = m_tiles * n_tiles * nsplit;
    const int atomic = nsplit > 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], kSplitThreads);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&accf[a], 1);
            mbar_init(&acce[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    auto a_hi = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES; };
    auto b_hi = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES + G::A_BYTES; };
    auto a_lo = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES + G::A_BYTES + G::B_BYTES; };
    auto b_lo = [&](int s) { return smem + (size_t)s * G::STAGE_BYTES + 2 * G::A_BYTES + G::B_BYTES; };
    // work item w -> (m tile fastest, n tile, split)
    auto decode = [&](int w, int& m0, int& n0, int& kb0, int& nkb, int& z) {
        const int mt = w % m_tiles, r = w / m_tiles, nt = r % n_tiles;
        z = r / n_tiles;
        m0 = mt * BM; n0 = nt * BN;
        kb0 = z * kb_per_split;
        const int kb1 = (kb0 + kb_per_split < total_kb) ? kb0 + kb_per_split : total_kb;
        nkb = kb1 - kb0;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
                int m0, n0, kb0, nkb, z;
                decode(w, m0, n0, kb0, nkb, z);
                for (int i = 0; i < nkb; ++i, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], G::A_BYTES + G::B_BYTES);
                    tma_load_2d(a_hi(s), &tmA, &full[s], (kb0 + i) * BK, m0);
                    tma_load_2d(b_hi(s), &tmB, &full[s], (kb0 + i) * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t g = 0, it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
                int m0, n0, kb0, nkb, z;
                decode(w, m0, n0, kb0, nkb, z);
                const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
                mbar_wait(&acce[as], aph ^ 1u);                       // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + as * (uint32_t)BN;
                for (int i = 0; i < nkb; ++i, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1u;
                    mbar_wait(&split[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t dah = umma_desc<G::ROWB>(smem_u32(a_hi(s))), dbh = umma_desc<G::ROWB>(smem_u32(b_hi(s)));
                    const uint64_t dal = umma_desc<G::ROWB>(smem_u32(a_lo(s))), dbl = umma_desc<G::ROWB>(smem_u32(b_lo(s)));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
                        umma_tf32(tacc, dah + adv, dbh + adv, G::IDESC, (i > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(tacc, dah + adv, dbl + adv, G::IDESC, 1u);
                        umma_tf32(tacc, dal + adv, dbh + adv, G::IDESC, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&accf[as]);
            }
        }
    } else if (warp < 6) {
        // ---- splitters: lo = x - trunc13(x) of every landed stage, across tiles without a break ---------------------------
        const int t = threadIdx.x - 64;
        uint32_t g = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
            int m0, n0, kb0, nkb, z;
            decode(w, m0, n0, kb0, nkb, z);
            for (int i = 0; i < nkb; ++i, ++g) {
                const int s = g % STAGES;
                const uint32_t ph = (g / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                const uint4* hi = reinterpret_cast<const uint4*>(a_hi(s));
                uint4* lo = reinterpret_cast<uint4*>(a_lo(s));
                constexpr int NV = (G::A_BYTES + G::B_BYTES) / 16 / kSplitThreads;
#pragma unroll 8
                for (int j = 0; j < NV; ++j) {
                    const uint4 v = hi[t + kSplitThreads * j];
                    uint4 l;
                    l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
                    l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
                    l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
                    l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
                    lo[t + kSplitThreads * j] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&split[s]);
            }
        }
    } else {
        // ---- epilogue warps 6..9: TMEM lane quadrant warp % 4 = rows m0 + 32 * (warp % 4) .. + 31 ---------------------------
        const int q = warp & 3;
        const bool vec_base = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        uint32_t it = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            int m0, n0, kb0, nkb, z;
            decode(w, m0, n0, kb0, nkb, z);
            const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
            mbar_wait(&accf[as], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + q * 32 + lane;
            float* crow = C + (long long)row * ldc;
            const bool vec_ok = !atomic && vec_base;
            const bool add_bias = bias != nullptr && z == 0;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + as * (uint32_t)BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
                if (row < M) {
                    const int col0 = n0 + c * 32;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const int col = col0 + 4 * j4;
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[j] = __uint_as_float(r[4 * j4 + j]);
                            if (add_bias && col + j < N) v[j] += bias[col + j];
                        }
                        if (vec_ok && col + 3 < N) {
                            *reinterpret_cast<float4*>(crow + col) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (col + j < N) {
                                    if (atomic) atomicAdd(crow + col + j, v[j]);
                                    else crow[col + j] = v[j];
                                }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acce[as]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x K fp32 matrix, row pitch ld (elements); box = bk columns (128 or 64 bytes) x box_rows, matching swizzle, zero fill
static int make_map(CUtensorMap* m, const float* p, long long rows, long long K, long long ld, int box_rows, int bk) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return PCD_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(launch_state().last_err, sizeof launch_state().last_err, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}

// The raw fp32 tile is fed as the hi operand: kind::tf32 ignores the low 13 mantissa bits (measured bit-identical to
// masking them explicitly, profiles/r01_gemm_3xtf32.txt; REWRITE_HI = true keeps the explicit variant compilable).
static int g_cfg = 0;

template <int BN, int BK, int STAGES, bool REWRITE_HI>
static int run(const CUtensorMap& ta, const CUtensorMap& tb, float* C, long long ldc, int M, int N, int K, const float* bias,
               int split_k, cudaStream_t st) {
    using G = Cfg<BN, BK, STAGES>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tn_3xtf32_kernel<BN, BK, STAGES, REWRITE_HI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES) != cudaSuccess)
            return PCD_ERR_CUDA;
        configured = true;
    }
    const int total_kb = (K + BK - 1) / BK;
    int kbps = (total_kb + split_k - 1) / split_k;
    if (kbps < 1) kbps = 1;
    const int gz = (total_kb + kbps - 1) / kbps;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, gz > 0 ? gz : 1);
    gemm_tn_3xtf32_kernel<BN, BK, STAGES, REWRITE_HI><<<grid, kGemmThreads, G::SMEM_BYTES, st>>>(ta, tb, C, ldc, M, N, K, bias, kbps, gz > 1 ? 1 : 0);
    LaunchState& L = launch_state();
    count_launch(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(L.last_err, sizeof L.last_err, "launch gemm_tn_3xtf32: %s", cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}

static bool persist_enabled() {
    static const bool on = !(getenv("PCD_GEMM_PERSIST") && getenv("PCD_GEMM_PERSIST")[0] == '0');
    return on;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaDeviceProp p;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        n = p.multiProcessorCount;
    }
    return n;
}

template <int BN, int BK, int STAGES>
static int run_persist(const CUtensorMap& ta, const CUtensorMap& tb, float* C, long long ldc, int M, int N, int K, const float* bias,
                       int split_k, cudaStream_t st) {
    using G = Cfg<BN, BK, STAGES>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tn_3xtf32_persist_kernel<BN, BK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES) != cudaSuccess)
            return PCD_ERR_CUDA;
        configured = true;
    }
    const int total_kb = (K + BK - 1) / BK;
    int kbps = (total_kb + split_k - 1) / split_k;
    if (kbps < 1) kbps = 1;
    const int nsplit = (total_kb + kbps - 1) / kbps;
    const int m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
    const long long nwork = (long long)m_tiles * n_tiles * nsplit;
    if (nwork > (1LL << 30)) return PCD_ERR_UNSUPPORTED;
    const int grid = nwork < sm_count() ? (int)nwork : sm_count();
    gemm_tn_3xtf32_persist_kernel<BN, BK, STAGES><<<grid, kPersistThreads, G::SMEM_BYTES, st>>>(ta, tb, C, ldc, M, N, K, bias, kbps, nsplit,
                                                                                            m_tiles, n_tiles);
    LaunchState& L = launch_state();
    count_launch(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(L.last_err, sizeof L.last_err, "launch gemm_tn_3xtf32_persist: %s", cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}

}  // namespace gemm
}  // namespace pcd

extern "C" int pcd_gemm_tn_3xtf32(const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, int M, int N,
                                  int K, const float* bias, int split_k, void* stream) {
    using namespace pcd;
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return PCD_ERR_ARG;
    if ((((uintptr_t)A) | ((uintptr_t)B)) & 15 || lda % 4 || ldb % 4 || lda < K || ldb < K || ldc < N) return PCD_ERR_ALIGN;
    if (split_k < 1) split_k = 1;
    cudaStream_t st = (cudaStream_t)stream;
    int cfg = gemm::g_cfg;
    if (cfg == 0) cfg = (N >= 256) ? 3 : 2;           // measured best: 128x256x16 with 4 stages (profiles/r01_gemm_3xtf32.txt)
    const int bn = (cfg == 1 || cfg == 3) ? 256 : 128, bk = (cfg <= 2) ? 32 : 16;
    CUtensorMap ta, tb;
    PCD_TRY(gemm::make_map(&ta, A, M, K, lda, gemm::BM, bk));
    PCD_TRY(gemm::make_map(&tb, B, N, K, ldb, bn, bk));
    const int total_kb = (K + bk - 1) / bk;
    if (split_k > total_kb) split_k = total_kb;
    if (split_k > 1) {      // partial sums are added atomically: start from zero
        if (cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st) != cudaSuccess) return PCD_ERR_CUDA;
    }
    if (gemm::persist_enabled()) {          // double-buffered accumulators need 2 * BN <= 512 tensor-memory columns: both tiles fit
        switch (cfg) {
            case 1: return gemm::run_persist<256, 32, 2>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
            case 2: return gemm::run_persist<128, 32, 3>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
            case 3: return gemm::run_persist<256, 16, 4>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
            default: return gemm::run_persist<128, 16, 6>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
        }
    }
    switch (cfg) {
        case 1: return gemm::run<256, 32, 2, false>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
        case 2: return gemm::run<128, 32, 3, false>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
        case 3: return gemm::run<256, 16, 4, false>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
        default: return gemm::run<128, 16, 6, false>(ta, tb, C, ldc, M, N, K, bias, split_k, st);
    }
}

/* experiment knob (not part of the reference-facing ABI): tile configuration 0 = auto, 1 = 128x256x32 (2 stages),
 * 2 = 128x128x32 (3 stages), 3 = 128x256x16 (4 stages), 4 = 128x128x16 (6 stages) */
extern "C" int pcd_gemm_debug_cfg(int cfg) { pcd::gemm::g_cfg = cfg; return 0; }

#else   // ---- CPU emulation build (tests only): plain fp32 loops ------------------------------------------------------

extern "C" int pcd_gemm_tn_3xtf32(const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, int M, int N,
                                  int K, const float* bias, int split_k, void*) {
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return PCD_ERR_ARG;
    (void)split_k;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = bias ? bias[n] : 0.0;
            for (int k = 0; k < K; ++k) s += (double)A[m * lda + k] * (double)B[n * ldb + k];
            C[m * ldc + n] = (float)s;
        }
    return PCD_OK;
}
extern "C" int pcd_gemm_debug_cfg(int) { return 0; }
#endif

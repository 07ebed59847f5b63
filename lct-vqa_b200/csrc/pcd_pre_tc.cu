// pcd_pre_tc.cu — the Cell preprocess op ReLUConvBN(C_in, C_out, 1, 1, 0, affine=False) (operations.py:22-33; call sites
// model_search.py:67-71) on the Blackwell tensor cores: the only dense contractions inside the search network (K = 48..256).
//
//   forward   z[co][p]   = sum_ci W[co][ci] * relu(x[ci][p])          + per-channel sum / sum^2 (BatchNorm statistics)
//   backward  dx[ci][p]  = relu'(x[ci][p]) * sum_co W[co][ci] * dz[co][p]
//             dW[co][ci] = sum_{n,p} dz[co][p] * relu(x[ci][p])          dz = rstd * (dy - mean(dy) - yhat * mean(dy * yhat))
//
// NCHW keeps the PIXEL index contiguous, so in the first two products the pixel (M) dimension of the A operand is the
// contiguous one ("MN-major").  Instead of an MN-major UMMA descriptor the splitter warps — which every 3xTF32 kernel here
// needs anyway to form lo = x - tf32(x) — also transpose: TMA lands the raw [16 channels][128 pixels] box unswizzled, each
// splitter thread owns one pixel, reads its 16 channel values (conflict-free: consecutive lanes = consecutive floats),
// applies the operand's transform (ReLU; BatchNorm-backward), and writes the hi / lo rows of a K-major SWIZZLE_64B tile
// (the swizzle XOR is done by hand, tc::store_row_sw64).  The MMA side is the ordinary K-major tcgen05.mma.kind::tf32
// issue (hi*hi + hi*lo + lo*hi, fp32 accumulation in tensor memory), the epilogue reads the accumulator with tcgen05.ld:
// one thread = one pixel row, so stores to the NCHW planes are coalesced across the warp.  In the dW product both operands
// are K-major as they lie in memory (K = pixels); the transforms are applied in place on the TMA-swizzled tiles.
//
// Roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-5 = splitters, then
// epilogue.  mbarriers: full[s] (TMA bytes landed) -> split[s] (128 arrivals) -> MMA -> empty[s] (tcgen05.commit) -> TMA.
#include "pcd_pre.cuh"
#include "pcd_kernels.h"
#include "pcd_tc.cuh"

#if PCD_CUDA

namespace pcd {
namespace pretc {

using namespace tc;

constexpr int kThreadsTC = 192, kSplit = 128, BM = 128;

struct Bars {
    uint64_t full[4], split[4], empty[4], accum;
    uint32_t tmem_slot, pad;
};

__device__ __forceinline__ void init_bars(Bars* b, int stages) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&b->full[s], 1);
            mbar_init(&b->split[s], kSplit);
            mbar_init(&b->empty[s], 1);
        }
        mbar_init(&b->accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
__device__ __forceinline__ void epilogue_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ======================================================================================================
// forward: z = W * relu(x) on a 128-pixel tile of one image; BN = C_out in {16, 32, 64}; K = C_in in blocks of 16
// ======================================================================================================
template <int BN, int STAGES>
struct FwdCfg {
    static constexpr uint32_t RAW = 16 * BM * 4, AT = BM * 64, BT = BN * 64;
    static constexpr uint32_t STAGE = RAW + 2 * AT + 2 * BT;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE + 1024 + sizeof(Bars) + 2 * BN * sizeof(float);
    static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

// one stage per CTA and five CTAs per SM: a tile has only 3..16 k-blocks, so the latency of its TMA -> split -> MMA -> epilogue
// chain is hidden by the OTHER resident tiles rather than by a deep ring inside the CTA
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreadsTC, 5)
pre_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* __restrict__ z,
                  double* __restrict__ stats, int Cin, int HW, int rows_per_tile) {
    using G = FwdCfg<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Bars* bars = reinterpret_cast<Bars*>(smem + (size_t)STAGES * G::STAGE);
    float* SS = reinterpret_cast<float*>(bars + 1);            // [2][BN] block sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, n = blockIdx.y;
    const int nkb = Cin / 16;
    init_bars(bars, STAGES);
    if (threadIdx.x < 2 * BN) SS[threadIdx.x] = 0.f;
    if (warp == 1) tmem_alloc(&bars->tmem_slot, G::TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = bars->tmem_slot;
    auto raw = [&](int s) { return smem + (size_t)s * G::STAGE; };
    auto a_hi = [&](int s) { return smem + (size_t)s * G::STAGE + G::RAW; };
    auto a_lo = [&](int s) { return smem + (size_t)s * G::STAGE + G::RAW + G::AT; };
    auto b_hi = [&](int s) { return smem + (size_t)s * G::STAGE + G::RAW + 2 * G::AT; };
    auto b_lo = [&](int s) { return smem + (size_t)s * G::STAGE + G::RAW + 2 * G::AT + G::BT; };

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&bars->empty[s], ph ^ 1u);
                mbar_expect_tx(&bars->full[s], G::RAW + G::BT);
                tma_load_3d(raw(s), &tmX, &bars->full[s], 0, tile * rows_per_tile, n * Cin + i * 16);
                tma_load_2d(b_hi(s), &tmW, &bars->full[s], i * 16, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC = idesc_tf32(BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&bars->split[s], ph);
                fence_after();
                const uint64_t dah = umma_desc<64>(smem_u32(a_hi(s))), dal = umma_desc<64>(smem_u32(a_lo(s)));
                const uint64_t dbh = umma_desc<64>(smem_u32(b_hi(s))), dbl = umma_desc<64>(smem_u32(b_lo(s)));
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                    umma_tf32(tmem_base, dah + adv, dbh + adv, IDESC, (i > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, dah + adv, dbl + adv, IDESC, 1u);
                    umma_tf32(tmem_base, dal + adv, dbh + adv, IDESC, 1u);
                }
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->accum);
        }
    } else {
        const int t = threadIdx.x - 64;                        // pixel of the tile
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(&bars->full[s], ph);
            const float* rw = reinterpret_cast<const float*>(raw(s));
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float v = fmaxf(rw[k * BM + t], 0.f);   // ReLU
                hi[k] = __float_as_uint(v);
                lo[k] = tf32_lo(hi[k]);
            }
            store_row_sw64(a_hi(s), t, hi);
            store_row_sw64(a_lo(s), t, lo);
            uint4* bh = reinterpret_cast<uint4*>(b_hi(s));
            uint4* bl = reinterpret_cast<uint4*>(b_lo(s));
            for (int j = t; j < BN * 4; j += kSplit) {
                const uint4 v = bh[j];
                bl[j] = make_uint4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
            fence_async_smem();
            mbar_arrive(&bars->split[s]);
        }
        // ---- epilogue: thread = pixel row of the accumulator; coalesced stores into the NCHW planes; BN sums ----------
        const int q = warp & 3, row = q * 32 + lane;
        mbar_wait(&bars->accum, 0);
        fence_after();
        float* zp = z + (long long)n * BN * HW + (long long)tile * BM + row;
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), r);
            float sv[16], qv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float v = __uint_as_float(r[j]);
                zp[(long long)(c * 16 + j) * HW] = v;
                sv[j] = v;
                qv[j] = v * v;
            }
            int idx;
            const float ts = warp_sum16(sv, idx);
            const float tq = warp_sum16(qv, idx);
            if ((lane & 1) == 0) {
                atomicAdd(&SS[c * 16 + idx], ts);
                atomicAdd(&SS[BN + c * 16 + idx], tq);
            }
        }
        epilogue_barrier();
        if (t < 2 * BN) atomicAdd(stats + t, (double)SS[t]);
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, G::TMEM_COLS);
}

// ======================================================================================================
// backward, data: dx = relu'(x) * (W^T dz) on a 128-pixel tile; N = C_in (multiple of 16, <= 256); K = C_out in blocks of 16
// ======================================================================================================
template <int STAGES>
struct DxCfg {
    static constexpr uint32_t RAW = 16 * BM * 4, AT = BM * 64;
    static constexpr uint32_t stage(int cin) { return 2 * RAW + 2 * AT + 16u * cin * 4 + 2u * cin * 64; }
    static size_t smem(int cin, int cout) { return (size_t)STAGES * stage(cin) + 1024 + sizeof(Bars) + 3 * cout * sizeof(float); }
};

template <int STAGES>
__global__ void __launch_bounds__(kThreadsTC, 3)
pre_tc_dx_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW,
                 const float* __restrict__ x, float* __restrict__ dx, const double* __restrict__ stats, const double* __restrict__ bstats,
                 float eps, double cnt, int Cin, int Cout, int HW, int rows_per_tile, uint32_t tmem_cols) {
    using G = DxCfg<STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t STAGE = G::stage(Cin), WRAW = 16u * Cin * 4, BT = (uint32_t)Cin * 64;
    Bars* bars = reinterpret_cast<Bars*>(smem + (size_t)STAGES * STAGE);
    float* COEF = reinterpret_cast<float*>(bars + 1);         // [Cout][3]: rstd, mean(dy), mean(dy*yhat)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, n = blockIdx.y;
    const int nkb = Cout / 16;
    init_bars(bars, STAGES);
    if ((int)threadIdx.x < Cout) {
        const int co = threadIdx.x;
        BnC b = bn_consts(stats, Cout, 0, co, cnt, eps);
        COEF[3 * co] = b.rstd;
        COEF[3 * co + 1] = (float)(bstats[co] / cnt);
        COEF[3 * co + 2] = (float)(bstats[Cout + co] / cnt);
    }
    if (warp == 1) tmem_alloc(&bars->tmem_slot, tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = bars->tmem_slot;
    auto rdy = [&](int s) { return smem + (size_t)s * STAGE; };
    auto ry = [&](int s) { return smem + (size_t)s * STAGE + G::RAW; };
    auto a_hi = [&](int s) { return smem + (size_t)s * STAGE + 2 * G::RAW; };
    auto a_lo = [&](int s) { return smem + (size_t)s * STAGE + 2 * G::RAW + G::AT; };
    auto wraw = [&](int s) { return smem + (size_t)s * STAGE + 2 * G::RAW + 2 * G::AT; };
    auto b_hi = [&](int s) { return smem + (size_t)s * STAGE + 2 * G::RAW + 2 * G::AT + WRAW; };
    auto b_lo = [&](int s) { return smem + (size_t)s * STAGE + 2 * G::RAW + 2 * G::AT + WRAW + BT; };

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&bars->empty[s], ph ^ 1u);
                mbar_expect_tx(&bars->full[s], 2 * G::RAW + WRAW);
                tma_load_3d(rdy(s), &tmDY, &bars->full[s], 0, tile * rows_per_tile, n * Cout + i * 16);
                tma_load_3d(ry(s), &tmY, &bars->full[s], 0, tile * rows_per_tile, n * Cout + i * 16);
                tma_load_2d(wraw(s), &tmW, &bars->full[s], 0, i * 16);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t IDESC = idesc_tf32(Cin);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&bars->split[s], ph);
                fence_after();
                const uint64_t dah = umma_desc<64>(smem_u32(a_hi(s))), dal = umma_desc<64>(smem_u32(a_lo(s)));
                const uint64_t dbh = umma_desc<64>(smem_u32(b_hi(s))), dbl = umma_desc<64>(smem_u32(b_lo(s)));
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                    umma_tf32(tmem_base, dah + adv, dbh + adv, IDESC, (i > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, dah + adv, dbl + adv, IDESC, 1u);
                    umma_tf32(tmem_base, dal + adv, dbh + adv, IDESC, 1u);
                }
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->accum);
        }
    } else {
        const int t = threadIdx.x - 64;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(&bars->full[s], ph);
            const float* dyr = reinterpret_cast<const float*>(rdy(s));
            const float* yr = reinterpret_cast<const float*>(ry(s));
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {                     // A row = pixel t: dz over the 16 output channels of this k-block
                const float* cf = COEF + 3 * (i * 16 + k);
                const float v = cf[0] * (dyr[k * BM + t] - cf[1] - yr[k * BM + t] * cf[2]);
                hi[k] = __float_as_uint(v);
                lo[k] = tf32_lo(hi[k]);
            }
            store_row_sw64(a_hi(s), t, hi);
            store_row_sw64(a_lo(s), t, lo);
            const float* wr = reinterpret_cast<const float*>(wraw(s));      // [16 co][Cin]: B row = input channel j
            for (int j = t; j < Cin; j += kSplit) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    hi[k] = __float_as_uint(wr[k * Cin + j]);
                    lo[k] = tf32_lo(hi[k]);
                }
                store_row_sw64(b_hi(s), j, hi);
                store_row_sw64(b_lo(s), j, lo);
            }
            fence_async_smem();
            mbar_arrive(&bars->split[s]);
        }
        const int q = warp & 3, row = q * 32 + lane;
        mbar_wait(&bars->accum, 0);
        fence_after();
        const long long o = (long long)n * Cin * HW + (long long)tile * BM + row;
        const float* xp = x + o;
        float* dp = dx + o;
#pragma unroll 1
        for (int c = 0; c < Cin / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), r);
            float xv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) xv[j] = xp[(long long)(c * 16 + j) * HW];
#pragma unroll
            for (int j = 0; j < 16; ++j) dp[(long long)(c * 16 + j) * HW] = xv[j] > 0.f ? __uint_as_float(r[j]) : 0.f;
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ======================================================================================================
// backward, weights: dW^T[ci][co] = sum_p relu(x)[ci][p] * dz[co][p]; M = 128 rows of (image, input channel), N = C_out,
// K = a chunk of the image's pixels in blocks of 32 (both operands K-major as they lie in memory, SWIZZLE_128B)
// ======================================================================================================
template <int BN, int STAGES>
struct DwCfg {
    static constexpr uint32_t AT = BM * 128, BT = BN * 128;
    static constexpr uint32_t STAGE = 2 * AT + 3 * BT;        // A hi | A lo | dy -> dz hi | yhat | dz lo
    static constexpr size_t SMEM = (size_t)STAGES * STAGE + 1024 + sizeof(Bars) + 3 * BN * sizeof(float);
    static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreadsTC, BN == 64 ? 1 : 2)
pre_tc_dw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmY,
                 float* __restrict__ gw, const double* __restrict__ stats, const double* __restrict__ bstats, float eps, double cnt,
                 int Cin, int HW, int kb_per_cta) {
    using G = DwCfg<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Bars* bars = reinterpret_cast<Bars*>(smem + (size_t)STAGES * G::STAGE);
    float* COEF = reinterpret_cast<float*>(bars + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x, n = blockIdx.y;
    const int total_kb = HW / 32;
    const int kb0 = blockIdx.z * kb_per_cta;
    const int nkb = (kb0 + kb_per_cta <= total_kb) ? kb_per_cta : total_kb - kb0;
    init_bars(bars, STAGES);
    if ((int)threadIdx.x < BN) {
        const int co = threadIdx.x;
        BnC b = bn_consts(stats, BN, 0, co, cnt, eps);
        COEF[3 * co] = b.rstd;
        COEF[3 * co + 1] = (float)(bstats[co] / cnt);
        COEF[3 * co + 2] = (float)(bstats[BN + co] / cnt);
    }
    if (warp == 1) tmem_alloc(&bars->tmem_slot, G::TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = bars->tmem_slot;
    auto a_hi = [&](int s) { return smem + (size_t)s * G::STAGE; };
    auto a_lo = [&](int s) { return smem + (size_t)s * G::STAGE + G::AT; };
    auto b_hi = [&](int s) { return smem + (size_t)s * G::STAGE + 2 * G::AT; };              // dy, rewritten as dz
    auto b_y = [&](int s) { return smem + (size_t)s * G::STAGE + 2 * G::AT + G::BT; };
    auto b_lo = [&](int s) { return smem + (size_t)s * G::STAGE + 2 * G::AT + 2 * G::BT; };

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&bars->empty[s], ph ^ 1u);
                mbar_expect_tx(&bars->full[s], G::AT + 2 * G::BT);
                tma_load_2d(a_hi(s), &tmX, &bars->full[s], (kb0 + i) * 32, n * Cin + mt * BM);
                tma_load_2d(b_hi(s), &tmDY, &bars->full[s], (kb0 + i) * 32, n * BN);
                tma_load_2d(b_y(s), &tmY, &bars->full[s], (kb0 + i) * 32, n * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t IDESC = idesc_tf32(BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&bars->split[s], ph);
                fence_after();
                const uint64_t dah = umma_desc<128>(smem_u32(a_hi(s))), dal = umma_desc<128>(smem_u32(a_lo(s)));
                const uint64_t dbh = umma_desc<128>(smem_u32(b_hi(s))), dbl = umma_desc<128>(smem_u32(b_lo(s)));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                    umma_tf32(tmem_base, dah + adv, dbh + adv, IDESC, (i > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, dah + adv, dbl + adv, IDESC, 1u);
                    umma_tf32(tmem_base, dal + adv, dbh + adv, IDESC, 1u);
                }
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->accum);
        }
    } else {
        const int t = threadIdx.x - 64;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(&bars->full[s], ph);
            uint4* ah = reinterpret_cast<uint4*>(a_hi(s));
            uint4* al = reinterpret_cast<uint4*>(a_lo(s));
#pragma unroll 8
            for (int j = 0; j < (int)(G::AT / 16 / kSplit); ++j) {      // A = relu(x): elementwise, in place on the swizzled tile
                uint4 v = ah[t + kSplit * j];
                v.x = __float_as_uint(fmaxf(__uint_as_float(v.x), 0.f)); v.y = __float_as_uint(fmaxf(__uint_as_float(v.y), 0.f));
                v.z = __float_as_uint(fmaxf(__uint_as_float(v.z), 0.f)); v.w = __float_as_uint(fmaxf(__uint_as_float(v.w), 0.f));
                ah[t + kSplit * j] = v;
                al[t + kSplit * j] = make_uint4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
            uint4* bh = reinterpret_cast<uint4*>(b_hi(s));
            const uint4* by = reinterpret_cast<const uint4*>(b_y(s));
            uint4* bl = reinterpret_cast<uint4*>(b_lo(s));
            for (int j = t; j < BN * 8; j += kSplit) {                  // B = dz: the row (= output channel) of a 16-byte chunk is j / 8
                const float* cf = COEF + 3 * (j >> 3);
                const uint4 d = bh[j], y = by[j];
                uint4 v;
                v.x = __float_as_uint(cf[0] * (__uint_as_float(d.x) - cf[1] - __uint_as_float(y.x) * cf[2]));
                v.y = __float_as_uint(cf[0] * (__uint_as_float(d.y) - cf[1] - __uint_as_float(y.y) * cf[2]));
                v.z = __float_as_uint(cf[0] * (__uint_as_float(d.z) - cf[1] - __uint_as_float(y.z) * cf[2]));
                v.w = __float_as_uint(cf[0] * (__uint_as_float(d.w) - cf[1] - __uint_as_float(y.w) * cf[2]));
                bh[j] = v;
                bl[j] = make_uint4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
            fence_async_smem();
            mbar_arrive(&bars->split[s]);
        }
        const int q = warp & 3, row = q * 32 + lane;
        const int ci = mt * BM + row;
        if (nkb > 0) {
            mbar_wait(&bars->accum, 0);
            fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 16; ++c) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), r);
                if (ci < Cin) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(gw + (long long)(c * 16 + j) * Cin + ci, __uint_as_float(r[j]));
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, G::TMEM_COLS);
}

// ---- host side -----------------------------------------------------------------------------------------------------
// (W, H, planes) view of an NCHW tensor; box = whole rows covering 128 pixels x 16 planes, unswizzled
static int map_planes_3d(CUtensorMap* m, const float* p, int W, int H, long long planes) {
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)(BM / W), 16};
    cuuint32_t es[3] = {1, 1, 1};
    return make_map_nd(m, p, 3, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE);
}
// (cols, rows) row-major matrix; box = bc x br with the given swizzle
static int map_2d(CUtensorMap* m, const float* p, long long cols, long long rows, int bc, int br, CUtensorMapSwizzle swz) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br};
    cuuint32_t es[2] = {1, 1};
    return make_map_nd(m, p, 2, dims, strides, box, es, swz);
}

static int launched(const char* what) {
    LaunchState& L = launch_state();
    count_launch(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(L.last_err, sizeof L.last_err, "launch %s: %s", what, cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}

struct ProfScope {          // per-launch CUDA events of the library's profiler (same records as launch<>)
    int rec;
    cudaStream_t st;
    ProfScope(const char* name, cudaStream_t s) : rec(-1), st(s) {
        LaunchState& L = launch_state();
        if (L.prof_on && L.prof_n < kMaxRecords) {
            rec = L.prof_n++;
            L.prof_kid[rec] = register_kernel(name);
            cudaEventRecord(L.ev[2 * rec], st);
        }
    }
    ~ProfScope() { if (rec >= 0) cudaEventRecord(launch_state().ev[2 * rec + 1], st); }
};

template <int BN>
static int fwd_go(const PreArgs& a, cudaStream_t st) {
    constexpr int STAGES = 1;
    using G = FwdCfg<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(pre_tc_fwd_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM) != cudaSuccess) return PCD_ERR_CUDA;
        configured = true;
    }
    CUtensorMap tx, tw;
    PCD_TRY(map_planes_3d(&tx, a.x, a.Win, a.Hin, (long long)a.B * a.Cin));
    PCD_TRY(map_2d(&tw, a.w, a.Cin, a.Cout, 16, BN, CU_TENSOR_MAP_SWIZZLE_64B));
    const int HW = a.Hin * a.Win;
    ProfScope prof("pre_tc_fwd", st);
    pre_tc_fwd_kernel<BN, STAGES><<<dim3(HW / BM, a.B), kThreadsTC, G::SMEM, st>>>(tx, tw, a.y, a.stats, a.Cin, HW, BM / a.Win);
    return launched("pre_tc_fwd");
}

template <int BN>
static int dw_go(const PreBwdArgs& a, cudaStream_t st) {
    constexpr int STAGES = BN == 64 ? 3 : 2;        // C_out <= 32: two CTAs per SM
    using G = DwCfg<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(pre_tc_dw_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM) != cudaSuccess) return PCD_ERR_CUDA;
        configured = true;
    }
    const int HW = a.Hin * a.Win;
    CUtensorMap tx, tdy, ty;
    PCD_TRY(map_2d(&tx, a.x, HW, (long long)a.B * a.Cin, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B));
    PCD_TRY(map_2d(&tdy, a.dy, HW, (long long)a.B * BN, 32, BN, CU_TENSOR_MAP_SWIZZLE_128B));
    PCD_TRY(map_2d(&ty, a.y, HW, (long long)a.B * BN, 32, BN, CU_TENSOR_MAP_SWIZZLE_128B));
    const int mt = (a.Cin + BM - 1) / BM, total_kb = HW / 32;
    int ksplit = (444 + mt * a.B - 1) / (mt * a.B);              // about three CTAs per SM over the grid
    if (ksplit < 1) ksplit = 1;
    if (ksplit > total_kb / 4) ksplit = total_kb / 4 > 0 ? total_kb / 4 : 1;
    const int kbps = (total_kb + ksplit - 1) / ksplit;
    const int gz = (total_kb + kbps - 1) / kbps;
    ProfScope prof("pre_tc_dw", st);
    pre_tc_dw_kernel<BN, STAGES><<<dim3(mt, a.B, gz), kThreadsTC, G::SMEM, st>>>(tx, tdy, ty, a.gw, a.stats, a.bstats, a.eps,
                                                                                 (double)a.B * HW, a.Cin, HW, kbps);
    return launched("pre_tc_dw");
}

}  // namespace pretc

// shapes the tensor-core preprocess kernels take: plain 1x1 (no FactorizedReduce), whole image rows per 128-pixel tile
bool pre_tc_supported(int B, int Cin, int Cout, int H, int W, int fr) {
    if (fr || B <= 0) return false;
    if (Cout != 16 && Cout != 32 && Cout != 64) return false;
    if (Cin % 16 || Cin < 16 || Cin > 256) return false;
    if (W != 16 && W != 32 && W != 64 && W != 128) return false;
    if ((H * W) % 128) return false;
    return getenv("PCD_NO_PRE_TC") == nullptr;
}

int launch_pre_tc_fwd(const PreArgs& a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (a.Cout == 16) return pretc::fwd_go<16>(a, st);
    if (a.Cout == 32) return pretc::fwd_go<32>(a, st);
    return pretc::fwd_go<64>(a, st);
}

int launch_pre_tc_bwd(const PreBwdArgs& a, void* stream) {
    using namespace pretc;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = a.Hin * a.Win;
    if (a.dx) {
        const int stages = 1;        // occupancy (3-5 CTAs per SM) hides the per-tile latency better than a second stage does
        const size_t sm = stages == 2 ? DxCfg<2>::smem(a.Cin, a.Cout) : DxCfg<1>::smem(a.Cin, a.Cout);
        static bool configured = false;
        if (!configured) {
            if (cudaFuncSetAttribute(pre_tc_dx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return PCD_ERR_CUDA;
            if (cudaFuncSetAttribute(pre_tc_dx_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return PCD_ERR_CUDA;
            configured = true;
        }
        CUtensorMap tdy, ty, tw;
        PCD_TRY(map_planes_3d(&tdy, a.dy, a.Win, a.Hin, (long long)a.B * a.Cout));
        PCD_TRY(map_planes_3d(&ty, a.y, a.Win, a.Hin, (long long)a.B * a.Cout));
        PCD_TRY(map_2d(&tw, a.w, a.Cin, a.Cout, a.Cin, 16, CU_TENSOR_MAP_SWIZZLE_NONE));
        uint32_t cols = 32;
        while ((int)cols < a.Cin) cols <<= 1;
        ProfScope prof("pre_tc_dx", st);
        if (stages == 2)
            pre_tc_dx_kernel<2><<<dim3(HW / BM, a.B), kThreadsTC, sm, st>>>(tdy, ty, tw, a.x, a.dx, a.stats, a.bstats, a.eps, (double)a.B * HW,
                                                                          a.Cin, a.Cout, HW, BM / a.Win, cols);
        else
            pre_tc_dx_kernel<1><<<dim3(HW / BM, a.B), kThreadsTC, sm, st>>>(tdy, ty, tw, a.x, a.dx, a.stats, a.bstats, a.eps, (double)a.B * HW,
                                                                          a.Cin, a.Cout, HW, BM / a.Win, cols);
        PCD_TRY(launched("pre_tc_dx"));
    }
    if (a.gw) {
        if (a.Cout == 16) return dw_go<16>(a, st);
        if (a.Cout == 32) return dw_go<32>(a, st);
        return dw_go<64>(a, st);
    }
    return PCD_OK;
}

}  // namespace pcd

#else   // ---- CPU emulation build: the tensor-core path does not exist there -------------------------------------------

namespace pcd {
bool pre_tc_supported(int, int, int, int, int, int) { return false; }
int launch_pre_tc_fwd(const PreArgs&, void*) { return PCD_ERR_UNSUPPORTED; }
int launch_pre_tc_bwd(const PreBwdArgs&, void*) { return PCD_ERR_UNSUPPORTED; }
}  // namespace pcd
#endif

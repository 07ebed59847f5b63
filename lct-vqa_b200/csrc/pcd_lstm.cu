// pcd_lstm.cu — the question encoder's single-layer LSTM recurrence (vqa_model.py:165,176-184: nn.LSTM(300, 512, 1) over
// T = 30 steps, h0 = c0 = image embedding) as two persistent cooperative kernels.
//
// cuDNN runs the recurrence as 30 (forward) + 30 (backward) tiny SIMT GEMM launches per pass; here the time loop lives
// inside one kernel: block g owns 4 hidden units (16 gate rows of W_hh, resident on chip for all steps), and the blocks
// exchange h_t / partial dh_t through L2 with one grid-wide barrier per step.
//   forward : gates_t[b][own rows] = gx_t + h_{t-1} W_hh^T  (h_{t-1} staged in shared memory), cell update, h_t -> global
//   backward: dgates_t[b][own rows] from (dh_t, dc_t, saved activations);  partial dh_{t-1}[b][:] = dgates_t[b][own rows]
//             W_hh[own rows][:] (own rows held in REGISTERS: thread = 4 output columns) -> per-block partial in global,
//             barrier, every block sums the 4 columns it owns over all blocks' partials
// The input projection gx = x W_ih^T + b_ih + b_hh and the weight / input gradients are dense GEMMs over all time steps
// at once and go through pcd_gemm_tn_3xtf32 (Python side, pcd_ops.LstmFunction).
#include "../../include/pcdarts_sm100.h"
#include "pcd_launch.cuh"

#if PCD_CUDA
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace pcd {
namespace lstm {

constexpr int kT = 256, kUnits = 4, kMaxB = 64;

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

struct FwdArgs {
    int T, B, H;
    const float* gx;      // [T][B][4H]  x W_ih^T + b_ih + b_hh
    const float* w_hh;    // [4H][H]
    const float* h0;      // [B][H]
    const float* c0;
    float* act;           // [T][B][4H]  i, f, g, o after the nonlinearities
    float* cs;            // [T][B][H]
    float* hs;            // [T][B][H]
};

__global__ void __launch_bounds__(kT, 1) lstm_fwd_kernel(FwdArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float sm[];
    const int H = a.H, B = a.B, H4 = H / 4, HP = H + 4;
    float* Ws = sm;                  // [16][HP]  row q*4+u  <-  W_hh row q*H + u0 + u
    float* Hs = Ws + 16 * HP;        // [B][HP]
    const int u0 = blockIdx.x * kUnits;
    for (int i = threadIdx.x; i < 16 * H4; i += kT) {
        const int rr = i / H4, k4 = i - rr * H4, q = rr >> 2, u = rr & 3;
        *reinterpret_cast<float4*>(Ws + rr * HP + 4 * k4) = *reinterpret_cast<const float4*>(a.w_hh + (long long)(q * H + u0 + u) * H + 4 * k4);
    }
    const int b = threadIdx.x & 63, u = threadIdx.x >> 6, unit = u0 + u;
    for (int t = 0; t < a.T; ++t) {
        const float* hprev = t ? a.hs + (long long)(t - 1) * B * H : a.h0;
        // h_{t-1} -> shared memory with 16-byte cp.async (L2 only: written by other SMs), all copies of a thread in flight at once
        for (int i = threadIdx.x; i < B * H4; i += kT) {
            const int bb = i / H4, k4 = i - bb * H4;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(Hs + bb * HP + 4 * k4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(hprev + (long long)bb * H + 4 * k4) : "memory");
        }
        // this thread's input-projection terms and previous cell state: issued before the wait so their latency overlaps
        float gxv[4] = {0.f, 0.f, 0.f, 0.f}, cprev = 0.f;
        if (b < B) {
            const long long row = (long long)t * B + b;
            const float* g = a.gx + row * 4 * H;
            gxv[0] = g[unit]; gxv[1] = g[H + unit]; gxv[2] = g[2 * H + unit]; gxv[3] = g[3 * H + unit];
            cprev = t ? a.cs[(row - B) * H + unit] : a.c0[(long long)b * H + unit];
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (b < B) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float* hr = Hs + b * HP;
#pragma unroll 4
            for (int k4 = 0; k4 < H4; ++k4) {
                const float4 h4 = *reinterpret_cast<const float4*>(hr + 4 * k4);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w4 = *reinterpret_cast<const float4*>(Ws + (q * 4 + u) * HP + 4 * k4);
                    acc[q] = fmaf(h4.x, w4.x, fmaf(h4.y, w4.y, fmaf(h4.z, w4.z, fmaf(h4.w, w4.w, acc[q]))));
                }
            }
            const long long row = (long long)t * B + b;
            const float gi = sigm(acc[0] + gxv[0]), gf = sigm(acc[1] + gxv[1]);
            const float gg = tanhf(acc[2] + gxv[2]), go = sigm(acc[3] + gxv[3]);
            const float c = fmaf(gf, cprev, gi * gg);
            float* ac = a.act + row * 4 * H;
            ac[unit] = gi; ac[H + unit] = gf; ac[2 * H + unit] = gg; ac[3 * H + unit] = go;
            a.cs[row * H + unit] = c;
            a.hs[row * H + unit] = go * tanhf(c);
        }
        grid.sync();
    }
}

struct BwdArgs {
    int T, B, H;
    const float* dhs;     // [T][B][H] grad w.r.t. every h_t (may be null)
    const float* dhT;     // [B][H] extra grad of the final hidden state (may be null)
    const float* dcT;     // [B][H] grad of the final cell state (may be null)
    const float* act;
    const float* cs;
    const float* c0;
    const float* w_hh;
    float* dgates;        // [T][B][4H]
    float* dh0;           // [B][H]
    float* dc0;
    float* pbuf;          // [2][gridDim.x][B][H] partial dh exchange
};

__global__ void __launch_bounds__(kT, 1) lstm_bwd_kernel(BwdArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) float DGs[kMaxB * 16];        // dgates of the block's 16 rows, [b][q*4+u]
    __shared__ __align__(16) float RED[4 * kMaxB * 4];
    const int H = a.H, B = a.B, H4 = H / 4, G = gridDim.x;
    const int u0 = blockIdx.x * kUnits;
    const int b = threadIdx.x & 63, u = threadIdx.x >> 6, unit = u0 + u;
    // register-resident W_hh[own 16 rows][4 columns of this thread]
    const int j4 = threadIdx.x % H4, bg = threadIdx.x / H4, nbg = kT / H4;       // H4 <= 256 and divides 256
    float4 Wr[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr)
        Wr[rr] = *reinterpret_cast<const float4*>(a.w_hh + (long long)((rr >> 2) * H + u0 + (rr & 3)) * H + 4 * j4);
    float dh_rec = 0.f, dc_rec = 0.f;
    if (b < B) {
        if (a.dhT) dh_rec = a.dhT[(long long)b * H + unit];
        if (a.dcT) dc_rec = a.dcT[(long long)b * H + unit];
    }
    int par = 0;
    for (int t = a.T - 1; t >= 0; --t) {
        if (b < B) {
            const long long row = (long long)t * B + b;
            const float dh = dh_rec + (a.dhs ? a.dhs[row * H + unit] : 0.f);
            const float* ac = a.act + row * 4 * H;
            const float gi = ac[unit], gf = ac[H + unit], gg = ac[2 * H + unit], go = ac[3 * H + unit];
            const float c = a.cs[row * H + unit];
            const float cprev = t ? a.cs[(row - B) * H + unit] : a.c0[(long long)b * H + unit];
            const float tc = tanhf(c);
            const float dc = fmaf(dh * go, 1.f - tc * tc, dc_rec);
            const float d_o = dh * tc * go * (1.f - go);
            const float d_i = dc * gg * gi * (1.f - gi);
            const float d_f = dc * cprev * gf * (1.f - gf);
            const float d_g = dc * gi * (1.f - gg * gg);
            dc_rec = dc * gf;
            float* dg = a.dgates + row * 4 * H;
            dg[unit] = d_i; dg[H + unit] = d_f; dg[2 * H + unit] = d_g; dg[3 * H + unit] = d_o;
            DGs[b * 16 + u] = d_i; DGs[b * 16 + 4 + u] = d_f; DGs[b * 16 + 8 + u] = d_g; DGs[b * 16 + 12 + u] = d_o;
        }
        __syncthreads();
        float* pb = a.pbuf + ((long long)par * G + blockIdx.x) * B * H;
        for (int bb = bg; bb < B; bb += nbg) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) {
                const float4 d = *reinterpret_cast<const float4*>(DGs + bb * 16 + 4 * r4);
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 w = Wr[4 * r4 + k];
                    o.x = fmaf(dv[k], w.x, o.x); o.y = fmaf(dv[k], w.y, o.y); o.z = fmaf(dv[k], w.z, o.z); o.w = fmaf(dv[k], w.w, o.w);
                }
            }
            __stcg(reinterpret_cast<float4*>(pb + (long long)bb * H + 4 * j4), o);
        }
        grid.sync();
        // this block's 4 columns of dh_{t-1}: sum over every block's partial
        {
            const int gq = threadIdx.x >> 6;                  // 4 slices of the block range
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < B) {
                const float* src = a.pbuf + (long long)par * G * B * H + (long long)b * H + u0;
#pragma unroll 8
                for (int g = gq; g < G; g += 4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (long long)g * B * H));
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
            }
            *reinterpret_cast<float4*>(RED + (gq * kMaxB + b) * 4) = s;
            __syncthreads();
            if (b < B) dh_rec = RED[(0 * kMaxB + b) * 4 + u] + RED[(1 * kMaxB + b) * 4 + u] + RED[(2 * kMaxB + b) * 4 + u] + RED[(3 * kMaxB + b) * 4 + u];
        }
        par ^= 1;
    }
    if (b < B) {
        a.dh0[(long long)b * H + unit] = dh_rec;
        a.dc0[(long long)b * H + unit] = dc_rec;
    }
}

// ---- v2 (H >= 64): block = (16 hidden units, 16 batch rows) -----------------------------------------------------------------
// v1 gives every block all of h_{t-1} (B x H) each step: 128 blocks x 128 KB = 16.8 MB through L2 per step, and its inner loop
// is bound by shared-memory wavefronts (one h float4 + four W float4 per 16 FMA).  v2 splits the batch over the grid as well:
// a block keeps the 64 gate rows of its 16 units resident (132 KB) and needs only its own 16 rows of h (4.2 MB per step over the
// grid).  The 8 warps split K; a thread accumulates a 4 gates x 8 rows register tile (12 LDS.128 per 128 FMA: FMA-bound), each
// warp stages just the K slice of h it reads (cp.async + __syncwarp, no block barrier), and the K-slice partials meet in shared
// memory laid out so that the thread that reduces a (row, unit) cell also owns its c_{t-1}.
constexpr int kUG = 16, kBB = 16;

__device__ __forceinline__ void fma4(float& acc, const float4& h, const float4& w) {
    acc = fmaf(h.x, w.x, fmaf(h.y, w.y, fmaf(h.z, w.z, fmaf(h.w, w.w, acc))));
}

__device__ __forceinline__ void stage_w64(float* Ws, const float* __restrict__ w_hh, int H, int u0) {
    const int H4 = H / 4, HP = H + 4;
    for (int i = threadIdx.x; i < 64 * H4; i += kT) {
        const int rr = i / H4, k4 = i - rr * H4, q = rr >> 4, u = rr & 15;       // local row q*16+u  <-  W_hh row q*H + u0 + u
        *reinterpret_cast<float4*>(Ws + rr * HP + 4 * k4) = *reinterpret_cast<const float4*>(w_hh + (long long)(q * H + u0 + u) * H + 4 * k4);
    }
}

__global__ void __launch_bounds__(kT, 1) lstm_fwd2_kernel(FwdArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float sm[];
    const int H = a.H, B = a.B, HP = H + 4;
    float* Ws = sm;                      // [64][HP]
    float* Hs = Ws + 64 * HP;            // [16][HP]   h_{t-1} of the block's batch rows
    float* Red = Hs + 16 * HP;           // [8 warps][32 tile elements][32 lanes]
    const int u0 = blockIdx.x * kUG, b0 = blockIdx.y * kBB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    stage_w64(Ws, a.w_hh, H, u0);
    for (int i = threadIdx.x; i < 16 * HP; i += kT) Hs[i] = 0.f;       // rows beyond the batch stay zero
    __syncthreads();
    const int KS = H / 8, KS4 = KS / 4, k0 = warp * KS;                 // this warp's K slice
    const int u = lane & 15, bsel = lane >> 4;                          // tile: gates q = 0..3 of unit u, rows 2i + bsel
    const int cb = b0 + 2 * warp + bsel, unit = u0 + u;                 // the cell this thread finishes (i = warp)
    const bool live = cb < B;
    float cprev = live ? a.c0[(long long)cb * H + unit] : 0.f;
    for (int t = 0; t < a.T; ++t) {
        const float* hprev = t ? a.hs + (long long)(t - 1) * B * H : a.h0;
        for (int i = lane; i < 16 * KS4; i += 32) {
            const int bb = i / KS4, k4 = i - bb * KS4;
            if (b0 + bb < B) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(Hs + bb * HP + k0 + 4 * k4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(hprev + (long long)(b0 + bb) * H + k0 + 4 * k4) : "memory");
            }
        }
        float gxv[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
            const float* g = a.gx + ((long long)t * B + cb) * 4 * H;
            gxv[0] = g[unit]; gxv[1] = g[H + unit]; gxv[2] = g[2 * H + unit]; gxv[3] = g[3 * H + unit];
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        float acc[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
        const float* wr = Ws + u * HP + k0;
        const float* hr = Hs + bsel * HP + k0;
#pragma unroll 2
        for (int k4 = 0; k4 < KS4; ++k4) {
            float4 w4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w4[q] = *reinterpret_cast<const float4*>(wr + q * 16 * HP + 4 * k4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 h4 = *reinterpret_cast<const float4*>(hr + 2 * i * HP + 4 * k4);
#pragma unroll
                for (int q = 0; q < 4; ++q) fma4(acc[q][i], h4, w4[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) Red[(warp * 32 + q * 8 + i) * 32 + lane] = acc[q][i];
        __syncthreads();
        float pre[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += Red[(w * 32 + q * 8 + warp) * 32 + lane];
            pre[q] = s + gxv[q];
        }
        if (live) {
            const long long row = (long long)t * B + cb;
            const float gi = sigm(pre[0]), gf = sigm(pre[1]), gg = tanhf(pre[2]), go = sigm(pre[3]);
            const float c = fmaf(gf, cprev, gi * gg);
            float* ac = a.act + row * 4 * H;
            ac[unit] = gi; ac[H + unit] = gf; ac[2 * H + unit] = gg; ac[3 * H + unit] = go;
            a.cs[row * H + unit] = c;
            a.hs[row * H + unit] = go * tanhf(c);
            cprev = c;
        }
        grid.sync();
    }
}

// backward: the block computes dgates of its (16 rows x 16 units) cells, then its partial of dh_{t-1} over its 64 gate rows for
// its 16 batch rows and ALL H columns (register tile: BPT rows x 4 columns, W rows resident in shared memory, dgates broadcast),
// publishes it (pbuf[par][unit group][b][:], 4.2 MB per step over the grid; v1: 16.8 MB), and after the barrier each cell sums
// its column over the H/16 unit groups.
template <int BPT>
__global__ void __launch_bounds__(kT, 1) lstm_bwd2_kernel(BwdArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float sm[];
    const int H = a.H, B = a.B, HP = H + 4, H4 = H / 4, G = gridDim.x;
    float* Ws = sm;                      // [64][HP]
    float* Dg = Ws + 64 * HP;            // [16][64]  dgates of the block's cells, [row][q*16+u]
    const int u0 = blockIdx.x * kUG, b0 = blockIdx.y * kBB;
    stage_w64(Ws, a.w_hh, H, u0);
    const int cbl = threadIdx.x >> 4, u = threadIdx.x & 15, cb = b0 + cbl, unit = u0 + u;
    const bool live = cb < B;
    const int k4 = threadIdx.x % H4, bg = threadIdx.x / H4;             // BPT * (256 / H4) == 16
    float dh_rec = 0.f, dc_rec = 0.f;
    if (live) {
        if (a.dhT) dh_rec = a.dhT[(long long)cb * H + unit];
        if (a.dcT) dc_rec = a.dcT[(long long)cb * H + unit];
    }
    int par = 0;
    for (int t = a.T - 1; t >= 0; --t) {
        float d_i = 0.f, d_f = 0.f, d_g = 0.f, d_o = 0.f;
        if (live) {
            const long long row = (long long)t * B + cb;
            const float dh = dh_rec + (a.dhs ? a.dhs[row * H + unit] : 0.f);
            const float* ac = a.act + row * 4 * H;
            const float gi = ac[unit], gf = ac[H + unit], gg = ac[2 * H + unit], go = ac[3 * H + unit];
            const float c = a.cs[row * H + unit];
            const float cprev = t ? a.cs[(row - B) * H + unit] : a.c0[(long long)cb * H + unit];
            const float tc = tanhf(c);
            const float dc = fmaf(dh * go, 1.f - tc * tc, dc_rec);
            d_o = dh * tc * go * (1.f - go);
            d_i = dc * gg * gi * (1.f - gi);
            d_f = dc * cprev * gf * (1.f - gf);
            d_g = dc * gi * (1.f - gg * gg);
            dc_rec = dc * gf;
            float* dg = a.dgates + row * 4 * H;
            dg[unit] = d_i; dg[H + unit] = d_f; dg[2 * H + unit] = d_g; dg[3 * H + unit] = d_o;
        }
        __syncthreads();                 // previous step's readers of Dg are done
        Dg[cbl * 64 + u] = d_i; Dg[cbl * 64 + 16 + u] = d_f; Dg[cbl * 64 + 32 + u] = d_g; Dg[cbl * 64 + 48 + u] = d_o;
        __syncthreads();
        float4 o[BPT];
#pragma unroll
        for (int i = 0; i < BPT; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int r4 = 0; r4 < 16; ++r4) {
            float4 d[BPT];
#pragma unroll
            for (int i = 0; i < BPT; ++i) d[i] = *reinterpret_cast<const float4*>(Dg + (bg * BPT + i) * 64 + 4 * r4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = *reinterpret_cast<const float4*>(Ws + (4 * r4 + j) * HP + 4 * k4);
#pragma unroll
                for (int i = 0; i < BPT; ++i) {
                    const float dv = j == 0 ? d[i].x : j == 1 ? d[i].y : j == 2 ? d[i].z : d[i].w;
                    o[i].x = fmaf(dv, w.x, o[i].x); o[i].y = fmaf(dv, w.y, o[i].y);
                    o[i].z = fmaf(dv, w.z, o[i].z); o[i].w = fmaf(dv, w.w, o[i].w);
                }
            }
        }
        float* pb = a.pbuf + ((long long)par * G + blockIdx.x) * B * H;
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int b = b0 + bg * BPT + i;
            if (b < B) __stcg(reinterpret_cast<float4*>(pb + (long long)b * H + 4 * k4), o[i]);
        }
        grid.sync();
        if (live) {
            const float* src = a.pbuf + (long long)par * G * B * H + (long long)cb * H + unit;
            float s = 0.f;
#pragma unroll 8
            for (int g = 0; g < G; ++g) s += __ldcg(src + (long long)g * B * H);
            dh_rec = s;
        }
        par ^= 1;
    }
    if (live) {
        a.dh0[(long long)cb * H + unit] = dh_rec;
        a.dc0[(long long)cb * H + unit] = dc_rec;
    }
}

static bool use_v2(int H) {
    static const bool off = getenv("PCD_LSTM_V1") != nullptr;
    return !off && H >= 64;
}
static size_t smem_fwd2(int H) { return ((size_t)80 * (H + 4) + 8 * 32 * 32) * sizeof(float); }
static size_t smem_bwd2(int H) { return ((size_t)64 * (H + 4) + 16 * 64) * sizeof(float); }

static int coop_launch(const void* fn, dim3 grid, size_t smem, void** args, cudaStream_t st, const char* what) {
    LaunchState& L = launch_state();
    cudaError_t e = cudaLaunchCooperativeKernel(fn, grid, dim3(kT), args, smem, st);
    count_launch(L);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(L.last_err, sizeof L.last_err, "cooperative launch %s: %s", what, cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}

static bool shape_ok(int T, int B, int H) {
    return T > 0 && B > 0 && B <= kMaxB && H >= 16 && H <= 512 && (H & (H - 1)) == 0;      // H/4 blocks (<= 128), H/4 divides 256
}

}  // namespace lstm
}  // namespace pcd

extern "C" {

size_t pcd_lstm_pbuf_floats(int B, int H) { return (size_t)2 * (H / (pcd::lstm::use_v2(H) ? pcd::lstm::kUG : pcd::lstm::kUnits)) * B * H; }

int pcd_lstm_forward(int T, int B, int H, const float* gx, const float* w_hh, const float* h0, const float* c0, float* act,
                     float* cs, float* hs, void* stream) {
    using namespace pcd;
    if (!gx || !w_hh || !h0 || !c0 || !act || !cs || !hs) return PCD_ERR_ARG;
    if (!lstm::shape_ok(T, B, H)) return PCD_ERR_UNSUPPORTED;
    if ((((uintptr_t)w_hh) | ((uintptr_t)h0) | ((uintptr_t)hs)) & 15) return PCD_ERR_ALIGN;
    lstm::FwdArgs a = {T, B, H, gx, w_hh, h0, c0, act, cs, hs};
    void* args[] = {&a};
    if (lstm::use_v2(H)) {
        static bool configured2 = false;
        if (!configured2) {
            if (cudaFuncSetAttribute(lstm::lstm_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return PCD_ERR_CUDA;
            configured2 = true;
        }
        return lstm::coop_launch((const void*)lstm::lstm_fwd2_kernel, dim3(H / lstm::kUG, (B + lstm::kBB - 1) / lstm::kBB),
                                 lstm::smem_fwd2(H), args, (cudaStream_t)stream, "lstm_fwd2");
    }
    const size_t smem = (size_t)(16 + B) * (H + 4) * sizeof(float);
    if (smem > 227 * 1024) return PCD_ERR_UNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(lstm::lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return PCD_ERR_CUDA;
        configured = true;
    }
    return lstm::coop_launch((const void*)lstm::lstm_fwd_kernel, dim3(H / lstm::kUnits), smem, args, (cudaStream_t)stream, "lstm_fwd");
}

int pcd_lstm_backward(int T, int B, int H, const float* dhs, const float* dhT, const float* dcT, const float* act, const float* cs,
                      const float* c0, const float* w_hh, float* dgates, float* dh0, float* dc0, float* pbuf, void* stream) {
    using namespace pcd;
    if (!act || !cs || !c0 || !w_hh || !dgates || !dh0 || !dc0 || !pbuf) return PCD_ERR_ARG;
    if (!lstm::shape_ok(T, B, H)) return PCD_ERR_UNSUPPORTED;
    if ((((uintptr_t)w_hh) | ((uintptr_t)pbuf)) & 15) return PCD_ERR_ALIGN;
    lstm::BwdArgs a = {T, B, H, dhs, dhT, dcT, act, cs, c0, w_hh, dgates, dh0, dc0, pbuf};
    void* args[] = {&a};
    if (lstm::use_v2(H)) {
        const void* fn = H == 512 ? (const void*)lstm::lstm_bwd2_kernel<8> : H == 256 ? (const void*)lstm::lstm_bwd2_kernel<4>
                       : H == 128 ? (const void*)lstm::lstm_bwd2_kernel<2> : (const void*)lstm::lstm_bwd2_kernel<1>;
        static bool configured2[4] = {false, false, false, false};
        const int slot = H == 512 ? 0 : H == 256 ? 1 : H == 128 ? 2 : 3;
        if (!configured2[slot]) {
            if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return PCD_ERR_CUDA;
            configured2[slot] = true;
        }
        return lstm::coop_launch(fn, dim3(H / lstm::kUG, (B + lstm::kBB - 1) / lstm::kBB), lstm::smem_bwd2(H), args,
                                 (cudaStream_t)stream, "lstm_bwd2");
    }
    return lstm::coop_launch((const void*)lstm::lstm_bwd_kernel, dim3(H / lstm::kUnits), 0, args, (cudaStream_t)stream, "lstm_bwd");
}

}  // extern "C"

#else   // ---- CPU emulation build (tests only) ----------------------------------------------------------------------

#include <math.h>
#include <vector>

static inline float sigm_(float x) { return 1.f / (1.f + expf(-x)); }

extern "C" {

size_t pcd_lstm_pbuf_floats(int B, int H) { return (size_t)2 * (H / 4) * B * H; }

int pcd_lstm_forward(int T, int B, int H, const float* gx, const float* w_hh, const float* h0, const float* c0, float* act,
                     float* cs, float* hs, void*) {
    for (int t = 0; t < T; ++t)
        for (int b = 0; b < B; ++b) {
            const float* hp = t ? hs + ((long long)(t - 1) * B + b) * H : h0 + (long long)b * H;
            const long long row = (long long)t * B + b;
            for (int j = 0; j < H; ++j) {
                float acc[4];
                for (int q = 0; q < 4; ++q) {
                    double s = 0.0;
                    for (int k = 0; k < H; ++k) s += (double)hp[k] * w_hh[(long long)(q * H + j) * H + k];
                    acc[q] = (float)s + gx[row * 4 * H + q * H + j];
                }
                const float gi = sigm_(acc[0]), gf = sigm_(acc[1]), gg = tanhf(acc[2]), go = sigm_(acc[3]);
                const float cprev = t ? cs[(row - B) * H + j] : c0[(long long)b * H + j];
                const float c = gf * cprev + gi * gg;
                act[row * 4 * H + j] = gi; act[row * 4 * H + H + j] = gf; act[row * 4 * H + 2 * H + j] = gg; act[row * 4 * H + 3 * H + j] = go;
                cs[row * H + j] = c;
                hs[row * H + j] = go * tanhf(c);
            }
        }
    return PCD_OK;
}

int pcd_lstm_backward(int T, int B, int H, const float* dhs, const float* dhT, const float* dcT, const float* act, const float* cs,
                      const float* c0, const float* w_hh, float* dgates, float* dh0, float* dc0, float*, void*) {
    std::vector<float> dh((size_t)B * H, 0.f), dc((size_t)B * H, 0.f), nh((size_t)B * H);
    for (long long i = 0; i < (long long)B * H; ++i) { dh[i] = dhT ? dhT[i] : 0.f; dc[i] = dcT ? dcT[i] : 0.f; }
    for (int t = T - 1; t >= 0; --t) {
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < H; ++j) {
                const long long row = (long long)t * B + b;
                const float d = dh[(size_t)b * H + j] + (dhs ? dhs[row * H + j] : 0.f);
                const float gi = act[row * 4 * H + j], gf = act[row * 4 * H + H + j], gg = act[row * 4 * H + 2 * H + j], go = act[row * 4 * H + 3 * H + j];
                const float c = cs[row * H + j], cprev = t ? cs[(row - B) * H + j] : c0[(long long)b * H + j];
                const float tc = tanhf(c);
                const float dcc = dc[(size_t)b * H + j] + d * go * (1.f - tc * tc);
                dgates[row * 4 * H + j] = dcc * gg * gi * (1.f - gi);
                dgates[row * 4 * H + H + j] = dcc * cprev * gf * (1.f - gf);
                dgates[row * 4 * H + 2 * H + j] = dcc * gi * (1.f - gg * gg);
                dgates[row * 4 * H + 3 * H + j] = d * tc * go * (1.f - go);
                dc[(size_t)b * H + j] = dcc * gf;
            }
        for (int b = 0; b < B; ++b)
            for (int k = 0; k < H; ++k) {
                double s = 0.0;
                const long long row = (long long)t * B + b;
                for (int r = 0; r < 4 * H; ++r) s += (double)dgates[row * 4 * H + r] * w_hh[(long long)r * H + k];
                nh[(size_t)b * H + k] = (float)s;
            }
        dh = nh;
    }
    for (long long i = 0; i < (long long)B * H; ++i) { dh0[i] = dh[i]; dc0[i] = dc[i]; }
    return PCD_OK;
}

}  // extern "C"
#endif

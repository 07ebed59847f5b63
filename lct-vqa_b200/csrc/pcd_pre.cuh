// pcd_pre.cuh — the Cell preprocess ops as pixel-major 1x1-convolution GEMM kernels (v3):
//   ReLUConvBN(C_in, C_out, 1, 1, 0, affine=False)   operations.py:22-33
//   FactorizedReduce(C_in, C_out, affine=False)      operations.py:90-104
//
// forward : y[co][p] = sum_ci W[co][ci] * relu(x[ci][p])  (+ per-channel sum / sum^2).
//           Block = 256 output pixels x (4*CPT) output channels; thread = 4 pixels x CPT channels.  The input
//           is streamed in chunks of 16 channels through a double-buffered shared-memory tile: the next chunk is
//           fetched into registers while the current one is multiplied (one barrier per chunk).
// backward: dz = BN-backward(dy, y) kept resident in shared memory for the block's 256 pixels, then per chunk of
//           input channels  dx[ci][p] = relu'(x) * sum_co W[co][ci] dz[co][p]   (thread = 4 pixels x CIT channels)
//           and             dW[co][ci] += sum_p dz[co][p] relu(x[ci][p])        (register tiles, block reduction).
//           The input-channel chunks are split over blockIdx.z when the layer has few pixels.
//
// Thread-private state that must survive a barrier is declared with PCD_TSTATE so that the CPU emulation
// build (one loop per phase) keeps one copy per emulated thread.
#pragma once
#include "pcd_edge.cuh"

namespace pcd {


struct PreArgs {
    int B, Cin, Cout, Hin, Win, Ho, Wo, fr;   // fr: 1 => FactorizedReduce (Ho = Hin/2)
    float eps, momentum;
    const float* x;      // (B, Cin, Hin, Win) contiguous
    const float* w;      // RCB: [Cout][Cin]; FR: conv_1 [Cout/2][Cin] then conv_2 [Cout/2][Cin]
    float* y;            // (B, Cout, Ho, Wo): conv output, normalised in place by the norm kernel
    double* stats;       // sum[Cout], sumsq[Cout]
    float* running;
    long long* nbt;
};

constexpr int kPrePx = 256;     // output pixels per block
constexpr int kPreKC = 16;      // input channels per forward chunk

PCD_HOSTDEV size_t pre_smem_floats(int cpt) { return (size_t)2 * kPreKC * kPrePx + (size_t)2 * kPreKC * 4 * cpt + 16; }

// 4 output pixels p..p+3 of channel plane `pl` (input geometry); FR samples (2oy+shift, 2ox+shift)
PCD_HD void pre_load4(const float* pl, int p, int HW, int fr, int shift, int Wo, int Win, float (&v)[4]) {
    if (!fr) {
        if (p + 3 < HW && (((uintptr_t)(pl + p)) & 15) == 0) {
            const F4 t = *reinterpret_cast<const F4*>(pl + p);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) v[t] = (p + t < HW) ? pl[p + t] : 0.f;
        }
    } else if (p + 3 < HW && ((Wo | Win) & 3) == 0 && (((uintptr_t)pl) & 15) == 0) {
        const int oy = p / Wo, ox = p - oy * Wo;           // Wo % 4 == 0: the 4 pixels share a row
        const F4* r = reinterpret_cast<const F4*>(pl + (long long)(2 * oy + shift) * Win + 2 * ox);
        const F4 a = r[0], b = r[1];
        if (shift) { v[0] = a.y; v[1] = a.w; v[2] = b.y; v[3] = b.w; }
        else { v[0] = a.x; v[1] = a.z; v[2] = b.x; v[3] = b.z; }
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int pp = p + t;
            if (pp < HW) {
                const int oy = pp / Wo, ox = pp - oy * Wo;
                v[t] = pl[(long long)(2 * oy + shift) * Win + 2 * ox + shift];
            } else {
                v[t] = 0.f;
            }
        }
    }
}

// z selects the block's 4*CPT output channels; with FactorizedReduce they lie in one half (one sampling grid)
template <int CPT>
PCD_HD void pre_conv_body(const PreArgs& a, int bx, int n, int z, float* smem) {
    constexpr int NC = 4 * CPT, NXR = kPreKC * 64 / kThreads, NWR = (kPreKC * NC + kThreads - 1) / kThreads;
    const int Cin = a.Cin, COUT = a.Cout, HW = a.Ho * a.Wo;
    float* XS = smem;                                  // [2][KC][256]
    float* WS = XS + 2 * kPreKC * kPrePx;              // [2][KC][NC]
    const int co_base = z * NC;
    const int shift = (a.fr && co_base >= COUT / 2) ? 1 : 0;
    const int p0 = bx * kPrePx;
    const long long cs = (long long)a.Hin * a.Win;
    const float* xb = a.x + (long long)n * Cin * cs;
    const int nchunks = (Cin + kPreKC - 1) / kPreKC;
    PCD_TSTATE(float, acc, [CPT][4]);
    PCD_TSTATE(float, xr, [NXR][4]);
    PCD_TSTATE(float, wr, [NWR]);

    auto fetch = [&](int task, int kc, float (&xq)[NXR][4], float (&wq)[NWR]) {
#pragma unroll
        for (int j = 0; j < NXR; ++j) {
            const int idx = task + kThreads * j, k = idx >> 6, s = idx & 63;
            const int ci = kc + k, p = p0 + 4 * s;
            if (ci < Cin && p < HW) {
                pre_load4(xb + ci * cs, p, HW, a.fr, shift, a.Wo, a.Win, xq[j]);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) xq[j][t] = 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < NWR; ++j) {
            const int idx = task + kThreads * j, col = idx % NC, k = idx / NC;
            wq[j] = (k < kPreKC && kc + k < Cin) ? a.w[(long long)(co_base + col) * Cin + kc + k] : 0.f;
        }
    };
    auto stash = [&](int task, int buf, const float (&xq)[NXR][4], const float (&wq)[NWR]) {
#pragma unroll
        for (int j = 0; j < NXR; ++j) {
            const int idx = task + kThreads * j;
            F4 o = {relu(xq[j][0]), relu(xq[j][1]), relu(xq[j][2]), relu(xq[j][3])};
            *reinterpret_cast<F4*>(XS + buf * kPreKC * kPrePx + idx * 4) = o;
        }
#pragma unroll
        for (int j = 0; j < NWR; ++j) {
            const int idx = task + kThreads * j;
            if (idx < kPreKC * NC) WS[buf * kPreKC * NC + idx] = wq[j];
        }
    };

    PCD_EACH(task) {
        auto& ac = PCD_TREF(acc, task);
#pragma unroll
        for (int i = 0; i < CPT; ++i)
#pragma unroll
            for (int t = 0; t < 4; ++t) ac[i][t] = 0.f;
        fetch(task, 0, PCD_TREF(xr, task), PCD_TREF(wr, task));
        stash(task, 0, PCD_TREF(xr, task), PCD_TREF(wr, task));
    }
    PCD_SYNC();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int cur = ch & 1;
        const bool more = ch + 1 < nchunks;
        PCD_EACH(task) {
            if (more) fetch(task, (ch + 1) * kPreKC, PCD_TREF(xr, task), PCD_TREF(wr, task));
        }
        PCD_EACH(task) {
            auto& ac = PCD_TREF(acc, task);
            const int strip = task & 63, cog = task >> 6;
            const float* xs = XS + cur * kPreKC * kPrePx + strip * 4;
            const float* ws = WS + cur * kPreKC * NC + cog * CPT;
#pragma unroll
            for (int k = 0; k < kPreKC; ++k) {
                const F4 x4 = *reinterpret_cast<const F4*>(xs + k * kPrePx);
#pragma unroll
                for (int i4 = 0; i4 < CPT / 4; ++i4) {
                    const F4 w = *reinterpret_cast<const F4*>(ws + k * NC + 4 * i4);
                    const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        ac[4 * i4 + q][0] = fmaf(wk[q], x4.x, ac[4 * i4 + q][0]);
                        ac[4 * i4 + q][1] = fmaf(wk[q], x4.y, ac[4 * i4 + q][1]);
                        ac[4 * i4 + q][2] = fmaf(wk[q], x4.z, ac[4 * i4 + q][2]);
                        ac[4 * i4 + q][3] = fmaf(wk[q], x4.w, ac[4 * i4 + q][3]);
                    }
                }
            }
        }
        PCD_EACH(task) {
            if (more) stash(task, cur ^ 1, PCD_TREF(xr, task), PCD_TREF(wr, task));
        }
        PCD_SYNC();
    }
    // ---- epilogue: store, per-channel sums (P / P2 alias the staging buffers) -------------------------------
    float* P = XS;                      // [2*CPT][256]
    float* P2 = WS;                     // 2*CPT*4*8 floats
    PCD_EACH(task) {
        auto& ac = PCD_TREF(acc, task);
        const int strip = task & 63, cog = task >> 6;
        const int p = p0 + 4 * strip;
        if (p < HW) {
            float* yb = a.y + ((long long)n * COUT + co_base + cog * CPT) * HW + p;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                if (p + 3 < HW && (((uintptr_t)(yb + (long long)i * HW)) & 15) == 0) {
                    F4 o = {ac[i][0], ac[i][1], ac[i][2], ac[i][3]};
                    *reinterpret_cast<F4*>(yb + (long long)i * HW) = o;
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (p + t < HW) yb[(long long)i * HW + t] = ac[i][t];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (p + t < HW) { s += ac[i][t]; q = fmaf(ac[i][t], ac[i][t], q); }
            P[(2 * i) * kThreads + task] = s;
            P[(2 * i + 1) * kThreads + task] = q;
        }
    }
    reduce_columns<8>(P, P2, 2 * CPT, 4, 64, kThreads, [&](int cog, int k, float v) {
        pcd_atomic_add(a.stats + (k & 1) * COUT + co_base + cog * CPT + (k >> 1), (double)v);
    });
}

struct PreBwdArgs {
    int B, Cin, Cout, Hin, Win, Ho, Wo, fr;
    int chunks_per_block;  // input-channel chunks handled by one block (blockIdx.z covers the rest)
    const float* x;        // cell input (B, Cin, Hin, Win)
    const float* w;
    const float* y;        // normalised preprocess output (B, Cout, Ho, Wo)
    const float* dy;       // its grad
    const double* stats;   // forward sums (for rstd)
    const double* bstats;  // sum dy, sum dy*y
    float eps;
    float* dx;             // (B, Cin, Hin, Win) written; may be null
    float* gw;             // [Cout][Cin] accumulated (atomics); may be null
};

constexpr int kPrePitch = kPrePx + 4;                    // shared-memory row pitch (bank spread for the dW tiles)
PCD_HOSTDEV int pre_bwd_kci(int fr) { return fr ? 16 : 32; }
PCD_HOSTDEV size_t pre_bwd_smem_floats(int Cout, int fr) {
    const int kci = pre_bwd_kci(fr);
    return (size_t)Cout * kPrePitch + (size_t)Cout * kci + (size_t)(fr ? 2 : 1) * kci * kPrePitch + 3 * Cout + 16;
}

template <int COUT, bool FR>
PCD_HD void pre_bwd_body(const PreBwdArgs& a, int bx, int n, int z, float* smem) {
    constexpr int KCI = FR ? 16 : 32, CIT = KCI / 4, HALF = COUT / 2;
    constexpr int COT = COUT >= 32 ? 8 : 4;                         // dW register tile: COT x 4
    constexpr int NOG = (COUT / COT) * (KCI / 4), NSL = kThreads / NOG;
    static_assert(NSL >= 1 && 64 % NSL == 0, "dW slices");
    const int Cin = a.Cin, HW = a.Ho * a.Wo;
    const long long cs = (long long)a.Hin * a.Win;
    float* DZ = smem;                                   // [COUT][pitch]
    float* WT = DZ + COUT * kPrePitch;                  // [COUT][KCI]
    float* R = WT + COUT * KCI;                         // [FR ? 2 : 1][KCI][pitch]   relu(x) samples
    float* COEF = R + (FR ? 2 : 1) * KCI * kPrePitch;
    float* P = R;                                       // [COT*4][256] (after the dW tiles are done with R)
    const double cnt = (double)a.B * HW;
    PCD_FOR(co, COUT) {
        BnC b = bn_consts(a.stats, COUT, 0, co, cnt, a.eps);
        COEF[3 * co] = b.rstd;
        COEF[3 * co + 1] = (float)(a.bstats[co] / cnt);
        COEF[3 * co + 2] = (float)(a.bstats[COUT + co] / cnt);
    }
    PCD_SYNC();
    const int p0 = bx * kPrePx;
    const float* xb = a.x + (long long)n * Cin * cs;
    float* dxb = a.dx ? a.dx + (long long)n * Cin * cs : nullptr;
    // ---- dz tile -------------------------------------------------------------------------------------
#pragma unroll 4
    for (int it = 0; it < COUT / 4; ++it) {
        PCD_EACH(task) {
            const int co = it * 4 + (task >> 6), s = task & 63;
            const int p = p0 + 4 * s;
            float dy[4] = {0.f, 0.f, 0.f, 0.f}, yy[4] = {0.f, 0.f, 0.f, 0.f}, dz[4];
            const long long o = ((long long)n * COUT + co) * HW;
            if (p < HW) {
                pre_load4(a.dy + o, p, HW, 0, 0, 0, 0, dy);
                pre_load4(a.y + o, p, HW, 0, 0, 0, 0, yy);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
                dz[t] = (p + t < HW) ? COEF[3 * co] * (dy[t] - COEF[3 * co + 1] - yy[t] * COEF[3 * co + 2]) : 0.f;
            F4 v = {dz[0], dz[1], dz[2], dz[3]};
            *reinterpret_cast<F4*>(DZ + co * kPrePitch + 4 * s) = v;
        }
    }
    const int nchunks = (Cin + KCI - 1) / KCI;
    const int ch0 = z * a.chunks_per_block;
    const int ch1 = (ch0 + a.chunks_per_block < nchunks) ? ch0 + a.chunks_per_block : nchunks;
    for (int ch = ch0; ch < ch1; ++ch) {
        const int kc = ch * KCI;
        PCD_SYNC();                                   // DZ ready / previous chunk's reduction done with P (= R)
        PCD_FOR(i, COUT * KCI) {
            const int co = i / KCI, j = i - co * KCI;
            WT[i] = (kc + j < Cin) ? a.w[(long long)co * Cin + kc + j] : 0.f;
        }
        PCD_SYNC();
        // ---- dx (and the relu(x) tile for dW) ----------------------------------------------------------
        PCD_EACH(task) {
            const int strip = task & 63, cig = task >> 6;
            const int p = p0 + 4 * strip, ci0 = kc + cig * CIT;
            if (!FR) {
                float xv[CIT][4], acc[CIT][4];
#pragma unroll
                for (int i = 0; i < CIT; ++i) {
                    if (ci0 + i < Cin && p < HW) {
                        pre_load4(xb + (ci0 + i) * cs, p, HW, 0, 0, 0, 0, xv[i]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t) xv[i][t] = 0.f;
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[i][t] = 0.f;
                }
                if (dxb) {
#pragma unroll 4
                    for (int co = 0; co < COUT; ++co) {
                        const F4 d = *reinterpret_cast<const F4*>(DZ + co * kPrePitch + 4 * strip);
                        const float* wr = WT + co * KCI + cig * CIT;
#pragma unroll
                        for (int i4 = 0; i4 < CIT / 4; ++i4) {
                            const F4 w = *reinterpret_cast<const F4*>(wr + 4 * i4);
                            const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                acc[4 * i4 + q][0] = fmaf(wk[q], d.x, acc[4 * i4 + q][0]);
                                acc[4 * i4 + q][1] = fmaf(wk[q], d.y, acc[4 * i4 + q][1]);
                                acc[4 * i4 + q][2] = fmaf(wk[q], d.z, acc[4 * i4 + q][2]);
                                acc[4 * i4 + q][3] = fmaf(wk[q], d.w, acc[4 * i4 + q][3]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < CIT; ++i) {
                    if (dxb && ci0 + i < Cin && p < HW) {
                        float* d = dxb + (ci0 + i) * cs + p;
                        if (p + 3 < HW && (((uintptr_t)d) & 15) == 0) {
                            F4 o = {xv[i][0] > 0.f ? acc[i][0] : 0.f, xv[i][1] > 0.f ? acc[i][1] : 0.f,
                                    xv[i][2] > 0.f ? acc[i][2] : 0.f, xv[i][3] > 0.f ? acc[i][3] : 0.f};
                            *reinterpret_cast<F4*>(d) = o;
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                if (p + t < HW) d[t] = xv[i][t] > 0.f ? acc[i][t] : 0.f;
                        }
                    }
                    if (a.gw) {
                        F4 r = {relu(xv[i][0]), relu(xv[i][1]), relu(xv[i][2]), relu(xv[i][3])};
                        *reinterpret_cast<F4*>(R + (cig * CIT + i) * kPrePitch + 4 * strip) = r;
                    }
                }
            } else {
                float s0[CIT][4], s1[CIT][4];
#pragma unroll
                for (int i = 0; i < CIT; ++i)
#pragma unroll
                    for (int t = 0; t < 4; ++t) { s0[i][t] = 0.f; s1[i][t] = 0.f; }
                if (dxb) {
#pragma unroll 4
                    for (int co = 0; co < HALF; ++co) {
                        const F4 d = *reinterpret_cast<const F4*>(DZ + co * kPrePitch + 4 * strip);
                        const F4 e = *reinterpret_cast<const F4*>(DZ + (co + HALF) * kPrePitch + 4 * strip);
                        const F4 w0 = *reinterpret_cast<const F4*>(WT + co * KCI + cig * CIT);
                        const F4 w1 = *reinterpret_cast<const F4*>(WT + (co + HALF) * KCI + cig * CIT);
                        const float wa[4] = {w0.x, w0.y, w0.z, w0.w}, wb[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            s0[q][0] = fmaf(wa[q], d.x, s0[q][0]); s0[q][1] = fmaf(wa[q], d.y, s0[q][1]);
                            s0[q][2] = fmaf(wa[q], d.z, s0[q][2]); s0[q][3] = fmaf(wa[q], d.w, s0[q][3]);
                            s1[q][0] = fmaf(wb[q], e.x, s1[q][0]); s1[q][1] = fmaf(wb[q], e.y, s1[q][1]);
                            s1[q][2] = fmaf(wb[q], e.z, s1[q][2]); s1[q][3] = fmaf(wb[q], e.w, s1[q][3]);
                        }
                    }
                }
                const bool fast = p + 3 < HW && ((a.Wo | a.Win) & 3) == 0 && (((uintptr_t)xb) & 15) == 0 &&
                                  (!dxb || (((uintptr_t)dxb) & 15) == 0);
#pragma unroll
                for (int i = 0; i < CIT; ++i) {
                    const int ci = ci0 + i;
                    float r0[4] = {0.f, 0.f, 0.f, 0.f}, r1[4] = {0.f, 0.f, 0.f, 0.f};
                    if (ci < Cin && p < HW) {
                        if (fast) {
                            const int oy = p / a.Wo, ox = p - oy * a.Wo;
                            const long long o0 = ci * cs + (long long)(2 * oy) * a.Win + 2 * ox;
                            const F4* x0 = reinterpret_cast<const F4*>(xb + o0);
                            const F4* x1 = reinterpret_cast<const F4*>(xb + o0 + a.Win);
                            const F4 a0 = x0[0], b0 = x0[1], a1 = x1[0], b1 = x1[1];
                            r0[0] = a0.x; r0[1] = a0.z; r0[2] = b0.x; r0[3] = b0.z;
                            r1[0] = a1.y; r1[1] = a1.w; r1[2] = b1.y; r1[3] = b1.w;
                            if (dxb) {
                                F4* d0 = reinterpret_cast<F4*>(dxb + o0);
                                F4* d1 = reinterpret_cast<F4*>(dxb + o0 + a.Win);
                                F4 u0 = {r0[0] > 0.f ? s0[i][0] : 0.f, 0.f, r0[1] > 0.f ? s0[i][1] : 0.f, 0.f};
                                F4 u1 = {r0[2] > 0.f ? s0[i][2] : 0.f, 0.f, r0[3] > 0.f ? s0[i][3] : 0.f, 0.f};
                                F4 v0 = {0.f, r1[0] > 0.f ? s1[i][0] : 0.f, 0.f, r1[1] > 0.f ? s1[i][1] : 0.f};
                                F4 v1 = {0.f, r1[2] > 0.f ? s1[i][2] : 0.f, 0.f, r1[3] > 0.f ? s1[i][3] : 0.f};
                                d0[0] = u0; d0[1] = u1; d1[0] = v0; d1[1] = v1;
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const int pp = p + t;
                                if (pp >= HW) continue;
                                const int oy = pp / a.Wo, ox = pp - oy * a.Wo;
                                const long long o0 = ci * cs + (long long)(2 * oy) * a.Win + 2 * ox;
                                r0[t] = xb[o0];
                                r1[t] = xb[o0 + a.Win + 1];
                                if (dxb) {
                                    dxb[o0] = r0[t] > 0.f ? s0[i][t] : 0.f;
                                    dxb[o0 + 1] = 0.f;
                                    dxb[o0 + a.Win] = 0.f;
                                    dxb[o0 + a.Win + 1] = r1[t] > 0.f ? s1[i][t] : 0.f;
                                }
                            }
                        }
                    }
                    if (a.gw) {
                        F4 q0 = {relu(r0[0]), relu(r0[1]), relu(r0[2]), relu(r0[3])};
                        F4 q1 = {relu(r1[0]), relu(r1[1]), relu(r1[2]), relu(r1[3])};
                        *reinterpret_cast<F4*>(R + (cig * CIT + i) * kPrePitch + 4 * strip) = q0;
                        *reinterpret_cast<F4*>(R + (KCI + cig * CIT + i) * kPrePitch + 4 * strip) = q1;
                    }
                }
            }
        }
        if (!a.gw) continue;
        PCD_SYNC();
        // ---- dW tile: COT x 4 outputs per task, strips interleaved over the NSL slices --------------------
        PCD_TSTATE(float, dwacc, [COT][4]);
        PCD_EACH(task) {
            auto& acc = PCD_TREF(dwacc, task);
            const int og = task / NSL, sl = task - og * NSL;
            const int co0 = (og / (KCI / 4)) * COT, cil = (og % (KCI / 4)) * 4;
            const float* Rr = R + ((FR && co0 >= HALF) ? KCI * kPrePitch : 0) + cil * kPrePitch;
            const float* Dr = DZ + co0 * kPrePitch;
#pragma unroll
            for (int i = 0; i < COT; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
#pragma unroll 2
            for (int st = sl; st < 64; st += NSL) {
                float rv[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const F4 r = *reinterpret_cast<const F4*>(Rr + k * kPrePitch + st * 4);
                    rv[k][0] = r.x; rv[k][1] = r.y; rv[k][2] = r.z; rv[k][3] = r.w;
                }
#pragma unroll
                for (int i = 0; i < COT; ++i) {
                    const F4 d = *reinterpret_cast<const F4*>(Dr + i * kPrePitch + st * 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[i][k] = fmaf(d.x, rv[k][0], acc[i][k]);
                        acc[i][k] = fmaf(d.y, rv[k][1], acc[i][k]);
                        acc[i][k] = fmaf(d.z, rv[k][2], acc[i][k]);
                        acc[i][k] = fmaf(d.w, rv[k][3], acc[i][k]);
                    }
                }
            }
        }
        PCD_SYNC();                                   // everyone is done reading R: reuse it as P
        PCD_EACH(task) {
            auto& acc = PCD_TREF(dwacc, task);
#pragma unroll
            for (int i = 0; i < COT; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) P[(i * 4 + k) * kThreads + task] = acc[i][k];
        }
        PCD_SYNC();
        PCD_FOR(kg, COT * 4 * NOG) {
            const int k = kg / NOG, og = kg - k * NOG;
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < NSL; ++t) s += P[k * kThreads + og * NSL + t];
            const int co = (og / (KCI / 4)) * COT + (k >> 2), ci = kc + (og % (KCI / 4)) * 4 + (k & 3);
            if (ci < Cin) pcd_atomic_add(a.gw + (long long)co * Cin + ci, s);
        }
    }
}

}  // namespace pcd

// pcd_opk.cuh — the candidate operations of operations.py as STAND-ALONE ops on all C channels of a tensor: what a network
// derived from a searched genotype runs (the reference stops at genotype(), model_search.py:218-263), and
// what `OPS[name](C, stride, affine)` modules execute when they are called on their own.
//
//   ReLU -> depthwise KxK (stride, dilation) -> 1x1 -> BatchNorm          operations.py:35-47 (DilConv), :50-66 (SepConv = two)
//   max_pool_3x3 / avg_pool_3x3 (pad 1, count_include_pad=False)          operations.py:6-7
//
// One "unit" is split where its data dependencies change shape:
//   dw_fwd   per (image, channel) plane : t = dw(relu(x))                               stencil, no channel mixing
//   pw_fwd   per (image, 128 pixels)    : z = W t  + per-channel sum / sum^2 (fp64)     channel mixing, no stencil
//   Norm     (pcd_fwd.cuh)              : y = gamma * (z - mean) * rstd + beta, running statistics
//   BnBwdStats (pcd_bwd.cuh)            : sum dy, sum dy * yhat
//   pw_bwd   per (image, 128 pixels)    : dz = BN backward on load; dt = W^T dz; dW += dz t^T
//   dw_bwd   per (image, channel) plane : dx = relu'(x) * dw^T(dt); dw weights' gradient
// These are first, straightforward versions (whole planes in shared memory, H, W <= 64): parity first.
#pragma once
#include "pcd_common.cuh"

namespace pcd {

constexpr int kOpMaxHW = 64;       // plane side the whole-plane kernels take
constexpr int kPwPx = 128;         // pixels per block of the pointwise kernels
constexpr int kPwPitch = kPwPx + 4; // row pitch of the [channel][pixel] tiles: lanes that walk channels hit different banks

struct DwArgs {
    int B, C, Hi, Wi, Ho, Wo, S, PAD, DIL, relu;
    const float* x;        // (B, C, Hi, Wi)
    const float* w;        // (C, KS*KS)
    float* t;              // fwd: (B, C, Ho, Wo) written
    const float* dt;       // bwd
    float* dx;             // bwd, may be null
    float* gw;             // bwd, accumulated (caller zeroes), may be null
};

PCD_HOSTDEV size_t dw_fwd_smem_floats(int Hi, int Wi, int PAD) { return (size_t)(Hi + 2 * PAD) * (Wi + 2 * PAD); }
PCD_HOSTDEV size_t dw_bwd_smem_floats(int Hi, int Wi, int Ho, int Wo, int PAD, int KS) {
    return (size_t)(Hi + 2 * PAD) * (Wi + 2 * PAD) + (size_t)Ho * Wo + KS * KS;
}

// zero-padded plane of f(x): padded row r / column q hold image row r - PAD / column q - PAD
PCD_HD void stage_plane(float* P, const float* x, int Hi, int Wi, int PAD, bool relu_in) {
    const int PH = Hi + 2 * PAD, PW = Wi + 2 * PAD;
    PCD_FOR(i, PH * PW) {
        const int r = i / PW, q = i - r * PW, iy = r - PAD, ix = q - PAD;
        float v = 0.f;
        if (iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) {
            v = x[iy * Wi + ix];
            if (relu_in) v = relu(v);
        }
        P[i] = v;
    }
}

template <int KS>
PCD_HD void dw_fwd_body(const DwArgs& a, int c, int n, float* smem) {
    float* P = smem;
    const int PW = a.Wi + 2 * a.PAD;
    const long long plane = (long long)n * a.C + c;
    stage_plane(P, a.x + plane * a.Hi * a.Wi, a.Hi, a.Wi, a.PAD, a.relu != 0);
    PCD_SYNC();
    const float* w = a.w + c * KS * KS;
    float* t = a.t + plane * a.Ho * a.Wo;
    PCD_FOR(o, a.Ho * a.Wo) {
        const int oy = o / a.Wo, ox = o - oy * a.Wo;
        const float* p = P + (oy * a.S) * PW + ox * a.S;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) acc = fmaf(w[ky * KS + kx], p[ky * a.DIL * PW + kx * a.DIL], acc);
        t[o] = acc;
    }
}

template <int KS>
PCD_HD void dw_bwd_body(const DwArgs& a, int c, int n, float* smem) {
    const int PH = a.Hi + 2 * a.PAD, PW = a.Wi + 2 * a.PAD, NO = a.Ho * a.Wo;
    float* P = smem;                 // f(x), zero padded
    float* DT = P + PH * PW;         // dt plane
    float* GW = DT + NO;             // [KS*KS]
    const long long plane = (long long)n * a.C + c;
    stage_plane(P, a.x + plane * a.Hi * a.Wi, a.Hi, a.Wi, a.PAD, a.relu != 0);
    const float* dt = a.dt + plane * NO;
    PCD_FOR(o, NO) DT[o] = dt[o];
    PCD_FOR(k, KS * KS) GW[k] = 0.f;
    PCD_SYNC();
    if (a.gw) {
        // dW[ky][kx] = sum_o dt[o] * f(x)[o*S - PAD + k*DIL]: every thread sums its share of the outputs for all taps
        PCD_FOR(task, kThreads) {
            float acc[KS * KS];
#pragma unroll
            for (int k = 0; k < KS * KS; ++k) acc[k] = 0.f;
            for (int o = task; o < NO; o += kThreads) {
                const int oy = o / a.Wo, ox = o - oy * a.Wo;
                const float* p = P + (oy * a.S) * PW + ox * a.S;
                const float d = DT[o];
#pragma unroll
                for (int ky = 0; ky < KS; ++ky)
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx) acc[ky * KS + kx] = fmaf(d, p[ky * a.DIL * PW + kx * a.DIL], acc[ky * KS + kx]);
            }
#pragma unroll
            for (int k = 0; k < KS * KS; ++k) {
                float v = acc[k];
#if PCD_CUDA
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);     // all kThreads lanes are in this loop
                if ((threadIdx.x & 31) == 0)
#endif
                    pcd_atomic_add(GW + k, v);
            }
        }
    }
    if (a.dx) {
        const float* w = a.w + c * KS * KS;
        float* dx = a.dx + plane * a.Hi * a.Wi;
        PCD_FOR(i, a.Hi * a.Wi) {
            const int iy = i / a.Wi, ix = i - iy * a.Wi;
            float acc = 0.f;
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const int ny = iy + a.PAD - ky * a.DIL;
                if (ny < 0 || ny % a.S) continue;
                const int oy = ny / a.S;
                if (oy >= a.Ho) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const int nx = ix + a.PAD - kx * a.DIL;
                    if (nx < 0 || nx % a.S) continue;
                    const int ox = nx / a.S;
                    if (ox >= a.Wo) continue;
                    acc = fmaf(w[ky * KS + kx], DT[oy * a.Wo + ox], acc);
                }
            }
            if (a.relu && !(P[(iy + a.PAD) * PW + ix + a.PAD] > 0.f)) acc = 0.f;
            dx[i] = acc;
        }
    }
    PCD_SYNC();
    if (a.gw) PCD_FOR(k, KS * KS) pcd_atomic_add(a.gw + c * KS * KS + k, GW[k]);
}

// ---- pointwise (1x1) conv + BatchNorm statistics ---------------------------------------------------------------------------
struct PwArgs {
    int B, Cin, Cout, HW;
    float eps;
    const float* t;        // (B, Cin, HW)
    const float* w;        // (Cout, Cin)
    float* z;              // fwd: (B, Cout, HW) written.  bwd: read
    double* stats;         // fwd: sum[Cout], sumsq[Cout] accumulated (caller zeroes).  bwd: read
    const float* g;        // bwd: gradient w.r.t. the BatchNorm output
    const float* gamma;    // bwd: null => 1
    const double* bstats;  // bwd: sum g [Cout], sum g*yhat [Cout]
    float* dt;             // bwd, may be null
    float* gw;             // bwd, accumulated (caller zeroes), may be null
    int cpb;               // bwd: 128-pixel chunks per block (launcher)
};

PCD_HOSTDEV size_t pw_fwd_smem_floats(int Cin, int Cout) { return (size_t)Cout * Cin + (size_t)Cin * kPwPitch + 2 * Cout; }
PCD_HOSTDEV size_t pw_bwd_smem_floats(int Cin, int Cout) { return 2 * (size_t)Cout * Cin + (size_t)(Cin + Cout) * kPwPitch + 4 * Cout; }

// tile [C][kPwPitch] of a (B, C, HW) tensor, pixels p0 .. p0 + kPwPx - 1 of image n, zero beyond HW
PCD_HD void stage_px_tile(float* T, const float* src, int C, int HW, int n, int p0) {
    PCD_FOR(i, C * (kPwPx / 4)) {
        const int ci = i / (kPwPx / 4), p4 = i - ci * (kPwPx / 4), p = p0 + 4 * p4;
        F4 v = {0.f, 0.f, 0.f, 0.f};
        if (p < HW) v = *reinterpret_cast<const F4*>(src + ((long long)n * C + ci) * HW + p);      // HW % 4 == 0
        *reinterpret_cast<F4*>(T + ci * kPwPitch + 4 * p4) = v;
    }
}

PCD_HD void pw_fwd_body(const PwArgs& a, int bx, int n, float* smem) {
    float* Ws = smem;                          // [Cout][Cin]
    float* T = Ws + a.Cout * a.Cin;            // [Cin][kPwPx]
    float* SACC = T + a.Cin * kPwPitch;        // [2][Cout]
    const int p0 = bx * kPwPx;
    PCD_FOR(i, a.Cout * a.Cin) Ws[i] = a.w[i];
    PCD_FOR(i, 2 * a.Cout) SACC[i] = 0.f;
    stage_px_tile(T, a.t, a.Cin, a.HW, n, p0);
    PCD_SYNC();
    PCD_FOR(task, (a.Cout / 4) * (kPwPx / 4)) {
        const int co4 = task / (kPwPx / 4), p4 = task - co4 * (kPwPx / 4), p = p0 + 4 * p4;
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
        for (int ci = 0; ci < a.Cin; ++ci) {
            const F4 tv = *reinterpret_cast<const F4*>(T + ci * kPwPitch + 4 * p4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float w = Ws[(4 * co4 + j) * a.Cin + ci];
                acc[j][0] = fmaf(w, tv.x, acc[j][0]); acc[j][1] = fmaf(w, tv.y, acc[j][1]);
                acc[j][2] = fmaf(w, tv.z, acc[j][2]); acc[j][3] = fmaf(w, tv.w, acc[j][3]);
            }
        }
        const bool live = p < a.HW;          // (tile columns beyond HW hold zeros: they add nothing to the sums)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = 4 * co4 + j;
            if (live) {
                F4 o = {acc[j][0], acc[j][1], acc[j][2], acc[j][3]};
                *reinterpret_cast<F4*>(a.z + ((long long)n * a.Cout + co) * a.HW + p) = o;
            }
            float s = (acc[j][0] + acc[j][1]) + (acc[j][2] + acc[j][3]);
            float q = fmaf(acc[j][0], acc[j][0], fmaf(acc[j][1], acc[j][1], fmaf(acc[j][2], acc[j][2], acc[j][3] * acc[j][3])));
#if PCD_CUDA
            // kPwPx / 4 == 32: a warp is one co4 and the 32 pixel strips; the task count is a multiple of 32, so warps are whole
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
            if ((threadIdx.x & 31) == 0)
#endif
            {
                pcd_atomic_add(SACC + co, s);
                pcd_atomic_add(SACC + a.Cout + co, q);
            }
        }
    }
    PCD_SYNC();
    PCD_FOR(i, 2 * a.Cout) pcd_atomic_add(a.stats + i, (double)SACC[i]);
}

PCD_HD void pw_bwd_body(const PwArgs& a, int bx, int n, float* smem) {
    float* Ws = smem;                          // [Cout][Cin]
    float* T = Ws + a.Cout * a.Cin;            // [Cin][kPwPitch]
    float* DZ = T + a.Cin * kPwPitch;          // [Cout][kPwPitch]
    float* K = DZ + a.Cout * kPwPitch;         // [4][Cout]: mean, rstd, mean(g), mean(g * yhat)
    float* GW = K + 4 * a.Cout;                // [Cout][Cin] weight-gradient partial of this block's chunks
    const double cnt = (double)a.B * a.HW;
    const int nchunks = (a.HW + kPwPx - 1) / kPwPx;
    PCD_FOR(i, a.Cout * a.Cin) { Ws[i] = a.w[i]; GW[i] = 0.f; }
    PCD_FOR(co, a.Cout) {
        BnC b = bn_consts(a.stats, a.Cout, 0, co, cnt, a.eps);
        K[co] = b.mean;
        K[a.Cout + co] = b.rstd;
        K[2 * a.Cout + co] = (float)(a.bstats[co] / cnt);
        K[3 * a.Cout + co] = (float)(a.bstats[a.Cout + co] / cnt);
    }
    // a block walks `cpb` consecutive 128-pixel chunks of its image: the weight gradient is summed in shared memory across
    // them (every (co, ci) pair belongs to one thread) and leaves with ONE global atomic per pair and block
    for (int k = 0; k < a.cpb; ++k) {
        const int chunk = bx * a.cpb + k;
        if (chunk >= nchunks) break;
        const int p0 = chunk * kPwPx;
        PCD_SYNC();                             // the previous chunk's readers of T / DZ are done
        stage_px_tile(T, a.t, a.Cin, a.HW, n, p0);
        PCD_SYNC();
        // dz = rstd * gamma * (g - mean(g) - yhat * mean(g * yhat)),  yhat = (z - mean) * rstd
        PCD_FOR(i, a.Cout * (kPwPx / 4)) {
            const int co = i / (kPwPx / 4), p4 = i - co * (kPwPx / 4), p = p0 + 4 * p4;
            F4 o = {0.f, 0.f, 0.f, 0.f};
            if (p < a.HW) {
                const long long off = ((long long)n * a.Cout + co) * a.HW + p;
                const F4 g = *reinterpret_cast<const F4*>(a.g + off), z = *reinterpret_cast<const F4*>(a.z + off);
                const float mean = K[co], rstd = K[a.Cout + co], m1 = K[2 * a.Cout + co], m2 = K[3 * a.Cout + co];
                const float kk = rstd * (a.gamma ? a.gamma[co] : 1.f);
                o.x = kk * (g.x - m1 - (z.x - mean) * rstd * m2); o.y = kk * (g.y - m1 - (z.y - mean) * rstd * m2);
                o.z = kk * (g.z - m1 - (z.z - mean) * rstd * m2); o.w = kk * (g.w - m1 - (z.w - mean) * rstd * m2);
            }
            *reinterpret_cast<F4*>(DZ + co * kPwPitch + 4 * p4) = o;
        }
        PCD_SYNC();
        if (a.dt) {
            PCD_FOR(task, (a.Cin / 4) * (kPwPx / 4)) {
                const int ci4 = task / (kPwPx / 4), p4 = task - ci4 * (kPwPx / 4), p = p0 + 4 * p4;
                float acc[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
                for (int co = 0; co < a.Cout; ++co) {
                    const F4 d = *reinterpret_cast<const F4*>(DZ + co * kPwPitch + 4 * p4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float w = Ws[co * a.Cin + 4 * ci4 + j];
                        acc[j][0] = fmaf(w, d.x, acc[j][0]); acc[j][1] = fmaf(w, d.y, acc[j][1]);
                        acc[j][2] = fmaf(w, d.z, acc[j][2]); acc[j][3] = fmaf(w, d.w, acc[j][3]);
                    }
                }
                if (p < a.HW) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        F4 o = {acc[j][0], acc[j][1], acc[j][2], acc[j][3]};
                        *reinterpret_cast<F4*>(a.dt + ((long long)n * a.Cin + 4 * ci4 + j) * a.HW + p) = o;
                    }
                }
            }
        }
        if (a.gw) {
            PCD_FOR(task, a.Cout * a.Cin) {
                const int co = task / a.Cin, ci = task - co * a.Cin;
                const F4* d = reinterpret_cast<const F4*>(DZ + co * kPwPitch);
                const F4* t = reinterpret_cast<const F4*>(T + ci * kPwPitch);
                float s = 0.f;
                for (int q = 0; q < kPwPx / 4; ++q)
                    s = fmaf(d[q].x, t[q].x, fmaf(d[q].y, t[q].y, fmaf(d[q].z, t[q].z, fmaf(d[q].w, t[q].w, s))));
                GW[task] += s;
            }
        }
    }
    PCD_SYNC();
    if (a.gw) PCD_FOR(task, a.Cout * a.Cin) pcd_atomic_add(a.gw + task, GW[task]);
}

// ---- 3x3 pools, padding 1 -------------------------------------------------------------------------------------------------
struct PoolArgs {
    int B, C, Hi, Wi, Ho, Wo, S, is_max;
    const float* x;
    float* y;              // fwd
    const float* dy;       // bwd
    float* dx;             // bwd
};

PCD_HOSTDEV size_t pool_smem_floats(int Hi, int Wi, int Ho, int Wo) { return (size_t)Hi * Wi + 2 * (size_t)Ho * Wo; }

// value and (for max) first-maximum input index of the window of output (oy, ox); avg: sum / number of in-bounds taps
PCD_HD float pool_window(const float* X, int Hi, int Wi, int S, int oy, int ox, bool is_max, int* arg, int* count) {
    float best = -INFINITY, sum = 0.f;
    int bi = -1, cnt = 0;
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * S - 1 + ky;
        if (iy < 0 || iy >= Hi) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * S - 1 + kx;
            if (ix < 0 || ix >= Wi) continue;
            const float v = X[iy * Wi + ix];
            if (v > best || bi < 0) { best = v; bi = iy * Wi + ix; }
            sum += v;
            ++cnt;
        }
    }
    *arg = bi;
    *count = cnt;
    return is_max ? best : sum / (float)cnt;
}

PCD_HD void pool_fwd_body(const PoolArgs& a, int c, int n, float* smem) {
    float* X = smem;
    const long long plane = (long long)n * a.C + c;
    const float* x = a.x + plane * a.Hi * a.Wi;
    PCD_FOR(i, a.Hi * a.Wi) X[i] = x[i];
    PCD_SYNC();
    float* y = a.y + plane * a.Ho * a.Wo;
    PCD_FOR(o, a.Ho * a.Wo) {
        int arg, cnt;
        y[o] = pool_window(X, a.Hi, a.Wi, a.S, o / a.Wo, o % a.Wo, a.is_max != 0, &arg, &cnt);
    }
}

PCD_HD void pool_bwd_body(const PoolArgs& a, int c, int n, float* smem) {
    const int NO = a.Ho * a.Wo;
    float* X = smem;                 // input plane
    float* DY = X + a.Hi * a.Wi;     // dy (avg: already divided by the tap count)
    float* ARG = DY + NO;            // max: input index of the first maximum (exact in fp32: < 2^24)
    const long long plane = (long long)n * a.C + c;
    const float* x = a.x + plane * a.Hi * a.Wi;
    PCD_FOR(i, a.Hi * a.Wi) X[i] = x[i];
    PCD_SYNC();
    const float* dy = a.dy + plane * NO;
    PCD_FOR(o, NO) {
        int arg, cnt;
        pool_window(X, a.Hi, a.Wi, a.S, o / a.Wo, o % a.Wo, a.is_max != 0, &arg, &cnt);
        DY[o] = a.is_max ? dy[o] : dy[o] / (float)cnt;
        ARG[o] = (float)arg;
    }
    PCD_SYNC();
    float* dx = a.dx + plane * a.Hi * a.Wi;
    PCD_FOR(i, a.Hi * a.Wi) {
        const int iy = i / a.Wi, ix = i - iy * a.Wi;
        float acc = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            const int ny = iy + 1 - ky;
            if (ny < 0 || ny % a.S) continue;
            const int oy = ny / a.S;
            if (oy >= a.Ho) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int nx = ix + 1 - kx;
                if (nx < 0 || nx % a.S) continue;
                const int ox = nx / a.S;
                if (ox >= a.Wo) continue;
                const int o = oy * a.Wo + ox;
                if (!a.is_max || ARG[o] == (float)i) acc += DY[o];
            }
        }
        dx[i] = acc;
    }
}

// ---- y = scale[c] * x + shift[c] (the affine half of BatchNorm(affine=True) after the preprocess kernels' normalisation) ----
struct AffineArgs {
    int B, C, HW;
    const float* x;
    const float* scale;    // null => 1
    const float* shift;    // null => 0
    float* y;
};

PCD_HD void affine_body(const AffineArgs& a, int bx, int ch) {
    const float sc = a.scale ? a.scale[ch] : 1.f, sh = a.shift ? a.shift[ch] : 0.f;
    const long long total = (long long)a.B * a.HW;
    PCD_FOR(k, 4096) {
        const long long t = (long long)bx * 4096 + k;
        if (t < total) {
            const long long n = t / a.HW, p = t - n * a.HW;
            const long long o = (n * a.C + ch) * a.HW + p;
            a.y[o] = fmaf(a.x[o], sc, sh);
        }
    }
}

}  // namespace pcd

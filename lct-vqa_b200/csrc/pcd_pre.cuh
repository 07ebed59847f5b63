// pcd_pre.cuh — the Cell preprocess ops as pixel-major 1x1-convolution kernels (v2):
//   ReLUConvBN(C_in, C_out, 1, 1, 0, affine=False)   operations.py:22-33
//   FactorizedReduce(C_in, C_out, affine=False)      operations.py:90-104
// forward : y[co][p] = sum_ci W[co][ci] * relu(x[ci][p])  (+ per-channel sum / sum^2), 256 pixels per block,
//           each thread 4 pixels x C_out/4 channels, W^T staged in shared memory
// backward: dz = BN-backward(dy, y);  dx[ci][p] = relu'(x) * sum_co W[co][ci] dz[co][p];
//           dW[co][ci] = sum_p dz[co][p] relu(x[ci][p])  (C_in processed in chunks of 16 through shared memory)
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

constexpr int kPrePx = 256;

PCD_HOSTDEV size_t pre_smem_floats(int Cin, int Cout) {
    return (size_t)Cin * Cout + (size_t)(Cout / 2) * 256 + (size_t)(Cout / 2) * 4 * 8 + 16;
}

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

template <int COUT>
PCD_HD void pre_conv_body(const PreArgs& a, int bx, int n, float* smem) {
    constexpr int CPT = COUT / 4;
    const int Cin = a.Cin, HW = a.Ho * a.Wo;
    float* Wt = smem;                      // [Cin][COUT]
    float* P = Wt + Cin * COUT;            // [2*CPT][256]
    float* P2 = P + 2 * CPT * 256;
    PCD_FOR(i, Cin * COUT) {
        const int co = i / Cin, ci = i - co * Cin;
        Wt[ci * COUT + co] = a.w[i];
    }
    PCD_SYNC();
    const int p0 = bx * kPrePx;
    const long long cs = (long long)a.Hin * a.Win;
    const float* xb = a.x + (long long)n * Cin * cs;
    PCD_FOR(task, 256) {
        const int pxg = task & 63, cog = task >> 6;
        const int p = p0 + pxg * 4;
        const int shift = (a.fr && cog >= 2) ? 1 : 0;
        float acc[CPT][4];
#pragma unroll
        for (int i = 0; i < CPT; ++i)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[i][t] = 0.f;
        if (p < HW) {
#pragma unroll 4
            for (int ci = 0; ci < Cin; ++ci) {
                float v[4];
                pre_load4(xb + ci * cs, p, HW, a.fr, shift, a.Wo, a.Win, v);
#pragma unroll
                for (int t = 0; t < 4; ++t) v[t] = relu(v[t]);
                const float* wr = Wt + ci * COUT + cog * CPT;
#pragma unroll
                for (int i4 = 0; i4 < CPT / 4; ++i4) {
                    const F4 w = *reinterpret_cast<const F4*>(wr + 4 * i4);
                    const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int t = 0; t < 4; ++t) acc[4 * i4 + k][t] = fmaf(wk[k], v[t], acc[4 * i4 + k][t]);
                }
            }
            float* yb = a.y + ((long long)n * COUT + cog * CPT) * HW + p;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                if (p + 3 < HW && (((uintptr_t)(yb + (long long)i * HW)) & 15) == 0) {
                    F4 o = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
                    *reinterpret_cast<F4*>(yb + (long long)i * HW) = o;
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (p + t < HW) yb[(long long)i * HW + t] = acc[i][t];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (p + t < HW) { s += acc[i][t]; q = fmaf(acc[i][t], acc[i][t], q); }
            P[(2 * i) * 256 + task] = s;
            P[(2 * i + 1) * 256 + task] = q;
        }
    }
    reduce_columns<8>(P, P2, 2 * CPT, 4, 64, 256, [&](int cog, int k, float v) {
        pcd_atomic_add(a.stats + (k & 1) * COUT + cog * CPT + (k >> 1), (double)v);
    });
}

struct PreBwdArgs {
    int B, Cin, Cout, Hin, Win, Ho, Wo, fr;
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

constexpr int kPreKC = 16;      // input channels per shared-memory chunk in the dW pass

PCD_HOSTDEV size_t pre_bwd_smem_floats(int Cout) {
    return (size_t)Cout * 256 + 2 * kPreKC * 256 + 16 * 256 + 3 * Cout + 16;
}

template <int COUT>
PCD_HD void pre_bwd_body(const PreBwdArgs& a, int bx, int n, float* smem) {
    constexpr int CPT = COUT / 4;
    const int Cin = a.Cin, HW = a.Ho * a.Wo;
    const long long cs = (long long)a.Hin * a.Win;
    float* DZ = smem;                       // [COUT][256]
    float* R = DZ + COUT * 256;             // [2][KC][256]
    float* P = R + 2 * kPreKC * 256;        // [16][256]
    float* COEF = P + 16 * 256;
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
    // ---- dz tile -------------------------------------------------------------------------------------
    PCD_FOR(task, 256) {
        const int pxg = task & 63, cog = task >> 6;
        const int p = p0 + pxg * 4;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int co = cog * CPT + i;
            float dy[4], yy[4], dz[4];
            const long long o = ((long long)n * COUT + co) * HW;
            pre_load4(a.dy + o, p, HW, 0, 0, 0, 0, dy);
            pre_load4(a.y + o, p, HW, 0, 0, 0, 0, yy);
#pragma unroll
            for (int t = 0; t < 4; ++t)
                dz[t] = (p + t < HW) ? COEF[3 * co] * (dy[t] - COEF[3 * co + 1] - yy[t] * COEF[3 * co + 2]) : 0.f;
            F4 v = {dz[0], dz[1], dz[2], dz[3]};
            *reinterpret_cast<F4*>(DZ + co * 256 + pxg * 4) = v;
        }
    }
    PCD_SYNC();
    // ---- dx ----------------------------------------------------------------------------------------------
    if (a.dx) {
        float* dxb = a.dx + (long long)n * Cin * cs;
        PCD_FOR(task, (Cin / 8) * 64) {
            const int cig = task >> 6, pxg = task & 63;
            const int p = p0 + pxg * 4;
            if (p >= HW) continue;
            float s0[8][4], s1[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int t = 0; t < 4; ++t) { s0[i][t] = 0.f; s1[i][t] = 0.f; }
            const int half = a.fr ? COUT / 2 : COUT;
            for (int co = 0; co < COUT; ++co) {
                const F4 d4 = *reinterpret_cast<const F4*>(DZ + co * 256 + pxg * 4);
                const float d[4] = {d4.x, d4.y, d4.z, d4.w};
                const float* wr = a.w + co * Cin + cig * 8;
                float wv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) wv[i] = wr[i];
                if (co < half) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int t = 0; t < 4; ++t) s0[i][t] = fmaf(wv[i], d[t], s0[i][t]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int t = 0; t < 4; ++t) s1[i][t] = fmaf(wv[i], d[t], s1[i][t]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ci = cig * 8 + i;
                if (!a.fr) {
                    float xv[4];
                    pre_load4(xb + ci * cs, p, HW, 0, 0, 0, 0, xv);
                    float* d = dxb + ci * cs + p;
                    if (p + 3 < HW && (((uintptr_t)d) & 15) == 0) {
                        F4 o = {xv[0] > 0.f ? s0[i][0] : 0.f, xv[1] > 0.f ? s0[i][1] : 0.f, xv[2] > 0.f ? s0[i][2] : 0.f,
                                xv[3] > 0.f ? s0[i][3] : 0.f};
                        *reinterpret_cast<F4*>(d) = o;
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (p + t < HW) d[t] = xv[t] > 0.f ? s0[i][t] : 0.f;
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int pp = p + t;
                        if (pp >= HW) continue;
                        const int oy = pp / a.Wo, ox = pp - oy * a.Wo;
                        const float* xs = xb + ci * cs + (long long)(2 * oy) * a.Win + 2 * ox;
                        float* d = dxb + ci * cs + (long long)(2 * oy) * a.Win + 2 * ox;
                        d[0] = xs[0] > 0.f ? s0[i][t] : 0.f;
                        d[1] = 0.f;
                        d[a.Win] = 0.f;
                        d[a.Win + 1] = xs[a.Win + 1] > 0.f ? s1[i][t] : 0.f;
                    }
                }
            }
        }
    }
    // ---- dW: chunks of kPreKC input channels through shared memory ------------------------------------
    if (a.gw) {
        constexpr int NOG = CPT * (kPreKC / 4), NSL = 256 / NOG, SPS = 64 / NSL;   // output groups, slices, strips/slice
        for (int kc = 0; kc < Cin; kc += kPreKC) {
            PCD_SYNC();
            PCD_FOR(i, kPreKC * 64) {
                const int cl = i >> 6, pxg = i & 63;
                const int p = p0 + pxg * 4;
                float v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f};
                if (kc + cl < Cin && p < HW) {
                    pre_load4(xb + (kc + cl) * cs, p, HW, a.fr, 0, a.Wo, a.Win, v0);
                    if (a.fr) pre_load4(xb + (kc + cl) * cs, p, HW, 1, 1, a.Wo, a.Win, v1);
                }
                F4 o0 = {relu(v0[0]), relu(v0[1]), relu(v0[2]), relu(v0[3])};
                *reinterpret_cast<F4*>(R + cl * 256 + pxg * 4) = o0;
                if (a.fr) {
                    F4 o1 = {relu(v1[0]), relu(v1[1]), relu(v1[2]), relu(v1[3])};
                    *reinterpret_cast<F4*>(R + (kPreKC + cl) * 256 + pxg * 4) = o1;
                }
            }
            PCD_SYNC();
            PCD_FOR(task, 256) {
                const int og = task / NSL, sl = task - og * NSL;
                const int co0 = (og / (kPreKC / 4)) * 4, ci0 = (og % (kPreKC / 4)) * 4;
                const float* Rr = R + ((a.fr && co0 >= COUT / 2) ? kPreKC * 256 : 0);
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
                for (int st = sl * SPS; st < (sl + 1) * SPS; ++st) {
                    float dz[4][4], rv[4][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const F4 d = *reinterpret_cast<const F4*>(DZ + (co0 + k) * 256 + st * 4);
                        dz[k][0] = d.x; dz[k][1] = d.y; dz[k][2] = d.z; dz[k][3] = d.w;
                        const F4 r = *reinterpret_cast<const F4*>(Rr + (ci0 + k) * 256 + st * 4);
                        rv[k][0] = r.x; rv[k][1] = r.y; rv[k][2] = r.z; rv[k][3] = r.w;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int t = 0; t < 4; ++t) acc[i][k] = fmaf(dz[i][t], rv[k][t], acc[i][k]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) P[(i * 4 + k) * 256 + task] = acc[i][k];
            }
            PCD_SYNC();
            PCD_FOR(kg, 16 * NOG) {
                const int k = kg / NOG, og = kg - k * NOG;
                float s = 0.f;
                for (int t = 0; t < NSL; ++t) s += P[k * 256 + og * NSL + t];
                const int co = (og / (kPreKC / 4)) * 4 + (k >> 2), ci = kc + (og % (kPreKC / 4)) * 4 + (k & 3);
                if (ci < Cin) pcd_atomic_add(a.gw + co * Cin + ci, s);
            }
        }
    }
}

}  // namespace pcd

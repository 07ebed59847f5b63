// pcd_bwd.cuh — backward kernel bodies.
//
// Backward of a node sum + its incoming MixedOp edges, split at the BatchNorm-backward reductions:
//   node_stats  : per node   sum h, sum h*z_k, sum h*xs, bypass dot  (h = dN[:, 0::4])
//   edge_bwdB   : second halves of the separable convs -> grad of the mid tensor (GA) + its sums,
//                 weight grads of the second dw/pw pair
//   edge_bwdA   : everything that reads the edge input xs: first sep halves, dilated convs, pools,
//                 skip / FactorizedReduce -> d xs, weight grads
//   source_grad : per source state: d state = [sum_e d xs_e | sum_e beta_e * unshuffled dN] (+ cell-output grad)
//   arch_grads  : d softmax(alpha) rows and d beta from the reduction scratch
#pragma once
#include "pcd_fwd.cuh"

namespace pcd {

// dz = c0 * (dy - a - (z - m) * c1)        (BatchNorm backward, affine=False, batch statistics)
struct DzC { float c0, a, m, c1; };

PCD_HD DzC dz_consts(const double* st, int c, int bn, int j, double n, float eps, double sum_dy,
                     double sum_dyz, float kappa) {
    BnC b = bn_consts(st, c, bn, j, n, eps);
    DzC r;
    r.c0 = b.rstd * kappa;
    r.a = (float)(sum_dy / n);
    r.m = b.mean;
    r.c1 = (float)((double)b.rstd * (double)b.rstd * (sum_dyz - (double)b.mean * sum_dy) / n);
    return r;
}

// ---- node_stats -------------------------------------------------------------------------------------
struct EdgeS {
    const float* x;
    long long x_ns;
    int stride, Hs, Ws;
    const float* saved;
    double* bstats;
};

struct NodeStatsArgs {
    int B, Ho, Wo, px_per_block;
    const float* dn;       // node grad (B, C, Ho, Wo) view
    long long dn_ns;
    int nin;
    EdgeS e[kMaxNodeIn];
};

constexpr int kStatK = 9;

PCD_HOSTDEV size_t node_stats_smem_floats(int C) { return (size_t)kStatK * 1024 + (size_t)kStatK * C * 32 + 16; }

template <int C>
PCD_HD void node_stats_body(const NodeStatsArgs& a, int bx, int n, int ez, float* smem) {
    const EdgeS& e = a.e[ez];
    const int HW = a.Ho * a.Wo, PXB = a.px_per_block, NSTRIP = PXB / 4, NT = C * NSTRIP;
    float* P = smem;
    float* P2 = P + kStatK * 1024;
    const int p0 = bx * PXB;
    const long long nslot = (long long)a.B * C * HW;
    const float* dnb = a.dn + (long long)n * a.dn_ns;
    const float* xb = e.x + (long long)n * e.x_ns;
    const int s = e.stride;
    PCD_FOR(task, NT) {
        const int j = task / NSTRIP, strip = task - j * NSTRIP;
        float acc[kStatK];
#pragma unroll
        for (int k = 0; k < kStatK; ++k) acc[k] = 0.f;
        const int slots[6] = {slot_p1(), slot_p2(), slot_z(1), slot_z(3), slot_z(4), slot_z(5)};
        for (int t = 0; t < 4; ++t) {
            const int p = p0 + strip * 4 + t;
            if (p >= HW) break;
            const float h = dnb[(long long)(4 * j) * HW + p];
            const long long so = ((long long)n * C + j) * HW + p;
            acc[0] += h;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[1 + k] = fmaf(h, e.saved[slots[k] * nslot + so], acc[1 + k]);
            if (s == 1) {
                acc[7] = fmaf(h, xb[(long long)j * HW + p], acc[7]);
#pragma unroll
                for (int q = 1; q < 4; ++q)
                    acc[8] = fmaf(dnb[(long long)(4 * j + q) * HW + p], xb[(long long)(q * C + j) * HW + p], acc[8]);
            } else {
                acc[7] = fmaf(h, e.saved[slot_f() * nslot + so], acc[7]);
                const int oy = p / a.Wo, ox = p - oy * a.Wo;
#pragma unroll
                for (int q = 1; q < 4; ++q) {
                    const float* pl = xb + (long long)(q * C + j) * e.Hs * e.Ws + (2 * oy) * e.Ws + 2 * ox;
                    float v = pl[0];
                    v = pl[1] > v ? pl[1] : v;
                    v = pl[e.Ws] > v ? pl[e.Ws] : v;
                    v = pl[e.Ws + 1] > v ? pl[e.Ws + 1] : v;
                    acc[8] = fmaf(dnb[(long long)(4 * j + q) * HW + p], v, acc[8]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kStatK; ++k) P[k * NT + task] = acc[k];
    }
    reduce_columns(P, P2, kStatK, C, NSTRIP, NT, [&](int j, int k, float v) {
        int row;
        if (k == 0) row = bs_s0();
        else if (k == 1) row = bs_sz(bn_p1());
        else if (k == 2) row = bs_sz(bn_p2());
        else if (k == 3) row = bs_sz(bn_unit(s, 1));
        else if (k == 4) row = bs_sz(bn_unit(s, 3));
        else if (k == 5) row = bs_sz(bn_unit(s, 4));
        else if (k == 6) row = bs_sz(bn_unit(s, 5));
        else if (k == 7) row = (s == 2) ? bs_sz(bn_f()) : bs_sx();
        else { pcd_atomic_add(e.bstats + 15 * C, (double)v); return; }
        pcd_atomic_add(e.bstats + row * C + j, (double)v);
    });
}

// ---- shared pieces of the edge backward kernels -----------------------------------------------------
struct EdgeG {
    const float* x;        // source state
    long long x_ns;
    const float* dn;       // grad of the node this edge feeds (B, C, Ho, Wo) view
    long long dn_ns;
    const float* saved;
    const double* stats;
    double* bstats;
    const float* par;
    float* gpar;           // may be null when need_wgrad == 0
    const float* alpha;
    const float* beta;     // null => 1
    float* ga;             // 2 slots: grad wrt BN(A3) / BN(A5) outputs (post ReLU mask)
    float* dxs;            // (B, c, Hs, Ws): grad wrt x[:, :c] from the 7 candidate ops
};

struct EdgeBwdArgs {
    int B, Hs, Ws, Ho, Wo, S;
    int TH, TW, tiles_x;
    float eps;
    int nedges, need_wgrad;
    EdgeG e[kMaxEdgesPerLaunch];
};

// dWpw[co][ci] += sum_p DZ[co][p] * t[ci][p] over the tile (t read from the saved slot).
// Tasks: (C/4)^2 output groups x NSL pixel slices; 16 partials per task.
template <int C>
PCD_HD void wgrad_pw(const float* DZ, const float* t_slot, float* gw, float* P, float* P2, const Geo& g) {
    constexpr int NOG = (C / 4) * (C / 4), NSL = 256 / NOG;
    const int NPIX = g.TH * g.TW, PPS = (NPIX + NSL - 1) / NSL;
    PCD_FOR(task, 256) {
        const int og = task / NSL, sl = task - og * NSL;
        const int co0 = (og / (C / 4)) * 4, ci0 = (og % (C / 4)) * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
        for (int pp = sl * PPS; pp < (sl + 1) * PPS && pp < NPIX; ++pp) {
            const int oyl = pp / g.TW, oxl = pp - oyl * g.TW;
            const int oy = g.oy0 + oyl, ox = g.ox0 + oxl;
            if (oy >= g.Ho || ox >= g.Wo) continue;
            float tv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) tv[k] = t_slot[out_index(g, C, ci0 + k, oy, ox)];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float d = DZ[(co0 + i) * NPIX + pp];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(d, tv[k], acc[i][k]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) P[(i * 4 + k) * 256 + task] = acc[i][k];
    }
    reduce_columns<8>(P, P2, 16, NOG, NSL, 256, [&](int og, int k, float v) {
        const int co = (og / (C / 4)) * 4 + (k >> 2), ci = (og % (C / 4)) * 4 + (k & 3);
        pcd_atomic_add(gw + co * C + ci, v);
    });
}

// dz on the haloed output tile -> dt = Wpw^T dz into DT[C][RH][IW]; centre dz into DZ[C][NPIX].
// dy comes from `dy_base` (+ channel stride dy_cs, channel index = dy_ch0 + j*dy_chm), z from z_slot.
template <int C>
PCD_HD void dz_dt_tile(float* DT, float* DZ, int RH, int IW, int halo_y, const float* dy_img /* image n */,
                       long long dy_cs, int dy_chm, const float* z_slot, const float* w_pw, const float* COEF,
                       const Geo& g) {
    const int NPIX = g.TH * g.TW;
    PCD_FOR(i, RH * IW) {
        const int r = i / IW, col = i - r * IW;
        const int oyl = r - halo_y, oxl = col - 4;
        const int oy = g.oy0 + oyl, ox = g.ox0 + oxl;
        float dz[C], dt[C];
        const bool in = (oy >= 0 && oy < g.Ho && ox >= 0 && ox < g.Wo);
#pragma unroll
        for (int j = 0; j < C; ++j) {
            float v = 0.f;
            if (in) {
                const float dy = dy_img[(long long)(j * dy_chm) * dy_cs + (long long)oy * g.Wo + ox];
                const float z = z_slot[out_index(g, C, j, oy, ox)];
                v = COEF[4 * j] * (dy - COEF[4 * j + 1] - (z - COEF[4 * j + 2]) * COEF[4 * j + 3]);
            }
            dz[j] = v;
        }
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
            float s = 0.f;
#pragma unroll
            for (int co = 0; co < C; ++co) s = fmaf(w_pw[co * C + ci], dz[co], s);
            dt[ci] = s;
        }
#pragma unroll
        for (int ci = 0; ci < C; ++ci) DT[(ci * RH + r) * IW + col] = dt[ci];
        if (oyl >= 0 && oyl < g.TH && oxl >= 0 && oxl < g.TW) {
#pragma unroll
            for (int j = 0; j < C; ++j) DZ[j * NPIX + oyl * g.TW + oxl] = dz[j];
        }
    }
}

// dWdw[ch][tap] += sum over tile patches of dt(centre of DT) * in(tile)
template <int C, int KS, int DIL, int S, bool RELU>
PCD_HD void wgrad_dw(const float* DT, int RH, int IW, int halo_y, const float* IN, int in_rows, int in_pitch,
                     int in_halo_y, float* gw, float* P, float* P2, const Geo& g) {
    constexpr int PAD = DIL * (KS - 1) / 2;
    const int PW4 = g.TW / 4, NPATCH = (g.TH / 4) * PW4, NT = C * NPATCH;
    PCD_FOR(task, NT) {
        const int ch = task / NPATCH, patch = task - ch * NPATCH;
        const int py = (patch / PW4) * 4, px = (patch % PW4) * 4;
        float dt[4][4], acc[KS * KS];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const F4 v = *reinterpret_cast<const F4*>(DT + (ch * RH + py + i + halo_y) * IW + px + 4);
            dt[i][0] = v.x; dt[i][1] = v.y; dt[i][2] = v.z; dt[i][3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < KS * KS; ++k) acc[k] = 0.f;
        dw_wgrad_patch<KS, DIL, S, RELU>(IN + ch * in_rows * in_pitch, in_pitch, S * py - PAD + in_halo_y, S * px,
                                         dt, acc);
#pragma unroll
        for (int k = 0; k < KS * KS; ++k) P[k * NT + task] = acc[k];
    }
    reduce_columns<8>(P, P2, KS * KS, C, NPATCH, NT, [&](int ch, int k, float v) {
        pcd_atomic_add(gw + ch * KS * KS + k, v);
    });
}

// ---- edge_bwdB ------------------------------------------------------------------------------------------
PCD_HOSTDEV size_t bwdB_smem_floats(int C, int TH, int TW) {
    const size_t tile = (size_t)C * (TH + 8) * (TW + 8);
    const int NPATCH = (TH / 4) * (TW / 4);
    size_t p = (size_t)25 * C * NPATCH;
    if (p < 16 * 256) p = 16 * 256;
    return 2 * tile + (size_t)C * TH * TW + p + 4096 + 6 * C + 64;
}

template <int C, int KS>
PCD_HD void bwdB_half(const EdgeBwdArgs& a, const EdgeG& e, const Geo& g, int half, float* smem) {
    constexpr int PAD = (KS - 1) / 2;
    const int S = a.S, TH = g.TH, TW = g.TW, NPIX = TH * TW, RH = TH + 8, IW = TW + 8;
    const int PW4 = TW / 4, NPATCH = (TH / 4) * PW4;
    float* DT = smem;
    float* Q = DT + C * RH * IW;
    float* DZ = Q + C * RH * IW;
    float* P = DZ + C * NPIX;
    size_t psz = (size_t)25 * C * NPATCH;
    if (psz < 16 * 256) psz = 16 * 256;
    float* P2 = P + psz;
    float* COEF = P2 + 4096;
    float* BNA = COEF + 4 * C;
    const int uA = half ? 2 : 0, uB = uA + 1;
    const long long nslot = (long long)a.B * C * a.Ho * a.Wo;
    const double cnt = (double)a.B * a.Ho * a.Wo;
    const float beta = e.beta ? e.beta[0] : 1.f;
    const float kappa = beta * e.alpha[half ? 5 : 4];
    PCD_FOR(j, C) {
        const int bnB = bn_unit(S, uB);
        DzC d = dz_consts(e.stats, C, bnB, j, cnt, a.eps, e.bstats[bs_s0() * C + j], e.bstats[bs_sz(bnB) * C + j], kappa);
        COEF[4 * j] = d.c0; COEF[4 * j + 1] = d.a; COEF[4 * j + 2] = d.m; COEF[4 * j + 3] = d.c1;
        BnC b = bn_consts(e.stats, C, bn_unit(S, uA), j, cnt, a.eps);
        BNA[2 * j] = b.mean; BNA[2 * j + 1] = b.rstd;
    }
    PCD_SYNC();
    const float* zA = e.saved + slot_z(uA) * nslot;
    const float* zB = e.saved + slot_z(uB) * nslot;
    const float* w_dw = e.par + edge_dw_off(C, S, uB);
    const float* w_pw = e.par + edge_pw_off(C, S, uB);
    dz_dt_tile<C>(DT, DZ, RH, IW, 4, e.dn + (long long)g.n * e.dn_ns, (long long)a.Ho * a.Wo, 4, zB, w_pw, COEF, g);
    PCD_FOR(i, C * RH * IW) {
        const int ch = i / (RH * IW), r = (i / IW) % RH, col = i % IW;
        const int oy = g.oy0 - 4 + r, ox = g.ox0 - 4 + col;
        float v = 0.f;
        if (oy >= 0 && oy < a.Ho && ox >= 0 && ox < a.Wo)
            v = relu((zA[out_index(g, C, ch, oy, ox)] - BNA[2 * ch]) * BNA[2 * ch + 1]);
        Q[i] = v;
    }
    PCD_SYNC();
    if (a.need_wgrad)
        wgrad_pw<C>(DZ, e.saved + slot_t(uB) * nslot, e.gpar + edge_pw_off(C, S, uB), P, P2, g);
    // grad wrt relu(bn(zA)) = flipped depthwise correlation of dt; mask by the ReLU; sums for BN-A backward
    float* ga = e.ga + half * nslot;
    PCD_FOR(task, C * NPATCH) {
        const int ch = task / NPATCH, patch = task - ch * NPATCH;
        const int py = (patch / PW4) * 4, px = (patch % PW4) * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        dw_patch<KS, 1, 1, true, false>(DT + ch * RH * IW, IW, py - PAD + 4, px, w_dw + ch * KS * KS, acc);
        float s = 0.f, sz = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int oy = g.oy0 + py + i;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ox = g.ox0 + px + j;
                const float q = Q[(ch * RH + py + i + 4) * IW + px + j + 4];
                o[j] = q > 0.f ? acc[i][j] : 0.f;
                if (oy < a.Ho && ox < a.Wo) {
                    s += o[j];
                    sz = fmaf(o[j], zA[out_index(g, C, ch, oy, ox)], sz);
                }
            }
            store4(ga, g, C, ch, oy, g.ox0 + px, o);
        }
        P[task] = s;
        P[C * NPATCH + task] = sz;
    }
    reduce_columns<8>(P, P2, 2, C, NPATCH, C * NPATCH, [&](int ch, int k, float v) {
        pcd_atomic_add(e.bstats + (bs_ga(half) + k) * C + ch, (double)v);
    });
    if (a.need_wgrad) {
        wgrad_dw<C, KS, 1, 1, false>(DT, RH, IW, 4, Q, RH, IW, 4, e.gpar + edge_dw_off(C, S, uB), P, P2, g);
    }
}

template <int C>
PCD_HD void bwdB_body(const EdgeBwdArgs& a, int bx, int n, int ez, float* smem) {
    const EdgeG& e = a.e[ez];
    Geo g;
    g.n = n; g.TH = a.TH; g.TW = a.TW; g.Ho = a.Ho; g.Wo = a.Wo;
    g.oy0 = (bx / a.tiles_x) * a.TH;
    g.ox0 = (bx % a.tiles_x) * a.TW;
    bwdB_half<C, 3>(a, e, g, 0, smem);
    PCD_SYNC();
    bwdB_half<C, 5>(a, e, g, 1, smem);
}

// ---- edge_bwdA ------------------------------------------------------------------------------------------
PCD_HOSTDEV size_t bwdA_smem_floats(int C, int S, int TH, int TW) {
    const size_t xin = (size_t)C * (S * TH + 8) * (S * TW + 8);
    const size_t tile = (size_t)C * (TH + 8) * (TW + 8);
    const int NPATCH = (TH / 4) * (TW / 4);
    size_t p = (size_t)25 * C * NPATCH;
    if (p < 16 * 256) p = 16 * 256;
    if (p < tile) p = tile;     // P aliases the second pooling tile
    return xin + tile + p + (size_t)C * TH * TW + (size_t)C * S * TH * S * TW + 4096 + 4 * C + 64;
}

// gather d relu(x) for one stride-2 depthwise conv: in pixel q gets sum_tap w[tap] * dt[(q + PAD - tap*DIL)/2]
template <int KS, int DIL>
PCD_HD void dw_bwd_data_s2(const float* dtp /* plane [RH][IW], halo_y 4, col halo 4 */, int IW, int qy0, int qx0,
                           const float* w, float (&acc)[4][4]) {
    constexpr int PAD = DIL * (KS - 1) / 2;
#pragma unroll
    for (int iy = 0; iy < 4; ++iy)
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
            const int ty = iy + PAD - ky * DIL;               // qy0 is a multiple of 4 (even)
            if ((ty & 1) != 0) continue;
            const int prow = (qy0 + ty) / 2 + 4;
#pragma unroll
            for (int ix = 0; ix < 4; ++ix)
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const int tx = ix + PAD - kx * DIL;
                    if ((tx & 1) != 0) continue;
                    const int pcol = (qx0 + tx) / 2 + 4;
                    acc[iy][ix] = fmaf(w[ky * KS + kx], dtp[prow * IW + pcol], acc[iy][ix]);
                }
        }
}

template <int C, int S, int KS, int DIL>
PCD_HD void bwdA_unit(const EdgeBwdArgs& a, const EdgeG& e, const Geo& g, int u, const float* dy_img, long long dy_cs,
                      int dy_chm, const float* XIN, float* DT, float* DZ, float* ACC, float* P, float* P2,
                      const float* COEF) {
    constexpr int PAD = DIL * (KS - 1) / 2;
    const int TH = g.TH, TW = g.TW, RH = TH + 8, IW = TW + 8, IH = S * TH + 8, XW = S * TW + 8;
    const long long nslot = (long long)a.B * C * a.Ho * a.Wo;
    const float* w_dw = e.par + edge_dw_off(C, S, u);
    const float* w_pw = e.par + edge_pw_off(C, S, u);
    dz_dt_tile<C>(DT, DZ, RH, IW, 4, dy_img, dy_cs, dy_chm, e.saved + slot_z(u) * nslot, w_pw, COEF, g);
    PCD_SYNC();
    if (a.need_wgrad) wgrad_pw<C>(DZ, e.saved + slot_t(u) * nslot, e.gpar + edge_pw_off(C, S, u), P, P2, g);
    // d relu(xs) accumulated into ACC[C][S*TH][S*TW]
    const int AH = S * TH, AW = S * TW;
    const int APW4 = AW / 4, ANP = (AH / 4) * APW4;
    PCD_FOR(task, C * ANP) {
        const int ch = task / ANP, patch = task - ch * ANP;
        const int qy = (patch / APW4) * 4, qx = (patch % APW4) * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        if (S == 1)
            dw_patch<KS, DIL, 1, true, false>(DT + ch * RH * IW, IW, qy - PAD + 4, qx, w_dw + ch * KS * KS, acc);
        else
            dw_bwd_data_s2<KS, DIL>(DT + ch * RH * IW, IW, qy, qx, w_dw + ch * KS * KS, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ACC[(ch * AH + qy + i) * AW + qx + j] += acc[i][j];
    }
    PCD_SYNC();
    if (a.need_wgrad)
        wgrad_dw<C, KS, DIL, S, true>(DT, RH, IW, 4, XIN, IH, XW, 4, e.gpar + edge_dw_off(C, S, u), P, P2, g);
}

template <int C, int S>
PCD_HD void bwdA_body(const EdgeBwdArgs& a, int bx, int n, int ez, float* smem) {
    const EdgeG& e = a.e[ez];
    Geo g;
    g.n = n; g.TH = a.TH; g.TW = a.TW; g.Ho = a.Ho; g.Wo = a.Wo;
    g.oy0 = (bx / a.tiles_x) * a.TH;
    g.ox0 = (bx % a.tiles_x) * a.TW;
    const int TH = a.TH, TW = a.TW, NPIX = TH * TW, RH = TH + 8, IW = TW + 8;
    const int IH = S * TH + 8, XW = S * TW + 8, AH = S * TH, AW = S * TW;
    const int NPATCH = (TH / 4) * (TW / 4);
    const size_t tile = (size_t)C * RH * IW;
    size_t psz = (size_t)25 * C * NPATCH;
    if (psz < 16 * 256) psz = 16 * 256;
    if (psz < tile) psz = tile;
    float* XIN = smem;
    float* DT = XIN + C * IH * XW;
    float* P = DT + tile;          // also the second pooling tile
    float* DT2 = P;
    float* DZ = P + psz;
    float* ACC = DZ + C * NPIX;
    float* P2 = ACC + C * AH * AW;
    float* COEF = P2 + 4096;
    const long long nslot = (long long)a.B * C * a.Ho * a.Wo;
    const long long HWo = (long long)a.Ho * a.Wo;
    const double cnt = (double)a.B * a.Ho * a.Wo;
    const float beta = e.beta ? e.beta[0] : 1.f;
    const int iy0 = S * g.oy0 - 4, ix0 = S * g.ox0 - 4;
    const float* xg = e.x + (long long)n * e.x_ns;
    const float* dn_img = e.dn + (long long)n * e.dn_ns;

    PCD_FOR(i, C * IH * XW) {
        const int ch = i / (IH * XW), r = (i / XW) % IH, col = i % XW;
        const int gy = iy0 + r, gx = ix0 + col;
        float v = 0.f;
        if (gy >= 0 && gy < a.Hs && gx >= 0 && gx < a.Ws) v = xg[((long long)ch * a.Hs + gy) * a.Ws + gx];
        XIN[i] = v;
    }
    PCD_FOR(i, C * AH * AW) ACC[i] = 0.f;
    PCD_SYNC();

    // ---- the four conv units that read relu(xs) ------------------------------------------------------
    for (int k = 0; k < 4; ++k) {
        const int u = (k == 0) ? 0 : (k == 1) ? 2 : (k == 2) ? 4 : 5;
        PCD_FOR(j, C) {
            DzC d;
            const int bn = bn_unit(S, u);
            if (k < 2)   // A units: dy = GA (already includes every upstream factor)
                d = dz_consts(e.stats, C, bn, j, cnt, a.eps, e.bstats[bs_ga(k) * C + j], e.bstats[(bs_ga(k) + 1) * C + j], 1.f);
            else
                d = dz_consts(e.stats, C, bn, j, cnt, a.eps, e.bstats[bs_s0() * C + j], e.bstats[bs_sz(bn) * C + j],
                              beta * e.alpha[k == 2 ? 6 : 7]);
            COEF[4 * j] = d.c0; COEF[4 * j + 1] = d.a; COEF[4 * j + 2] = d.m; COEF[4 * j + 3] = d.c1;
        }
        PCD_SYNC();
        const float* ga_img = e.ga + (k < 2 ? k : 0) * nslot + (long long)n * C * HWo;
        if (k == 0) bwdA_unit<C, S, 3, 1>(a, e, g, u, ga_img, HWo, 1, XIN, DT, DZ, ACC, P, P2, COEF);
        else if (k == 1) bwdA_unit<C, S, 5, 1>(a, e, g, u, ga_img, HWo, 1, XIN, DT, DZ, ACC, P, P2, COEF);
        else if (k == 2) bwdA_unit<C, S, 3, 2>(a, e, g, u, dn_img, HWo, 4, XIN, DT, DZ, ACC, P, P2, COEF);
        else bwdA_unit<C, S, 5, 2>(a, e, g, u, dn_img, HWo, 4, XIN, DT, DZ, ACC, P, P2, COEF);
        PCD_SYNC();
    }

    // ---- skip_connect at stride 2: FactorizedReduce backward ----------------------------------------
    if (S == 2) {
        PCD_FOR(j, C) {
            DzC d = dz_consts(e.stats, C, bn_f(), j, cnt, a.eps, e.bstats[bs_s0() * C + j],
                              e.bstats[bs_sz(bn_f()) * C + j], beta * e.alpha[3]);
            COEF[4 * j] = d.c0; COEF[4 * j + 1] = d.a; COEF[4 * j + 2] = d.m; COEF[4 * j + 3] = d.c1;
        }
        PCD_SYNC();
        const float* F = e.saved + slot_f() * nslot;
        PCD_FOR(pp, NPIX) {
            const int oyl = pp / TW, oxl = pp - oyl * TW;
            const int oy = g.oy0 + oyl, ox = g.ox0 + oxl;
            const bool in = oy < a.Ho && ox < a.Wo;
            float dz[C];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                float v = 0.f;
                if (in) {
                    const float h = dn_img[(long long)(4 * j) * HWo + (long long)oy * a.Wo + ox];
                    v = COEF[4 * j] * (h - COEF[4 * j + 1] - (F[out_index(g, C, j, oy, ox)] - COEF[4 * j + 2]) * COEF[4 * j + 3]);
                }
                dz[j] = v;
                DZ[j * NPIX + pp] = v;
            }
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int co = 0; co < C / 2; ++co) {
                    s0 = fmaf(e.par[co * C + ci], dz[co], s0);
                    s1 = fmaf(e.par[(co + C / 2) * C + ci], dz[co + C / 2], s1);
                }
                ACC[(ci * AH + 2 * oyl) * AW + 2 * oxl] += s0;
                ACC[(ci * AH + 2 * oyl + 1) * AW + 2 * oxl + 1] += s1;
            }
        }
        PCD_SYNC();
        if (a.need_wgrad) {
            // dW_fr[co][ci] += sum_p dz[co][p] * relu(x[ci][2p + off(co)])
            constexpr int NOG = (C / 4) * (C / 4), NSL = 256 / NOG;
            const int PPS = (NPIX + NSL - 1) / NSL;
            PCD_FOR(task, 256) {
                const int og = task / NSL, sl = task - og * NSL;
                const int co0 = (og / (C / 4)) * 4, ci0 = (og % (C / 4)) * 4;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
                for (int pp = sl * PPS; pp < (sl + 1) * PPS && pp < NPIX; ++pp) {
                    const int oyl = pp / TW, oxl = pp - oyl * TW;
                    float rv[2][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        rv[0][k] = relu(XIN[((ci0 + k) * IH + 2 * oyl + 4) * XW + 2 * oxl + 4]);
                        rv[1][k] = relu(XIN[((ci0 + k) * IH + 2 * oyl + 5) * XW + 2 * oxl + 5]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float d = DZ[(co0 + i) * NPIX + pp];
                        const int off = (co0 + i) >= C / 2 ? 1 : 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(d, rv[off][k], acc[i][k]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) P[(i * 4 + k) * 256 + task] = acc[i][k];
            }
            reduce_columns<8>(P, P2, 16, NOG, NSL, 256, [&](int og, int k, float v) {
                const int co = (og / (C / 4)) * 4 + (k >> 2), ci = (og % (C / 4)) * 4 + (k & 3);
                pcd_atomic_add(e.gpar + co * C + ci, v);
            });
        }
    }

    // ---- ReLU mask on the conv paths; identity skip at stride 1 -----------------------------------------
    PCD_FOR(i, C * AH * AW) {
        const int ch = i / (AH * AW), qy = (i / AW) % AH, qx = i % AW;
        const float xv = XIN[(ch * IH + qy + 4) * XW + qx + 4];
        float v = xv > 0.f ? ACC[i] : 0.f;
        if (S == 1) {
            const int oy = g.oy0 + qy, ox = g.ox0 + qx;
            if (oy < a.Ho && ox < a.Wo)
                v = fmaf(beta * e.alpha[3], dn_img[(long long)(4 * ch) * HWo + (long long)oy * a.Wo + ox], v);
        }
        ACC[i] = v;
    }
    PCD_SYNC();

    // ---- pools: dz on the 1-haloed output tile, then gather over the windows containing each input px --
    for (int pool = 0; pool < 2; ++pool) {
        const int bn = pool ? bn_p2() : bn_p1();
        PCD_FOR(j, C) {
            DzC d = dz_consts(e.stats, C, bn, j, cnt, a.eps, e.bstats[bs_s0() * C + j], e.bstats[bs_sz(bn) * C + j],
                              beta * e.alpha[pool ? 2 : 1]);
            COEF[4 * j] = d.c0; COEF[4 * j + 1] = d.a; COEF[4 * j + 2] = d.m; COEF[4 * j + 3] = d.c1;
        }
        PCD_SYNC();
        const float* Z = e.saved + (pool ? slot_p2() : slot_p1()) * nslot;
        PCD_FOR(i, C * RH * IW) {
            const int ch = i / (RH * IW), r = (i / IW) % RH, col = i % IW;
            const int oyl = r - 4, oxl = col - 4;
            const int oy = g.oy0 + oyl, ox = g.ox0 + oxl;
            float dzv = 0.f, code = -1.f;
            if (oyl >= -1 && oyl <= TH && oxl >= -1 && oxl <= TW && oy >= 0 && oy < a.Ho && ox >= 0 && ox < a.Wo) {
                const float h = dn_img[(long long)(4 * ch) * HWo + (long long)oy * a.Wo + ox];
                dzv = COEF[4 * ch] * (h - COEF[4 * ch + 1] - (Z[out_index(g, C, ch, oy, ox)] - COEF[4 * ch + 2]) * COEF[4 * ch + 3]);
                float m = -INFINITY;
                int cntv = 0, best = -1;
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) {
                        const int gy = S * oy + dy - 1, gx = S * ox + dx - 1;
                        if (gy >= 0 && gy < a.Hs && gx >= 0 && gx < a.Ws) {
                            const float v = XIN[(ch * IH + S * oyl + dy + 3) * XW + S * oxl + dx + 3];
                            if (v > m || best < 0) { m = v; best = dy * 3 + dx; }
                            ++cntv;
                        }
                    }
                if (pool) dzv = dzv / (float)cntv;
                code = (float)best;
            }
            DT[i] = dzv;
            if (!pool) DT2[i] = code;
        }
        PCD_SYNC();
        PCD_FOR(i, C * AH * AW) {
            const int ch = i / (AH * AW), qy = (i / AW) % AH, qx = i % AW;
            float s = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int ty = qy + 1 - dy;
                if (ty % S != 0) continue;
                const int pr = ty / S + 4;          // ty >= -1; -1 only when S == 1
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int tx = qx + 1 - dx;
                    if (tx % S != 0) continue;
                    const int pc = tx / S + 4;
                    const int idx = (ch * RH + pr) * IW + pc;
                    if (pool) s += DT[idx];
                    else if (DT2[idx] == (float)(dy * 3 + dx)) s += DT[idx];
                }
            }
            ACC[i] += s;
        }
        PCD_SYNC();
    }

    PCD_FOR(i, C * AH * AW) {
        const int ch = i / (AH * AW), qy = (i / AW) % AH, qx = i % AW;
        const int gy = S * g.oy0 + qy, gx = S * g.ox0 + qx;
        if (gy < a.Hs && gx < a.Ws) e.dxs[(((long long)n * C + ch) * a.Hs + gy) * a.Ws + gx] = ACC[i];
    }
}

// ---- source_grad ----------------------------------------------------------------------------------------
struct SrcEdge {
    const float* dxs;      // (B, c, Hs, Ws)
    const float* dn;       // node grad of the consumer (B, C, Ho, Wo) view
    long long dn_ns;
    const float* beta;     // null => 1
    int stride;
};

constexpr int kMaxSrcEdges = 4;

struct SourceGradArgs {
    int B, C, Hs, Ws;
    const float* x;        // the source state itself (needed for the 2x2 max-pool argmax at stride 2)
    long long x_ns;
    const float* g0;       // optional initial grad (cell-output slice), same view geometry as `out`
    long long g0_ns;
    float* out;            // d source (B, C, Hs, Ws) view
    long long out_ns;
    int nedges;
    SrcEdge e[kMaxSrcEdges];
};

PCD_HD void source_grad_body(const SourceGradArgs& a, int bx, int ch, int n) {
    const int HW = a.Hs * a.Ws, c = a.C / 4;
    const int p0 = bx * 4096;
    const int npx = (HW - p0) < 4096 ? (HW - p0) : 4096;
    float* ob = a.out + (long long)n * a.out_ns + (long long)ch * HW;
    PCD_FOR(i, npx) {
        const int p = p0 + i;
        float v = a.g0 ? a.g0[(long long)n * a.g0_ns + (long long)ch * HW + p] : 0.f;
        for (int k = 0; k < a.nedges; ++k) {
            const SrcEdge& e = a.e[k];
            if (ch < c) {
                v += e.dxs[((long long)n * c + ch) * HW + p];
            } else {
                const int q = ch / c, j = ch - q * c;     // x channel q*c + j  <->  node channel 4j + q
                const float beta = e.beta ? e.beta[0] : 1.f;
                if (e.stride == 1) {
                    v = fmaf(beta, e.dn[(long long)n * e.dn_ns + (long long)(4 * j + q) * HW + p], v);
                } else {
                    const int y = p / a.Ws, x = p - y * a.Ws;
                    const int oy = y >> 1, ox = x >> 1, Ho = a.Hs >> 1, Wo = a.Ws >> 1;
                    const float* pl = a.x + (long long)n * a.x_ns + (long long)ch * HW + (2 * oy) * a.Ws + 2 * ox;
                    int best = 0;
                    float m = pl[0];
                    if (pl[1] > m) { m = pl[1]; best = 1; }
                    if (pl[a.Ws] > m) { m = pl[a.Ws]; best = 2; }
                    if (pl[a.Ws + 1] > m) { m = pl[a.Ws + 1]; best = 3; }
                    if (best == (y & 1) * 2 + (x & 1))
                        v = fmaf(beta, e.dn[(long long)n * e.dn_ns + ((long long)(4 * j + q) * Ho + oy) * Wo + ox], v);
                }
            }
        }
        ob[p] = v;
    }
}

// ---- arch_grads -----------------------------------------------------------------------------------------
struct ArchEdge {
    const double* stats;
    const double* bstats;
    const float* alpha;
    const float* beta;
    int stride;
    double count;
    float* gw;     // 8 floats
    float* gw2;    // 1 float or null
};

struct ArchGradArgs {
    int c, nedges;
    float eps;
    ArchEdge e[PCD_MAX_EDGES_CONST];
};

PCD_HD void arch_grads_body(const ArchGradArgs& a) {
    PCD_FOR(ei, a.nedges) {
        const ArchEdge& e = a.e[ei];
        const int c = a.c, s = e.stride;
        const float beta = e.beta ? e.beta[0] : 1.f;
        const int bns[7] = {bn_p1(), bn_p2(), bn_unit(s, 1), bn_unit(s, 3), bn_unit(s, 4), bn_unit(s, 5), bn_f()};
        const int prim[7] = {1, 2, 4, 5, 6, 7, 3};
        double D[8];
        for (int k = 0; k < 8; ++k) D[k] = 0.0;
        for (int k = 0; k < 7; ++k) {
            if (k == 6 && s != 2) break;
            double acc = 0.0;
            for (int j = 0; j < c; ++j) {
                BnC b = bn_consts(e.stats, c, bns[k], j, e.count, a.eps);
                acc += (double)b.rstd * (e.bstats[bs_sz(bns[k]) * c + j] - (double)b.mean * e.bstats[bs_s0() * c + j]);
            }
            D[prim[k]] = acc;
        }
        if (s == 1) {
            double acc = 0.0;
            for (int j = 0; j < c; ++j) acc += e.bstats[bs_sx() * c + j];
            D[3] = acc;
        }
        double db = e.bstats[15 * c];
        for (int k = 0; k < 8; ++k) {
            e.gw[k] = (float)(beta * D[k]);
            db += (double)e.alpha[k] * D[k];
        }
        if (e.gw2) e.gw2[0] = (float)db;
    }
}

// ---- preprocess backward ---------------------------------------------------------------------------------
struct BnBwdStatArgs {
    int B, C, HW;
    const float* dy;       // (B, C, HW)
    const float* y;        // normalised output (affine=False) or pre-BN z when `stats` is given
    const double* stats;   // null => y is already normalised
    float eps;
    double* bstats;        // sum dy [C], sum dy*yhat [C]
};

PCD_HOSTDEV size_t bn_bwd_stats_smem_floats() { return 2 * 1024 + 2 * 32 + 16; }

PCD_HD void bn_bwd_stats_body(const BnBwdStatArgs& a, int bx, int ch, int n, float* smem) {
    float* P = smem;
    float* P2 = P + 2 * 1024;
    const long long base = ((long long)n * a.C + ch) * a.HW;
    const int p0 = bx * 4096;
    float mean = 0.f, rstd = 1.f;
    if (a.stats) {
        BnC b = bn_consts(a.stats, a.C, 0, ch, (double)a.B * a.HW, a.eps);
        mean = b.mean; rstd = b.rstd;
    }
    PCD_FOR(task, 1024) {
        float s = 0.f, q = 0.f;
        for (int t = 0; t < 4; ++t) {
            const int p = p0 + task * 4 + t;
            if (p < a.HW) {
                const float d = a.dy[base + p];
                s += d;
                q = fmaf(d, (a.y[base + p] - mean) * rstd, q);
            }
        }
        P[task] = s;
        P[1024 + task] = q;
    }
    reduce_columns(P, P2, 2, 1, 1024, 1024, [&](int, int k, float v) {
        pcd_atomic_add(a.bstats + k * a.C + ch, (double)v);
    });
}

struct PreBwdArgs {
    int B, Cin, Cout, Hin, Win, Ho, Wo, fr, PXB, nblocks_px, nblocks_launch;
    const float* x;        // cell input (B, Cin, Hin, Win)
    const float* w;
    const float* y;        // normalised preprocess output (B, Cout, Ho, Wo)
    const float* dy;       // its grad
    const double* stats;   // forward sums (for rstd)
    const double* bstats;  // sum dy, sum dy*y
    float eps;
    float* dx;             // (B, Cin, Hin, Win) written; may be null
    float* gw;             // [Cout][Cin] accumulated; may be null
};

PCD_HOSTDEV size_t pre_bwd_smem_floats(int Cin, int Cout, int PXB, int fr) {
    return (size_t)Cout * Cin + (size_t)Cout * PXB + (size_t)(fr ? 2 : 1) * Cin * PXB + 3 * Cout + 16;
}

PCD_HD void pre_bwd_body(const PreBwdArgs& a, int bx, int nblk, float* smem) {
    const int Cin = a.Cin, Cout = a.Cout, PXB = a.PXB, HWo = a.Ho * a.Wo;
    const long long HWi = (long long)a.Hin * a.Win;
    float* WACC = smem;
    float* DZ = WACC + Cout * Cin;
    float* R = DZ + Cout * PXB;                 // [fr?2:1][Cin][PXB]
    float* COEF = R + (a.fr ? 2 : 1) * Cin * PXB;
    const double cnt = (double)a.B * HWo;
    PCD_FOR(i, Cout * Cin) WACC[i] = 0.f;
    PCD_FOR(co, Cout) {
        BnC b = bn_consts(a.stats, Cout, 0, co, cnt, a.eps);
        COEF[3 * co] = b.rstd;
        COEF[3 * co + 1] = (float)(a.bstats[co] / cnt);
        COEF[3 * co + 2] = (float)(a.bstats[Cout + co] / cnt);
    }
    PCD_SYNC();
    const int per_img = (HWo + PXB - 1) / PXB;
    for (int blk = bx; blk < a.nblocks_px; blk += nblk) {
        const int n = blk / per_img, p0 = (blk - n * per_img) * PXB;
        const float* xb = a.x + (long long)n * Cin * HWi;
        PCD_FOR(i, Cout * PXB) {
            const int co = i / PXB, t = i - co * PXB, p = p0 + t;
            float v = 0.f;
            if (p < HWo) {
                const long long o = ((long long)n * Cout + co) * HWo + p;
                v = COEF[3 * co] * (a.dy[o] - COEF[3 * co + 1] - a.y[o] * COEF[3 * co + 2]);
            }
            DZ[i] = v;
        }
        PCD_FOR(i, Cin * PXB) {
            const int ci = i / PXB, t = i - ci * PXB, p = p0 + t;
            float v0 = 0.f, v1 = 0.f;
            if (p < HWo) {
                if (a.fr) {
                    const int oy = p / a.Wo, ox = p - oy * a.Wo;
                    v0 = relu(xb[ci * HWi + (long long)(2 * oy) * a.Win + 2 * ox]);
                    v1 = relu(xb[ci * HWi + (long long)(2 * oy + 1) * a.Win + 2 * ox + 1]);
                } else {
                    v0 = relu(xb[ci * HWi + p]);
                }
            }
            R[i] = v0;
            if (a.fr) R[Cin * PXB + i] = v1;
        }
        PCD_SYNC();
        if (a.gw) {
            const int ncig = Cin / 4;
            PCD_FOR(task, (Cout / 4) * ncig) {
                const int co0 = (task / ncig) * 4, ci0 = (task % ncig) * 4;
                const float* Rr = R + ((a.fr && co0 >= Cout / 2) ? Cin * PXB : 0);
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
                for (int t = 0; t < PXB; ++t) {
                    float rv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) rv[k] = Rr[(ci0 + k) * PXB + t];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float d = DZ[(co0 + i) * PXB + t];
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(d, rv[k], acc[i][k]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) WACC[(co0 + i) * Cin + ci0 + k] += acc[i][k];
            }
        }
        if (a.dx) {
            float* dxb = a.dx + (long long)n * Cin * HWi;
            PCD_FOR(i, Cin * PXB) {
                const int ci = i / PXB, t = i - ci * PXB, p = p0 + t;
                if (p >= HWo) continue;
                if (!a.fr) {
                    float s = 0.f;
                    for (int co = 0; co < Cout; ++co) s = fmaf(a.w[co * Cin + ci], DZ[co * PXB + t], s);
                    dxb[ci * HWi + p] = R[i] > 0.f ? s : 0.f;
                } else {
                    float s0 = 0.f, s1 = 0.f;
                    for (int co = 0; co < Cout / 2; ++co) {
                        s0 = fmaf(a.w[co * Cin + ci], DZ[co * PXB + t], s0);
                        s1 = fmaf(a.w[(co + Cout / 2) * Cin + ci], DZ[(co + Cout / 2) * PXB + t], s1);
                    }
                    const int oy = p / a.Wo, ox = p - oy * a.Wo;
                    float* d = dxb + ci * HWi + (long long)(2 * oy) * a.Win + 2 * ox;
                    d[0] = R[i] > 0.f ? s0 : 0.f;
                    d[1] = 0.f;
                    d[a.Win] = 0.f;
                    d[a.Win + 1] = R[Cin * PXB + i] > 0.f ? s1 : 0.f;
                }
            }
        }
        PCD_SYNC();
    }
    if (a.gw) {
        PCD_FOR(i, Cout * Cin) pcd_atomic_add(a.gw + i, WACC[i]);
    }
}

// ---- stem backward ---------------------------------------------------------------------------------------
struct StemBwdArgs {
    int B, Cout, H, W, PXB, nblocks_px, nblocks_launch;
    const float* x;
    const float* z;        // pre-BN conv output
    const float* dy;
    const float* gamma;
    const double* stats;
    const double* bstats;  // sum dy, sum dy*yhat
    float eps;
    float* gw;             // conv weight grad [Cout][27]
    float* ggamma;
    float* gbias;
};

PCD_HOSTDEV size_t stem_bwd_smem_floats(int Cout, int PXB) { return (size_t)Cout * 27 + (size_t)Cout * PXB + 5 * Cout + 16; }

PCD_HD void stem_bwd_body(const StemBwdArgs& a, int bx, int nblk, float* smem) {
    const int Cout = a.Cout, PXB = a.PXB, HW = a.H * a.W;
    float* WACC = smem;
    float* DZ = WACC + Cout * 27;
    float* COEF = DZ + Cout * PXB;
    const double cnt = (double)a.B * HW;
    PCD_FOR(i, Cout * 27) WACC[i] = 0.f;
    PCD_FOR(co, Cout) {
        BnC b = bn_consts(a.stats, Cout, 0, co, cnt, a.eps);
        COEF[5 * co] = b.rstd * a.gamma[co];
        COEF[5 * co + 1] = (float)(a.bstats[co] / cnt);
        COEF[5 * co + 2] = (float)(a.bstats[Cout + co] / cnt);
        COEF[5 * co + 3] = b.mean;
        COEF[5 * co + 4] = b.rstd;
        if (bx == 0) {
            a.gbias[co] = (float)a.bstats[co];
            a.ggamma[co] = (float)a.bstats[Cout + co];
        }
    }
    PCD_SYNC();
    const int per_img = (HW + PXB - 1) / PXB;
    for (int blk = bx; blk < a.nblocks_px; blk += nblk) {
        const int n = blk / per_img, p0 = (blk - n * per_img) * PXB;
        PCD_FOR(i, Cout * PXB) {
            const int co = i / PXB, t = i - co * PXB, p = p0 + t;
            float v = 0.f;
            if (p < HW) {
                const long long o = ((long long)n * Cout + co) * HW + p;
                const float yh = (a.z[o] - COEF[5 * co + 3]) * COEF[5 * co + 4];
                v = COEF[5 * co] * (a.dy[o] - COEF[5 * co + 1] - yh * COEF[5 * co + 2]);
            }
            DZ[i] = v;
        }
        PCD_SYNC();
        const float* xb = a.x + (long long)n * 3 * HW;
        PCD_FOR(task, (Cout / 4) * 27) {
            const int co0 = (task / 27) * 4, tap = task % 27;
            const int ci = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int t = 0; t < PXB; ++t) {
                const int p = p0 + t;
                if (p >= HW) break;
                const int oy = p / a.W, ox = p - oy * a.W;
                const int gy = oy + ky - 1, gx = ox + kx - 1;
                if (gy < 0 || gy >= a.H || gx < 0 || gx >= a.W) continue;
                const float xv = xb[(long long)ci * HW + gy * a.W + gx];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(DZ[(co0 + i) * PXB + t], xv, acc[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) WACC[(co0 + i) * 27 + tap] += acc[i];
        }
        PCD_SYNC();
    }
    PCD_FOR(i, Cout * 27) pcd_atomic_add(a.gw + i, WACC[i]);
}

}  // namespace pcd

// pcd_fwd.cuh — forward kernel bodies (see pcd_common.cuh for the phase discipline).
//
// Forward of one MixedOp edge (model_search.py:44-58) is split at the BatchNorm batch-statistic
// barriers into
//   passA   : xs tile -> max/avg pool, FactorizedReduce, first halves of the separable convs and both
//             dilated convs (depthwise -> pointwise), + per-channel sum / sum^2 of every pre-BN tensor
//   passB   : BN+ReLU of the sep-conv mid tensors on load -> second depthwise/pointwise pair + stats
//   combine : per NODE: sum over incoming edges of beta * [ sum_k w_k BN(op_k) | bypass ] written
//             channel-shuffled straight into the node's slice of the cell output (model_search.py:90)
#pragma once
#include "pcd_edge.cuh"

namespace pcd {

// ---- node combine ---------------------------------------------------------------------------------
struct EdgeC {
    const float* x;
    long long x_ns;
    int stride, Hs, Ws;
    const float* saved;
    const double* stats;
    const float* alpha;    // 8 softmaxed op weights (device)
    const float* beta;     // 1 scalar (device) or null => 1
    float* running;        // BN running stats of this edge (updated once per forward)
    long long* nbt;
};

constexpr int kMaxNodeIn = 5;

struct CombineArgs {
    int B, Ho, Wo;
    float eps, momentum;
    int nin, update_running;
    int px_per_block;      // pixels of one image per block (multiple of 4)
    float* out;            // node tensor (B, C, Ho, Wo) view
    long long out_ns;
    EdgeC e[kMaxNodeIn];
};

// pixels per block: one float4 strip per thread and channel where the image is large enough to still fill the GPU
// (a node of cell 2/3 is only 64 images x 256 pixels: 64-pixel chunks give 256 blocks instead of 64)
PCD_HOSTDEV int combine_px(int C) { return 1024 / C; }

PCD_HOSTDEV size_t combine_smem_floats(int C) { return (size_t)kMaxNodeIn * (8 + 7) * C + 16; }

// coefficient rows per edge: 0 P1, 1 P2, 2 B3, 3 B5, 4 D3, 5 D5, 6 F (stride 2) or identity (stride 1), 7 shift
template <int C>
PCD_HD void combine_body(const CombineArgs& a, int bx, int n, float* smem) {
    float* COEF = smem;
    const double cnt = (double)a.B * a.Ho * a.Wo;
    const int HW = a.Ho * a.Wo;
    // one thread per (edge, BN, channel): the fp64 statistics loads of all coefficients are in flight together
    float* MEAN = COEF + kMaxNodeIn * 8 * C;
    PCD_FOR(i, a.nin * 7 * C) {
        const int ei = i / (7 * C), r = i - ei * 7 * C, k = r / C, j = r - k * C;
        const EdgeC& e = a.e[ei];
        const float beta = e.beta ? e.beta[0] : 1.f;
        const int s = e.stride;
        float* co = COEF + ei * 8 * C;
        if (k < 6) {
            const int bn = k == 0 ? bn_p1() : k == 1 ? bn_p2() : bn_unit(s, k == 2 ? 1 : k + 0);
            const int prim = k == 0 ? 1 : k == 1 ? 2 : k + 2;
            BnC b = bn_consts(e.stats, C, bn, j, cnt, a.eps);
            co[k * C + j] = beta * e.alpha[prim] * b.rstd;
            MEAN[(ei * 7 + k) * C + j] = b.mean;
        } else if (s == 2) {
            BnC b = bn_consts(e.stats, C, bn_f(), j, cnt, a.eps);
            co[6 * C + j] = beta * e.alpha[3] * b.rstd;
            MEAN[(ei * 7 + 6) * C + j] = b.mean;
        } else {
            co[6 * C + j] = beta * e.alpha[3];
            MEAN[(ei * 7 + 6) * C + j] = 0.f;
        }
    }
    PCD_SYNC();
    PCD_FOR(i, a.nin * C) {
        const int ei = i / C, j = i - ei * C;
        const float* co = COEF + ei * 8 * C;
        float shift = 0.f;
        for (int k = 0; k < 6; ++k) shift -= co[k * C + j] * MEAN[(ei * 7 + k) * C + j];
        if (a.e[ei].stride == 2) shift -= co[6 * C + j] * MEAN[(ei * 7 + 6) * C + j];
        COEF[ei * 8 * C + 7 * C + j] = shift;
    }
    PCD_SYNC();
    const int p0 = bx * a.px_per_block;
    const int npx = (HW - p0) < a.px_per_block ? (HW - p0) : a.px_per_block;
    const int nstrip = (npx + 3) / 4;
    const long long nslot = (long long)a.B * C * HW;
    bool vec = (HW % 4 == 0) && ((((uintptr_t)a.out) & 15) == 0) && (a.out_ns % 4 == 0);
    for (int ei = 0; ei < a.nin; ++ei)
        vec = vec && ((((uintptr_t)a.e[ei].x) | ((uintptr_t)a.e[ei].saved)) & 15) == 0 && a.e[ei].x_ns % 4 == 0 &&
              (a.e[ei].stride == 1 || (a.Wo % 4 == 0 && a.e[ei].Ws % 4 == 0));
    if (vec) {
        PCD_FOR(task, C * nstrip) {
            const int j = task / nstrip, strip = task - j * nstrip;
            const int p = p0 + strip * 4;
            float m[4] = {0.f, 0.f, 0.f, 0.f}, by[3][4];
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int t = 0; t < 4; ++t) by[q][t] = 0.f;
            // stride-1 edges: the 10 float4 loads of edge ei + 1 are issued before the arithmetic of edge ei (the small
            // cells have too few threads in flight to hide the latency otherwise)
            const int slots[6] = {slot_p1(), slot_p2(), slot_z(1), slot_z(3), slot_z(4), slot_z(5)};
            F4 cz[6], cx[4], nz[6], nx[4];
            auto issue = [&](int ei, F4* z, F4* x) {
                const EdgeC& e = a.e[ei];
                const float* sv = e.saved + ((long long)n * C + j) * HW + p;
                const float* xb = e.x + (long long)n * e.x_ns;
#pragma unroll
                for (int k = 0; k < 6; ++k) z[k] = *reinterpret_cast<const F4*>(sv + slots[k] * nslot);
#pragma unroll
                for (int q = 0; q < 4; ++q) x[q] = *reinterpret_cast<const F4*>(xb + (long long)(q * C + j) * HW + p);
            };
            if (a.e[0].stride == 1) issue(0, nz, nx);
            for (int ei = 0; ei < a.nin; ++ei) {
                const EdgeC& e = a.e[ei];
                const float* co = COEF + ei * 8 * C;
                const float beta = e.beta ? e.beta[0] : 1.f;
                float acc[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) acc[t] = co[7 * C + j];
                const float c6 = co[6 * C + j];
                if (e.stride == 1) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) cz[k] = nz[k];
#pragma unroll
                    for (int q = 0; q < 4; ++q) cx[q] = nx[q];
                    if (ei + 1 < a.nin && a.e[ei + 1].stride == 1) issue(ei + 1, nz, nx);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const float c = co[k * C + j];
                        acc[0] = fmaf(c, cz[k].x, acc[0]); acc[1] = fmaf(c, cz[k].y, acc[1]);
                        acc[2] = fmaf(c, cz[k].z, acc[2]); acc[3] = fmaf(c, cz[k].w, acc[3]);
                    }
                    acc[0] = fmaf(c6, cx[0].x, acc[0]); acc[1] = fmaf(c6, cx[0].y, acc[1]);
                    acc[2] = fmaf(c6, cx[0].z, acc[2]); acc[3] = fmaf(c6, cx[0].w, acc[3]);
#pragma unroll
                    for (int q = 1; q < 4; ++q) {
                        by[q - 1][0] = fmaf(beta, cx[q].x, by[q - 1][0]); by[q - 1][1] = fmaf(beta, cx[q].y, by[q - 1][1]);
                        by[q - 1][2] = fmaf(beta, cx[q].z, by[q - 1][2]); by[q - 1][3] = fmaf(beta, cx[q].w, by[q - 1][3]);
                    }
                } else {
                    if (ei + 1 < a.nin && a.e[ei + 1].stride == 1) issue(ei + 1, nz, nx);
                    const float* sv = e.saved + ((long long)n * C + j) * HW + p;
                    const float* xb = e.x + (long long)n * e.x_ns;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const F4 v = *reinterpret_cast<const F4*>(sv + slots[k] * nslot);
                        const float c = co[k * C + j];
                        acc[0] = fmaf(c, v.x, acc[0]); acc[1] = fmaf(c, v.y, acc[1]);
                        acc[2] = fmaf(c, v.z, acc[2]); acc[3] = fmaf(c, v.w, acc[3]);
                    }
                    const F4 v = *reinterpret_cast<const F4*>(sv + slot_f() * nslot);
                    acc[0] = fmaf(c6, v.x, acc[0]); acc[1] = fmaf(c6, v.y, acc[1]);
                    acc[2] = fmaf(c6, v.z, acc[2]); acc[3] = fmaf(c6, v.w, acc[3]);
                    const int oy = p / a.Wo, ox = p - oy * a.Wo;
#pragma unroll
                    for (int q = 1; q < 4; ++q) {
                        const float* pl = xb + (long long)(q * C + j) * e.Hs * e.Ws + (long long)(2 * oy) * e.Ws + 2 * ox;
                        const F4 r0a = *reinterpret_cast<const F4*>(pl), r0b = *reinterpret_cast<const F4*>(pl + 4);
                        const F4 r1a = *reinterpret_cast<const F4*>(pl + e.Ws), r1b = *reinterpret_cast<const F4*>(pl + e.Ws + 4);
                        const float w0 = fmaxf(fmaxf(r0a.x, r0a.y), fmaxf(r1a.x, r1a.y));
                        const float w1 = fmaxf(fmaxf(r0a.z, r0a.w), fmaxf(r1a.z, r1a.w));
                        const float w2 = fmaxf(fmaxf(r0b.x, r0b.y), fmaxf(r1b.x, r1b.y));
                        const float w3 = fmaxf(fmaxf(r0b.z, r0b.w), fmaxf(r1b.z, r1b.w));
                        by[q - 1][0] = fmaf(beta, w0, by[q - 1][0]); by[q - 1][1] = fmaf(beta, w1, by[q - 1][1]);
                        by[q - 1][2] = fmaf(beta, w2, by[q - 1][2]); by[q - 1][3] = fmaf(beta, w3, by[q - 1][3]);
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) m[t] += acc[t];
            }
            float* ob = a.out + (long long)n * a.out_ns + p;
            F4 o = {m[0], m[1], m[2], m[3]};
            *reinterpret_cast<F4*>(ob + (long long)(4 * j) * HW) = o;
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                F4 b = {by[q - 1][0], by[q - 1][1], by[q - 1][2], by[q - 1][3]};
                *reinterpret_cast<F4*>(ob + (long long)(4 * j + q) * HW) = b;
            }
        }
    } else
    PCD_FOR(task, C * nstrip) {
        const int j = task / nstrip, strip = task - j * nstrip;
        const int p = p0 + strip * 4;
        float m[4] = {0.f, 0.f, 0.f, 0.f}, by[3][4];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int t = 0; t < 4; ++t) by[q][t] = 0.f;
        for (int ei = 0; ei < a.nin; ++ei) {
            const EdgeC& e = a.e[ei];
            const float* co = COEF + ei * 8 * C;
            const float beta = e.beta ? e.beta[0] : 1.f;
            const long long so = ((long long)n * C + j) * HW + p;
            const int slots[6] = {slot_p1(), slot_p2(), slot_z(1), slot_z(3), slot_z(4), slot_z(5)};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (p + t >= HW) break;
                float acc = co[7 * C + j];
#pragma unroll
                for (int k = 0; k < 6; ++k) acc = fmaf(co[k * C + j], e.saved[slots[k] * nslot + so + t], acc);
                if (e.stride == 1) {
                    const float* xb = e.x + (long long)n * e.x_ns;
                    acc = fmaf(co[6 * C + j], xb[(long long)j * HW + p + t], acc);
#pragma unroll
                    for (int q = 1; q < 4; ++q) by[q - 1][t] = fmaf(beta, xb[(long long)(q * C + j) * HW + p + t], by[q - 1][t]);
                } else {
                    acc = fmaf(co[6 * C + j], e.saved[slot_f() * nslot + so + t], acc);
                    const int oy = (p + t) / a.Wo, ox = (p + t) - oy * a.Wo;
                    const float* xb = e.x + (long long)n * e.x_ns;
#pragma unroll
                    for (int q = 1; q < 4; ++q) {
                        const float* pl = xb + (long long)(q * C + j) * e.Hs * e.Ws + (2 * oy) * e.Ws + 2 * ox;
                        float v = pl[0];
                        v = pl[1] > v ? pl[1] : v;
                        v = pl[e.Ws] > v ? pl[e.Ws] : v;
                        v = pl[e.Ws + 1] > v ? pl[e.Ws + 1] : v;
                        by[q - 1][t] = fmaf(beta, v, by[q - 1][t]);
                    }
                }
                m[t] += acc;
            }
        }
        float* ob = a.out + (long long)n * a.out_ns;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (p + t >= HW) break;
            ob[(long long)(4 * j + 0) * HW + p + t] = m[t];
#pragma unroll
            for (int q = 1; q < 4; ++q) ob[(long long)(4 * j + q) * HW + p + t] = by[q - 1][t];
        }
    }
    if (a.update_running && bx == 0 && n == 0) {
        PCD_FOR(i, a.nin * 9 * C) {
            const int ei = i / (9 * C), bn = (i / C) % 9, j = i % C;
            const EdgeC& e = a.e[ei];
            if (bn < edge_nbn(e.stride))
                bn_running_update(e.stats + bn * 2 * C, C, j, cnt, a.momentum, e.running + bn * 2 * C);
        }
        PCD_FOR(i, a.nin * 9) {
            const int ei = i / 9, bn = i % 9;
            if (bn < edge_nbn(a.e[ei].stride)) a.e[ei].nbt[bn] += 1;
        }
    }
}

// y = (y - mean) * rstd in place (+ optional affine), running-stat update by block (0,0)
struct NormArgs {
    int B, C, HW;
    float eps, momentum;
    const float* src;    // pre-BN tensor
    float* dst;          // may alias src
    const double* stats;
    const float* gamma;  // null => 1
    const float* bias;   // null => 0
    float* running;
    long long* nbt;
};

// block (bx, ch): float4 tasks bx*1024 .. +1023 of channel ch, flattened over (image, pixel) when the planes allow float4
PCD_HD void norm_body(const NormArgs& a, int bx, int ch, int) {
    const double cnt = (double)a.B * a.HW;
    BnC b = bn_consts(a.stats, a.C, 0, ch, cnt, a.eps);
    const float sc = b.rstd * (a.gamma ? a.gamma[ch] : 1.f);
    const float sh = (a.bias ? a.bias[ch] : 0.f) - b.mean * sc;
    const bool vec = (a.HW % 4 == 0) && ((((uintptr_t)a.src) | ((uintptr_t)a.dst)) & 15) == 0;
    if (vec) {
        const int HW4 = a.HW / 4, total = a.B * HW4;
        PCD_FOR(k, 1024) {
            const int t = bx * 1024 + k;
            if (t < total) {
                const int n = t / HW4, p4 = t - n * HW4;
                const long long o = ((long long)n * a.C + ch) * a.HW + 4 * p4;
                const F4 v = *reinterpret_cast<const F4*>(a.src + o);
                F4 r = {fmaf(v.x, sc, sh), fmaf(v.y, sc, sh), fmaf(v.z, sc, sh), fmaf(v.w, sc, sh)};
                *reinterpret_cast<F4*>(a.dst + o) = r;
            }
        }
    } else {
        const long long total = (long long)a.B * a.HW;
        PCD_FOR(k, 4096) {
            const long long t = (long long)bx * 4096 + k;
            if (t < total) {
                const long long n = t / a.HW, p = t - n * a.HW;
                const long long o = (n * a.C + ch) * a.HW + p;
                a.dst[o] = fmaf(a.src[o], sc, sh);
            }
        }
    }
    if (bx == 0 && a.running) {
        PCD_FOR(i, 1) {
            bn_running_update(a.stats, a.C, ch, cnt, a.momentum, a.running);
            if (ch == 0) a.nbt[0] += 1;
        }
    }
}
PCD_HOSTDEV int norm_grid_x(int B, int HW) { return (int)(((long long)B * HW + 4095) / 4096); }

// ---- stem: Conv2d(3, Cout, 3, padding=1) (model_search.py:110-113); BN applied by norm_body ----------
struct StemArgs {
    int B, Cout, H, W;
    const float* x;   // (B,3,H,W)
    const float* w;   // (Cout,3,3,3)
    float* z;         // (B,Cout,H,W)
    double* stats;
};

constexpr int kStemPx = 512;

PCD_HOSTDEV size_t stem_smem_floats(int Cout) {
    return (size_t)Cout * 27 + (size_t)16 * (Cout / 8) * (kStemPx / 4) + 16 * (Cout / 8) * 32 + 16 + 3 * 1542;     // last term: max over W | 512 of (512 / W + 2) * (W + 2)
}

// (r2) the 3-channel input rows of the block's pixels (+ one halo row / column each side, zero outside the image) are staged in
// shared memory when the block covers whole rows — every tap was a bounds-checked global load before (98 us per launch for
// 50 MB of output) — and the outputs leave as float4.
PCD_HD void stem_conv_body(const StemArgs& a, int bx, int n, float* smem) {
    const int Cout = a.Cout, HW = a.H * a.W, NCG = Cout / 8, NSTRIPMAX = kStemPx / 4;
    float* Wt = smem;
    float* P = Wt + Cout * 27;
    float* P2 = P + 16 * NCG * NSTRIPMAX;
    float* XS = P2 + 16 * NCG * 32 + 16;         // [3][rows + 2][W + 2]
    PCD_FOR(i, Cout * 27) Wt[i] = a.w[i];
    const int p0 = bx * kStemPx;
    const int NT = NCG * NSTRIPMAX;
    const float* xb = a.x + (long long)n * 3 * HW;
    // whole rows per block, strips inside one row, the tile fits the buffer sized for W >= 4
    const bool tiled = (a.W % 4 == 0) && (kStemPx % a.W == 0) && (a.W <= kStemPx) && ((((uintptr_t)a.z) & 15) == 0);
    const int rows = tiled ? kStemPx / a.W : 0, PW = a.W + 2, y0 = tiled ? p0 / a.W : 0;
    if (tiled) {
        PCD_FOR(i, 3 * (rows + 2) * PW) {
            const int ci = i / ((rows + 2) * PW), r = (i - ci * (rows + 2) * PW) / PW, q = i - ci * (rows + 2) * PW - r * PW;
            const int gy = y0 + r - 1, gx = q - 1;
            XS[i] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? xb[(long long)ci * HW + gy * a.W + gx] : 0.f;
        }
    }
    PCD_SYNC();
    PCD_FOR(task, NT) {
        const int cg = task / NSTRIPMAX, strip = task - cg * NSTRIPMAX;
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[i][t] = 0.f;
        const int p = p0 + strip * 4;
        if (tiled) {
            if (p < HW) {
                const int oy = p / a.W, ox = p - oy * a.W;
                for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const float* row = XS + (ci * (rows + 2) + (oy - y0) + ky) * PW + ox;      // image column ox - 1
                        float v[6];
#pragma unroll
                        for (int j = 0; j < 6; ++j) v[j] = row[j];
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float w = Wt[(cg * 8 + i) * 27 + ci * 9 + ky * 3 + kx];
#pragma unroll
                                for (int t = 0; t < 4; ++t) acc[i][t] = fmaf(w, v[t + kx], acc[i][t]);
                            }
                    }
            }
        } else {
            for (int t = 0; t < 4; ++t) {
                if (p + t >= HW) break;
                const int oy = (p + t) / a.W, ox = (p + t) - oy * a.W;
                for (int ci = 0; ci < 3; ++ci)
                    for (int ky = 0; ky < 3; ++ky)
                        for (int kx = 0; kx < 3; ++kx) {
                            const int gy = oy + ky - 1, gx = ox + kx - 1;
                            if (gy < 0 || gy >= a.H || gx < 0 || gx >= a.W) continue;
                            const float v = xb[(long long)ci * HW + gy * a.W + gx];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                acc[i][t] = fmaf(Wt[(cg * 8 + i) * 27 + ci * 9 + ky * 3 + kx], v, acc[i][t]);
                        }
            }
        }
        float* zb = a.z + ((long long)n * Cout + cg * 8) * HW;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = 0.f, q = 0.f;
            if (tiled && p < HW) {
                F4 o = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
                *reinterpret_cast<F4*>(zb + (long long)i * HW + p) = o;
#pragma unroll
                for (int t = 0; t < 4; ++t) { s += acc[i][t]; q = fmaf(acc[i][t], acc[i][t], q); }
            } else if (!tiled) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (p + t < HW) {
                        zb[(long long)i * HW + p + t] = acc[i][t];
                        s += acc[i][t];
                        q = fmaf(acc[i][t], acc[i][t], q);
                    }
            }
            P[(2 * i) * NT + task] = s;
            P[(2 * i + 1) * NT + task] = q;
        }
    }
    reduce_columns(P, P2, 16, NCG, NSTRIPMAX, NT, [&](int grp, int k, float v) {
        pcd_atomic_add(a.stats + (k & 1) * Cout + grp * 8 + (k >> 1), (double)v);
    });
}

// ---- AdaptiveAvgPool2d (model_search.py:129,176) -----------------------------------------------------
struct GapArgs {
    int B, C, H, W, OH, OW, ny;      // ny: planes are strided over gridDim.y (backward)
    const float* x;
    float* y;
};

PCD_HD void gap_fwd_body(const GapArgs& a, int bx) {
    const long long total = (long long)a.B * a.C * a.OH * a.OW;
    PCD_FOR(t, kThreads) {
        const long long i = (long long)bx * kThreads + t;
        if (i < total) {
            const int ox = (int)(i % a.OW), oy = (int)((i / a.OW) % a.OH);
            const long long nc = i / (a.OW * a.OH);
            const int y0 = (oy * a.H) / a.OH, y1 = ((oy + 1) * a.H + a.OH - 1) / a.OH;
            const int x0 = (ox * a.W) / a.OW, x1 = ((ox + 1) * a.W + a.OW - 1) / a.OW;
            float s = 0.f;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) s += a.x[(nc * a.H + y) * a.W + x];
            a.y[i] = s / (float)((y1 - y0) * (x1 - x0));
        }
    }
}

// gx[nc][y][x] = sum over output windows containing (y,x) of gy / window size
// block (bx, by): pixels bx*256.. of planes by, by + nplanes_stride, ...  (32-bit index arithmetic only)
PCD_HD void gap_bwd_body(const GapArgs& a /* x = gy (OHxOW), y = gx (HxW) */, int bx, int by, int ny) {
    const int HW = a.H * a.W, planes = a.B * a.C;
    PCD_FOR(t, kThreads) {
        const int p = bx * kThreads + t;
        if (p < HW) {
            const int y = p / a.W, x = p - y * a.W;
            const int oyc = (y * a.OH) / a.H, oxc = (x * a.OW) / a.W;     // windows overlap by at most one neighbour
            int oyA = -1, oyB = -1, hyA = 1, hyB = 1, oxA = -1, oxB = -1, hxA = 1, hxB = 1;
            for (int oy = (oyc > 0 ? oyc - 1 : 0); oy <= oyc + 1 && oy < a.OH; ++oy) {
                const int y0 = (oy * a.H) / a.OH, y1 = ((oy + 1) * a.H + a.OH - 1) / a.OH;
                if (y < y0 || y >= y1) continue;
                if (oyA < 0) { oyA = oy; hyA = y1 - y0; } else if (oyB < 0) { oyB = oy; hyB = y1 - y0; }
            }
            for (int ox = (oxc > 0 ? oxc - 1 : 0); ox <= oxc + 1 && ox < a.OW; ++ox) {
                const int x0 = (ox * a.W) / a.OW, x1 = ((ox + 1) * a.W + a.OW - 1) / a.OW;
                if (x < x0 || x >= x1) continue;
                if (oxA < 0) { oxA = ox; hxA = x1 - x0; } else if (oxB < 0) { oxB = ox; hxB = x1 - x0; }
            }
            const float dAA = (float)(hyA * hxA), dAB = (float)(hyA * hxB), dBA = (float)(hyB * hxA), dBB = (float)(hyB * hxB);
            for (int nc = by; nc < planes; nc += ny) {
                const float* g = a.x + (long long)nc * a.OH * a.OW;
                float s = 0.f;                 // same order as a scan over (oy, ox)
                if (oyA >= 0 && oxA >= 0) s += g[oyA * a.OW + oxA] / dAA;
                if (oyA >= 0 && oxB >= 0) s += g[oyA * a.OW + oxB] / dAB;
                if (oyB >= 0 && oxA >= 0) s += g[oyB * a.OW + oxA] / dBA;
                if (oyB >= 0 && oxB >= 0) s += g[oyB * a.OW + oxB] / dBB;
                a.y[(long long)nc * HW + p] = s;
            }
        }
    }
}

// ---- channel_shuffle (model_search.py:14-28) ----------------------------------------------------------
struct ShuffleArgs {
    int B, C, HW, groups;
    const float* x;
    float* y;
};

PCD_HD void shuffle_body(const ShuffleArgs& a, int bx, int oc, int n) {
    const int per = a.C / a.groups;
    const int ic = (oc % a.groups) * per + oc / a.groups;
    const float* src = a.x + ((long long)n * a.C + ic) * a.HW;
    float* dst = a.y + ((long long)n * a.C + oc) * a.HW;
    const int p0 = bx * 4096;
    const int npx = (a.HW - p0) < 4096 ? (a.HW - p0) : 4096;
    PCD_FOR(i, npx) dst[p0 + i] = src[p0 + i];
}

}  // namespace pcd

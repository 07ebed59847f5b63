// pcd_edge_v4.cuh — forward kernels of the MixedOp edges (model_search.py:44-58) for the production geometries, v4.
//
// What changed against v2/v3 (pcd_edge.cuh), driven by the round-1 ncu captures (FFMA 13-20 of 128 lanes/clk, 80 % of the
// issued instructions integer addressing / loads / reductions, shared-memory wavefronts 1.5x the FFMA cycles):
//   * ONE block per (edge, image, row tile) runs every stage-A job from a single staged input tile (v2: five blocks, five
//     tile loads); the ReLU is applied once, in place, after the pool job has used the raw values;
//   * depthwise patches are 8 rows x 4 columns (stride 1) so that the halo rows/columns are amortised: shared-memory
//     wavefronts drop from 25.5 to 18.7 per output pixel and channel, below the 17 FFMA cycles they feed;  two units run
//     concurrently on the two halves of the block (5x5 | dilated 5x5, then 3x3 | dilated 3x3: equal tap counts);
//   * stride-2 edges keep the input tile as four parity planes (even/odd rows x even/odd columns): the dilated stride-2
//     convolutions become plain 3x3 / 5x5 stencils on the even/even plane and every access is unit-stride;
//   * every index is a compile-time function of the task id (all five geometries have exactly 128 depthwise and 256
//     pointwise tasks per unit), tap tables are constexpr, loops are fully unrolled: the inner loops are FFMA + LDS;
//   * BatchNorm sums: registers -> warp butterfly -> shared fp32 accumulators -> ONE fp64 atomic per channel and block
//     (v2: three block-wide column reductions with barriers per unit);
//   * the saved depthwise outputs t[] (only the weight-grad jobs read them) are not written in activation-only passes.
// Written in the phase style of pcd_common.cuh (for_tasks + PCD_SYNC), so the CPU emulation build runs the same bodies.
#pragma once
#include "pcd_edge.cuh"

namespace pcd {

struct alignas(8) F2 { float x, y; };

// constexpr helpers usable from host and device code (and from the g++ emulation build)
#if PCD_CUDA
#define PCD_CX __host__ __device__
#else
#define PCD_CX
#endif

// v[i] = p[LO + i], i in [0, HI - LO]; p is 16-byte aligned; LO in {0,-1,-2,-4}, HI in {3,4,5,7}
template <int LO, int HI>
PCD_HD void load_seg(const float* PCD_RESTRICT p, float (&v)[HI - LO + 1]) {
    static_assert(LO == 0 || LO == -1 || LO == -2 || LO == -4, "left extent");
    static_assert(HI == 3 || HI == 4 || HI == 5 || HI == 7, "right extent");
    if (LO == -4) { const F4 t = ld4(p - 4); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if (LO == -2) { const F2 t = *reinterpret_cast<const F2*>(p - 2); v[0] = t.x; v[1] = t.y; }
    else if (LO == -1) v[0] = p[-1];
    { const F4 t = ld4(p); v[-LO] = t.x; v[-LO + 1] = t.y; v[-LO + 2] = t.z; v[-LO + 3] = t.w; }
    if (HI == 7) { const F4 t = ld4(p + 4); v[4 - LO] = t.x; v[5 - LO] = t.y; v[6 - LO] = t.z; v[7 - LO] = t.w; }
    else if (HI == 5) { const F2 t = *reinterpret_cast<const F2*>(p + 4); v[4 - LO] = t.x; v[5 - LO] = t.y; }
    else if (HI == 4) v[4 - LO] = p[4];
}
PCD_CX constexpr int seg_lo(int lo) { return lo >= 0 ? 0 : lo == -1 ? -1 : lo == -2 ? -2 : -4; }
PCD_CX constexpr int seg_hi(int hi) { return hi <= 3 ? 3 : hi == 4 ? 4 : hi == 5 ? 5 : 7; }

// Taps of a KS x KS depthwise stencil with dilation DIL and stride S along one axis.  Tap k reads input coordinate
// S*o + off(k).  For S == 2 the input is stored as parity planes: coordinate 2*i + par lives at index i of plane `par`, so
// tap k reads plane par(k) at index o + pos(k).  For S == 1 there is one plane and pos(k) = off(k).
template <int KS, int DIL, int S>
struct TapGeo {
    static constexpr int PAD = DIL * (KS - 1) / 2;
    PCD_CX static constexpr int off(int k) { return k * DIL - PAD; }
    PCD_CX static constexpr int par(int k) { return S == 1 ? 0 : (off(k) & 1); }
    PCD_CX static constexpr int pos(int k) { return S == 1 ? off(k) : (off(k) - par(k)) / 2; }
    PCD_CX static constexpr bool any(int a) {
        for (int k = 0; k < KS; ++k) if (par(k) == a) return true;
        return false;
    }
    PCD_CX static constexpr int pmin(int a) {
        int m = 99;
        for (int k = 0; k < KS; ++k) if (par(k) == a && pos(k) < m) m = pos(k);
        return m;
    }
    PCD_CX static constexpr int pmax(int a) {
        int m = -99;
        for (int k = 0; k < KS; ++k) if (par(k) == a && pos(k) > m) m = pos(k);
        return m;
    }
    PCD_CX static constexpr int find(int a, int f) {
        for (int k = 0; k < KS; ++k) if (par(k) == a && pos(k) == f) return k;
        return -1;
    }
};

// acc[oy][j] += sum over the taps (ky, kx) that live on plane (A, B) of w[ky][kx] * plane[oy + pos(ky)][j + pos(kx)];
// `base` points at the plane element of (patch row 0, patch column 0) — 16-byte aligned; P = plane pitch.
template <int KS, int DIL, int S, int PR, int A, int B, int P>
PCD_HD void dw_plane(const float* PCD_RESTRICT base, const float (&w)[KS * KS], float (&acc)[PR][4]) {
    using G = TapGeo<KS, DIL, S>;
    if constexpr (G::any(A) && G::any(B)) {
        constexpr int RMIN = G::pmin(A), RMAX = G::pmax(A) + PR - 1;
        constexpr int LO = seg_lo(G::pmin(B)), HI = seg_hi(3 + G::pmax(B));
#pragma unroll
        for (int rr = RMIN; rr <= RMAX; ++rr) {
            float v[HI - LO + 1];
            load_seg<LO, HI>(base + rr * P, v);
#pragma unroll
            for (int oy = 0; oy < PR; ++oy) {
                const int ky = G::find(A, rr - oy);
                if (ky < 0) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    if (G::par(kx) != B) continue;
                    const int g = G::pos(kx) - LO;
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[oy][j] = fmaf(w[ky * KS + kx], v[j + g], acc[oy][j]);
                }
            }
        }
    }
}

// z[i][t] = sum_ci Wm[(cg*4 + i)*C + ci] * tin(ci)[t]      (4 output channels x 4 pixels; tin(ci) -> 16-byte aligned strip)
template <int C, class TF>
PCD_HD void pw_tile(const float* PCD_RESTRICT Wm, int cg, TF tin, float (&z)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t) z[i][t] = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
        F4 tv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) tv[k] = ld4(tin(c4 * 4 + k));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const F4 w = ld4(Wm + (cg * 4 + i) * C + c4 * 4);
            const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                z[i][0] = fmaf(wk[k], tv[k].x, z[i][0]);
                z[i][1] = fmaf(wk[k], tv[k].y, z[i][1]);
                z[i][2] = fmaf(wk[k], tv[k].z, z[i][2]);
                z[i][3] = fmaf(wk[k], tv[k].w, z[i][3]);
            }
        }
    }
}

// ---- block-level accumulation of per-task partial sums --------------------------------------------------------------
// Every lane of an aligned group of GROUP lanes contributes v[0..NV) to the SAME NV shared accumulators acc[idx(k)].
// CUDA: butterfly over the group, then one shared-memory atomic per value from one lane of the group.  Emulation: the
// tasks run one after the other, so a plain add.
template <int NV, int GROUP, class IdxF>
PCD_HD void group_accumulate(float* acc, IdxF idx, float (&v)[NV]) {
#if PCD_CUDA
    static_assert(GROUP == 4 || GROUP == 8 || GROUP == 16 || GROUP == 32, "group");
    if constexpr (NV == 8 && GROUP == 32) {
        // transposed butterfly: 8 -> 4 -> 2 -> 1 live values (7 shuffles), then two plain steps: 9 shuffles instead of 40
        const unsigned lane = threadIdx.x & 31u;
        float a4[4], a2[2], a1;
        {
            const bool hi = (lane & 16u) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float keep = hi ? v[i + 4] : v[i], send = hi ? v[i] : v[i + 4];
                a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool hi = (lane & 8u) != 0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float keep = hi ? a4[i + 2] : a4[i], send = hi ? a4[i] : a4[i + 2];
                a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
        {
            const bool hi = (lane & 4u) != 0;
            const float keep = hi ? a2[1] : a2[0], send = hi ? a2[0] : a2[1];
            a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        if ((lane & 3u) == 0) {
            const int k = (int)(((lane >> 4) & 1u) * 4 + ((lane >> 3) & 1u) * 2 + ((lane >> 2) & 1u));
            atomicAdd(acc + idx(k), a1);
        }
    } else {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float s = v[k];
#pragma unroll
            for (int o = GROUP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & (GROUP - 1)) == 0) atomicAdd(acc + idx(k), s);
        }
    }
#else
    for (int k = 0; k < NV; ++k) acc[idx(k)] += v[k];
#endif
}

// ---- geometry of one (C, S, W) production edge shape -------------------------------------------------------------------
template <int C, int S, int W>
struct V4Geo {
    static constexpr int TH = (S == 1) ? 16 : 8;            // output rows per tile
    static constexpr int PR = (S == 1) ? 8 : 4;             // output rows per depthwise patch
    static constexpr int NPIX = TH * W, NSTRIP = NPIX / 4;
    static constexpr int WS4 = W / 4;                       // 4-pixel strips per output row
    static constexpr int DW_TASKS = C * WS4 * (TH / PR);    // per unit
    static constexpr int PW_TASKS = (C / 4) * NSTRIP;       // per unit
    static_assert(DW_TASKS == 128, "two units run on the two halves of a 256-thread block");
    // ---- stride 1: one plane per channel, rows oy0-4 .. oy0+TH+3, data columns at [4, 4+W) ----
    static constexpr int P1 = W + 8 + (W == 16 ? 4 : 0);    // pitch
    static constexpr int IH1 = TH + 8;
    // W == 16: 4 strips per row, so 8 consecutive lanes span two channels; +16 floats between channels keeps them on
    // different banks (rows of one channel are 0 mod 32 banks apart whatever the pitch)
    static constexpr int CS1 = IH1 * P1 + (W == 16 ? 16 : 0);
    // ---- stride 2: four parity planes per channel ----
    //   even rows: plane rows oy0-2 .. oy0+TH+1 (RE), odd rows: oy0-1 .. oy0+TH-1 (RO)
    //   even cols: data at [4, 4+W), halo 4 both sides (PE); odd cols: data at [4, 4+W), left halo only (PO)
    static constexpr int RE = TH + 4, RO = TH + 1;
    static constexpr int PE = W + 8 + (W == 16 ? 4 : 0), PO = W + 4;
    static constexpr int OFF_EE = 0, OFF_EO = RE * PE, OFF_OE = OFF_EO + RE * PO, OFF_OO = OFF_OE + RO * PE;
    static constexpr int CS2 = OFF_OO + RO * PO;
    static constexpr int XIN_FLOATS = C * (S == 1 ? CS1 : CS2);
    static constexpr int T_FLOATS = C * NPIX;               // one unit's depthwise output [C][TH][W]
    static constexpr int PAR_FLOATS = (S == 2 ? C * C : 0) + 102 * C + 6 * C * C;      // edge_param_floats(C, S)
    static constexpr int NBN = 8 + (S == 2);
    static constexpr int SACC_FLOATS = NBN * 2 * C;
    static constexpr size_t SMEM_FLOATS = (size_t)XIN_FLOATS + 2 * T_FLOATS + PAR_FLOATS + SACC_FLOATS;

    // depthwise task -> (channel, row block, strip).  W == 16: channel inside the row block (see CS1)
    PCD_CX static inline void dw_task(int t, int& ch, int& rb, int& strip) {
        strip = t % WS4;
        if (W == 16 && S == 1) { ch = (t / WS4) % C; rb = t / (WS4 * C); }
        else { rb = (t / WS4) % (TH / PR); ch = t / (WS4 * (TH / PR)); }
    }
    // plane (A, B) of channel ch at local output row r, output column x   (stride 1: A = B = 0)
    PCD_CX static inline int plane_off(int A, int B, int ch, int r, int x) {
        if (S == 1) return ch * CS1 + (r + 4) * P1 + 4 + x;
        const int base = ch * CS2 + (A ? (B ? OFF_OO : OFF_OE) : (B ? OFF_EO : OFF_EE));
        return base + (r + (A ? 1 : 2)) * (B ? PO : PE) + 4 + x;
    }
};

struct FwdV4Args {
    int B, Hs, Ws, Ho, Wo;
    float eps;
    int nedges, save_t, jobs;      // jobs: 1 = one block runs every stage-A job; 2 = {5x5 pair} | {pools, FR, 3x3 pair}
    EdgeF e[kMaxEdgesPerLaunch];
};

// one depthwise unit on this thread's patch: acc -> T (shared, [C][TH][W]) and, when asked, the saved slot of the image
// (t_img = slot + n*C*Ho*W, rows oy0 + ...)
template <int C, int S, int W, int KS, int DIL, int U>
PCD_HD void v4_dw_unit(const float* XIN, const float* PAR, int tt, float* T, float* t_img, int Ho, int oy0) {
    using G = V4Geo<C, S, W>;
    constexpr int PR = G::PR;
    int ch, rb, strip;
    G::dw_task(tt, ch, rb, strip);
    const int py = rb * PR, px = strip * 4;
    float w[KS * KS];
    const float* wp = PAR + edge_dw_off(C, S, U) + ch * KS * KS;
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) w[i] = wp[i];
    float acc[PR][4];
#pragma unroll
    for (int i = 0; i < PR; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (S == 1) {
        dw_plane<KS, DIL, 1, PR, 0, 0, G::P1>(XIN + G::plane_off(0, 0, ch, py, px), w, acc);
    } else {
        dw_plane<KS, DIL, 2, PR, 0, 0, G::PE>(XIN + G::plane_off(0, 0, ch, py, px), w, acc);
        dw_plane<KS, DIL, 2, PR, 0, 1, G::PO>(XIN + G::plane_off(0, 1, ch, py, px), w, acc);
        dw_plane<KS, DIL, 2, PR, 1, 0, G::PE>(XIN + G::plane_off(1, 0, ch, py, px), w, acc);
        dw_plane<KS, DIL, 2, PR, 1, 1, G::PO>(XIN + G::plane_off(1, 1, ch, py, px), w, acc);
    }
#pragma unroll
    for (int i = 0; i < PR; ++i) {
        st4(T + (ch * G::TH + py + i) * W + px, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (t_img) st4(t_img + ((long long)ch * Ho + oy0 + py + i) * W + px, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// pointwise conv of one unit on a 4-channel x 4-pixel tile of T -> pre-BN output z (saved slot) + its channel sums
template <int C, int W, int TH, class TF>
PCD_HD void v4_pw_task(const float* Wm, int tt, TF tin, float* z_img, int Ho, int oy0, float* sacc /* [2][C] of this BN */) {
    constexpr int NSTRIP = TH * W / 4, WS4 = W / 4;
    const int cg = tt / NSTRIP, strip = tt % NSTRIP;
    float z[4][4];
    pw_tile<C>(Wm, cg, [&](int ci) { return tin(ci, strip); }, z);
    const int oy = oy0 + strip / WS4, ox = (strip % WS4) * 4;
    float sq[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        st4(z_img + ((long long)(cg * 4 + i) * Ho + oy) * W + ox, z[i][0], z[i][1], z[i][2], z[i][3]);
        sq[i] = (z[i][0] + z[i][1]) + (z[i][2] + z[i][3]);
        sq[4 + i] = fmaf(z[i][0], z[i][0], fmaf(z[i][1], z[i][1], fmaf(z[i][2], z[i][2], z[i][3] * z[i][3])));
    }
    static_assert(NSTRIP % 32 == 0, "a warp must stay inside one channel group");
    group_accumulate<8, 32>(sacc, [&](int k) { return (k >> 2) * C + cg * 4 + (k & 3); }, sq);
}

// max / avg pool partial of one parity plane (3x3, pad 1, stride S; operations.py:6-7)
template <int S, int PRP, int A, int B, int P>
PCD_HD void pool_plane(const float* PCD_RESTRICT base, const bool (&rok)[PRP][3], bool left, bool right,
                       float (&mx)[PRP][4], float (&sm)[PRP][4]) {
    using G = TapGeo<3, 1, S>;
    if constexpr (G::any(A) && G::any(B)) {
        constexpr int RMIN = G::pmin(A), RMAX = G::pmax(A) + PRP - 1;
        constexpr int LO = seg_lo(G::pmin(B)), HI = seg_hi(3 + G::pmax(B));
#pragma unroll
        for (int rr = RMIN; rr <= RMAX; ++rr) {
            float v[HI - LO + 1];
            load_seg<LO, HI>(base + rr * P, v);
#pragma unroll
            for (int oy = 0; oy < PRP; ++oy) {
                const int ky = G::find(A, rr - oy);
                if (ky < 0) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    if (G::par(kx) != B) continue;
                    const int g = G::pos(kx) - LO;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float val = v[j + g];
                        bool ok = rok[oy][ky];
                        if (kx == 0 && j == 0) ok = ok && !left;
                        if (kx == 2 && j == 3) ok = ok && !right;
                        sm[oy][j] += val;                         // out-of-image taps read the zero halo
                        mx[oy][j] = fmaxf(mx[oy][j], ok ? val : -INFINITY);
                    }
                }
            }
        }
    }
}

// ======================================================================================================
// stage A: pools (+FactorizedReduce), first halves of the separable convs, both dilated convs
// ======================================================================================================
template <int C, int S, int W>
PCD_HD void fwdA4_body(const FwdV4Args& a, int tile, int n, int z, float* smem) {
    using G = V4Geo<C, S, W>;
    constexpr int TH = G::TH, WS4 = G::WS4;
    const int ez = z / a.jobs, job = z - ez * a.jobs;
    const bool do_big = (a.jobs == 1) || job == 0;           // 5x5 | dilated 5x5
    const bool do_small = (a.jobs == 1) || job == 1;         // pools, FactorizedReduce, 3x3 | dilated 3x3
    const EdgeF& e = a.e[ez];
    float* XIN = smem;
    float* T0 = XIN + G::XIN_FLOATS;
    float* T1 = T0 + G::T_FLOATS;
    float* PAR = T1 + G::T_FLOATS;
    float* SACC = PAR + G::PAR_FLOATS;
    const int oy0 = tile * TH, Ho = a.Ho;
    const long long HWo = (long long)Ho * W, nslot = (long long)a.B * C * HWo;
    const float* src = e.x + (long long)n * e.x_ns;
    const long long scs = (long long)a.Hs * a.Ws;
    // ---- 1. stage the edge's parameters and the input tile --------------------------------------------------------
    constexpr int NPARF = G::PAR_FLOATS;
    static_assert(NPARF % 4 == 0, "parameter block is a whole number of float4");
    for_tasks<NPARF / 4>([&](int i) { cp16(PAR + 4 * i, e.par + 4 * i, true); });
    for_tasks<G::SACC_FLOATS>([&](int i) { SACC[i] = 0.f; });
    if (S == 1) {
        for_tasks<C * G::IH1 * WS4>([&](int i) {
            const int x4 = i % WS4, r = (i / WS4) % G::IH1, ch = i / (WS4 * G::IH1);
            const int gy = oy0 - 4 + r;
            const bool ok = gy >= 0 && gy < a.Hs;
            cp16(XIN + ch * G::CS1 + r * G::P1 + 4 + 4 * x4, ok ? src + ch * scs + (long long)gy * W + 4 * x4 : src, ok);
        });
        constexpr int NH = (G::P1 - W) / 4;                  // halo float4 per row: 1 left, the rest right
        for_tasks<C * G::IH1 * NH>([&](int i) {
            const int h = i % NH, row = i / NH, r = row % G::IH1, ch = row / G::IH1;
            st4(XIN + ch * G::CS1 + r * G::P1 + (h == 0 ? 0 : W + 4 * h), 0.f, 0.f, 0.f, 0.f);
        });
    } else {
        // parity planes: input row 2*i + A -> row i of the planes (A, .); an 8-float span of it -> 4 even + 4 odd columns
        constexpr int NR = G::RE + G::RO;
        for_tasks<C * NR * WS4>([&](int i) {
            const int x8 = i % WS4, rr = (i / WS4) % NR, ch = i / (WS4 * NR);
            const int A = rr >= G::RE ? 1 : 0, lr = A ? rr - G::RE : rr;
            const int gy = A ? 2 * (oy0 - 1 + lr) + 1 : 2 * (oy0 - 2 + lr);
            F4 v0 = {0.f, 0.f, 0.f, 0.f}, v1 = {0.f, 0.f, 0.f, 0.f};
            if (gy >= 0 && gy < a.Hs) {
                const float* p = src + ch * scs + (long long)gy * (2 * W) + 8 * x8;
                v0 = ld4(p);
                v1 = ld4(p + 4);
            }
            float* pe = XIN + ch * G::CS2 + (A ? G::OFF_OE : G::OFF_EE) + lr * G::PE + 4 + 4 * x8;
            float* po = XIN + ch * G::CS2 + (A ? G::OFF_OO : G::OFF_EO) + lr * G::PO + 4 + 4 * x8;
            st4(pe, v0.x, v0.z, v1.x, v1.z);
            st4(po, v0.y, v0.w, v1.y, v1.w);
        });
        constexpr int NHE = (G::PE - W) / 4;                 // even-column planes: 1 left + the rest right; odd: 1 left
        for_tasks<C * NR * (NHE + 1)>([&](int i) {
            const int h = i % (NHE + 1), row = i / (NHE + 1), rr = row % NR, ch = row / NR;
            const int A = rr >= G::RE ? 1 : 0, lr = A ? rr - G::RE : rr;
            float* pe = XIN + ch * G::CS2 + (A ? G::OFF_OE : G::OFF_EE) + lr * G::PE;
            float* po = XIN + ch * G::CS2 + (A ? G::OFF_OO : G::OFF_EO) + lr * G::PO;
            if (h == NHE) st4(po, 0.f, 0.f, 0.f, 0.f);
            else st4(pe + (h == 0 ? 0 : W + 4 * h), 0.f, 0.f, 0.f, 0.f);
        });
    }
    cp16_wait();
    PCD_SYNC();
    // ---- 2. max / avg pool on the raw tile ---------------------------------------------------------------------------
    if (do_small) {
        constexpr int PRP = (S == 1) ? 4 : 2, NRB = TH / PRP;
        static_assert(C * NRB * WS4 == kThreads, "one pool patch per thread");
        constexpr int LANES_PER_CH = NRB * WS4;
        constexpr int GROUP = LANES_PER_CH >= 32 ? 32 : LANES_PER_CH;
        float* p1_img = e.saved + slot_p1() * nslot + (long long)n * C * HWo;
        float* p2_img = e.saved + slot_p2() * nslot + (long long)n * C * HWo;
        for_tasks<kThreads>([&](int t) {
            const int strip = t % WS4, rb = (t / WS4) % NRB, ch = t / (WS4 * NRB);
            const int py = rb * PRP, px = strip * 4;
            const bool left = px == 0, right = (S == 1) && (px == W - 4);
            bool rok[PRP][3];
            float mx[PRP][4], sm[PRP][4];
#pragma unroll
            for (int i = 0; i < PRP; ++i) {
                const int oy = oy0 + py + i;
                rok[i][0] = oy > 0;
                rok[i][1] = true;
                rok[i][2] = (S == 2) || (oy < Ho - 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) { mx[i][j] = -INFINITY; sm[i][j] = 0.f; }
            }
            if (S == 1) {
                pool_plane<1, PRP, 0, 0, G::P1>(XIN + G::plane_off(0, 0, ch, py, px), rok, left, right, mx, sm);
            } else {
                pool_plane<2, PRP, 0, 0, G::PE>(XIN + G::plane_off(0, 0, ch, py, px), rok, left, right, mx, sm);
                pool_plane<2, PRP, 0, 1, G::PO>(XIN + G::plane_off(0, 1, ch, py, px), rok, left, right, mx, sm);
                pool_plane<2, PRP, 1, 0, G::PE>(XIN + G::plane_off(1, 0, ch, py, px), rok, left, right, mx, sm);
                pool_plane<2, PRP, 1, 1, G::PO>(XIN + G::plane_off(1, 1, ch, py, px), rok, left, right, mx, sm);
            }
            float st[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < PRP; ++i) {
                const int nrow = (int)rok[i][0] + 1 + (int)rok[i][2];
                float av[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ncol = 3 - ((j == 0 && left) ? 1 : 0) - ((j == 3 && right) ? 1 : 0);
                    av[j] = sm[i][j] / (float)(nrow * ncol);
                    st[0] += mx[i][j]; st[1] = fmaf(mx[i][j], mx[i][j], st[1]);
                    st[2] += av[j]; st[3] = fmaf(av[j], av[j], st[3]);
                }
                const long long o = ((long long)ch * Ho + oy0 + py + i) * W + px;
                st4(p1_img + o, mx[i][0], mx[i][1], mx[i][2], mx[i][3]);
                st4(p2_img + o, av[0], av[1], av[2], av[3]);
            }
            // bn_p1 = 0, bn_p2 = 1: [sum P1 | sumsq P1 | sum P2 | sumsq P2] rows of C
            group_accumulate<4, GROUP>(SACC, [&](int k) { return k * C + ch; }, st);
        });
    }
    PCD_SYNC();
    // ---- 3. ReLU in place (every conv candidate starts with it; halo zeros stay zeros) ------------------------------
    for_tasks<G::XIN_FLOATS / 4>([&](int i) {
        const F4 v = ld4(XIN + 4 * i);
        st4(XIN + 4 * i, relu(v.x), relu(v.y), relu(v.z), relu(v.w));
    });
    PCD_SYNC();
    // ---- 4. skip_connect at stride 2 = FactorizedReduce (operations.py:90-104): 1x1 convs on the EE / OO planes ----------
    if (S == 2 && do_small) {
        float* f_img = e.saved + slot_f() * nslot + (long long)n * C * HWo;
        for_tasks<G::PW_TASKS>([&](int tt) {
            const int cg = tt / G::NSTRIP;
            const int A = (cg * 4 >= C / 2) ? 1 : 0;          // conv_2 samples x[:, :, 1:, 1:] (odd rows, odd columns)
            v4_pw_task<C, W, TH>(PAR, tt, [&](int ci, int strip) {
                return XIN + G::plane_off(A, A, ci, strip / WS4, (strip % WS4) * 4);
            }, f_img, Ho, oy0, SACC + bn_f() * 2 * C);
        });
    }
    // ---- 5. depthwise -> pointwise units, two at a time ---------------------------------------------------------------
    auto tslot = [&](int u) -> float* { return a.save_t ? e.saved + slot_t(u) * nslot + (long long)n * C * HWo : nullptr; };
    auto zslot = [&](int u) -> float* { return e.saved + slot_z(u) * nslot + (long long)n * C * HWo; };
    if (do_big) {
        for_tasks<kThreads>([&](int t) {
            if (t < 128) v4_dw_unit<C, S, W, 5, 1, 2>(XIN, PAR, t, T0, tslot(2), Ho, oy0);
            else v4_dw_unit<C, S, W, 5, 2, 5>(XIN, PAR, t - 128, T1, tslot(5), Ho, oy0);
        });
        PCD_SYNC();
        for_tasks<2 * G::PW_TASKS>([&](int t) {
            const int which = t / G::PW_TASKS, tt = t % G::PW_TASKS;
            const float* T = which ? T1 : T0;
            const int u = which ? 5 : 2;
            v4_pw_task<C, W, TH>(PAR + (which ? edge_pw_off(C, S, 5) : edge_pw_off(C, S, 2)), tt,
                                 [&](int ci, int strip) { return T + ci * G::NPIX + strip * 4; }, zslot(u), Ho, oy0,
                                 SACC + bn_unit(S, u) * 2 * C);
        });
        PCD_SYNC();
    }
    if (do_small) {
        for_tasks<kThreads>([&](int t) {
            if (t < 128) v4_dw_unit<C, S, W, 3, 1, 0>(XIN, PAR, t, T0, tslot(0), Ho, oy0);
            else v4_dw_unit<C, S, W, 3, 2, 4>(XIN, PAR, t - 128, T1, tslot(4), Ho, oy0);
        });
        PCD_SYNC();
        for_tasks<2 * G::PW_TASKS>([&](int t) {
            const int which = t / G::PW_TASKS, tt = t % G::PW_TASKS;
            const float* T = which ? T1 : T0;
            const int u = which ? 4 : 0;
            v4_pw_task<C, W, TH>(PAR + (which ? edge_pw_off(C, S, 4) : edge_pw_off(C, S, 0)), tt,
                                 [&](int ci, int strip) { return T + ci * G::NPIX + strip * 4; }, zslot(u), Ho, oy0,
                                 SACC + bn_unit(S, u) * 2 * C);
        });
        PCD_SYNC();
    }
    // ---- 6. one fp64 atomic per (BN, moment, channel) this block produced ---------------------------------------------
    for_tasks<G::SACC_FLOATS>([&](int i) {
        const int bn = i / (2 * C);
        const bool small_bn = bn == bn_p1() || bn == bn_p2() || (S == 2 && bn == bn_f()) || bn == bn_unit(S, 0) || bn == bn_unit(S, 4);
        const bool big_bn = bn == bn_unit(S, 2) || bn == bn_unit(S, 5);
        if ((small_bn && do_small) || (big_bn && do_big)) pcd_atomic_add(e.stats + i, (double)SACC[i]);
    });
}

// ======================================================================================================
// stage B: second half of a separable conv: BN -> ReLU -> depthwise (stride 1) -> pointwise   (operations.py:58-62)
// one block = (edge, image, 16-row tile, half); 4x4 depthwise patches (256 tasks)
// ======================================================================================================
template <int C, int W>
struct V4GeoB {
    static constexpr int TH = 16, PR = 4, WS4 = W / 4, NPIX = TH * W, NSTRIP = NPIX / 4;
    static constexpr int P = W + 8 + (W == 16 ? 4 : 0);      // 4 rows * 28 = 16 mod 32 banks: the two row blocks a quarter warp
    static constexpr int RH = TH + 4;                         // spans (W == 16) sit on different banks
    static constexpr int CS = RH * P;
    static constexpr int Q_FLOATS = C * CS, T_FLOATS = C * NPIX;
    static constexpr int PAR_FLOATS = 25 * C + C * C;
    static constexpr size_t SMEM_FLOATS = (size_t)Q_FLOATS + T_FLOATS + PAR_FLOATS + 2 * C + 2 * C;
    static_assert(C * (TH / PR) * WS4 == kThreads, "one depthwise patch per thread");
};

template <int C, int W, int KS, int UB>
PCD_HD void fwdB4_job(const FwdV4Args& a, const EdgeF& e, int tile, int n, float* smem) {
    using G = V4GeoB<C, W>;
    constexpr int TH = G::TH, WS4 = G::WS4, PR = G::PR, UA = UB - 1, HY = 2;
    float* Q = smem;
    float* T = Q + G::Q_FLOATS;
    float* PAR = T + G::T_FLOATS;                 // [dw KS*KS*C][pw C*C]
    float* BNC = PAR + G::PAR_FLOATS;             // mean, rstd of BN-A
    float* SACC = BNC + 2 * C;                    // [2][C]
    const int oy0 = tile * TH, Ho = a.Ho;
    const long long HWo = (long long)Ho * W, nslot = (long long)a.B * C * HWo;
    const int S = (a.Hs == a.Ho) ? 1 : 2;         // stride of the EDGE (parameter / BN numbering); stage B itself is stride 1
    const float* src = e.saved + slot_z(UA) * nslot + (long long)n * C * HWo;
    // raw zA rows start their way into shared memory before the BN constants are derived
    for_tasks<C * G::RH * WS4>([&](int i) {
        const int x4 = i % WS4, r = (i / WS4) % G::RH, ch = i / (WS4 * G::RH);
        const int gy = oy0 - HY + r;
        const bool ok = gy >= 0 && gy < Ho;
        cp16(Q + ch * G::CS + r * G::P + 4 + 4 * x4, ok ? src + ch * HWo + (long long)gy * W + 4 * x4 : src, ok);
    });
    constexpr int NH = (G::P - W) / 4;
    for_tasks<C * G::RH * NH>([&](int i) {
        const int h = i % NH, row = i / NH;
        st4(Q + row * G::P + (h == 0 ? 0 : W + 4 * h), 0.f, 0.f, 0.f, 0.f);
    });
    const float* par = e.par + edge_dw_off(C, S, UB);
    for_tasks<(KS * KS * C + C * C) / 4>([&](int i) { cp16(PAR + 4 * i, par + 4 * i, true); });
    const double cnt = (double)a.B * Ho * W;
    PCD_FOR(j, C) {
        BnC b = bn_consts(e.stats, C, bn_unit(S, UA), j, cnt, a.eps);
        BNC[2 * j] = b.mean;
        BNC[2 * j + 1] = b.rstd;
        SACC[j] = 0.f;
        SACC[C + j] = 0.f;
    }
    cp16_wait();
    PCD_SYNC();
    for_tasks<C * G::RH * WS4>([&](int i) {       // BN + ReLU in place on the rows inside the image (the rest stay zero padding)
        const int x4 = i % WS4, r = (i / WS4) % G::RH, ch = i / (WS4 * G::RH);
        const int gy = oy0 - HY + r;
        if (gy >= 0 && gy < Ho) {
            float* p = Q + ch * G::CS + r * G::P + 4 + 4 * x4;
            const F4 v = ld4(p);
            const float m = BNC[2 * ch], rs = BNC[2 * ch + 1];
            st4(p, relu((v.x - m) * rs), relu((v.y - m) * rs), relu((v.z - m) * rs), relu((v.w - m) * rs));
        }
    });
    PCD_SYNC();
    float* t_img = a.save_t ? e.saved + slot_t(UB) * nslot + (long long)n * C * HWo : nullptr;
    for_tasks<kThreads>([&](int t) {
        const int strip = t % WS4, rb = (t / WS4) % (TH / PR), ch = t / (WS4 * (TH / PR));
        const int py = rb * PR, px = strip * 4;
        float w[KS * KS];
#pragma unroll
        for (int i = 0; i < KS * KS; ++i) w[i] = PAR[ch * KS * KS + i];
        float acc[PR][4];
#pragma unroll
        for (int i = 0; i < PR; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        dw_plane<KS, 1, 1, PR, 0, 0, G::P>(Q + ch * G::CS + (py + HY) * G::P + 4 + px, w, acc);
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            st4(T + (ch * TH + py + i) * W + px, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (t_img) st4(t_img + ((long long)ch * Ho + oy0 + py + i) * W + px, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    });
    PCD_SYNC();
    float* z_img = e.saved + slot_z(UB) * nslot + (long long)n * C * HWo;
    for_tasks<(C / 4) * G::NSTRIP>([&](int tt) {
        v4_pw_task<C, W, TH>(PAR + KS * KS * C, tt, [&](int ci, int strip) { return T + ci * G::NPIX + strip * 4; }, z_img, Ho, oy0, SACC);
    });
    PCD_SYNC();
    PCD_FOR(i, 2 * C) pcd_atomic_add(e.stats + bn_unit(S, UB) * 2 * C + i, (double)SACC[i]);
}

template <int C, int W>
PCD_HD void fwdB4_body(const FwdV4Args& a, int tile, int n, int z, float* smem) {
    if ((z & 1) == 0) fwdB4_job<C, W, 3, 1>(a, a.e[z >> 1], tile, n, smem);
    else fwdB4_job<C, W, 5, 3>(a, a.e[z >> 1], tile, n, smem);
}

}  // namespace pcd

// pcd_common.cuh — shared definitions for the PC-DARTS sm_100a kernels.
//
// Every kernel body in pcd_fwd.cuh / pcd_bwd.cuh is written as a sequence of block-wide "phases":
//     PCD_FOR(task, n) { ... }   PCD_SYNC();
// with all cross-phase state in shared memory.  nvcc compiles the phases as strided thread loops
// separated by __syncthreads(); the test-only CPU build (tests/emu, -DPCD_EMU) runs each phase as a
// plain loop over tasks, which lets the index arithmetic be checked against the oracle without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__) && !defined(PCD_EMU)
#define PCD_CUDA 1
#include <cuda_runtime.h>
#define PCD_HD __device__ __forceinline__
#define PCD_HOSTDEV __host__ __device__ __forceinline__
#define PCD_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define PCD_SYNC() __syncthreads()
#define PCD_RESTRICT __restrict__
PCD_HD void pcd_atomic_add(double* p, double v) { atomicAdd(p, v); }
PCD_HD void pcd_atomic_add(float* p, float v) { atomicAdd(p, v); }
#else
#define PCD_CUDA 0
#include <string.h>
#define PCD_HD static inline
#define PCD_HOSTDEV static inline
#define PCD_FOR(i, n) for (int i = 0; i < (n); ++i)
#define PCD_SYNC() ((void)0)
#define PCD_RESTRICT
PCD_HD void pcd_atomic_add(double* p, double v) { *p += v; }
PCD_HD void pcd_atomic_add(float* p, float v) { *p += v; }
#endif

namespace pcd {

struct alignas(16) F4 { float x, y, z, w; };

constexpr int kThreads = 256;
}  // namespace pcd

// Thread-private state that must survive a barrier: one copy per emulated thread in the CPU emulation build
#if PCD_CUDA
#define PCD_TSTATE(type, name, dims) type name dims
#define PCD_TREF(name, tid) name
#define PCD_TPASS(name) name                    // the whole state, as a function argument / parameter name
#else
#define PCD_TSTATE(type, name, dims) type name##_all[pcd::kThreads] dims
#define PCD_TREF(name, tid) name##_all[tid]
#define PCD_TPASS(name) name##_all
#endif
#define PCD_EACH(task) PCD_FOR(task, pcd::kThreads)

namespace pcd {
constexpr int kMaxEdgesPerLaunch = 8;
constexpr int kUnits = 6;
constexpr int PCD_MAX_EDGES_CONST = 14;   // A3 B3 A5 B5 D3 D5  (sep3 first/second half, sep5, dil3, dil5)

// ---- per-edge (MixedOp) layouts; c = C/4 partial channels, s = stride -------------------------
PCD_HOSTDEV int unit_ks(int u) { return (u == 0 || u == 1 || u == 4) ? 3 : 5; }
PCD_HOSTDEV int edge_fr_floats(int c, int s) { return s == 2 ? c * c : 0; }
PCD_HOSTDEV int edge_dw_off(int c, int s, int u) {
    int o = edge_fr_floats(c, s);
    for (int v = 0; v < u; ++v) o += unit_ks(v) * unit_ks(v) * c + c * c;
    return o;
}
PCD_HOSTDEV int edge_pw_off(int c, int s, int u) { return edge_dw_off(c, s, u) + unit_ks(u) * unit_ks(u) * c; }
PCD_HOSTDEV int edge_param_floats(int c, int s) { return edge_fr_floats(c, s) + 102 * c + 6 * c * c; }
PCD_HOSTDEV int edge_nbn(int s) { return 8 + (s == 2); }
// BN ids follow the registration order of MixedOp buffers (model_search.py:37-41):
//   P1 (max_pool), P2 (avg_pool), [F (FactorizedReduce, stride 2)], A3, B3, A5, B5, D3, D5
PCD_HOSTDEV int bn_p1() { return 0; }
PCD_HOSTDEV int bn_p2() { return 1; }
PCD_HOSTDEV int bn_f() { return 2; }
PCD_HOSTDEV int bn_unit(int s, int u) { return 2 + (s == 2) + u; }
// saved activation slots (each B*c*Ho*Wo floats): P1 P2 z[6] t[6] [F]
PCD_HOSTDEV int slot_p1() { return 0; }
PCD_HOSTDEV int slot_p2() { return 1; }
PCD_HOSTDEV int slot_z(int u) { return 2 + u; }
PCD_HOSTDEV int slot_t(int u) { return 8 + u; }
PCD_HOSTDEV int slot_f() { return 14; }
PCD_HOSTDEV int edge_nslots(int s) { return 14 + (s == 2); }
PCD_HOSTDEV int edge_stats_doubles(int c, int s) { return edge_nbn(s) * 2 * c; }
// backward reduction scratch per edge (doubles): rows of c values, then one scalar
//   0: S0 = sum h          1..9: SZ[bn] = sum h*z_bn      10: SX = sum h*xs (identity)
//   11,12: sum GA, sum GA*zA  (A3)   13,14: same (A5)     [15*c]: bypass dot product
PCD_HOSTDEV int bs_s0() { return 0; }
PCD_HOSTDEV int bs_sz(int bn) { return 1 + bn; }
PCD_HOSTDEV int bs_sx() { return 10; }
PCD_HOSTDEV int bs_ga(int which) { return 11 + 2 * which; }
PCD_HOSTDEV int edge_bstats_doubles(int c) { return 15 * c + 1; }

struct BnC { float mean, rstd; };

PCD_HD BnC bn_consts(const double* st, int c, int bn, int j, double n, float eps) {
    double m = st[(bn * 2 + 0) * c + j] / n;
    double v = st[(bn * 2 + 1) * c + j] / n - m * m;
    if (v < 0.0) v = 0.0;
    BnC r;
    r.mean = (float)m;
    r.rstd = (float)(1.0 / sqrt(v + (double)eps));
    return r;
}

PCD_HD void bn_running_update(const double* st2c /*sum[c], sumsq[c]*/, int c, int j, double n,
                              float momentum, float* running /*mean[c], var[c]*/) {
    double m = st2c[j] / n;
    double v = st2c[c + j] / n - m * m;
    if (v < 0.0) v = 0.0;
    double unb = n > 1.0 ? v * n / (n - 1.0) : v;
    running[j] = (1.f - momentum) * running[j] + momentum * (float)m;
    running[c + j] = (1.f - momentum) * running[c + j] + momentum * (float)unb;
}

PCD_HD float relu(float v) { return v > 0.f ? v : 0.f; }

// ---- register-blocked depthwise stencils ---------------------------------------------------------
// A "patch" is 4x4 output pixels of one channel.  The shared-memory plane has its column origin at
// image column S*ox0 - 4 (so col0 = S*px is 16-byte aligned) and its row origin chosen by the caller:
// row0 = smem row holding image row S*(oy0+py) - PAD.
template <int KS, int DIL, int S, bool FLIP, bool RELU>
PCD_HD void dw_patch(const float* PCD_RESTRICT plane, int pitch, int row0, int col0,
                     const float* PCD_RESTRICT w, float (&acc)[4][4]) {
    constexpr int PAD = DIL * (KS - 1) / 2;
    constexpr int NR = 3 * S + (KS - 1) * DIL + 1;
    constexpr int NV4 = (S == 1) ? 3 : 4;
    float wr[KS * KS];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) wr[i] = FLIP ? w[KS * KS - 1 - i] : w[i];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const F4* rp = reinterpret_cast<const F4*>(plane + (row0 + r) * pitch + col0);
        float v[4 * NV4];
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            F4 t = rp[q];
            v[4 * q + 0] = RELU ? relu(t.x) : t.x;
            v[4 * q + 1] = RELU ? relu(t.y) : t.y;
            v[4 * q + 2] = RELU ? relu(t.z) : t.z;
            v[4 * q + 3] = RELU ? relu(t.w) : t.w;
        }
#pragma unroll
        for (int oy = 0; oy < 4; ++oy) {
            const int d = r - oy * S;
            if (d >= 0 && d % DIL == 0 && d / DIL < KS) {
                const int ky = d / DIL;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[oy][j] = fmaf(wr[ky * KS + kx], v[4 - PAD + S * j + kx * DIL], acc[oy][j]);
            }
        }
    }
}

// dW[ky][kx] += sum over the 4x4 patch of dt[oy][j] * in[S*oy + ky*DIL - PAD][S*j + kx*DIL - PAD]
template <int KS, int DIL, int S, bool RELU>
PCD_HD void dw_wgrad_patch(const float* PCD_RESTRICT plane, int pitch, int row0, int col0,
                           const float (&dt)[4][4], float (&acc)[KS * KS]) {
    constexpr int PAD = DIL * (KS - 1) / 2;
    constexpr int NR = 3 * S + (KS - 1) * DIL + 1;
    constexpr int NV4 = (S == 1) ? 3 : 4;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const F4* rp = reinterpret_cast<const F4*>(plane + (row0 + r) * pitch + col0);
        float v[4 * NV4];
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            F4 t = rp[q];
            v[4 * q + 0] = RELU ? relu(t.x) : t.x;
            v[4 * q + 1] = RELU ? relu(t.y) : t.y;
            v[4 * q + 2] = RELU ? relu(t.z) : t.z;
            v[4 * q + 3] = RELU ? relu(t.w) : t.w;
        }
#pragma unroll
        for (int oy = 0; oy < 4; ++oy) {
            const int d = r - oy * S;
            if (d >= 0 && d % DIL == 0 && d / DIL < KS) {
                const int ky = d / DIL;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[ky * KS + kx] = fmaf(dt[oy][j], v[4 - PAD + S * j + kx * DIL], acc[ky * KS + kx]);
            }
        }
    }
}

// ---- block-level column reduction ----------------------------------------------------------------
// P[k][t] holds value k of task t (K values, G groups of TPG consecutive tasks).  Adds, for every
// (g,k), sum_t P[k][g*TPG + t] into dst via `sink(g, k, value)`.  Two syncs.  P2 needs K*G*NP floats.
template <int NP = 32, class Sink>
PCD_HD void reduce_columns(const float* P, float* P2, int K, int G, int TPG, int NT, Sink sink) {
    PCD_SYNC();
    PCD_FOR(q, K * G * NP) {
        const int part = q % NP, kg = q / NP, k = kg / G, g = kg - k * G;
        float s = 0.f;
        for (int t = part; t < TPG; t += NP) s += P[k * NT + g * TPG + t];
        P2[q] = s;
    }
    PCD_SYNC();
    PCD_FOR(kg, K * G) {
        float s = 0.f;
        for (int p = 0; p < NP; ++p) s += P2[kg * NP + p];
        const int k = kg / G, g = kg - k * G;
        sink(g, k, s);
    }
    PCD_SYNC();
}

PCD_HOSTDEV int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Tile geometry shared by host launchers and kernels.
struct Tile {
    int TH, TW, tiles_x, tiles_y;
};

inline Tile pick_tile(int Ho, int Wo, int c, int target_px_times_c) {
    Tile t;
    t.TW = round_up(Wo < 64 ? Wo : 64, 4);
    int px = target_px_times_c / c;
    int th = px / t.TW;
    if (th < 4) th = 4;
    th = th / 4 * 4;
    int hmax = round_up(Ho, 4);
    if (th > hmax) th = hmax;
    t.TH = th;
    t.tiles_x = (Wo + t.TW - 1) / t.TW;
    t.tiles_y = (Ho + t.TH - 1) / t.TH;
    return t;
}

}  // namespace pcd

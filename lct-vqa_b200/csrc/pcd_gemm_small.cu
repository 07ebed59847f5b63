// pcd_gemm_small.cu — FP32-FMA GEMM for the small dense products of the answer head (vqa_model.py:308-316: fc1 512->1000,
// fc2 1000->1000) and the question encoder's fc2 (:186-190, 1024->512) at batch 64, and their backward products.
//
//     C[i][j] = sum_l A(i, l) * B(j, l) (+ bias[j])        A(i, l) = A[i*a_i + l*a_l],  B(j, l) = B[j*b_j + l*b_l]
//
// The generic strides express the three products of a Linear layer without operand transposes:
//     y  = x W^T + b   : A = x  (a_i = K, a_l = 1)   B = W  (b_j = K, b_l = 1)      I = M, J = N, L = K
//     dx = dy W        : A = dy (a_i = N, a_l = 1)   B = W  (b_j = 1, b_l = K)      I = M, J = K, L = N
//     dW = dy^T x      : A = dy (a_i = 1, a_l = N)   B = x  (b_j = 1, b_l = K)      I = N, J = K, L = M
// These products are 0.07 - 0.13 GFLOP with a 2 - 4 MB weight matrix: bound by launch latency and by reading W once, far below
// the point where the tcgen05 path (pcd_gemm_sm100.cu) pays — and the search network's gradients amplify the 3xTF32 rounding
// noise of the head ~100x (measured: the weight gradients' median error against float64 goes from 9e-5 to 5e-4 when these
// run as 3xTF32), so they stay exact fp32.  Round 1 sent them to cuBLAS; this is the library's own kernel instead.
// 64 x 64 (or, for skinny grids, 32 x 32) output tile per block, 16-deep chunks through shared memory, register tiles.
#include "../../include/pcdarts_sm100.h"
#include "pcd_launch.cuh"

namespace pcd {

struct SmallGemmArgs {
    const float* A; const float* B; float* C; const float* bias;
    long long a_i, a_l, b_j, b_l, ldc;
    int I, J, L;
};

constexpr int kSgK = 32;

// T x T output tile per block (T = 64: 4 x 4 register tile per thread; T = 32: 2 x 2, four times as many blocks for the
// skinny products whose 64 x 64 grid would leave most SMs idle).  The next 32-deep chunk of both operands is fetched into
// registers while the current one is multiplied out of shared memory: these products are a serial chain of up to 32 chunks,
// so an exposed global-load latency per chunk (~1 us) was the whole run time of the first version (98 us per launch).
template <int T>
struct KSmallGemm {
    static constexpr int kMinBlocks = 2, MT = T / 16, P = T + 4, NL = T * kSgK / kThreads;
    static const char* name() { return T == 64 ? "gemm_small_f32_t64" : "gemm_small_f32_t32"; }

    static PCD_D float fetch(const float* M, long long s_r, long long s_l, int n_r, int L, int r0, int l0, int e) {
        int r, l;
        if (s_l == 1) { l = e % kSgK; r = e / kSgK; } else { r = e % T; l = e / T; }
        const int gr = r0 + r, gl = l0 + l;
        return (gr < n_r && gl < L) ? M[gr * s_r + gl * s_l] : 0.f;
    }
    static PCD_D void put(float* S, long long s_l, int e, float v) {
        int r, l;
        if (s_l == 1) { l = e % kSgK; r = e / kSgK; } else { r = e % T; l = e / T; }
        S[l * P + r] = v;
    }

    static PCD_D void run(const SmallGemmArgs& a, int bx, int by, int, float* sm) {
        float* As = sm;                       // [kSgK][P]
        float* Bs = sm + kSgK * P;
        const int i0 = by * T, j0 = bx * T;
        PCD_TSTATE(float, acc, [MT][MT]);
        PCD_TSTATE(float, pa, [NL]);
        PCD_TSTATE(float, pb, [NL]);
        PCD_EACH(t) {
            auto& c = PCD_TREF(acc, t);
            auto& ra = PCD_TREF(pa, t);
            auto& rb = PCD_TREF(pb, t);
#pragma unroll
            for (int p = 0; p < MT; ++p)
#pragma unroll
                for (int q = 0; q < MT; ++q) c[p][q] = 0.f;
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                ra[k] = fetch(a.A, a.a_i, a.a_l, a.I, a.L, i0, 0, t + k * kThreads);
                rb[k] = fetch(a.B, a.b_j, a.b_l, a.J, a.L, j0, 0, t + k * kThreads);
            }
        }
        for (int l0 = 0; l0 < a.L; l0 += kSgK) {
            PCD_SYNC();                       // the previous chunk's readers are done
            PCD_EACH(t) {
                auto& ra = PCD_TREF(pa, t);
                auto& rb = PCD_TREF(pb, t);
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    put(As, a.a_l, t + k * kThreads, ra[k]);
                    put(Bs, a.b_l, t + k * kThreads, rb[k]);
                }
                if (l0 + kSgK < a.L) {        // next chunk: in flight while this one is multiplied
#pragma unroll
                    for (int k = 0; k < NL; ++k) {
                        ra[k] = fetch(a.A, a.a_i, a.a_l, a.I, a.L, i0, l0 + kSgK, t + k * kThreads);
                        rb[k] = fetch(a.B, a.b_j, a.b_l, a.J, a.L, j0, l0 + kSgK, t + k * kThreads);
                    }
                }
            }
            PCD_SYNC();
            PCD_EACH(t) {
                auto& c = PCD_TREF(acc, t);
                const int ti = (t >> 4) * MT, tj = (t & 15) * MT;
#pragma unroll
                for (int l = 0; l < kSgK; ++l) {
                    float ar[MT], br[MT];
#pragma unroll
                    for (int p = 0; p < MT; ++p) { ar[p] = As[l * P + ti + p]; br[p] = Bs[l * P + tj + p]; }
#pragma unroll
                    for (int p = 0; p < MT; ++p)
#pragma unroll
                        for (int q = 0; q < MT; ++q) c[p][q] = fmaf(ar[p], br[q], c[p][q]);
                }
            }
        }
        PCD_EACH(t) {
            auto& c = PCD_TREF(acc, t);
            const int ti = (t >> 4) * MT, tj = (t & 15) * MT;
#pragma unroll
            for (int p = 0; p < MT; ++p) {
                const int gi = i0 + ti + p;
                if (gi >= a.I) continue;
#pragma unroll
                for (int q = 0; q < MT; ++q) {
                    const int gj = j0 + tj + q;
                    if (gj < a.J) a.C[gi * a.ldc + gj] = c[p][q] + (a.bias ? a.bias[gj] : 0.f);
                }
            }
        }
    }
};

}  // namespace pcd

using namespace pcd;

extern "C" int pcd_gemm_small_f32(const float* A, long long a_i, long long a_l, const float* B, long long b_j, long long b_l, float* C,
                                  long long ldc, int I, int J, int L, const float* bias, void* stream) {
    if (!A || !B || !C || I <= 0 || J <= 0 || L <= 0 || ldc < J) return PCD_ERR_ARG;
    SmallGemmArgs a;
    a.A = A; a.B = B; a.C = C; a.bias = bias; a.a_i = a_i; a.a_l = a_l; a.b_j = b_j; a.b_l = b_l; a.ldc = ldc; a.I = I; a.J = J; a.L = L;
    if ((long long)((J + 63) / 64) * ((I + 63) / 64) >= 128)
        return launch<KSmallGemm<64>, SmallGemmArgs>(a, (J + 63) / 64, (I + 63) / 64, 1, 2 * kSgK * (64 + 4), stream);
    return launch<KSmallGemm<32>, SmallGemmArgs>(a, (J + 31) / 32, (I + 31) / 32, 1, 2 * kSgK * (32 + 4), stream);
}

// pcd_decode.cu — greedy question decode (vqa_model.py:103-136, models_lct.py:124-157) as ONE persistent cooperative
// kernel.  The reference runs, 30 times in a row, embedding -> one LSTM step -> tanh -> Linear(H, V) -> argmax: ~10 small
// launches per word whose next input depends on the previous argmax.  On the LCT path that loop runs three times per
// alpha-step (architect_lct.py:54,69).  Here the time loop lives inside the kernel; per word
//   phase A  (blocks 0 .. H/4-1, as in pcd_lstm.cu: block g owns 4 hidden units, its 16 rows of W_hh AND W_ih stay in
//             shared memory for all steps, the cell state in a register)
//             x = emb[word] (tanh only for <start>, vqa_model.py:119 vs :133);  gates = x W_ih^T + h W_hh^T + b;  h_t -> L2
//   -- grid barrier --
//   phase B  (every block: one 64 x 128 tile of the vocabulary)  logits = tanh(h_t) W_out^T + b_out: K chunks of 32 of
//             (tanh(h_t) rows | W_out tile rows; W_out is L2-resident, 36.6 MB at V = 17858) stream through a 4-stage
//             cp.async ring; 4 x 8 register tiles; row-wise argmax of the tile -> one packed (value, index) candidate per
//             row and tile
//   -- grid barrier --
//   the next phase A starts by reducing the candidates (lowest index wins ties, like torch.argmax) to the new words;
//   the input projection is split four ways along the embedding row (all ~19 float4 loads of a thread independent).
// The logits are never written anywhere.  Sampling (deterministic=False, torch.multinomial) is not covered: the Python side
// keeps that on stock torch ops.
#include "../../include/pcdarts_sm100.h"
#include "pcd_launch.cuh"

#include <limits.h>

namespace pcd {
namespace decode {
constexpr int kT = 256, kUnits = 4, kMaxB = 64, kTileN = 128, kKC = 32, kWP = kKC + 4, kStages = 4, kPickJ = 5;
constexpr int kStageFloats = (kMaxB + kTileN) * kWP;
PCD_HOSTDEV int u_floats(int H) { return kMaxB * (H + 4) > kStages * kStageFloats ? kMaxB * (H + 4) : kStages * kStageFloats; }
static inline int tiles_of(int V) { return (V + kTileN - 1) / kTileN; }
static inline bool shape_ok(int T, int B, int H, int E, int V) {
    return T > 0 && B > 0 && B <= kMaxB && H >= 32 && H <= 512 && H % 32 == 0 && E > 0 && E % 4 == 0 && V > 0;
}
}  // namespace decode
}  // namespace pcd

extern "C" size_t pcd_decode_work_floats(int B, int H, int V) {
    return (size_t)3 * B * H + (size_t)2 * B * pcd::decode::tiles_of(V);
}

#if PCD_CUDA
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace pcd {
namespace decode {

struct Args {
    int T, B, H, E, V, start, ntiles;
    const float *emb, *w_ih, *w_hh, *b_ih, *b_hh, *h0, *c0, *w_out, *b_out;
    long long* tokens;    // [B][T]
    float* hbuf;          // [2][B][H]  h_t, double-buffered by step parity
    float* tbuf;          // [B][H]     tanh(h_t): the input of the vocabulary projection
    float2* cand;         // [B][ntiles]  (logit, index bits): one 8-byte word per row and vocabulary tile
};

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}
__device__ __forceinline__ bool better(float v, int i, float best, int bi) { return v > best || (v == best && i < bi); }

// [B][H] (L2: written by other SMs) -> shared memory rows of pitch H + 4, 16-byte cp.async, no commit
__device__ __forceinline__ void stage_rows(float* Hs, const float* src, int B, int H) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, HP = H + 4;
    for (int bb = warp; bb < B; bb += kT / 32) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(Hs + bb * HP);
        const float* s = src + (long long)bb * H;
        for (int k = 4 * lane; k < H; k += 128)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 4u * k), "l"(s + k) : "memory");
    }
}

// candidates of the previous step -> TOK[row].  Warp w takes rows 8w .. 8w+7, lane l the tiles l, l+32, ...: the first
// kPickJ x 8 loads of a lane are independent (one L2 round trip for up to 32 kPickJ tiles), then 8 warp reductions.
__device__ __forceinline__ void pick_words(const Args& a, int* TOK) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 cv[8][kPickJ];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = warp * 8 + r;
        const float2* cr = a.cand + (long long)(row < a.B ? row : 0) * a.ntiles;
#pragma unroll
        for (int j = 0; j < kPickJ; ++j) {
            const int k = lane + 32 * j;
            cv[r][j] = (row < a.B && k < a.ntiles) ? __ldcg(cr + k) : make_float2(-INFINITY, __int_as_float(INT_MAX));
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = warp * 8 + r;
        float best = -INFINITY;
        int bi = INT_MAX;
#pragma unroll
        for (int j = 0; j < kPickJ; ++j) {
            const int i = __float_as_int(cv[r][j].y);
            if (better(cv[r][j].x, i, best, bi)) { best = cv[r][j].x; bi = i; }
        }
        if (row < a.B)
            for (int k = lane + 32 * kPickJ; k < a.ntiles; k += 32) {          // only for V > 4096 kPickJ
                const float2 c = __ldcg(a.cand + (long long)row * a.ntiles + k);
                if (better(c.x, __float_as_int(c.y), best, bi)) { best = c.x; bi = __float_as_int(c.y); }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
        }
        if (lane == 0 && row < a.B) TOK[row] = (bi >= 0 && bi < a.V) ? bi : 0;
    }
}

__global__ void __launch_bounds__(kT, 1) decode_kernel(Args a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) float sm[];
    const int H = a.H, E = a.E, B = a.B, HP = H + 4, EP = E + 4, H4 = H / 4, E4 = E / 4;
    float* Ws = sm;                       // [16][HP]  row q*4+u <- W_hh row q*H + u0 + u
    float* Wi = Ws + 16 * HP;             // [16][EP]  same rows of W_ih
    float* U = Wi + 16 * EP;              // phase A: Hs [B][HP] = h_{t-1};  phase B: kStages x ([64][kWP] | [128][kWP]) chunk ring
    float* Hs = U;
    int* TOK = reinterpret_cast<int*>(U + u_floats(H));            // [64]
    float* XP = reinterpret_cast<float*>(TOK + kMaxB);             // [4 quarters][16 gate rows][64] input-projection partials
    const int tid = threadIdx.x;
    const bool gate_block = (int)blockIdx.x < H / kUnits;
    const int u0 = blockIdx.x * kUnits;
    const int b = tid & 63, u = tid >> 6, unit = u0 + u;
    const int ty = tid >> 4, tx = tid & 15;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f}, c = 0.f;
    if (gate_block) {
        for (int i = tid; i < 16 * H4; i += kT) {
            const int rr = i / H4, k4 = i - rr * H4;
            *reinterpret_cast<float4*>(Ws + rr * HP + 4 * k4) =
                *reinterpret_cast<const float4*>(a.w_hh + (long long)((rr >> 2) * H + u0 + (rr & 3)) * H + 4 * k4);
        }
        for (int i = tid; i < 16 * E4; i += kT) {
            const int rr = i / E4, k4 = i - rr * E4;
            *reinterpret_cast<float4*>(Wi + rr * EP + 4 * k4) =
                *reinterpret_cast<const float4*>(a.w_ih + (long long)((rr >> 2) * H + u0 + (rr & 3)) * E + 4 * k4);
        }
        if (b < B) {
#pragma unroll
            for (int q = 0; q < 4; ++q) bsum[q] = a.b_ih[q * H + unit] + a.b_hh[q * H + unit];
            c = a.c0[(long long)b * H + unit];
        }
    }
    const int nk = H / kKC;
    for (int t = 0; t < a.T; ++t) {
        // ---- the word chosen at the previous step ------------------------------------------------------------------------
        if (t == 0) { if (tid < kMaxB) TOK[tid] = a.start; }
        else pick_words(a, TOK);
        __syncthreads();
        if (t > 0 && blockIdx.x == 0 && tid < B) a.tokens[(long long)tid * a.T + t - 1] = TOK[tid];
        // ---- phase A: one LSTM step for this block's 4 hidden units ----------------------------------------------------------
        if (gate_block) {
            const float* hprev = t ? a.hbuf + (long long)((t - 1) & 1) * B * H : a.h0;
            stage_rows(Hs, hprev, B, H);
            asm volatile("cp.async.commit_group;" ::: "memory");
            // input projection while h_{t-1} is in flight: thread (b, u) takes quarter u of x_b against all 16 gate rows of
            // the block (its ~19 float4 of the embedding row are independent loads), partial sums meet in shared memory
            if (b < B) {
                const float4* xr = reinterpret_cast<const float4*>(a.emb + (long long)TOK[b] * E);
                const int kq0 = (E4 * u) >> 2, kq1 = (E4 * (u + 1)) >> 2;
                float p[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) p[r] = 0.f;
                if (t == 0) {                             // <start>: tanh(emb[2]) (vqa_model.py:119)
                    for (int k4 = kq0; k4 < kq1; ++k4) {
                        float4 x = __ldg(xr + k4);
                        x.x = tanhf(x.x); x.y = tanhf(x.y); x.z = tanhf(x.z); x.w = tanhf(x.w);
#pragma unroll
                        for (int r = 0; r < 16; ++r) p[r] = dot4(x, *reinterpret_cast<const float4*>(Wi + r * EP + 4 * k4), p[r]);
                    }
                } else {
#pragma unroll 10
                    for (int k4 = kq0; k4 < kq1; ++k4) {
                        const float4 x = __ldg(xr + k4);
#pragma unroll
                        for (int r = 0; r < 16; ++r) p[r] = dot4(x, *reinterpret_cast<const float4*>(Wi + r * EP + 4 * k4), p[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 16; ++r) XP[(u * 16 + r) * kMaxB + b] = p[r];
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (b < B) {
                float acc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    acc[q] = bsum[q] + ((XP[(0 * 16 + q * 4 + u) * kMaxB + b] + XP[(1 * 16 + q * 4 + u) * kMaxB + b]) +
                                        (XP[(2 * 16 + q * 4 + u) * kMaxB + b] + XP[(3 * 16 + q * 4 + u) * kMaxB + b]));
                const float* hr = Hs + b * HP;
#pragma unroll 4
                for (int k4 = 0; k4 < H4; ++k4) {
                    const float4 h4 = *reinterpret_cast<const float4*>(hr + 4 * k4);
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] = dot4(h4, *reinterpret_cast<const float4*>(Ws + (q * 4 + u) * HP + 4 * k4), acc[q]);
                }
                const float gi = sigm(acc[0]), gf = sigm(acc[1]), gg = tanhf(acc[2]), go = sigm(acc[3]);
                c = fmaf(gf, c, gi * gg);
                const float h = go * tanhf(c);
                __stcg(a.hbuf + (long long)(t & 1) * B * H + (long long)b * H + unit, h);
                __stcg(a.tbuf + (long long)b * H + unit, tanhf(h));
            }
        }
        grid.sync();
        // ---- phase B: vocabulary logits of this block's tiles, row-wise argmax candidates ----------------------------------------
        // K chunks of 32: (tanh(h_t) rows | W_out tile rows) pairs through a kStages-deep cp.async ring in the U region
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            {
                const int col0 = tile * kTileN;
                auto issue = [&](int kc) {
                    float* sbuf = U + (kc % kStages) * kStageFloats;
                    float* wbuf = sbuf + kMaxB * kWP;
#pragma unroll
                    for (int i = tid; i < kMaxB * (kKC / 4); i += kT) {
                        const int row = i >> 3, j = i & 7;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(sbuf + row * kWP + 4 * j);
                        const float* src = a.tbuf + (long long)(row < B ? row : B - 1) * H + kc * kKC + 4 * j;
                        const int nbytes = row < B ? 16 : 0;           // rows past the batch: zero fill
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
                    }
#pragma unroll
                    for (int i = tid; i < kTileN * (kKC / 4); i += kT) {
                        const int col = i >> 3, j = i & 7;
                        const int gc = col0 + col;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(wbuf + col * kWP + 4 * j);
                        const float* src = a.w_out + (long long)(gc < a.V ? gc : a.V - 1) * H + kc * kKC + 4 * j;
                        const int nbytes = gc < a.V ? 16 : 0;          // columns past V: zero fill
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
                    }
                };
                float acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll
                for (int s = 0; s < kStages - 1; ++s) {
                    if (s < nk) issue(s);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                for (int kc = 0; kc < nk; ++kc) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
                    __syncthreads();              // chunk kc landed for everyone; everyone is done with chunk kc - 1
                    if (kc + kStages - 1 < nk) issue(kc + kStages - 1);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    const float* sb = U + (kc % kStages) * kStageFloats;
                    const float* wb = sb + kMaxB * kWP;
#pragma unroll
                    for (int k4 = 0; k4 < kKC / 4; ++k4) {
                        float4 s[4], w[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) s[i] = *reinterpret_cast<const float4*>(sb + (4 * ty + i) * kWP + 4 * k4);
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const float4*>(wb + (tx + 16 * j) * kWP + 4 * k4);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = dot4(s[i], w[j], acc[i][j]);
                    }
                }
                __syncthreads();                  // ring free again (next tile's prologue / next step's h staging)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float best = -INFINITY;
                    int bi = INT_MAX;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = col0 + tx + 16 * j;
                        if (col < a.V) {
                            const float v = acc[i][j] + __ldg(a.b_out + col);
                            if (v > best) { best = v; bi = col; }
                        }
                    }
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
                    }
                    const int row = 4 * ty + i;
                    if (tx == 0 && row < B) {
                        __stcg(a.cand + (long long)row * a.ntiles + tile, make_float2(best, __int_as_float(bi)));
                    }
                }
            }
        }
        grid.sync();
    }
    if (blockIdx.x == 0) {
        pick_words(a, TOK);
        __syncthreads();
        if (tid < B) a.tokens[(long long)tid * a.T + a.T - 1] = TOK[tid];
    }
}

}  // namespace decode
}  // namespace pcd

extern "C" int pcd_decode_greedy(int T, int B, int H, int E, int V, int start_token, const float* emb, const float* w_ih,
                                 const float* w_hh, const float* b_ih, const float* b_hh, const float* h0, const float* c0,
                                 const float* w_out, const float* b_out, long long* tokens, float* work, void* stream) {
    using namespace pcd;
    if (!emb || !w_ih || !w_hh || !b_ih || !b_hh || !h0 || !c0 || !w_out || !b_out || !tokens || !work) return PCD_ERR_ARG;
    if (!decode::shape_ok(T, B, H, E, V) || start_token < 0 || start_token >= V) return PCD_ERR_UNSUPPORTED;
    if ((((uintptr_t)emb) | ((uintptr_t)w_ih) | ((uintptr_t)w_hh) | ((uintptr_t)h0) | ((uintptr_t)w_out) | ((uintptr_t)work)) & 15) return PCD_ERR_ALIGN;
    LaunchState& L = launch_state();
    const size_t smem = ((size_t)16 * (H + 4) + (size_t)16 * (E + 4) + (size_t)decode::u_floats(H) + decode::kMaxB + 64 * decode::kMaxB) * sizeof(float);
    if (smem > 227 * 1024) return PCD_ERR_UNSUPPORTED;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaFuncSetAttribute(decode::decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            sms = 0;
            snprintf(L.last_err, sizeof L.last_err, "decode setup: %s", cudaGetErrorString(cudaGetLastError()));
            return PCD_ERR_CUDA;
        }
    }
    const int ntiles = decode::tiles_of(V);
    int grid = ntiles < sms ? ntiles : sms;                 // one block per SM: all of them co-resident (cooperative launch)
    if (grid < H / decode::kUnits) grid = H / decode::kUnits;
    if (grid > sms) return PCD_ERR_UNSUPPORTED;
    decode::Args a;
    a.T = T; a.B = B; a.H = H; a.E = E; a.V = V; a.start = start_token; a.ntiles = ntiles;
    a.emb = emb; a.w_ih = w_ih; a.w_hh = w_hh; a.b_ih = b_ih; a.b_hh = b_hh; a.h0 = h0; a.c0 = c0; a.w_out = w_out; a.b_out = b_out;
    a.tokens = tokens;
    a.hbuf = work;
    a.tbuf = work + (size_t)2 * B * H;
    a.cand = reinterpret_cast<float2*>(work + (size_t)3 * B * H);
    void* args[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)decode::decode_kernel, dim3(grid), dim3(decode::kT), args, smem, (cudaStream_t)stream);
    count_launch(L);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(L.last_err, sizeof L.last_err, "cooperative launch decode: %s", cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
    return PCD_OK;
}

#else   // ---- CPU emulation build (tests only) ----------------------------------------------------------------------

#include <math.h>
#include <vector>

static inline float sigm_(float x) { return 1.f / (1.f + expf(-x)); }

extern "C" int pcd_decode_greedy(int T, int B, int H, int E, int V, int start_token, const float* emb, const float* w_ih,
                                 const float* w_hh, const float* b_ih, const float* b_hh, const float* h0, const float* c0,
                                 const float* w_out, const float* b_out, long long* tokens, float* work, void*) {
    if (!emb || !w_ih || !w_hh || !b_ih || !b_hh || !h0 || !c0 || !w_out || !b_out || !tokens || !work) return PCD_ERR_ARG;
    if (!pcd::decode::shape_ok(T, B, H, E, V) || start_token < 0 || start_token >= V) return PCD_ERR_UNSUPPORTED;
    std::vector<float> h(h0, h0 + (size_t)B * H), c(c0, c0 + (size_t)B * H), hn((size_t)H), x((size_t)E), s((size_t)H);
    for (int b = 0; b < B; ++b) {
        int word = start_token;
        for (int t = 0; t < T; ++t) {
            for (int k = 0; k < E; ++k) x[k] = t ? emb[(long long)word * E + k] : tanhf(emb[(long long)word * E + k]);
            float* hb = h.data() + (size_t)b * H;
            float* cb = c.data() + (size_t)b * H;
            for (int j = 0; j < H; ++j) {
                float g[4];
                for (int q = 0; q < 4; ++q) {
                    double acc = (double)b_ih[q * H + j] + b_hh[q * H + j];
                    for (int k = 0; k < E; ++k) acc += (double)x[k] * w_ih[(long long)(q * H + j) * E + k];
                    for (int k = 0; k < H; ++k) acc += (double)hb[k] * w_hh[(long long)(q * H + j) * H + k];
                    g[q] = (float)acc;
                }
                const float gi = sigm_(g[0]), gf = sigm_(g[1]), gg = tanhf(g[2]), go = sigm_(g[3]);
                cb[j] = gf * cb[j] + gi * gg;
                hn[j] = go * tanhf(cb[j]);
            }
            for (int j = 0; j < H; ++j) { hb[j] = hn[j]; s[j] = tanhf(hn[j]); }
            float best = -INFINITY;
            int bi = 0;
            for (int v = 0; v < V; ++v) {
                double acc = b_out[v];
                for (int k = 0; k < H; ++k) acc += (double)s[k] * w_out[(long long)v * H + k];
                if ((float)acc > best) { best = (float)acc; bi = v; }
            }
            word = bi;
            tokens[(long long)b * T + t] = bi;
        }
    }
    return PCD_OK;
}

#endif

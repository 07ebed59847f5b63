// pcd_flat.cu — optimizer-side elementwise kernels over a short table of flat runs.
//
// The search step touches all 732 parameter tensors several times per step outside the network: w' = w - eta * g
// (architect_vqa.py:35-38), w +- R v of the finite-difference Hessian-vector product (:106-118), |v|, clip_grad_norm_ and
// Adam (experiment.py:196-198).  The parameters of the search network live back to back in one arena and their gradients
// come back as one flat buffer per cell, so those 732 tensors are ~20 contiguous runs: one launch per operation instead
// of ~12 multi-tensor launches.  The host side (pcd_flat.py) finds the runs; these kernels take them as a by-value table.
#include "../../include/pcdarts_sm100.h"
#include "pcd_launch.cuh"

namespace pcd {
namespace flat {

constexpr int kMaxRuns = 48, kChunk = 4096;

struct Table {
    int n;
    long long start[kMaxRuns + 1];       // prefix sums of the run lengths (elements)
    float* a[kMaxRuns];
    float* b[kMaxRuns];
    float* c[kMaxRuns];
    float* d[kMaxRuns];
};

PCD_HD int find_run(const Table& t, long long i) {
    int lo = 0, hi = t.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t.start[mid] <= i) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// f(run, offset, count<=4 contiguous elements inside the run) for every element of the chunk handled by this block
template <class F>
PCD_HD void for_chunk(const Table& t, long long chunk, F f) {
    const long long total = t.start[t.n];
    const long long c0 = chunk * kChunk;
    PCD_FOR(k, kChunk / 4) {
        long long i = c0 + 4 * (long long)k;
        if (i >= total) continue;
        int r = find_run(t, i);
        int left = 4;
        while (left > 0 && i < total) {
            const long long off = i - t.start[r], room = t.start[r + 1] - i;
            const int cnt = room < left ? (int)room : left;
            f(r, off, cnt);
            i += cnt; left -= cnt;
            if (left > 0) ++r;
        }
    }
}

struct AxpyArgs { Table t; const float* alpha_dev; float alpha; };
struct KAxpy { static constexpr int kMinBlocks = 4; static const char* name() { return "flat_axpy"; }
    static PCD_D void run(const AxpyArgs& a, int x, int, int, float*) {
        const float al = a.alpha_dev ? a.alpha * a.alpha_dev[0] : a.alpha;
        for_chunk(a.t, x, [&](int r, long long off, int cnt) {
            float* y = a.t.a[r] + off;
            const float* xx = a.t.b[r] + off;
            if (cnt == 4 && ((((uintptr_t)y) | ((uintptr_t)xx)) & 15) == 0) {
                F4 yv = *reinterpret_cast<F4*>(y);
                const F4 xv = *reinterpret_cast<const F4*>(xx);
                yv.x = fmaf(al, xv.x, yv.x); yv.y = fmaf(al, xv.y, yv.y); yv.z = fmaf(al, xv.z, yv.z); yv.w = fmaf(al, xv.w, yv.w);
                *reinterpret_cast<F4*>(y) = yv;
            } else {
                for (int j = 0; j < cnt; ++j) y[j] = fmaf(al, xx[j], y[j]);
            }
        });
    }
};

struct ScaleArgs { Table t; const float* scale_dev; };
struct KScale { static constexpr int kMinBlocks = 4; static const char* name() { return "flat_scale"; }
    static PCD_D void run(const ScaleArgs& a, int x, int, int, float*) {
        const float s = a.scale_dev[0];
        for_chunk(a.t, x, [&](int r, long long off, int cnt) {
            float* y = a.t.a[r] + off;
            for (int j = 0; j < cnt; ++j) y[j] *= s;
        });
    }
};

// Sum of squares, DETERMINISTIC (no atomics) and accumulated in fp64 from the first add: under data parallelism every rank
// derives the clip coefficient from the same averaged gradients, and replicas only stay bit-identical if they all get the same
// last bit here.  Pass 1: one fp64 partial per 4096-element chunk (fixed tree inside the block); pass 2: one block adds the
// partials in a fixed order.  The squares are exact in fp64 and every add rounds at 1e-16, so the fp32 norm does not depend on
// where the chunk boundaries fall — they move with the number of runs per launch, i.e. with which gradient buffers the
// allocator happened to place back to back on a rank (fp32 partials made 4 replicas drift apart by ulps:
// profiles/r02_bench_dp4_drift.json).
struct SumsqArgs { Table t; double* partials; };
struct KSumsq { static constexpr int kMinBlocks = 4; static const char* name() { return "flat_sumsq"; }
    static PCD_D void run(const SumsqArgs& a, int x, int, int, float* sm) {
#if PCD_CUDA
        double s = 0.0;
        const Table& t = a.t;
        const long long total = t.start[t.n], c0 = (long long)x * kChunk;
        for (int k = threadIdx.x; k < kChunk / 4; k += blockDim.x) {
            long long i = c0 + 4 * (long long)k;
            if (i >= total) continue;
            int r = find_run(t, i);
            for (int j = 0; j < 4 && i < total; ++j, ++i) {
                while (i >= t.start[r + 1]) ++r;
                const double v = (double)t.a[r][i - t.start[r]];
                s = fma(v, v, s);
            }
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        double* smd = reinterpret_cast<double*>(sm);
        if ((threadIdx.x & 31) == 0) smd[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) tot += smd[w];
            a.partials[x] = tot;
        }
#else
        (void)sm;
        double tot = 0.0;
        for_chunk(a.t, x, [&](int r, long long off, int cnt) {
            for (int j = 0; j < cnt; ++j) tot += (double)a.t.a[r][off + j] * (double)a.t.a[r][off + j];
        });
        a.partials[x] = tot;
#endif
    }
};
struct SumFinalArgs { const double* partials; int n; double* out; };
struct KSumFinal { static constexpr int kMinBlocks = 1; static const char* name() { return "flat_sumsq_final"; }
    static PCD_D void run(const SumFinalArgs& a, int, int, int, float* sm) {
        double* part = reinterpret_cast<double*>(sm);        // [kThreads]
        PCD_FOR(t, kThreads) {
            double s = 0.0;
            for (int i = t; i < a.n; i += kThreads) s += a.partials[i];
            part[t] = s;
        }
        PCD_SYNC();
        PCD_FOR(t, 1) {
            double s = 0.0;
            for (int i = 0; i < kThreads; ++i) s += part[i];
            a.out[0] += s;
        }
    }
};

// torch.optim.Adam (L2-style weight decay, no amsgrad): a = param, b = grad, c = exp_avg, d = exp_avg_sq
struct AdamArgs { Table t; float lr, b1, b2, eps, wd; const float* step_dev; };
struct KAdam { static constexpr int kMinBlocks = 4; static const char* name() { return "flat_adam"; }
    static PCD_D void run(const AdamArgs& a, int x, int, int, float*) {
        const float t = a.step_dev[0];
        const float bc1 = 1.f - powf(a.b1, t), bc2 = 1.f - powf(a.b2, t);
        const float step_size = a.lr / bc1, inv_bc2_sqrt = 1.f / sqrtf(bc2);
        for_chunk(a.t, x, [&](int r, long long off, int cnt) {
            float* p = a.t.a[r] + off;
            const float* g = a.t.b[r] + off;
            float* m = a.t.c[r] + off;
            float* v = a.t.d[r] + off;
            for (int j = 0; j < cnt; ++j) {
                float gj = g[j];
                if (a.wd != 0.f) gj = fmaf(a.wd, p[j], gj);
                const float mj = m[j] + (gj - m[j]) * (1.f - a.b1);
                const float vj = fmaf(1.f - a.b2, gj * gj, v[j] * a.b2);
                m[j] = mj;
                v[j] = vj;
                p[j] -= step_size * mj / (sqrtf(vj) * inv_bc2_sqrt + a.eps);
            }
        });
    }
};

static int fill(Table& t, int n, const long long* sizes, float* const* a, float* const* b, float* const* c, float* const* d) {
    if (n <= 0 || n > kMaxRuns || !sizes || !a) return PCD_ERR_ARG;
    t.n = n;
    t.start[0] = 0;
    for (int i = 0; i < n; ++i) {
        if (sizes[i] <= 0) return PCD_ERR_ARG;
        t.start[i + 1] = t.start[i] + sizes[i];
        t.a[i] = a[i];
        t.b[i] = b ? b[i] : nullptr;
        t.c[i] = c ? c[i] : nullptr;
        t.d[i] = d ? d[i] : nullptr;
    }
    return PCD_OK;
}
static int chunks(const Table& t) { return (int)((t.start[t.n] + kChunk - 1) / kChunk); }

}  // namespace flat
}  // namespace pcd

using namespace pcd;
using namespace pcd::flat;

extern "C" {

int pcd_flat_max_runs(void) { return kMaxRuns; }

int pcd_flat_axpy(int n, const long long* sizes, float* const* y, float* const* x, const float* alpha_dev, float alpha, void* stream) {
    AxpyArgs a;
    PCD_TRY(fill(a.t, n, sizes, y, x, nullptr, nullptr));
    a.alpha_dev = alpha_dev; a.alpha = alpha;
    return launch<KAxpy, AxpyArgs>(a, chunks(a.t), 1, 1, 0, stream);
}

int pcd_flat_scale(int n, const long long* sizes, float* const* y, const float* scale_dev, void* stream) {
    if (!scale_dev) return PCD_ERR_ARG;
    ScaleArgs a;
    PCD_TRY(fill(a.t, n, sizes, y, nullptr, nullptr, nullptr));
    a.scale_dev = scale_dev;
    return launch<KScale, ScaleArgs>(a, chunks(a.t), 1, 1, 0, stream);
}

/* *out += sum of squares over the runs (the caller zeroes *out); deterministic; work: pcd_flat_sumsq_work(total elements) doubles */
long long pcd_flat_sumsq_work(long long total_elements) { return (total_elements + kChunk - 1) / kChunk; }
int pcd_flat_sumsq(int n, const long long* sizes, float* const* x, double* out, double* work, void* stream) {
    if (!out || !work) return PCD_ERR_ARG;
    SumsqArgs a;
    PCD_TRY(fill(a.t, n, sizes, x, nullptr, nullptr, nullptr));
    a.partials = work;
    const int nc = chunks(a.t);
    PCD_TRY((launch<KSumsq, SumsqArgs>(a, nc, 1, 1, 64, stream)));
    SumFinalArgs f;
    f.partials = work; f.n = nc; f.out = out;
    return launch<KSumFinal, SumFinalArgs>(f, 1, 1, 1, 2 * kThreads, stream);
}

int pcd_flat_adam(int n, const long long* sizes, float* const* p, float* const* g, float* const* m, float* const* v, float lr, float beta1,
                  float beta2, float eps, float weight_decay, const float* step_dev, void* stream) {
    if (!step_dev || !g || !m || !v) return PCD_ERR_ARG;
    AdamArgs a;
    PCD_TRY(fill(a.t, n, sizes, p, g, m, v));
    a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay; a.step_dev = step_dev;
    return launch<KAdam, AdamArgs>(a, chunks(a.t), 1, 1, 0, stream);
}

}  // extern "C"

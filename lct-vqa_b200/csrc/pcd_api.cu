// pcd_api.cu — launch orchestration and the extern "C" surface declared in include/pcdarts_sm100.h.
//
// Built by nvcc (-gencode arch=compute_100a,code=sm_100a) together with pcd_k_*.cu into
// libpcdarts_sm100.so.  The same sources are compiled by g++ with -DPCD_EMU into tests/emu/libpcd_emu.so,
// where every kernel body runs as nested host loops over (block, task): test infrastructure for the index
// arithmetic, never a product path (pcd_is_cuda_build() returns 0 there and the product loader refuses it).
#include "pcd_bwd.cuh"
#include "pcd_pre.cuh"
#include "pcd_edge_v4.cuh"
#include "pcd_edge_bwd4.cuh"
#include "pcd_opk.cuh"
#include "pcd_kernels.h"
#include "pcd_launch.cuh"

namespace pcd {

static LaunchState g_state;
LaunchState& launch_state() { return g_state; }

int register_kernel(const char* name) {
    LaunchState& L = g_state;
    for (int i = 0; i < L.num_kernels; ++i)
        if (!strcmp(L.names[i], name)) return i;
    if (L.num_kernels >= kMaxKernels) return kMaxKernels - 1;
    L.names[L.num_kernels] = name;
    return L.num_kernels++;
}

static int zero_async(void* p, size_t bytes, void* stream) {
    if (!bytes) return PCD_OK;
#if PCD_CUDA
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        snprintf(g_state.last_err, sizeof g_state.last_err, "memset: %s", cudaGetErrorString(e));
        return PCD_ERR_CUDA;
    }
#else
    (void)stream;
    memset(p, 0, bytes);
#endif
    return PCD_OK;
}

// ---- optional auxiliary stream for work off the critical path (weight-grad jobs) -------------------------------
struct OverlapState {
    int on, pending;
#if PCD_CUDA
    cudaStream_t aux;
    cudaEvent_t fork_ev, join_ev;
#endif
};
static OverlapState g_overlap;

static void* overlap_stream() {
#if PCD_CUDA
    return g_overlap.on ? (void*)g_overlap.aux : nullptr;
#else
    return nullptr;
#endif
}
// aux waits for everything enqueued on `stream` so far
static int stream_fork(void* stream, void* aux) {
#if PCD_CUDA
    if (cudaEventRecord(g_overlap.fork_ev, (cudaStream_t)stream) != cudaSuccess) return PCD_ERR_CUDA;
    if (cudaStreamWaitEvent((cudaStream_t)aux, g_overlap.fork_ev, 0) != cudaSuccess) return PCD_ERR_CUDA;
#endif
    return PCD_OK;
}
// `stream` waits for everything enqueued on aux so far
static int stream_join(void* stream, void* aux) {
#if PCD_CUDA
    if (cudaEventRecord(g_overlap.join_ev, (cudaStream_t)aux) != cudaSuccess) return PCD_ERR_CUDA;
    if (cudaStreamWaitEvent((cudaStream_t)stream, g_overlap.join_ev, 0) != cudaSuccess) return PCD_ERR_CUDA;
#endif
    return PCD_OK;
}

// ---- kernel body adaptors -----------------------------------------------------------------------------
template <int C> struct KCombine { static constexpr int kMinBlocks = 2; static const char* name() { return C == 4 ? "combine_c4" : C == 8 ? "combine_c8" : "combine_c16"; } static PCD_D void run(const CombineArgs& a, int x, int y, int, float* sm) { combine_body<C>(a, x, y, sm); } };
struct KNorm { static constexpr int kMinBlocks = 1; static const char* name() { return "Norm"; } static PCD_D void run(const NormArgs& a, int x, int y, int z, float*) { norm_body(a, x, y, z); } };
struct KStem { static constexpr int kMinBlocks = 1; static const char* name() { return "Stem"; } static PCD_D void run(const StemArgs& a, int x, int y, int, float* sm) { stem_conv_body(a, x, y, sm); } };
struct KGapF { static constexpr int kMinBlocks = 1; static const char* name() { return "GapF"; } static PCD_D void run(const GapArgs& a, int x, int, int, float*) { gap_fwd_body(a, x); } };
struct KGapB { static constexpr int kMinBlocks = 1; static const char* name() { return "GapB"; } static PCD_D void run(const GapArgs& a, int x, int y, int, float*) { gap_bwd_body(a, x, y, a.ny); } };
struct KShuffle { static constexpr int kMinBlocks = 1; static const char* name() { return "Shuffle"; } static PCD_D void run(const ShuffleArgs& a, int x, int y, int z, float*) { shuffle_body(a, x, y, z); } };
template <int C> struct KNodeStats { static constexpr int kMinBlocks = 3; static const char* name() { return C == 4 ? "node_stats_c4" : C == 8 ? "node_stats_c8" : "node_stats_c16"; } static PCD_D void run(const NodeStatsArgs& a, int x, int y, int z, float* sm) { node_stats_body<C>(a, x, y, z, sm); } };
struct KSourceGrad { static constexpr int kMinBlocks = 4; static const char* name() { return "SourceGrad"; } static PCD_D void run(const SourceGradArgs& a, int x, int y, int z, float*) { source_grad_body(a, x, y, z); } };
struct KArchGrads { static constexpr int kMinBlocks = 1; static const char* name() { return "ArchGrads"; } static PCD_D void run(const ArchGradArgs& a, int, int, int, float* sm) { arch_grads_body(a, sm); } };
struct KBnBwdStats { static constexpr int kMinBlocks = 1; static const char* name() { return "BnBwdStats"; } static PCD_D void run(const BnBwdStatArgs& a, int x, int y, int z, float* sm) { bn_bwd_stats_body(a, x, y, z, sm); } };
struct KStemBwd { static constexpr int kMinBlocks = 1; static const char* name() { return "StemBwd"; } static PCD_D void run(const StemBwdArgs& a, int x, int, int, float* sm) { stem_bwd_body(a, x, a.nblocks_launch, sm); } };
struct KStemBwd2 { static constexpr int kMinBlocks = 2; static const char* name() { return "StemBwd"; } static PCD_D void run(const StemBwdArgs& a, int x, int, int, float* sm) { stem_bwd2_body(a, x, a.nblocks_launch, sm); } };

#define PCD_DISPATCH_C(c, EXPR)                         \
    do {                                                \
        if ((c) == 4) { constexpr int CC = 4; EXPR; }   \
        else if ((c) == 8) { constexpr int CC = 8; EXPR; } \
        else if ((c) == 16) { constexpr int CC = 16; EXPR; } \
        else return PCD_ERR_UNSUPPORTED;                \
    } while (0)

// ---- geometry of one homogeneous group of edges ---------------------------------------------------------
struct EdgeGeom {
    int B, c, S, Hs, Ws, Ho, Wo;
    long long nslot() const { return (long long)B * c * Ho * Wo; }
};

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// preconditions of the compile-time-tile ("FAST") specialisations: the tile picked for this geometry is one of
// the fixed ones, tiles are full and span the whole image width, every row of every tensor touched starts 16-byte aligned
static bool fast_ok(const EdgeGeom& q, const Tile& t, bool stageB) {
    return edge_tile_is_fixed(q.c, q.S, t.TH, t.TW, stageB) && q.Ho % t.TH == 0 && q.Wo == t.TW && q.Ws == q.S * q.Wo &&
           q.Hs == q.S * q.Ho && q.Ws % 4 == 0;
}

static int run_passAB(const EdgeGeom& q, const EdgeF* edges, int n, float eps, int save_t, void* stream) {
    if (n == 0) return PCD_OK;
    if (n > kMaxEdgesPerLaunch) return PCD_ERR_ARG;
    PassArgs a;
    memset(&a, 0, sizeof a);
    a.B = q.B; a.Hs = q.Hs; a.Ws = q.Ws; a.Ho = q.Ho; a.Wo = q.Wo; a.S = q.S; a.eps = eps; a.nedges = n;
    bool al = true;
    for (int i = 0; i < n; ++i) {
        a.e[i] = edges[i];
        al = al && aligned16(edges[i].x) && aligned16(edges[i].saved) && edges[i].x_ns % 4 == 0;
    }
    // production geometries: the v4 kernels (one block runs every stage-A job from one staged tile)
    bool al4 = al;
    for (int i = 0; i < n; ++i) al4 = al4 && aligned16(edges[i].par);
    if (al4 && q.Hs == q.S * q.Ho && q.Ws == q.S * q.Wo && fwd4_supported(q.c, q.S, q.Ho, q.Wo) && !getenv("PCD_NO_V4")) {
        FwdV4Args v;
        memset(&v, 0, sizeof v);
        v.B = q.B; v.Hs = q.Hs; v.Ws = q.Ws; v.Ho = q.Ho; v.Wo = q.Wo; v.eps = eps; v.nedges = n; v.save_t = save_t; v.jobs = 1;
        for (int i = 0; i < n; ++i) v.e[i] = edges[i];
        return launch_fwd4(v, q.c, q.S, stream);
    }
    Tile t = pick_tile(q.Ho, q.Wo, q.c, q.S == 1 ? 4096 : 2048);
    a.TH = t.TH; a.TW = t.TW; a.tiles_x = t.tiles_x;
    PCD_TRY(launch_fwdA(a, q.c, al && fast_ok(q, t, false), t.tiles_x * t.tiles_y, q.B, n * kFwdAJobs, stream));
    Tile tb = pick_tile(q.Ho, q.Wo, q.c, 4096);
    a.TH = tb.TH; a.TW = tb.TW; a.tiles_x = tb.tiles_x;
    PCD_TRY(launch_fwdB(a, q.c, al && fast_ok(q, tb, true), tb.tiles_x * tb.tiles_y, q.B, n * 2, stream));
    return PCD_OK;
}

static int run_combine(int B, int c, int Ho, int Wo, float eps, float mom, float* out, long long out_ns,
                       const EdgeC* edges, int n, void* stream) {
    if (n > kMaxNodeIn) return PCD_ERR_ARG;
    CombineArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.eps = eps; a.momentum = mom; a.nin = n; a.update_running = 1;
    a.out = out; a.out_ns = out_ns;
    for (int i = 0; i < n; ++i) a.e[i] = edges[i];
    a.px_per_block = combine_px(c);
    const int chunks = (Ho * Wo + a.px_per_block - 1) / a.px_per_block;
    PCD_DISPATCH_C(c, PCD_TRY((launch<KCombine<CC>, CombineArgs>(a, chunks, B, 1, combine_smem_floats(CC), stream))));
    return PCD_OK;
}

static int run_node_stats(int B, int c, int Ho, int Wo, const float* dn, long long dn_ns, const EdgeS* edges, int n,
                          void* stream) {
    NodeStatsArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.px_per_block = 4096 / c; a.dn = dn; a.dn_ns = dn_ns; a.nin = n;
    for (int i = 0; i < n; ++i) a.e[i] = edges[i];
    const int chunks = (Ho * Wo + a.px_per_block - 1) / a.px_per_block;
    PCD_DISPATCH_C(c, PCD_TRY((launch<KNodeStats<CC>, NodeStatsArgs>(a, chunks, B, n, node_stats_smem_floats(CC), stream))));
    return PCD_OK;
}

static void fill_edge_bwd_args(EdgeBwdArgs& a, const EdgeGeom& q, const EdgeG* edges, int n, float eps, int need_wgrad, bool* al) {
    memset(&a, 0, sizeof a);
    a.B = q.B; a.Hs = q.Hs; a.Ws = q.Ws; a.Ho = q.Ho; a.Wo = q.Wo; a.S = q.S; a.eps = eps; a.nedges = n;
    a.need_wgrad = need_wgrad;
    *al = true;
    for (int i = 0; i < n; ++i) {
        a.e[i] = edges[i];
        *al = *al && aligned16(edges[i].x) && aligned16(edges[i].saved) && aligned16(edges[i].dn) && aligned16(edges[i].ga) &&
              aligned16(edges[i].pd) && edges[i].x_ns % 4 == 0 && edges[i].dn_ns % 4 == 0;
    }
}

// does this edge geometry run the v3 backward kernels (merged partial-grad slots)?  Pointer alignment is checked again
// at launch; the arenas handed to the cell / MixedOp entry points are 16-byte aligned by contract.
static bool edge_bwd_is_v3(const EdgeGeom& q) {
    int thA = 0, thB = 0;
    return q.Ws % 4 == 0 && q.Hs == q.S * q.Ho && q.Ws == q.S * q.Wo && bwd2_tile(q.c, q.S, q.Ho, q.Wo, &thA) &&
           bwd2_tile(q.c, 1, q.Ho, q.Wo, &thB);
}

// v4 stage-A data kernels (plain stores into the partial-grad slots: no memset, no reductions)
static bool edge_bwd_is_v4(const EdgeGeom& q) {
    return edge_bwd_is_v3(q) && fwd4_supported(q.c, q.S, q.Ho, q.Wo) && !getenv("PCD_NO_V4");
}

// v3 weight-gradient jobs for edges whose data jobs already ran (any number of edges, one stride)
static int run_edge_wgrad2(const EdgeGeom& q, const EdgeG* edges, int n, float eps, void* stream) {
    int thA = 0;
    if (!bwd2_tile(q.c, q.S, q.Ho, q.Wo, &thA)) return PCD_ERR_UNSUPPORTED;
    for (int i0 = 0; i0 < n; i0 += kMaxEdgesPerLaunch) {
        const int m = (n - i0 < kMaxEdgesPerLaunch) ? n - i0 : kMaxEdgesPerLaunch;
        EdgeBwdArgs a;
        bool al;
        fill_edge_bwd_args(a, q, edges + i0, m, eps, 1, &al);
        a.TH = thA; a.TW = q.Wo; a.tiles_x = 1;
        PCD_TRY(launch_wgrad2(a, q.c, m * bwdA_njobs(q.S), stream));
    }
    return PCD_OK;
}

// Backward of a group of edges with one geometry.  Production geometries run the v3 data kernels; their weight
// gradients are either launched here (defer == nullptr) or left to the caller (*defer set to 1), which batches
// them over the whole cell.  Other geometries run the generic v2 kernels (weight grads inline).
static int run_edge_bwd(const EdgeGeom& q, const EdgeG* edges, int n, float eps, int need_wgrad, void* stream,
                        int* defer = nullptr, int* merged = nullptr) {
    if (defer) *defer = 0;
    if (merged) *merged = 0;
    if (n == 0) return PCD_OK;
    if (n > kMaxEdgesPerLaunch) return PCD_ERR_ARG;
    EdgeBwdArgs a;
    bool al;
    fill_edge_bwd_args(a, q, edges, n, eps, need_wgrad, &al);
    int thA = 0, thB = 0;
    if (edge_bwd_is_v3(q)) {
        if (!al) return PCD_ERR_ALIGN;
        bwd2_tile(q.c, q.S, q.Ho, q.Wo, &thA);
        bwd2_tile(q.c, 1, q.Ho, q.Wo, &thB);
        a.need_wgrad = 0;
        a.TH = thB; a.TW = q.Wo; a.tiles_x = 1;
        PCD_TRY(launch_bwdB2(a, q.c, n * 2, stream));
        a.TH = thA;
        if (edge_bwd_is_v4(q)) {
            if (merged) *merged = 2;     // edges[i].pd: two slots written with plain stores, ReLU mask already applied to slot 0
            PCD_TRY(launch_bwdA4(a, q.c, n, stream));
        } else {
            if (merged) *merged = 1;     // edges[i].pd: two zeroed slots accumulated with reductions (caller's job)
            PCD_TRY(launch_bwdA2(a, q.c, n * bwdA_njobs(q.S), stream));
        }
        if (need_wgrad) {
            if (defer) *defer = 1;
            else PCD_TRY(run_edge_wgrad2(q, edges, n, eps, stream));
        }
        return PCD_OK;
    }
    Tile tb = pick_tile(q.Ho, q.Wo, q.c, 4096);
    a.TH = tb.TH; a.TW = tb.TW; a.tiles_x = tb.tiles_x;
    PCD_TRY(launch_bwdB(a, q.c, al && fast_ok(q, tb, true), tb.tiles_x * tb.tiles_y, q.B, n * 2, stream));
    Tile t = pick_tile(q.Ho, q.Wo, q.c, q.S == 1 ? 4096 : 2048);
    a.TH = t.TH; a.TW = t.TW; a.tiles_x = t.tiles_x;
    PCD_TRY(launch_bwdA(a, q.c, al && fast_ok(q, t, false), t.tiles_x * t.tiles_y, q.B, n * bwdA_njobs(q.S), stream));
    return PCD_OK;
}

static int run_source_grad(const SourceGradArgs& a0, void* stream) {
    SourceGradArgs a = a0;
    const int HW = a.Hs * a.Ws, chunks = (HW + 4095) / 4096;
    a.nimg = HW >= 4096 ? 1 : (4096 / HW < a.B ? 4096 / HW : a.B);       // ~4096 pixels (4 float4 per thread) per block
    return launch<KSourceGrad, SourceGradArgs>(a, chunks, a.C, (a.B + a.nimg - 1) / a.nimg, 0, stream);
}

// ---- preprocess ------------------------------------------------------------------------------------------
static int run_pre_forward(int B, int Cin, int Cout, int Hin, int Win, int fr, float eps, float mom, const float* x,
                           const float* w, float* y, double* stats, float* running, long long* nbt, void* stream) {
    if (Cout != 16 && Cout != 32 && Cout != 64) return PCD_ERR_UNSUPPORTED;
    if (fr && ((Hin | Win) & 1)) return PCD_ERR_UNSUPPORTED;
    PreArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.Cin = Cin; a.Cout = Cout; a.Hin = Hin; a.Win = Win; a.fr = fr;
    a.Ho = fr ? Hin / 2 : Hin; a.Wo = fr ? Win / 2 : Win;
    a.eps = eps; a.momentum = mom; a.x = x; a.w = w; a.y = y; a.stats = stats; a.running = running; a.nbt = nbt;
    const int HW = a.Ho * a.Wo;
    // The tcgen05 forward exists and is accurate as an op (y within 6e-7..1.4e-6 of a float64 evaluation, the FP32-FMA
    // kernel 3e-7..7e-7: tests/test_gpu_parity.py::test_preprocess_tensor_core_path_full_batch), but the search network's
    // weight gradients amplify forward rounding noise ~100x (cancellation behind BatchNorm): with it, 163 of 714 weight-grad
    // tensors of the full-size network miss rel 1e-4 against float64 instead of 27 (profiles/r02_pre_tc_parity_ab.txt).
    // Parity comes first: the forward stays on the FP32-FMA kernel unless PCD_PRE_TC_FWD=1; the two backward products,
    // whose rounding noise is not amplified, run on the tensor cores.
    const char* tcf = getenv("PCD_PRE_TC_FWD");
    if (tcf && tcf[0] == '1' && pre_tc_supported(B, Cin, Cout, Hin, Win, fr) && aligned16(x) && aligned16(w) && aligned16(y))
        PCD_TRY(launch_pre_tc_fwd(a, stream));
    else PCD_TRY(launch_pre_conv(a, stream));
    NormArgs nrm;
    memset(&nrm, 0, sizeof nrm);
    nrm.B = B; nrm.C = Cout; nrm.HW = HW; nrm.eps = eps; nrm.momentum = mom; nrm.src = y; nrm.dst = y;
    nrm.stats = stats; nrm.running = running; nrm.nbt = nbt;
    return launch<KNorm, NormArgs>(nrm, norm_grid_x(B, HW), Cout, 1, 0, stream);
}

static int run_pre_backward(int B, int Cin, int Cout, int Hin, int Win, int fr, float eps, const float* x, const float* w,
                            const float* y, const float* dy, const double* stats, double* bstats, float* dx, float* gw,
                            void* stream, int bstats_done = 0) {
    const int Ho = fr ? Hin / 2 : Hin, Wo = fr ? Win / 2 : Win, HW = Ho * Wo;
    if (!bstats_done) {        // (the cell backward accumulates these two sums in the kernel that produces dy: source_grad)
        BnBwdStatArgs s;
        memset(&s, 0, sizeof s);
        s.B = B; s.C = Cout; s.HW = HW; s.dy = dy; s.y = y; s.stats = nullptr; s.eps = eps; s.bstats = bstats;
        PCD_TRY((launch<KBnBwdStats, BnBwdStatArgs>(s, norm_grid_x(B, HW), Cout, 1, bn_bwd_stats_smem_floats(), stream)));
    }
    if (!dx && !gw) return PCD_OK;
    PreBwdArgs a;
    memset(&a, 0, sizeof a);
    a.B = B; a.Cin = Cin; a.Cout = Cout; a.Hin = Hin; a.Win = Win; a.Ho = Ho; a.Wo = Wo; a.fr = fr;
    a.x = x; a.w = w; a.y = y; a.dy = dy; a.stats = stats; a.bstats = bstats; a.eps = eps; a.dx = dx; a.gw = gw;
    a.chunks_per_block = 0;      // chosen by the launcher
    if (pre_tc_supported(B, Cin, Cout, Hin, Win, fr) && aligned16(x) && aligned16(w) && aligned16(y) && aligned16(dy) && (!dx || aligned16(dx)))
        return launch_pre_tc_bwd(a, stream);
    return launch_pre_bwd(a, stream);
}

// ---- cell layout -------------------------------------------------------------------------------------------
struct CellLayout {
    int B, C, c, Cpp, Cp, Hs, Ws, Ho, Wo, H0, W0, red, redp;
    int node_of[PCD_MAX_EDGES], src_of[PCD_MAX_EDGES], stride[PCD_MAX_EDGES], first_edge[PCD_MAX_STEPS];
    long long par[PCD_MAX_EDGES], run[PCD_MAX_EDGES], nbt[PCD_MAX_EDGES], stats[PCD_MAX_EDGES], bstats[PCD_MAX_EDGES];
    long long saved[PCD_MAX_EDGES], ga[PCD_MAX_EDGES], dxs[PCD_MAX_EDGES];
    long long pd2[PCD_MAX_EDGES], pd2_begin, pd2_floats;     // v3: two merged partial-grad slots per edge, contiguous over the cell
    long long pre_par[2], pre_run[2], pre_nbt[2], pre_stats[2], pre_bstats[2], pre_out[2], pre_dout[2];
    long long dn[PCD_MAX_STEPS];
    pcd_cell_sizes tot;
};

static int cell_layout(const pcd_cell_shape& s, CellLayout& L) {
    if (s.steps != 4) return PCD_ERR_UNSUPPORTED;
    if (s.channels % 16 || s.batch <= 0 || s.height <= 0 || s.width <= 0) return PCD_ERR_UNSUPPORTED;
    const int c = s.channels / 4;
    if (c != 4 && c != 8 && c != 16) return PCD_ERR_UNSUPPORTED;
    if (s.reduction && ((s.height | s.width) & 1)) return PCD_ERR_UNSUPPORTED;
    memset(&L, 0, sizeof L);
    L.B = s.batch; L.C = s.channels; L.c = c; L.Cpp = s.c_prev_prev; L.Cp = s.c_prev;
    L.Hs = s.height; L.Ws = s.width; L.red = s.reduction; L.redp = s.reduction_prev;
    L.Ho = s.reduction ? s.height / 2 : s.height;
    L.Wo = s.reduction ? s.width / 2 : s.width;
    L.H0 = s.reduction_prev ? 2 * s.height : s.height;
    L.W0 = s.reduction_prev ? 2 * s.width : s.width;
    long long par = 0, run = 0, nbt = 0, st = 0, bst = 0, sv = 0, wk = 0;
    const long long state = (long long)L.B * L.C * L.Hs * L.Ws;
    const int cin[2] = {L.Cpp, L.Cp};
    for (int i = 0; i < 2; ++i) {
        L.pre_par[i] = par; par += (long long)L.C * cin[i];
        L.pre_run[i] = run; run += 2 * L.C;
        L.pre_nbt[i] = nbt; nbt += 1;
        L.pre_stats[i] = st; st += 2 * L.C;
        L.pre_bstats[i] = bst; bst += 2 * L.C;
        L.pre_out[i] = sv; sv += state;
        L.pre_dout[i] = wk; wk += state;
    }
    int e = 0;
    for (int i = 0; i < 4; ++i) {
        L.first_edge[i] = e;
        for (int j = 0; j < 2 + i; ++j, ++e) {
            const int sd = (s.reduction && j < 2) ? 2 : 1;
            L.node_of[e] = i; L.src_of[e] = j; L.stride[e] = sd;
            const long long nslot = (long long)L.B * c * L.Ho * L.Wo;
            const long long in_px = (long long)L.B * c * (sd == 2 ? L.Hs * L.Ws : L.Ho * L.Wo);
            L.par[e] = par; par += edge_param_floats(c, sd);
            L.run[e] = run; run += edge_nbn(sd) * 2 * c;
            L.nbt[e] = nbt; nbt += edge_nbn(sd);
            L.stats[e] = st; st += edge_stats_doubles(c, sd);
            L.bstats[e] = bst; bst += edge_bstats_doubles(c);
            L.saved[e] = sv; sv += edge_nslots(sd) * nslot;
            L.ga[e] = wk; wk += 2 * nslot;
            L.dxs[e] = wk; wk += edge_npd(sd) * in_px;
        }
    }
    const long long node = (long long)L.B * L.C * L.Ho * L.Wo;
    for (int i = 0; i < 3; ++i) { L.dn[i] = wk; wk += node; }
    L.pd2_begin = wk;
    for (int k = 0; k < PCD_MAX_EDGES; ++k) {
        const long long in_px = (long long)L.B * c * (L.stride[k] == 2 ? L.Hs * L.Ws : L.Ho * L.Wo);
        L.pd2[k] = wk; wk += 2 * in_px;
    }
    L.pd2_floats = wk - L.pd2_begin;
    L.tot.param_floats = par; L.tot.running_floats = run; L.tot.nbt_int64 = nbt;
    L.tot.out_floats = 4 * node; L.tot.saved_floats = sv; L.tot.stats_doubles = st;
    L.tot.bwd_work_floats = wk; L.tot.bwd_stats_doubles = bst;
    L.tot.out_height = L.Ho; L.tot.out_width = L.Wo;
    return PCD_OK;
}

struct StateRef { const float* p; long long ns; int H, W; };

static StateRef state_ref(const CellLayout& L, const float* saved, const float* out, int j) {
    StateRef r;
    if (j < 2) { r.p = saved + L.pre_out[j]; r.ns = (long long)L.C * L.Hs * L.Ws; r.H = L.Hs; r.W = L.Ws; }
    else { r.p = out + (long long)(j - 2) * L.C * L.Ho * L.Wo; r.ns = 4LL * L.C * L.Ho * L.Wo; r.H = L.Ho; r.W = L.Wo; }
    return r;
}

}  // namespace pcd

using namespace pcd;

extern "C" {

int pcd_version(void) { return PCD_VERSION; }
long long pcd_launch_count(void) { return __atomic_load_n(&g_state.launches, __ATOMIC_RELAXED); }

int pcd_profile_enable(int on) {
#if PCD_CUDA
    if (on && !g_state.ev) {
        g_state.ev = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * 2 * kMaxRecords);
        for (int i = 0; i < 2 * kMaxRecords; ++i)
            if (cudaEventCreate(&g_state.ev[i]) != cudaSuccess) return PCD_ERR_CUDA;
    }
#endif
    if (on) g_state.prof_n = 0;
    g_state.prof_on = on;
    return PCD_OK;
}

int pcd_profile_num_kernels(void) { return g_state.num_kernels; }
const char* pcd_profile_kernel_name(int id) { return (id >= 0 && id < g_state.num_kernels) ? g_state.names[id] : ""; }

/* Synchronises the recorded events; adds each launch's duration to ms[kernel id] and bumps count[kernel id]. */
int pcd_profile_collect(double* ms, long long* count, int max_kernels) {
    if (!ms || !count) return PCD_ERR_ARG;
#if PCD_CUDA
    for (int r = 0; r < g_state.prof_n; ++r) {
        float t = 0.f;
        if (cudaEventSynchronize(g_state.ev[2 * r + 1]) != cudaSuccess) return PCD_ERR_CUDA;
        if (cudaEventElapsedTime(&t, g_state.ev[2 * r], g_state.ev[2 * r + 1]) != cudaSuccess) return PCD_ERR_CUDA;
        const int k = g_state.prof_kid[r];
        if (k < max_kernels) { ms[k] += t; count[k] += 1; }
    }
#endif
    const int n = g_state.prof_n;
    g_state.prof_n = 0;
    return n;
}
/* Overlap of the weight-gradient jobs with the rest of the backward pass.  With overlap on, pcd_cell_backward enqueues
 * its deferred weight-grad jobs on a library-owned low-priority stream (forked from the caller's stream with an event;
 * no host synchronisation; capturable in a CUDA graph) and returns WITHOUT joining: grad_params, and every buffer
 * passed to that call, must stay alive and unread until pcd_overlap_join(stream) has been enqueued.  Call
 * pcd_set_overlap outside any stream capture.  Off by default. */
int pcd_set_overlap(int on) {
#if PCD_CUDA
    if (on && !g_overlap.aux) {
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return PCD_ERR_CUDA;
        if (cudaStreamCreateWithPriority(&g_overlap.aux, cudaStreamNonBlocking, lo) != cudaSuccess) return PCD_ERR_CUDA;
        if (cudaEventCreateWithFlags(&g_overlap.fork_ev, cudaEventDisableTiming) != cudaSuccess) return PCD_ERR_CUDA;
        if (cudaEventCreateWithFlags(&g_overlap.join_ev, cudaEventDisableTiming) != cudaSuccess) return PCD_ERR_CUDA;
    }
#endif
    g_overlap.on = on ? 1 : 0;
    return PCD_OK;
}
int pcd_overlap_join(void* stream) {
    if (!g_overlap.pending) return PCD_OK;
    g_overlap.pending = 0;
#if PCD_CUDA
    return stream_join(stream, (void*)g_overlap.aux);
#else
    return PCD_OK;
#endif
}
int pcd_is_cuda_build(void) { return PCD_CUDA; }
const char* pcd_last_cuda_error(void) { return g_state.last_err; }
const char* pcd_strerror(int s) {
    switch (s) {
        case PCD_OK: return "ok";
        case PCD_ERR_ARG: return "invalid argument";
        case PCD_ERR_UNSUPPORTED: return "shape not supported by the compiled kernels";
        case PCD_ERR_CUDA: return "CUDA launch failed";
        case PCD_ERR_ALIGN: return "pointer not 16-byte aligned";
        default: return "unknown error";
    }
}

int pcd_channel_shuffle(const float* x, float* y, int batch, int channels, int hw, int groups, void* stream) {
    if (!x || !y || groups <= 0 || channels % groups) return PCD_ERR_ARG;
    ShuffleArgs a;
    a.B = batch; a.C = channels; a.HW = hw; a.groups = groups; a.x = x; a.y = y;
    return launch<KShuffle, ShuffleArgs>(a, (hw + 4095) / 4096, channels, batch, 0, stream);
}

int pcd_cell_sizes_of(const pcd_cell_shape* shape, pcd_cell_sizes* out) {
    if (!shape || !out) return PCD_ERR_ARG;
    CellLayout L;
    PCD_TRY(cell_layout(*shape, L));
    *out = L.tot;
    return PCD_OK;
}

int pcd_cell_forward(const pcd_cell_fwd_args* a, void* stream) {
    if (!a || !a->s0 || !a->s1 || !a->weights || !a->weights2 || !a->params || !a->running || !a->nbt || !a->out ||
        !a->saved || !a->stats)
        return PCD_ERR_ARG;
    CellLayout L;
    PCD_TRY(cell_layout(a->shape, L));
    if ((((uintptr_t)a->saved) | ((uintptr_t)a->out) | ((uintptr_t)a->params)) & 15) return PCD_ERR_ALIGN;
    const float eps = a->shape.bn_eps, mom = a->shape.bn_momentum;
    PCD_TRY(zero_async(a->stats, L.tot.stats_doubles * sizeof(double), stream));
    PCD_TRY(run_pre_forward(L.B, L.Cpp, L.C, L.H0, L.W0, L.redp, eps, mom, a->s0, a->params + L.pre_par[0],
                            a->saved + L.pre_out[0], a->stats + L.pre_stats[0], a->running + L.pre_run[0],
                            (long long*)a->nbt + L.pre_nbt[0], stream));
    PCD_TRY(run_pre_forward(L.B, L.Cp, L.C, L.Hs, L.Ws, 0, eps, mom, a->s1, a->params + L.pre_par[1],
                            a->saved + L.pre_out[1], a->stats + L.pre_stats[1], a->running + L.pre_run[1],
                            (long long*)a->nbt + L.pre_nbt[1], stream));
    // wave w: edges whose source is state (w == 0 ? {s0,s1} : node w-1); then node w is complete
    for (int w = 0; w < 4; ++w) {
        EdgeF ef[kMaxEdgesPerLaunch];
        int n = 0;
        EdgeGeom q;
        q.B = L.B; q.c = L.c;
        for (int e = 0; e < PCD_MAX_EDGES; ++e) {
            const int j = L.src_of[e];
            if ((w == 0) ? (j >= 2) : (j != w + 1)) continue;
            StateRef s = state_ref(L, a->saved, a->out, j);
            ef[n].x = s.p; ef[n].x_ns = s.ns;
            ef[n].par = a->params + L.par[e];
            ef[n].saved = a->saved + L.saved[e];
            ef[n].stats = a->stats + L.stats[e];
            q.S = L.stride[e]; q.Hs = s.H; q.Ws = s.W;
            ++n;
        }
        q.Ho = L.Ho; q.Wo = L.Wo;
        PCD_TRY(run_passAB(q, ef, n, eps, a->skip_dw_outputs ? 0 : 1, stream));
        EdgeC ec[kMaxNodeIn];
        const int e0 = L.first_edge[w];
        for (int j = 0; j < 2 + w; ++j) {
            const int e = e0 + j;
            StateRef s = state_ref(L, a->saved, a->out, j);
            ec[j].x = s.p; ec[j].x_ns = s.ns; ec[j].stride = L.stride[e]; ec[j].Hs = s.H; ec[j].Ws = s.W;
            ec[j].saved = a->saved + L.saved[e];
            ec[j].stats = a->stats + L.stats[e];
            ec[j].alpha = a->weights + e * PCD_NUM_PRIMITIVES;
            ec[j].beta = a->weights2 + e;
            ec[j].running = a->running + L.run[e];
            ec[j].nbt = (long long*)a->nbt + L.nbt[e];
        }
        PCD_TRY(run_combine(L.B, L.c, L.Ho, L.Wo, eps, mom, a->out + (long long)w * L.C * L.Ho * L.Wo,
                            4LL * L.C * L.Ho * L.Wo, ec, 2 + w, stream));
    }
    return PCD_OK;
}

int pcd_cell_backward(const pcd_cell_bwd_args* a, void* stream) {
    if (!a || !a->s0 || !a->s1 || !a->weights || !a->weights2 || !a->params || !a->out || !a->saved || !a->stats ||
        !a->grad_out || !a->grad_weights || !a->grad_weights2 || !a->work || !a->bstats)
        return PCD_ERR_ARG;
    if (a->need_param_grads && !a->grad_params) return PCD_ERR_ARG;
    if (a->need_input_grads && (!a->grad_s0 || !a->grad_s1)) return PCD_ERR_ARG;
    CellLayout L;
    PCD_TRY(cell_layout(a->shape, L));
    const float eps = a->shape.bn_eps;
    const long long node = (long long)L.B * L.C * L.Ho * L.Wo;
    PCD_TRY(zero_async(a->bstats, L.tot.bwd_stats_doubles * sizeof(double), stream));
    if (a->need_param_grads) PCD_TRY(zero_async(a->grad_params, L.tot.param_floats * sizeof(float), stream));
    {   // merged partial-grad slots: the v3 data kernels accumulate into them (zero first); the v4 kernels store
        bool all_v4 = true;
        for (int e = 0; e < PCD_MAX_EDGES; ++e) {
            StateRef s = state_ref(L, a->saved, a->out, L.src_of[e]);
            EdgeGeom q;
            q.B = L.B; q.c = L.c; q.S = L.stride[e]; q.Hs = s.H; q.Ws = s.W; q.Ho = L.Ho; q.Wo = L.Wo;
            all_v4 = all_v4 && edge_bwd_is_v4(q);
        }
        if (!all_v4) PCD_TRY(zero_async(a->work + L.pd2_begin, L.pd2_floats * sizeof(float), stream));
    }
    auto dn_ptr = [&](int i, long long& ns) -> const float* {
        if (i == 3) { ns = 4 * (node / L.B); return a->grad_out + 3 * (node / L.B); }
        ns = node / L.B;
        return a->work + L.dn[i];
    };
    int edge_merged[PCD_MAX_EDGES] = {0};
    EdgeG wq[2][PCD_MAX_EDGES];          // edges whose weight-grad jobs are deferred, by stride
    EdgeGeom wgeo[2];
    int nwq[2] = {0, 0};
    for (int i = 3; i >= 0; --i) {
        long long dn_ns;
        const float* dn = dn_ptr(i, dn_ns);
        // 1. reductions over node i's incoming edges
        EdgeS es[kMaxNodeIn];
        const int e0 = L.first_edge[i];
        for (int j = 0; j < 2 + i; ++j) {
            const int e = e0 + j;
            StateRef s = state_ref(L, a->saved, a->out, j);
            es[j].x = s.p; es[j].x_ns = s.ns; es[j].stride = L.stride[e]; es[j].Hs = s.H; es[j].Ws = s.W;
            es[j].saved = a->saved + L.saved[e];
            es[j].bstats = a->bstats + L.bstats[e];
        }
        PCD_TRY(run_node_stats(L.B, L.c, L.Ho, L.Wo, dn, dn_ns, es, 2 + i, stream));
        // 2. every edge whose source is state (i == 0 ? {s0,s1} : node i-1) now has its consumer's sums
        EdgeG eg[kMaxEdgesPerLaunch];
        int n = 0;
        EdgeGeom q;
        q.B = L.B; q.c = L.c; q.Ho = L.Ho; q.Wo = L.Wo;
        for (int e = 0; e < PCD_MAX_EDGES; ++e) {
            const int j = L.src_of[e];
            if ((i == 0) ? (j >= 2) : (j != i + 1)) continue;
            StateRef s = state_ref(L, a->saved, a->out, j);
            long long cns;
            const float* cdn = dn_ptr(L.node_of[e], cns);
            eg[n].x = s.p; eg[n].x_ns = s.ns; eg[n].dn = cdn; eg[n].dn_ns = cns;
            eg[n].saved = a->saved + L.saved[e];
            eg[n].stats = a->stats + L.stats[e];
            eg[n].bstats = a->bstats + L.bstats[e];
            eg[n].par = a->params + L.par[e];
            eg[n].gpar = a->need_param_grads ? a->grad_params + L.par[e] : nullptr;
            eg[n].alpha = a->weights + e * PCD_NUM_PRIMITIVES;
            eg[n].beta = a->weights2 + e;
            eg[n].ga = a->work + L.ga[e];
            q.S = L.stride[e]; q.Hs = s.H; q.Ws = s.W;
            eg[n].pd = a->work + (edge_bwd_is_v3(q) ? L.pd2[e] : L.dxs[e]);
            ++n;
        }
        int deferred = 0, merged = 0;
        PCD_TRY(run_edge_bwd(q, eg, n, eps, a->need_param_grads, stream, &deferred, &merged));
        for (int e = 0; e < PCD_MAX_EDGES; ++e) {
            const int j = L.src_of[e];
            if (!((i == 0) ? (j >= 2) : (j != i + 1))) edge_merged[e] = merged;
        }
        if (deferred && n > 0) {
            const int g = q.S - 1;
            wgeo[g] = q;
            for (int k = 0; k < n; ++k) wq[g][nwq[g]++] = eg[k];
        }
        // 3. gradient of the source state(s)
        for (int j = (i == 0 ? 0 : i + 1); j <= (i == 0 ? 1 : i + 1); ++j) {
            StateRef s = state_ref(L, a->saved, a->out, j);
            SourceGradArgs sg;
            memset(&sg, 0, sizeof sg);
            sg.B = L.B; sg.C = L.C; sg.Hs = s.H; sg.Ws = s.W; sg.x = s.p; sg.x_ns = s.ns;
            if (j >= 2) {
                sg.g0 = a->grad_out + (long long)(j - 2) * (node / L.B); sg.g0_ns = 4 * (node / L.B);
                sg.out = a->work + L.dn[j - 2]; sg.out_ns = node / L.B;
            } else {
                sg.g0 = nullptr;
                sg.out = a->work + L.pre_dout[j]; sg.out_ns = (long long)L.C * L.Hs * L.Ws;
                sg.bn_sums = a->bstats + L.pre_bstats[j];      // sum dy, sum dy * yhat of preprocess j, fused (bstats is zeroed above)
            }
            int m = 0;
            for (int e = 0; e < PCD_MAX_EDGES; ++e) {
                if (L.src_of[e] != j) continue;
                if (m >= kMaxSrcEdges) return PCD_ERR_ARG;
                long long cns;
                const float* cdn = dn_ptr(L.node_of[e], cns);
                sg.e[m].pd = a->work + (edge_merged[e] ? L.pd2[e] : L.dxs[e]); sg.e[m].dn = cdn; sg.e[m].dn_ns = cns;
                sg.e[m].beta = a->weights2 + e; sg.e[m].stride = L.stride[e]; sg.e[m].merged = edge_merged[e];
                ++m;
            }
            sg.nedges = m;
            PCD_TRY(run_source_grad(sg, stream));
        }
    }
    // deferred weight-grad jobs: on the auxiliary stream when overlap is on (joined later by pcd_overlap_join: the
    // caller keeps every buffer of this call alive until then), else right here
    if (nwq[0] + nwq[1] > 0) {
        void* ws = stream;
        if (void* aux = overlap_stream()) {
            PCD_TRY(stream_fork(stream, aux));
            g_overlap.pending = 1;
            ws = aux;
        }
        for (int g = 0; g < 2; ++g)
            if (nwq[g]) PCD_TRY(run_edge_wgrad2(wgeo[g], wq[g], nwq[g], eps, ws));
    }
    // 4. d softmax(alpha) rows, d beta
    ArchGradArgs ag;
    memset(&ag, 0, sizeof ag);
    ag.c = L.c; ag.nedges = PCD_MAX_EDGES; ag.eps = eps;
    for (int e = 0; e < PCD_MAX_EDGES; ++e) {
        ag.e[e].stats = a->stats + L.stats[e];
        ag.e[e].bstats = a->bstats + L.bstats[e];
        ag.e[e].alpha = a->weights + e * PCD_NUM_PRIMITIVES;
        ag.e[e].beta = a->weights2 + e;
        ag.e[e].stride = L.stride[e];
        ag.e[e].count = (double)L.B * L.Ho * L.Wo;
        ag.e[e].gw = a->grad_weights + e * PCD_NUM_PRIMITIVES;
        ag.e[e].gw2 = a->grad_weights2 + e;
    }
    PCD_TRY((launch<KArchGrads, ArchGradArgs>(ag, 1, 1, 1, arch_grads_smem_floats(), stream)));
    // 5. preprocess backward
    PCD_TRY(run_pre_backward(L.B, L.Cpp, L.C, L.H0, L.W0, L.redp, eps, a->s0, a->params + L.pre_par[0],
                             a->saved + L.pre_out[0], a->work + L.pre_dout[0], a->stats + L.pre_stats[0],
                             a->bstats + L.pre_bstats[0], a->need_input_grads ? a->grad_s0 : nullptr,
                             a->need_param_grads ? a->grad_params + L.pre_par[0] : nullptr, stream, 1));
    PCD_TRY(run_pre_backward(L.B, L.Cp, L.C, L.Hs, L.Ws, 0, eps, a->s1, a->params + L.pre_par[1],
                             a->saved + L.pre_out[1], a->work + L.pre_dout[1], a->stats + L.pre_stats[1],
                             a->bstats + L.pre_bstats[1], a->need_input_grads ? a->grad_s1 : nullptr,
                             a->need_param_grads ? a->grad_params + L.pre_par[1] : nullptr, stream, 1));
    return PCD_OK;
}

// ---- MixedOp on its own --------------------------------------------------------------------------------
static int mixedop_geom(const pcd_mixedop_shape& s, EdgeGeom& q) {
    if (s.channels % 16 || (s.stride != 1 && s.stride != 2) || s.batch <= 0) return PCD_ERR_UNSUPPORTED;
    const int c = s.channels / 4;
    if (c != 4 && c != 8 && c != 16) return PCD_ERR_UNSUPPORTED;
    if (s.stride == 2 && ((s.height | s.width) & 1)) return PCD_ERR_UNSUPPORTED;
    q.B = s.batch; q.c = c; q.S = s.stride; q.Hs = s.height; q.Ws = s.width;
    q.Ho = s.height / s.stride; q.Wo = s.width / s.stride;
    return PCD_OK;
}

int pcd_mixedop_sizes_of(const pcd_mixedop_shape* s, pcd_mixedop_sizes* o) {
    if (!s || !o) return PCD_ERR_ARG;
    EdgeGeom q;
    PCD_TRY(mixedop_geom(*s, q));
    o->param_floats = edge_param_floats(q.c, q.S);
    o->running_floats = edge_nbn(q.S) * 2 * q.c;
    o->nbt_int64 = edge_nbn(q.S);
    o->out_floats = 4 * q.nslot();
    o->saved_floats = edge_nslots(q.S) * q.nslot();
    o->stats_doubles = edge_stats_doubles(q.c, q.S);
    o->bwd_work_floats = 2 * q.nslot() + (long long)edge_npd(q.S) * q.B * q.c * q.Hs * q.Ws;
    o->bwd_stats_doubles = edge_bstats_doubles(q.c);
    o->out_height = q.Ho; o->out_width = q.Wo;
    return PCD_OK;
}

int pcd_mixedop_forward(const pcd_mixedop_fwd_args* a, void* stream) {
    if (!a || !a->x || !a->weights || !a->params || !a->running || !a->nbt || !a->out || !a->saved || !a->stats)
        return PCD_ERR_ARG;
    EdgeGeom q;
    PCD_TRY(mixedop_geom(a->shape, q));
    if ((((uintptr_t)a->saved) | ((uintptr_t)a->out) | ((uintptr_t)a->params)) & 15) return PCD_ERR_ALIGN;
    const long long C = 4LL * q.c;
    PCD_TRY(zero_async(a->stats, edge_stats_doubles(q.c, q.S) * sizeof(double), stream));
    EdgeF ef;
    ef.x = a->x; ef.x_ns = C * q.Hs * q.Ws; ef.par = a->params; ef.saved = a->saved; ef.stats = a->stats;
    PCD_TRY(run_passAB(q, &ef, 1, a->shape.bn_eps, 1, stream));
    EdgeC ec;
    memset(&ec, 0, sizeof ec);
    ec.x = a->x; ec.x_ns = ef.x_ns; ec.stride = q.S; ec.Hs = q.Hs; ec.Ws = q.Ws; ec.saved = a->saved; ec.stats = a->stats;
    ec.alpha = a->weights; ec.beta = nullptr; ec.running = a->running; ec.nbt = (long long*)a->nbt;
    return run_combine(q.B, q.c, q.Ho, q.Wo, a->shape.bn_eps, a->shape.bn_momentum, a->out, C * q.Ho * q.Wo, &ec, 1, stream);
}

int pcd_mixedop_backward(const pcd_mixedop_bwd_args* a, void* stream) {
    if (!a || !a->x || !a->weights || !a->params || !a->saved || !a->stats || !a->grad_out || !a->grad_x ||
        !a->grad_weights || !a->work || !a->bstats)
        return PCD_ERR_ARG;
    if (a->need_param_grads && !a->grad_params) return PCD_ERR_ARG;
    EdgeGeom q;
    PCD_TRY(mixedop_geom(a->shape, q));
    const long long C = 4LL * q.c;
    PCD_TRY(zero_async(a->bstats, edge_bstats_doubles(q.c) * sizeof(double), stream));
    if (a->need_param_grads) PCD_TRY(zero_async(a->grad_params, edge_param_floats(q.c, q.S) * sizeof(float), stream));
    EdgeS es;
    memset(&es, 0, sizeof es);
    es.x = a->x; es.x_ns = C * q.Hs * q.Ws; es.stride = q.S; es.Hs = q.Hs; es.Ws = q.Ws; es.saved = a->saved; es.bstats = a->bstats;
    PCD_TRY(run_node_stats(q.B, q.c, q.Ho, q.Wo, a->grad_out, C * q.Ho * q.Wo, &es, 1, stream));
    EdgeG eg;
    memset(&eg, 0, sizeof eg);
    eg.x = a->x; eg.x_ns = es.x_ns; eg.dn = a->grad_out; eg.dn_ns = C * q.Ho * q.Wo; eg.saved = a->saved; eg.stats = a->stats;
    eg.bstats = a->bstats; eg.par = a->params; eg.gpar = a->need_param_grads ? a->grad_params : nullptr;
    eg.alpha = a->weights; eg.beta = nullptr; eg.ga = a->work; eg.pd = a->work + 2 * q.nslot();
    int merged = 0;
    if (edge_bwd_is_v3(q) && !edge_bwd_is_v4(q)) PCD_TRY(zero_async(eg.pd, (size_t)2 * q.B * q.c * q.Hs * q.Ws * sizeof(float), stream));
    PCD_TRY(run_edge_bwd(q, &eg, 1, a->shape.bn_eps, a->need_param_grads, stream, nullptr, &merged));
    SourceGradArgs sg;
    memset(&sg, 0, sizeof sg);
    sg.B = q.B; sg.C = (int)C; sg.Hs = q.Hs; sg.Ws = q.Ws; sg.x = a->x; sg.x_ns = es.x_ns; sg.g0 = nullptr;
    sg.out = a->grad_x; sg.out_ns = es.x_ns; sg.nedges = 1;
    sg.e[0].pd = eg.pd; sg.e[0].dn = a->grad_out; sg.e[0].dn_ns = eg.dn_ns; sg.e[0].beta = nullptr; sg.e[0].stride = q.S;
    sg.e[0].merged = merged;
    PCD_TRY(run_source_grad(sg, stream));
    ArchGradArgs ag;
    memset(&ag, 0, sizeof ag);
    ag.c = q.c; ag.nedges = 1; ag.eps = a->shape.bn_eps;
    ag.e[0].stats = a->stats; ag.e[0].bstats = a->bstats; ag.e[0].alpha = a->weights; ag.e[0].beta = nullptr;
    ag.e[0].stride = q.S; ag.e[0].count = (double)q.B * q.Ho * q.Wo; ag.e[0].gw = a->grad_weights; ag.e[0].gw2 = nullptr;
    return launch<KArchGrads, ArchGradArgs>(ag, 1, 1, 1, arch_grads_smem_floats(), stream);
}

// ---- stem ---------------------------------------------------------------------------------------------------
int pcd_stem_forward(const pcd_stem_args* a, void* stream) {
    if (!a || !a->x || !a->params || !a->running || !a->nbt || !a->out || !a->saved_z || !a->stats) return PCD_ERR_ARG;
    if (a->c_out % 8) return PCD_ERR_UNSUPPORTED;
    const int HW = a->height * a->width, Co = a->c_out;
    PCD_TRY(zero_async(a->stats, 2 * Co * sizeof(double), stream));
    StemArgs s;
    s.B = a->batch; s.Cout = Co; s.H = a->height; s.W = a->width; s.x = a->x; s.w = a->params; s.z = a->saved_z; s.stats = a->stats;
    PCD_TRY((launch<KStem, StemArgs>(s, (HW + kStemPx - 1) / kStemPx, a->batch, 1, stem_smem_floats(Co), stream)));
    NormArgs n;
    memset(&n, 0, sizeof n);
    n.B = a->batch; n.C = Co; n.HW = HW; n.eps = a->bn_eps; n.momentum = a->bn_momentum; n.src = a->saved_z; n.dst = a->out;
    n.stats = a->stats; n.gamma = a->params + Co * 27; n.bias = a->params + Co * 28; n.running = a->running;
    n.nbt = (long long*)a->nbt;
    return launch<KNorm, NormArgs>(n, norm_grid_x(a->batch, HW), Co, 1, 0, stream);
}

int pcd_stem_backward(const pcd_stem_args* a, void* stream) {
    if (!a || !a->x || !a->params || !a->saved_z || !a->stats || !a->grad_out || !a->bstats) return PCD_ERR_ARG;
    if (a->grad_x) return PCD_ERR_UNSUPPORTED;   // the image never requires grad on this path
    const int HW = a->height * a->width, Co = a->c_out;
    PCD_TRY(zero_async(a->bstats, 2 * Co * sizeof(double), stream));
    BnBwdStatArgs s;
    memset(&s, 0, sizeof s);
    s.B = a->batch; s.C = Co; s.HW = HW; s.dy = a->grad_out; s.y = a->saved_z; s.stats = a->stats; s.eps = a->bn_eps;
    s.bstats = a->bstats;
    PCD_TRY((launch<KBnBwdStats, BnBwdStatArgs>(s, norm_grid_x(a->batch, HW), Co, 1, bn_bwd_stats_smem_floats(), stream)));
    if (!a->grad_params) return PCD_OK;
    PCD_TRY(zero_async(a->grad_params, (size_t)Co * 29 * sizeof(float), stream));
    StemBwdArgs b;
    memset(&b, 0, sizeof b);
    b.B = a->batch; b.Cout = Co; b.H = a->height; b.W = a->width; b.PXB = 64;
    b.nblocks_px = a->batch * ((HW + 63) / 64);
    b.nblocks_launch = b.nblocks_px < 296 ? b.nblocks_px : 296;
    b.x = a->x; b.z = a->saved_z; b.dy = a->grad_out; b.gamma = a->params + Co * 27; b.stats = a->stats; b.bstats = a->bstats;
    b.eps = a->bn_eps; b.gw = a->grad_params; b.ggamma = a->grad_params + Co * 27; b.gbias = a->grad_params + Co * 28;
    if (stem_bwd2_ok(Co, a->height, a->width) && !((((uintptr_t)a->x) | ((uintptr_t)a->saved_z) | ((uintptr_t)a->grad_out)) & 15)) {
        const int ntiles = a->batch * (a->height / kStemTR);
        b.nblocks_launch = ntiles < 592 ? ntiles : 592;
        return launch<KStemBwd2, StemBwdArgs>(b, b.nblocks_launch, 1, 1, stem_bwd2_smem_floats(Co, a->width), stream);
    }
    return launch<KStemBwd, StemBwdArgs>(b, b.nblocks_launch, 1, 1, stem_bwd_smem_floats(Co, 64), stream);
}

int pcd_preprocess_forward(const pcd_pre_args* a, void* stream) {
    if (!a || !a->x || !a->weight || !a->running || !a->nbt || !a->y || !a->stats) return PCD_ERR_ARG;
    PCD_TRY(zero_async(a->stats, 2 * a->c_out * sizeof(double), stream));
    return run_pre_forward(a->batch, a->c_in, a->c_out, a->height, a->width, a->factorized, a->bn_eps, a->bn_momentum,
                           a->x, a->weight, a->y, a->stats, a->running, (long long*)a->nbt, stream);
}

int pcd_preprocess_backward(const pcd_pre_args* a, void* stream) {
    if (!a || !a->x || !a->weight || !a->y || !a->stats || !a->grad_y || !a->bstats) return PCD_ERR_ARG;
    PCD_TRY(zero_async(a->bstats, 2 * a->c_out * sizeof(double), stream));
    if (a->grad_weight) PCD_TRY(zero_async(a->grad_weight, (size_t)a->c_out * a->c_in * sizeof(float), stream));
    return run_pre_backward(a->batch, a->c_in, a->c_out, a->height, a->width, a->factorized, a->bn_eps, a->x, a->weight,
                            a->y, a->grad_y, a->stats, a->bstats, a->grad_x, a->grad_weight, stream);
}

int pcd_adaptive_avgpool_forward(const float* x, float* y, int batch, int channels, int h, int w, int oh, int ow, void* stream) {
    if (!x || !y) return PCD_ERR_ARG;
    GapArgs a;
    a.B = batch; a.C = channels; a.H = h; a.W = w; a.OH = oh; a.OW = ow; a.ny = 1; a.x = x; a.y = y;
    const long long total = (long long)batch * channels * oh * ow;
    return launch<KGapF, GapArgs>(a, (int)((total + kThreads - 1) / kThreads), 1, 1, 0, stream);
}

int pcd_adaptive_avgpool_backward(const float* gy, float* gx, int batch, int channels, int h, int w, int oh, int ow, void* stream) {
    if (!gy || !gx) return PCD_ERR_ARG;
    GapArgs a;
    a.B = batch; a.C = channels; a.H = h; a.W = w; a.OH = oh; a.OW = ow; a.ny = 1; a.x = gy; a.y = gx;
    const int planes = batch * channels;
    a.ny = planes < 4096 ? planes : 4096;
    return launch<KGapB, GapArgs>(a, (h * w + kThreads - 1) / kThreads, a.ny, 1, 0, stream);
}

// ---- stand-alone candidate operations (pcd_opk.cuh) ---------------------------------------------------------------------
static int dw_args(const pcd_dwconv_args* a, DwArgs& d) {
    memset(&d, 0, sizeof d);
    d.B = a->batch; d.C = a->channels; d.Hi = a->height; d.Wi = a->width; d.S = a->stride; d.PAD = a->padding; d.DIL = a->dilation;
    if (d.S < 1 || d.DIL < 1) return PCD_ERR_ARG;
    d.Ho = (d.Hi + 2 * d.PAD - d.DIL * (a->kernel - 1) - 1) / d.S + 1;
    d.Wo = (d.Wi + 2 * d.PAD - d.DIL * (a->kernel - 1) - 1) / d.S + 1;
    d.relu = a->relu_input; d.x = a->x; d.w = a->weight;
    return PCD_OK;
}

int pcd_dwconv_forward(const pcd_dwconv_args* a, void* stream) {
    if (!a || !a->x || !a->weight || !a->out) return PCD_ERR_ARG;
    DwArgs d;
    PCD_TRY(dw_args(a, d));
    d.t = a->out;
    return launch_dw_fwd(d, a->kernel, stream);
}

int pcd_dwconv_backward(const pcd_dwconv_args* a, void* stream) {
    if (!a || !a->x || !a->weight || !a->grad_out) return PCD_ERR_ARG;
    DwArgs d;
    PCD_TRY(dw_args(a, d));
    d.dt = a->grad_out; d.dx = a->grad_x; d.gw = a->grad_weight;
    if (!d.dx && !d.gw) return PCD_OK;
    if (d.gw) PCD_TRY(zero_async(d.gw, (size_t)a->channels * a->kernel * a->kernel * sizeof(float), stream));
    return launch_dw_bwd(d, a->kernel, stream);
}

static void pw_args(const pcd_pwconv_args* a, PwArgs& p) {
    memset(&p, 0, sizeof p);
    p.B = a->batch; p.Cin = a->c_in; p.Cout = a->c_out; p.HW = a->hw; p.eps = a->bn_eps;
    p.t = a->x; p.w = a->weight; p.z = a->z; p.stats = a->stats;
}

int pcd_pwconv_forward(const pcd_pwconv_args* a, void* stream) {
    if (!a || !a->x || !a->weight || !a->z || !a->stats) return PCD_ERR_ARG;
    if ((((uintptr_t)a->x) | ((uintptr_t)a->z)) & 15) return PCD_ERR_ALIGN;
    PwArgs p;
    pw_args(a, p);
    PCD_TRY(zero_async(a->stats, (size_t)2 * a->c_out * sizeof(double), stream));
    return launch_pw_fwd(p, stream);
}

int pcd_pwconv_backward(const pcd_pwconv_args* a, void* stream) {
    if (!a || !a->x || !a->weight || !a->z || !a->stats || !a->grad_y || !a->bstats) return PCD_ERR_ARG;
    if ((((uintptr_t)a->x) | ((uintptr_t)a->z) | ((uintptr_t)a->grad_y) | ((uintptr_t)a->grad_x)) & 15) return PCD_ERR_ALIGN;
    PwArgs p;
    pw_args(a, p);
    p.g = a->grad_y; p.gamma = a->gamma; p.bstats = a->bstats; p.dt = a->grad_x; p.gw = a->grad_weight;
    if (!p.dt && !p.gw) return PCD_OK;
    if (p.gw) PCD_TRY(zero_async(p.gw, (size_t)a->c_out * a->c_in * sizeof(float), stream));
    return launch_pw_bwd(p, stream);
}

int pcd_bn_apply(const pcd_bn_args* a, void* stream) {
    if (!a || !a->z || !a->stats || !a->y || (a->running && !a->nbt)) return PCD_ERR_ARG;
    if (a->batch <= 0 || a->channels <= 0 || a->channels > 65535 || a->hw <= 0) return PCD_ERR_UNSUPPORTED;
    NormArgs nrm;
    memset(&nrm, 0, sizeof nrm);
    nrm.B = a->batch; nrm.C = a->channels; nrm.HW = a->hw; nrm.eps = a->bn_eps; nrm.momentum = a->bn_momentum;
    nrm.src = a->z; nrm.dst = a->y; nrm.stats = a->stats; nrm.gamma = a->gamma; nrm.bias = a->beta;
    nrm.running = a->running; nrm.nbt = (long long*)a->nbt;
    return launch<KNorm, NormArgs>(nrm, norm_grid_x(a->batch, a->hw), a->channels, 1, 0, stream);
}

int pcd_bn_backward_stats(const pcd_bn_args* a, void* stream) {
    if (!a || !a->z || !a->stats || !a->grad_y || !a->bstats) return PCD_ERR_ARG;
    if (a->batch <= 0 || a->channels <= 0 || a->channels > 65535 || a->hw <= 0) return PCD_ERR_UNSUPPORTED;
    PCD_TRY(zero_async(a->bstats, (size_t)2 * a->channels * sizeof(double), stream));
    BnBwdStatArgs s;
    memset(&s, 0, sizeof s);
    s.B = a->batch; s.C = a->channels; s.HW = a->hw; s.dy = a->grad_y; s.y = a->z; s.stats = a->stats; s.eps = a->bn_eps;
    s.bstats = a->bstats;
    return launch<KBnBwdStats, BnBwdStatArgs>(s, norm_grid_x(a->batch, a->hw), a->channels, 1, bn_bwd_stats_smem_floats(), stream);
}

static int pool_args(const pcd_pool_args* a, PoolArgs& p) {
    memset(&p, 0, sizeof p);
    if (a->stride < 1) return PCD_ERR_ARG;
    p.B = a->batch; p.C = a->channels; p.Hi = a->height; p.Wi = a->width; p.S = a->stride; p.is_max = a->is_max;
    p.Ho = (p.Hi - 1) / p.S + 1; p.Wo = (p.Wi - 1) / p.S + 1;
    p.x = a->x;
    return PCD_OK;
}

int pcd_pool3x3_forward(const pcd_pool_args* a, void* stream) {
    if (!a || !a->x || !a->y) return PCD_ERR_ARG;
    PoolArgs p;
    PCD_TRY(pool_args(a, p));
    p.y = a->y;
    return launch_pool_fwd(p, stream);
}

int pcd_pool3x3_backward(const pcd_pool_args* a, void* stream) {
    if (!a || !a->x || !a->grad_y || !a->grad_x) return PCD_ERR_ARG;
    PoolArgs p;
    PCD_TRY(pool_args(a, p));
    p.dy = a->grad_y; p.dx = a->grad_x;
    return launch_pool_bwd(p, stream);
}

int pcd_channel_affine(const float* x, const float* scale, const float* shift, float* y, int batch, int channels, int hw, void* stream) {
    if (!x || !y) return PCD_ERR_ARG;
    AffineArgs f;
    memset(&f, 0, sizeof f);
    f.B = batch; f.C = channels; f.HW = hw; f.x = x; f.scale = scale; f.shift = shift; f.y = y;
    return launch_affine(f, stream);
}

}  // extern "C"

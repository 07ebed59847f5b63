// v4 stage-A backward data kernels (pcd_edge_bwd4.cuh): instantiations for the five production edge shapes + host dispatch
#include "pcd_edge_bwd4.cuh"
#include "pcd_kernels.h"
#include "pcd_launch.cuh"

namespace pcd {

template <int C, int S, int W> struct KBwdA4 {
    static constexpr int kMinBlocks = 3;
    static const char* name() {
        return S == 1 ? (C == 4 ? "bwdA4_c4_s1" : C == 8 ? "bwdA4_c8_s1" : "bwdA4_c16_s1") : (C == 8 ? "bwdA4_c8_s2" : "bwdA4_c16_s2");
    }
    static PCD_D void run(const EdgeBwdArgs& a, int x, int y, int z, float* sm) { bwdA4_body<C, S, W>(a, x, y, z, sm); }
};

template <int C, int S, int W>
static int go(const EdgeBwdArgs& a, int nedges, void* stream) {
    using G = V4GeoBwd<C, S, W>;
    return launch<KBwdA4<C, S, W>, EdgeBwdArgs>(a, a.Ho / G::TH, a.B, nedges * 2, G::SMEM_FLOATS, stream);
}

// blockIdx.z = edge * 2 + {0: conv block, 1: pool block}; a.e[i].pd = two slots per edge, written with plain stores
int launch_bwdA4(const EdgeBwdArgs& a, int c, int nedges, void* stream) {
    if (!fwd4_supported(c, a.S, a.Ho, a.Wo)) return PCD_ERR_UNSUPPORTED;
    if (a.S == 1) {
        if (c == 4) return go<4, 1, 64>(a, nedges, stream);
        if (c == 8) return go<8, 1, 32>(a, nedges, stream);
        return go<16, 1, 16>(a, nedges, stream);
    }
    if (c == 8) return go<8, 2, 32>(a, nedges, stream);
    return go<16, 2, 16>(a, nedges, stream);
}

}  // namespace pcd

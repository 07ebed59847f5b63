// v4 forward edge kernels (pcd_edge_v4.cuh): instantiations for the five production edge shapes + host dispatch
#include "pcd_edge_v4.cuh"
#include "pcd_kernels.h"
#include "pcd_launch.cuh"

namespace pcd {

template <int C, int S, int W> struct KFwdA4 {
    static constexpr int kMinBlocks = (V4Geo<C, S, W>::SMEM_FLOATS * 4 > 76 * 1024) ? 2 : 3;
    static const char* name() {
        return S == 1 ? (C == 4 ? "fwdA4_c4_s1" : C == 8 ? "fwdA4_c8_s1" : "fwdA4_c16_s1") : (C == 8 ? "fwdA4_c8_s2" : "fwdA4_c16_s2");
    }
    static PCD_D void run(const FwdV4Args& a, int x, int y, int z, float* sm) { fwdA4_body<C, S, W>(a, x, y, z, sm); }
};
template <int C, int W> struct KFwdB4 {
    static constexpr int kMinBlocks = 3;
    static const char* name() { return C == 4 ? "fwdB4_c4" : C == 8 ? "fwdB4_c8" : "fwdB4_c16"; }
    static PCD_D void run(const FwdV4Args& a, int x, int y, int z, float* sm) { fwdB4_body<C, W>(a, x, y, z, sm); }
};

bool fwd4_supported(int c, int S, int Ho, int Wo) {
    if (Ho % 16) return false;
    if (S == 1) return (c == 4 && Wo == 64) || (c == 8 && Wo == 32) || (c == 16 && Wo == 16);
    return (c == 8 && Wo == 32) || (c == 16 && Wo == 16);
}

template <int C, int S, int W>
static int go(FwdV4Args a, void* stream) {
    using G = V4Geo<C, S, W>;
    const int tilesA = a.Ho / G::TH;
    // few blocks (the last waves of a cell): split the stage-A jobs over two blocks
    a.jobs = ((long long)tilesA * a.B * a.nedges < 296) ? 2 : 1;
    if (const char* f = getenv("PCD_V4_JOBS")) a.jobs = (f[0] == '2') ? 2 : 1;      // test hook: force either job layout
    PCD_TRY((launch<KFwdA4<C, S, W>, FwdV4Args>(a, tilesA, a.B, a.nedges * a.jobs, G::SMEM_FLOATS, stream)));
    return launch<KFwdB4<C, W>, FwdV4Args>(a, a.Ho / 16, a.B, a.nedges * 2, V4GeoB<C, W>::SMEM_FLOATS, stream);
}

int launch_fwd4(const FwdV4Args& a, int c, int S, void* stream) {
    if (!fwd4_supported(c, S, a.Ho, a.Wo)) return PCD_ERR_UNSUPPORTED;
    if (S == 1) {
        if (c == 4) return go<4, 1, 64>(a, stream);
        if (c == 8) return go<8, 1, 32>(a, stream);
        return go<16, 1, 16>(a, stream);
    }
    if (c == 8) return go<8, 2, 32>(a, stream);
    return go<16, 2, 16>(a, stream);
}

}  // namespace pcd

// preprocess 1x1-conv kernels — instantiations + host dispatch
#include "pcd_pre.cuh"
#include "pcd_kernels.h"
#include "pcd_launch.cuh"

namespace pcd {

template <int CPT> struct KPreConv {
    static constexpr int kMinBlocks = CPT == 4 ? 3 : 2;      // t4: 78 registers without spills, 35 KB
    static const char* name() { return CPT == 4 ? "pre_conv_t4" : CPT == 8 ? "pre_conv_t8" : "pre_conv_t16"; }
    static PCD_D void run(const PreArgs& a, int x, int y, int z, float* sm) { pre_conv_body<CPT>(a, x, y, z, sm); }
};
template <int COUT, bool FR> struct KPreBwd {
    static constexpr int kMinBlocks = 2;
    static const char* name() {
        return FR ? (COUT == 32 ? "pre_bwd_fr32" : "pre_bwd_fr64")
                  : (COUT == 16 ? "pre_bwd_16" : COUT == 32 ? "pre_bwd_32" : "pre_bwd_64");
    }
    static PCD_D void run(const PreBwdArgs& a, int x, int y, int z, float* sm) { pre_bwd_body<COUT, FR>(a, x, y, z, sm); }
};

// enough blocks to fill 148 SMs x 2 resident blocks twice over
constexpr int kPreTargetBlocks = 592;

int launch_pre_conv(const PreArgs& a, void* stream) {
    if (a.Cout != 16 && a.Cout != 32 && a.Cout != 64) return PCD_ERR_UNSUPPORTED;
    if (a.fr && a.Cout < 32) return PCD_ERR_UNSUPPORTED;
    const int gx = (a.Ho * a.Wo + kPrePx - 1) / kPrePx;
    const int span = a.fr ? a.Cout / 2 : a.Cout;        // channels that share one sampling grid
    int cpt = 4;
    for (int c = 16; c >= 4; c /= 2)
        if (4 * c <= span && (long long)gx * a.B * (a.Cout / (4 * c)) >= kPreTargetBlocks) { cpt = c; break; }
    const int gz = a.Cout / (4 * cpt);
    const size_t sm = pre_smem_floats(cpt);
    if (cpt == 4) return launch<KPreConv<4>, PreArgs>(a, gx, a.B, gz, sm, stream);
    if (cpt == 8) return launch<KPreConv<8>, PreArgs>(a, gx, a.B, gz, sm, stream);
    return launch<KPreConv<16>, PreArgs>(a, gx, a.B, gz, sm, stream);
}

int launch_pre_bwd(const PreBwdArgs& a0, void* stream) {
    PreBwdArgs a = a0;
    if (a.Cin % 4) return PCD_ERR_UNSUPPORTED;
    const int gx = (a.Ho * a.Wo + kPrePx - 1) / kPrePx;
    const int kci = pre_bwd_kci(a.fr), nchunks = (a.Cin + kci - 1) / kci;
    int gz = (kPreTargetBlocks + gx * a.B - 1) / (gx * a.B);
    if (gz > nchunks) gz = nchunks;
    if (gz < 1) gz = 1;
    a.chunks_per_block = (nchunks + gz - 1) / gz;
    gz = (nchunks + a.chunks_per_block - 1) / a.chunks_per_block;
    const size_t sm = pre_bwd_smem_floats(a.Cout, a.fr);
    if (a.fr) {
        if (a.Cout == 32) return launch<KPreBwd<32, true>, PreBwdArgs>(a, gx, a.B, gz, sm, stream);
        if (a.Cout == 64) return launch<KPreBwd<64, true>, PreBwdArgs>(a, gx, a.B, gz, sm, stream);
        return PCD_ERR_UNSUPPORTED;
    }
    if (a.Cout == 16) return launch<KPreBwd<16, false>, PreBwdArgs>(a, gx, a.B, gz, sm, stream);
    if (a.Cout == 32) return launch<KPreBwd<32, false>, PreBwdArgs>(a, gx, a.B, gz, sm, stream);
    if (a.Cout == 64) return launch<KPreBwd<64, false>, PreBwdArgs>(a, gx, a.B, gz, sm, stream);
    return PCD_ERR_UNSUPPORTED;
}

}  // namespace pcd

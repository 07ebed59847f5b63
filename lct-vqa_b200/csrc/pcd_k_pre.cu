// preprocess 1x1-conv kernels — instantiations + host dispatch
#include "pcd_pre.cuh"
#include "pcd_kernels.h"
#include "pcd_launch.cuh"

namespace pcd {

template <int COUT> struct KPreConv {
    static constexpr int kMinBlocks = 1; static const char* name() { return COUT == 16 ? "pre_conv_16" : COUT == 32 ? "pre_conv_32" : "pre_conv_64"; }
    static PCD_D void run(const PreArgs& a, int x, int y, int, float* sm) { pre_conv_body<COUT>(a, x, y, sm); }
};
template <int COUT> struct KPreBwd {
    static constexpr int kMinBlocks = 1; static const char* name() { return COUT == 16 ? "pre_bwd_16" : COUT == 32 ? "pre_bwd_32" : "pre_bwd_64"; }
    static PCD_D void run(const PreBwdArgs& a, int x, int y, int, float* sm) { pre_bwd_body<COUT>(a, x, y, sm); }
};

int launch_pre_conv(const PreArgs& a, void* stream) {
    const int gx = (a.Ho * a.Wo + kPrePx - 1) / kPrePx;
    const size_t sm = pre_smem_floats(a.Cin, a.Cout);
    if (a.Cout == 16) return launch<KPreConv<16>, PreArgs>(a, gx, a.B, 1, sm, stream);
    if (a.Cout == 32) return launch<KPreConv<32>, PreArgs>(a, gx, a.B, 1, sm, stream);
    if (a.Cout == 64) return launch<KPreConv<64>, PreArgs>(a, gx, a.B, 1, sm, stream);
    return PCD_ERR_UNSUPPORTED;
}

int launch_pre_bwd(const PreBwdArgs& a, void* stream) {
    if (a.Cin % 8) return PCD_ERR_UNSUPPORTED;
    const int gx = (a.Ho * a.Wo + kPrePx - 1) / kPrePx;
    const size_t sm = pre_bwd_smem_floats(a.Cout);
    if (a.Cout == 16) return launch<KPreBwd<16>, PreBwdArgs>(a, gx, a.B, 1, sm, stream);
    if (a.Cout == 32) return launch<KPreBwd<32>, PreBwdArgs>(a, gx, a.B, 1, sm, stream);
    if (a.Cout == 64) return launch<KPreBwd<64>, PreBwdArgs>(a, gx, a.B, 1, sm, stream);
    return PCD_ERR_UNSUPPORTED;
}

}  // namespace pcd

// stand-alone candidate operations (pcd_opk.cuh) — instantiations + host dispatch
#include "pcd_opk.cuh"
#include "pcd_kernels.h"
#include "pcd_launch.cuh"
#include <stdlib.h>

namespace pcd {

template <int KS> struct KDwFwd {
    static constexpr int kMinBlocks = 2;
    static const char* name() { return KS == 3 ? "op_dw_fwd_k3" : KS == 5 ? "op_dw_fwd_k5" : "op_dw_fwd_k7"; }
    static PCD_D void run(const DwArgs& a, int x, int y, int, float* sm) { dw_fwd_body<KS>(a, x, y, sm); }
};
template <int KS> struct KDwBwd {
    static constexpr int kMinBlocks = 2;
    static const char* name() { return KS == 3 ? "op_dw_bwd_k3" : KS == 5 ? "op_dw_bwd_k5" : "op_dw_bwd_k7"; }
    static PCD_D void run(const DwArgs& a, int x, int y, int, float* sm) { dw_bwd_body<KS>(a, x, y, sm); }
};
struct KPwFwd {
    static constexpr int kMinBlocks = 2;
    static const char* name() { return "op_pw_fwd"; }
    static PCD_D void run(const PwArgs& a, int x, int y, int, float* sm) { pw_fwd_body(a, x, y, sm); }
};
struct KPwBwd {
    static constexpr int kMinBlocks = 2;
    static const char* name() { return "op_pw_bwd"; }
    static PCD_D void run(const PwArgs& a, int x, int y, int, float* sm) { pw_bwd_body(a, x, y, sm); }
};
struct KPoolFwd {
    static constexpr int kMinBlocks = 2;
    static const char* name() { return "op_pool_fwd"; }
    static PCD_D void run(const PoolArgs& a, int x, int y, int, float* sm) { pool_fwd_body(a, x, y, sm); }
};
struct KPoolBwd {
    static constexpr int kMinBlocks = 2;
    static const char* name() { return "op_pool_bwd"; }
    static PCD_D void run(const PoolArgs& a, int x, int y, int, float* sm) { pool_bwd_body(a, x, y, sm); }
};
struct KAffine {
    static constexpr int kMinBlocks = 2;
    static const char* name() { return "op_affine"; }
    static PCD_D void run(const AffineArgs& a, int x, int y, int, float*) { affine_body(a, x, y); }
};

static bool dw_ok(const DwArgs& a, int KS) {
    return (KS == 3 || KS == 5 || KS == 7) && a.B > 0 && a.C > 0 && a.C <= 65535 && a.B <= 65535 && a.Hi > 0 && a.Wi > 0 &&
           a.Hi <= kOpMaxHW && a.Wi <= kOpMaxHW && (a.S == 1 || a.S == 2) && a.DIL >= 1 && a.PAD >= 0 &&
           a.Ho == (a.Hi + 2 * a.PAD - a.DIL * (KS - 1) - 1) / a.S + 1 && a.Wo == (a.Wi + 2 * a.PAD - a.DIL * (KS - 1) - 1) / a.S + 1 &&
           a.Ho > 0 && a.Wo > 0;
}

int launch_dw_fwd(const DwArgs& a, int KS, void* stream) {
    if (!dw_ok(a, KS)) return PCD_ERR_UNSUPPORTED;
    const size_t sm = dw_fwd_smem_floats(a.Hi, a.Wi, a.PAD);
    if (KS == 3) return launch<KDwFwd<3>, DwArgs>(a, a.C, a.B, 1, sm, stream);
    if (KS == 5) return launch<KDwFwd<5>, DwArgs>(a, a.C, a.B, 1, sm, stream);
    return launch<KDwFwd<7>, DwArgs>(a, a.C, a.B, 1, sm, stream);
}

int launch_dw_bwd(const DwArgs& a, int KS, void* stream) {
    if (!dw_ok(a, KS)) return PCD_ERR_UNSUPPORTED;
    const size_t sm = dw_bwd_smem_floats(a.Hi, a.Wi, a.Ho, a.Wo, a.PAD, KS);
    if (KS == 3) return launch<KDwBwd<3>, DwArgs>(a, a.C, a.B, 1, sm, stream);
    if (KS == 5) return launch<KDwBwd<5>, DwArgs>(a, a.C, a.B, 1, sm, stream);
    return launch<KDwBwd<7>, DwArgs>(a, a.C, a.B, 1, sm, stream);
}

static bool pw_ok(const PwArgs& a) {
    return a.B > 0 && a.B <= 65535 && a.Cin > 0 && a.Cout > 0 && a.Cin % 4 == 0 && a.Cout % 4 == 0 && a.Cin <= 128 && a.Cout <= 128 &&
           a.HW > 0 && a.HW % 4 == 0;
}

int launch_pw_fwd(const PwArgs& a, void* stream) {
    if (!pw_ok(a)) return PCD_ERR_UNSUPPORTED;
    return launch<KPwFwd, PwArgs>(a, (a.HW + kPwPx - 1) / kPwPx, a.B, 1, pw_fwd_smem_floats(a.Cin, a.Cout), stream);
}

int launch_pw_bwd(const PwArgs& a0, void* stream) {
    if (!pw_ok(a0)) return PCD_ERR_UNSUPPORTED;
    PwArgs a = a0;
    const int nchunks = (a.HW + kPwPx - 1) / kPwPx;
    // as many chunks per block as still leaves ~4 blocks per SM: fewer global atomics on the (Cout x Cin) weight gradient
    a.cpb = (int)((long long)a.B * nchunks / 592);
    if (const char* e = getenv("PCD_PW_CPB")) a.cpb = atoi(e);      // test hook: the multi-chunk walk at small sizes
    if (a.cpb < 1) a.cpb = 1;
    if (a.cpb > nchunks) a.cpb = nchunks;
    return launch<KPwBwd, PwArgs>(a, (nchunks + a.cpb - 1) / a.cpb, a.B, 1, pw_bwd_smem_floats(a.Cin, a.Cout), stream);
}

static bool pool_ok(const PoolArgs& a) {
    return a.B > 0 && a.B <= 65535 && a.C > 0 && a.C <= 65535 && a.Hi > 0 && a.Wi > 0 && a.Hi <= kOpMaxHW && a.Wi <= kOpMaxHW &&
           (a.S == 1 || a.S == 2) && a.Ho == (a.Hi - 1) / a.S + 1 && a.Wo == (a.Wi - 1) / a.S + 1;
}

int launch_pool_fwd(const PoolArgs& a, void* stream) {
    if (!pool_ok(a)) return PCD_ERR_UNSUPPORTED;
    return launch<KPoolFwd, PoolArgs>(a, a.C, a.B, 1, pool_smem_floats(a.Hi, a.Wi, a.Ho, a.Wo), stream);
}

int launch_pool_bwd(const PoolArgs& a, void* stream) {
    if (!pool_ok(a)) return PCD_ERR_UNSUPPORTED;
    return launch<KPoolBwd, PoolArgs>(a, a.C, a.B, 1, pool_smem_floats(a.Hi, a.Wi, a.Ho, a.Wo), stream);
}

int launch_affine(const AffineArgs& a, void* stream) {
    if (a.B <= 0 || a.C <= 0 || a.C > 65535 || a.HW <= 0) return PCD_ERR_UNSUPPORTED;
    return launch<KAffine, AffineArgs>(a, (int)(((long long)a.B * a.HW + 4095) / 4096), a.C, 1, 0, stream);
}

}  // namespace pcd

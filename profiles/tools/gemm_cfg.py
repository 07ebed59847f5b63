"""Time pcd_gemm_tn_3xtf32 for the step's three vocabulary-projection products under every tile configuration."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "lct-vqa_b200"))
import pcd_native as N
lib = N.load_cuda()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 << 20, device=dev)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return 1e3 * tot / iters
shapes = {"fwd  M=1920 N=17858 K=512 ": (1920, 17858, 512, 1), "dX   M=1920 N=512 K=17860  ": (1920, 512, 17860, 17), "dW   M=17858 N=512 K=1920  ": (17858, 512, 1920, 1),
          "imfc M=64 N=512 K=12544    ": (64, 512, 12544, 32), "lstm M=1920 N=2048 K=300   ": (1920, 2048, 300, 1)}
for name, (M, Nn, K, split) in shapes.items():
    A = torch.randn(M, K, device=dev); B = torch.randn(Nn, K, device=dev); ldc = (Nn + 3) // 4 * 4
    Cc = torch.empty(M, ldc, device=dev)
    row = []
    for cfg in (1, 2, 3, 4):
        lib.pcd_gemm_debug_cfg(cfg)
        def fn():
            N.check(lib, lib.pcd_gemm_tn_3xtf32(N.ptr(A), K, N.ptr(B), K, N.ptr(Cc), ldc, M, Nn, K, None, split, st), "gemm")
        row.append(timeit(fn))
    print(name, " ".join(f"cfg{c}={t:7.1f}us" for c, t in zip((1, 2, 3, 4), row)), f"  best TF/s {2.0 * M * Nn * K / min(row) / 1e6:6.1f}", flush=True)
lib.pcd_gemm_debug_cfg(0)

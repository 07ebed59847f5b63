"""Time pcd_lstm_forward / pcd_lstm_backward alone (CUDA events) at the benchmark shape; PCD_LSTM_V1=1 selects the old kernels."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "lct-vqa_b200"))
import pcd_native as N
dev = torch.device("cuda:0")
T, B, H = 30, 64, 512
lib = N.load_cuda()
g = lambda *s: torch.randn(*s, device=dev)
gx, wh, h0 = g(T * B, 4 * H), g(4 * H, H) / H ** 0.5, 0.5 * g(B, H)
act, cs, hs = torch.empty(T, B, 4 * H, device=dev), torch.empty(T, B, H, device=dev), torch.empty(T, B, H, device=dev)
dhs, dg, dh0, dc0 = g(T, B, H), torch.empty(T * B, 4 * H, device=dev), torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
pbuf = torch.empty(lib.pcd_lstm_pbuf_floats(B, H), device=dev)
st = torch.cuda.current_stream().cuda_stream
def fwd():
    N.check(lib, lib.pcd_lstm_forward(T, B, H, N.ptr(gx), N.ptr(wh), N.ptr(h0), N.ptr(h0), N.ptr(act), N.ptr(cs), N.ptr(hs), st), "f")
def bwd():
    N.check(lib, lib.pcd_lstm_backward(T, B, H, N.ptr(dhs), None, None, N.ptr(act), N.ptr(cs), N.ptr(h0), N.ptr(wh), N.ptr(dg), N.ptr(dh0),
                                       N.ptr(dc0), N.ptr(pbuf), st), "b")
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"lstm {name} {'v1' if os.environ.get('PCD_LSTM_V1') else 'v2'}: {1e3 * e0.elapsed_time(e1) / 20:.1f} us per launch, {1e3 * e0.elapsed_time(e1) / 20 / T:.2f} us per step")

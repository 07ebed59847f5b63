"""QstEncoder.generate at the production size (B=64, H=512, E=300, V=17858, T=30): persistent decode kernel vs the stock
torch loop (eager wall time per call and summed kernel time)."""
import sys, os, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'lct-vqa_b200')); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from vqa_model import QstEncoder
dev = 'cuda'
torch.manual_seed(0)
q = QstEncoder(17858, 300, 512, 1, 512).to(dev)
img = torch.randn(64, 512, device=dev) * 0.1


def fast():
    return q.generate(img)


def slow():
    q.deterministic = None
    q.sample = lambda prob: torch.argmax(prob, 2)
    try:
        return q.generate(img)
    finally:
        q.deterministic = True


a, b = fast(), slow()
print("words equal to the torch loop's:", float((a == b).float().mean()))
for name, fn in (("decode kernel", fast), ("torch loop", slow)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            agg[e.name][0] += e.device_time; agg[e.name][1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms per generate (eager, device-timed), kernel time {tot / 1e3:.3f} ms, {sum(v[1] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:5]:
        print(f"   {v[0]:8.1f} us n={v[1]:3d} {k[:100]}")

"""Are all_reduce results bit-identical on every rank?  (torchrun, one process per GPU)"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "lct-vqa_b200"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda")
def bits(t):
    v = t.contiguous().view(torch.int32).to(torch.int64)
    return torch.stack([v.sum(), (v * (torch.arange(v.numel(), device=dev) % 1000003)).sum()])
res = {}
for op_name, op in (("avg", dist.ReduceOp.AVG), ("sum", dist.ReduceOp.SUM)):
    for n in (252, 4096, 1 << 16, 1 << 18, 1 << 20, 6 << 20, 25 << 20):
        g = torch.Generator(device=dev).manual_seed(1000 * rank + n % 977)
        t = torch.randn(n, device=dev, generator=g)
        dist.all_reduce(t, op=op)
        h = bits(t)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        res[f"{op_name}_{n}"] = all(torch.equal(x, hs[0]) for x in hs)
# the reducer itself, side stream, mixed sizes
from pcd_dist import GradReducer
red = GradReducer()
g = torch.Generator(device=dev).manual_seed(77 + rank)
ts = [torch.randn(s, device=dev, generator=g) for s in (17858 * 512, 17858, 2048 * 512, 2048, 300 * 17858, 1000, 12544 * 512, 64, 252)]
red.start(ts[:5], overlapped=True)
x = torch.randn(4096, 4096, device=dev) @ torch.randn(4096, 4096, device=dev)
red.start(ts[5:])
red.finish()
h = torch.stack([bits(t) for t in ts])
hs = [torch.zeros_like(h) for _ in range(world)]
dist.all_gather(hs, h)
res["reducer"] = [bool(all(torch.equal(x[i], hs[0][i]) for x in hs)) for i in range(len(ts))]
if rank == 0:
    print("NCCL_ALGO", os.environ.get("NCCL_ALGO"), "NCCL_PROTO", os.environ.get("NCCL_PROTO"), res, flush=True)
dist.barrier()
dist.destroy_process_group()

"""Where do data-parallel replicas stop being bit-identical?  torchrun, one process per GPU; prints per phase whether
weights / alphas / scalars are bit-equal on all ranks."""
import os, sys, torch, torch.distributed as dist
from argparse import Namespace
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lct-vqa_b200"))
import bench, config, pcd_dist as pdist, pcd_native, pcd_ops, pcd_flat
rank, world = pdist.init_from_env()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local); config.DEVICE = dev
pcd_native.load_cuda(); pcd_ops.set_wgrad_overlap(True)
from pcdarts.architect_vqa import Architect
from search import SearchStep, GraphedSearchStep
from vqa_model import VqaModel
torch.manual_seed(10)
model = VqaModel(qst_vocab_size=17858, img_encoder_type="darts", **bench.DIMS).to(dev).train()
for t in list(model.parameters()) + list(model.buffers()) + list(model.arch_parameters()):
    dist.broadcast(t.data, 0)
reducer = pdist.GradReducer()
opt = pcd_flat.FlatAdam(model.parameters(), lr=1e-3)
architect = Architect(model, Namespace(arch_learn_rate=6e-4, arch_wt_decay=1e-3, qst_only=False), reducer=reducer)
architect.optimizer = torch.optim.Adam(model.arch_parameters(), lr=6e-4, betas=(0.5, 0.999), weight_decay=1e-3, capturable=True)
stepper = SearchStep(model, architect, opt, reducer=reducer)
train = [t.to(dev) for t in bench.synth_batch(10 + rank, 64, 17858, 64)]
valid = [t.to(dev) for t in bench.synth_batch(1010 + rank, 64, 17858, 64)]

def h(ts):
    acc = torch.zeros(2, dtype=torch.int64, device=dev)
    for t in ts:
        v = t.detach().contiguous().view(-1).view(torch.int32).to(torch.int64)
        acc[0] += v.sum(); acc[1] += (v * (torch.arange(v.numel(), device=dev) % 1000003)).sum()
    return acc
def same(x):
    xs = [torch.zeros_like(x) for _ in range(world)]
    dist.all_gather(xs, x.contiguous())
    return all(torch.equal(y, xs[0]) for y in xs)
names = [n for n, _ in model.named_parameters()]
def report(tag):
    torch.cuda.synchronize()
    P = list(model.parameters())
    out = {"weights": same(h(P)), "arch": same(h(model.arch_parameters()))}
    if not out["weights"]:
        bad = [n for n, p in zip(names, P) if not same(h([p]))]
        out["n_bad"] = len(bad); out["bad"] = bad[:6]
    for k in ("vnorm", "R"):
        v = architect.last.get(k)
        if torch.is_tensor(v): out[k] = same(v.detach().reshape(1).float())
    if rank == 0: print(tag, out, flush=True)
report("init")
for i in range(2):
    stepper.alpha_step(train, valid, 1e-3, True); report(f"eager{i} after alpha_step")
    stepper.w_step(*train); report(f"eager{i} after w_step")
    gs = [p.grad for p in model.parameters() if p.grad is not None]
    if rank == 0: print("   grads identical:", end=" ")
    r = same(h(gs))
    if rank == 0: print(r, flush=True)
g = GraphedSearchStep(stepper, train, valid, 1e-3, unrolled=True)
report("after capture")
for i in range(3):
    g(); report(f"graph replay {i}")
dist.barrier()
os._exit(0)

import sys, os, copy, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lct-vqa_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcd_build, pcd_native
pcd_native.enable_emulation(pcd_build.build_emu())
import parity_cases as P
from helpers import rel_err
from pcdarts.genotypes import Genotype
from pcdarts.model import NetworkDerived
B, img = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(11)
geno = Genotype(normal=P.TEST_GENOTYPE["normal"], normal_concat=range(2, 6), reduce=P.TEST_GENOTYPE["reduce"], reduce_concat=range(2, 6))
net = NetworkDerived(16, 4, geno).train()
with torch.no_grad():
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d) and m.affine:
            m.weight.copy_(1.0 + 0.2 * torch.randn_like(m.weight)); m.bias.copy_(0.2 * torch.randn_like(m.bias))
ref64 = copy.deepcopy(net).double(); ref32 = copy.deepcopy(net)
x = torch.randn(B, 3, img, img); gy = None
y = net(x); gy = torch.randn(y.shape)
g_ours = torch.autograd.grad(y, list(net.parameters()), gy)
y32 = ref32(x, stock=True); g32 = torch.autograd.grad(y32, list(ref32.parameters()), gy)
y64 = ref64(x.double(), stock=True); g64 = torch.autograd.grad(y64, list(ref64.parameters()), gy.double())
print("y: ours", rel_err(y, y64), "stock32", rel_err(y32, y64))
eo = sorted([(rel_err(a, b), n) for a, b, (n, _) in zip(g_ours, g64, net.named_parameters())], reverse=True)
es = sorted([(rel_err(a, b), n) for a, b, (n, _) in zip(g32, g64, net.named_parameters())], reverse=True)
import statistics
print("ours  worst", eo[:4], "median", statistics.median(e for e, _ in eo), "share<=1e-4", sum(e <= 1e-4 for e, _ in eo) / len(eo))
print("stock worst", es[:4], "median", statistics.median(e for e, _ in es), "share<=1e-4", sum(e <= 1e-4 for e, _ in es) / len(es))

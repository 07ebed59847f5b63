"""One whole search step (unrolled alpha-step + w-step, eager, weight-grad overlap off) inside a cudaProfilerStart/Stop
range, for ncu launch lists:  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ..."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'lct-vqa_b200'))
sys.path.insert(0, ROOT)
import torch
from argparse import Namespace
import config
config.DEVICE = torch.device('cuda')
import bench
from pcdarts.architect_vqa import Architect
from search import SearchStep
from vqa_model import VqaModel
dev = torch.device('cuda')
torch.manual_seed(10)
first_order = len(sys.argv) > 1 and sys.argv[1] == 'first'
model = VqaModel(qst_vocab_size=17858, img_encoder_type='darts', **bench.DIMS).to(dev).train()
import pcd_flat
opt = pcd_flat.FlatAdam(model.parameters(), lr=1e-3)          # as bench.py: Adam / clip / axpy over the flat runs
arch = Architect(model, Namespace(arch_learn_rate=6e-4, arch_wt_decay=1e-3, qst_only=False))
arch.device_scalars = True
st = SearchStep(model, arch, opt)
tr = [t.to(dev) for t in bench.synth_batch(10, 64, 17858, 64)]
va = [t.to(dev) for t in bench.synth_batch(1010, 64, 17858, 64)]
for _ in range(2):
    st.step(tr, va, 1e-3, unrolled=not first_order)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
st.step(tr, va, 1e-3, unrolled=not first_order)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")

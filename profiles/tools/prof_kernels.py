import sys, os, time, collections
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,os.path.join(ROOT,'lct-vqa_b200')); sys.path.insert(0,ROOT)
import torch
from argparse import Namespace
import config; config.DEVICE=torch.device('cuda')
import bench
from pcdarts.architect_vqa import Architect
from search import SearchStep
from vqa_model import VqaModel
dev=torch.device('cuda')
torch.manual_seed(10)
model=VqaModel(qst_vocab_size=17858,img_encoder_type='darts',**bench.DIMS).to(dev).train()
opt=torch.optim.Adam(model.parameters(),lr=1e-3)
arch=Architect(model,Namespace(arch_learn_rate=6e-4,arch_wt_decay=1e-3,qst_only=False))
st=SearchStep(model,arch,opt)
tr=[t.to(dev) for t in bench.synth_batch(10,64,17858,64)]; va=[t.to(dev) for t in bench.synth_batch(11,64,17858,64)]
for _ in range(3): st.step(tr,va,1e-3)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    st.step(tr,va,1e-3); torch.cuda.synchronize()
agg=collections.defaultdict(lambda:[0.0,0])
for e in prof.events():
    if e.device_type==torch.autograd.DeviceType.CUDA:
        agg[e.name][0]+=e.device_time; agg[e.name][1]+=1
tot=sum(v[0] for v in agg.values()); ours=sum(v[0] for k,v in agg.items() if 'pcd::' in k)
print(f"total kernel ms {tot/1e3:.2f}  ours {ours/1e3:.2f}  other {(tot-ours)/1e3:.2f}")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    if 'pcd::pcd_kernel' in k: continue
    if v[0]<30: continue
    print(f"{v[0]/1e3:8.3f} ms n={v[1]:5d}  {k[:150]}")

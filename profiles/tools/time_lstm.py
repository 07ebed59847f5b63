import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,os.path.join(ROOT,'lct-vqa_b200')); sys.path.insert(0,ROOT)
import torch
from pcd_ops import lstm_forward
dev='cuda'
T,B,E,H=30,64,300,512
lstm=torch.nn.LSTM(E,H,1).to(dev)
x=torch.randn(T,B,E,device=dev,requires_grad=True); h0=torch.randn(1,B,H,device=dev,requires_grad=True)
G=torch.randn(T,B,H,device=dev)
torch.backends.cudnn.allow_tf32=False
def ours():
    out,(h,c)=lstm_forward(lstm,x,h0,h0); (out*G).sum().backward()
def ref():
    out,(h,c)=lstm(x,(h0,h0)); (out*G).sum().backward()
from torch.profiler import profile, ProfilerActivity
import collections
for name,fn in (("ours",ours),("cudnn",ref)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    agg=collections.defaultdict(lambda:[0.0,0])
    for e in prof.events():
        if e.device_type==torch.autograd.DeviceType.CUDA: agg[e.name][0]+=e.device_time; agg[e.name][1]+=1
    tot=sum(v[0] for v in agg.values())
    print(name,"total kernel us",round(tot))
    for k,v in sorted(agg.items(),key=lambda kv:-kv[1][0])[:8]: print(f"   {v[0]:8.1f} us n={v[1]:3d} {k[:90]}")

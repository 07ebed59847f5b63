import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,os.path.join(ROOT,'lct-vqa_b200')); sys.path.insert(0,ROOT)
import torch, pcd_native as N
lib=N.load_cuda()
dev=torch.device('cuda')
def run(M,Nn,K,lda=None,ldb=None,ldc=None,bias=True,split=1,reps=5):
    lda=lda or K; ldb=ldb or K; ldc=ldc or Nn
    g=torch.Generator(device='cpu').manual_seed(M+Nn+K)
    A=torch.randn(M,lda,generator=g).to(dev); B=torch.randn(Nn,ldb,generator=g).to(dev); b=torch.randn(Nn,generator=g).to(dev) if bias else None
    C=torch.full((M,ldc),float('nan'),device=dev)
    st=torch.cuda.current_stream().cuda_stream
    def call():
        rc=lib.pcd_gemm_tn_3xtf32(A.data_ptr(),lda,B.data_ptr(),ldb,C.data_ptr(),ldc,M,Nn,K,b.data_ptr() if bias else None,split,st)
        N.check(lib,rc,'gemm')
    call(); torch.cuda.synchronize()
    ref=(A[:,:K].double()@B[:,:K].double().T)+(b.double() if bias else 0)
    got=C[:,:Nn].double()
    err=(got-ref).abs().max().item()/ref.abs().max().item()
    ref32=torch.addmm(b,A[:,:K],B[:,:K].T) if bias else A[:,:K]@B[:,:K].T
    err32=(ref32.double()-ref).abs().max().item()/ref.abs().max().item()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    for _ in range(2): call()
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    e0.record()
    for _ in range(reps): ref32=A[:,:K]@B[:,:K].T
    e1.record(); torch.cuda.synchronize()
    ms32=e0.elapsed_time(e1)/reps
    print(f"M={M} N={Nn} K={K} split={split}: rel err {err:.2e} (cuBLAS fp32 {err32:.2e})  {ms*1e3:.0f} us = {2*M*Nn*K/ms/1e9:.1f} TFLOP/s   cuBLAS sgemm {ms32*1e3:.0f} us = {2*M*Nn*K/ms32/1e9:.1f} TFLOP/s", flush=True)
    assert err<6e-5, err
for cfg in (1,2,3,4):
  lib.pcd_gemm_debug_cfg(cfg); print("cfg",cfg)
  run(1920,17858,512,ldc=17860)
  run(1920,512,17860,split=16,bias=False)
  run(17858,512,1920,bias=False)
lib.pcd_gemm_debug_cfg(0)
run(128,128,32,bias=False)
run(128,128,64)
run(256,256,512)
run(100,70,36)            # ragged everything
run(1920,17858,512,ldc=17860)          # vocab projection forward
run(1920,512,17860,split=8,bias=False) # dH  (K = padded vocab)
run(1920,512,17858,lda=17860,ldb=17860,split=16,bias=False)
run(1920,512,17858,lda=17860,ldb=17860,split=32,bias=False)
run(17858,512,1920,bias=False)         # dW
print("gemm ok")

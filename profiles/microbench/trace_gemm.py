import torch, sys
sys.path.insert(0,'/root/repo')
from graphvqa_b200 import _cabi
DEV='cuda:0'
g=torch.Generator().manual_seed(1)
a=torch.randn(7680,512,generator=g).to(DEV); b=(torch.randn(2048,512,generator=g)*0.05).to(DEV)
hi,lo=_cabi.split_tf32(b); out=torch.empty(7680,2048,device=DEV)
for _ in range(3): _cabi.proj_gemm_3xtf32(a,hi,lo,out=out)
flags=int(sys.argv[1]) if len(sys.argv)>1 else 0
_cabi.lib().gvqa_debug_set_gemm_flags(flags)
tr=torch.zeros(1100*8,dtype=torch.int64,device=DEV)
_cabi.lib().gvqa_debug_set_gemm_trace(tr.data_ptr())
_cabi.proj_gemm_3xtf32(a,hi,lo,out=out); torch.cuda.synchronize()
_cabi.lib().gvqa_debug_set_gemm_trace(None)
t=tr.cpu().view(1100,8)
t0=int(t[0,0])
kb=512//32
print("it  prod  conv0 conv1 mma0 mma1 looptop waited  (cycles since first producer issue)")
for it in list(range(0,24))+list(range(60,72)):
    r=[int(x)-t0 if int(x)>0 else -1 for x in t[it,:7]]
    print(it, r)
print("epilogue tiles:")
for i in range(7):
    r=[int(x)-t0 for x in t[1024+i,:2]]
    print(i, r, 'dur', r[1]-r[0])

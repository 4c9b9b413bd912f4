"""Eight launches of the fp16-split projection GEMM at the cfg2 shape (target of `ncu --set full -k regex:proj_gemm_3xf16`)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200 import _cabi
DEV = 'cuda:0'
g = torch.Generator().manual_seed(1)
a = torch.randn(7680, 512, generator=g).to(DEV); b = (torch.randn(2064, 512, generator=g) * 0.05).to(DEV)
hi, lo = _cabi.split_f16(b); out = torch.empty(7680, 2064, device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for _ in range(8):
    flush.zero_()
    _cabi.proj_gemm_3xf16(a, hi, lo, out=out)
torch.cuda.synchronize()

"""Attribute the projection GEMM's time: run it with producer / converter / epilogue stages disabled."""
import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200 import _cabi
DEV = 'cuda:0'
g = torch.Generator().manual_seed(1)
M, N, K = 7680, 2064, 512
a = torch.randn(M, K, generator=g).to(DEV); b = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
hi, lo = _cabi.split_tf32(b); out = torch.empty(M, N, device=DEV)
def t(flags, reps=20):
    _cabi.lib().gvqa_debug_set_gemm_flags(flags)
    for _ in range(3): _cabi.proj_gemm_3xtf32(a, hi, lo, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _cabi.proj_gemm_3xtf32(a, hi, lo, out=out)
    e1.record(); torch.cuda.synchronize()
    _cabi.lib().gvqa_debug_set_gemm_flags(0)
    return e0.elapsed_time(e1) / reps * 1e3
names = {0: "full", 1: "no TMA", 2: "no converters", 4: "no epilogue", 3: "no TMA+conv (MMA issue + epilogue)",
         5: "no TMA, no epi", 6: "no conv, no epi", 7: "MMA issue only"}
for f in (0, 1, 2, 4, 3, 5, 6, 7):
    print("flags=%d %-36s %8.1f us" % (f, names[f], t(f)))

"""Attribute the fp16-split projection GEMM's time: run it with stages disabled (gvqa_debug_set_gemm_flags:
1 no TMA loads, 2 converters idle, 4 no epilogue, 8 no MMAs, 32 one CTA per tile instead of two-CTA pairs,
64 double-buffered hi*hi accumulator (one per tile) instead of the three accumulators per tile)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200 import _cabi
DEV = 'cuda:0'
g = torch.Generator().manual_seed(1)
M, N, K = 7680, 2064, 512
a = torch.randn(M, K, generator=g).to(DEV); b = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
hi, lo = _cabi.split_f16(b); out = torch.empty(M, N, device=DEV)
want = a.double() @ b.double().t()
def t(flags, reps=30):
    _cabi.lib().gvqa_debug_set_gemm_flags(flags)
    for _ in range(3): _cabi.proj_gemm_3xf16(a, hi, lo, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _cabi.proj_gemm_3xf16(a, hi, lo, out=out)
    e1.record(); torch.cuda.synchronize()
    _cabi.lib().gvqa_debug_set_gemm_flags(0)
    return e0.elapsed_time(e1) / reps * 1e3
names = {0: "full", 1: "no TMA", 2: "no converters", 4: "no epilogue", 8: "no MMA", 3: "no TMA+conv (MMA + epilogue)",
         7: "MMA issue only", 9: "conv + epilogue only (no TMA, no MMA)", 13: "converters only", 14: "TMA only", 11: "epilogue only"}
for base, label in ((0, "CTA pairs, three accumulators per tile (default)"),
                    (64, "CTA pairs, hi*hi accumulator per tile (DB, opt-in)"), (32, "single CTA, three accumulators")):
    _cabi.lib().gvqa_debug_set_gemm_flags(base)
    got = _cabi.proj_gemm_3xf16(a, hi, lo)
    _cabi.lib().gvqa_debug_set_gemm_flags(0)
    print("== %s: max rel err vs fp64 %.2e" % (label, float((got.double() - want).abs().max() / want.abs().max())), flush=True)
    for f in ((0, 0, 0) if len(sys.argv) > 1 and sys.argv[1] == 'quick' else (0, 1, 2, 4, 8, 3, 7, 9, 13, 14, 11)):
        print("flags=%2d %-40s %8.1f us" % (f, names[f], t(base | f)), flush=True)

"""Accuracy of the fp16-split GEMM with ONE hi*hi accumulator chain over the whole K (debug flag 64) against the
default two K-half chains: error vs float64 relative to the test bar max(4 x fp32 error, 2e-6 x scale)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200 import _cabi
DEV = 'cuda:0'
shapes = [(128, 256, 64), (256, 512, 128), (7680, 2064, 512), (59, 1200, 300), (1000, 2048, 512), (130, 260, 36),
          (7680, 1216, 300), (300, 32, 512), (2048, 512, 1024), (1024, 512, 2048)]
for m, n, k in shapes:
    for trial in range(2):
        g = torch.Generator().manual_seed(m + n + k + 1 + trial)
        a = torch.randn(m, k, generator=g) * (3.0 if trial == 0 else 1.0)
        if trial == 1:
            a = a.abs()                    # all-positive products: the worst case for a truncating accumulator
        b = torch.randn(n, k, generator=g) * 0.05
        if trial == 1:
            b = b.abs()
        hi, lo = _cabi.split_f16(b.to(DEV))
        want = a.double() @ b.double().t()
        err32 = float(((a @ b.t()).double() - want).abs().max())
        scale = float(want.abs().max())
        bar = max(4 * err32, 2e-6 * scale)
        res = []
        for flags in (0, 64):
            _cabi.lib().gvqa_debug_set_gemm_flags(flags)
            out = _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo).cpu()
            d = out.double() - want
            res.append((float(d.abs().max()) / bar, float(d.mean()) / scale))
        _cabi.lib().gvqa_debug_set_gemm_flags(0)
        print("m=%5d n=%5d k=%5d %s  two chains: err/bar %.3f bias/scale %+.2e | one chain: err/bar %.3f bias/scale %+.2e"
              % (m, n, k, "pos" if trial else "rnd", res[0][0], res[0][1], res[1][0], res[1][1]), flush=True)

"""GPU test of the tcgen05 3xTF32 projection GEMM against a float64 reference: fp32-level accuracy
(the 1e-4 parity bar of the hop stack needs ~1e-6 relative here), ragged M/N/K tails, strided A."""
import pytest
import torch

from graphvqa_b200 import _cabi

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("m,n,k", [(128, 256, 32), (256, 512, 64), (7680, 2048, 512), (59, 1200, 300),
                                   (1000, 2048, 512), (130, 260, 36), (7680, 1200, 300)])
def test_proj_gemm_3xtf32_matches_fp64(m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g)
    b = torch.randn(n, k, generator=g) * 0.05
    hi, lo = _cabi.split_tf32(b.to(DEV))
    assert torch.equal((hi.double() + lo.double()).float().cpu(), b) or \
        ((hi.double() + lo.double()).cpu() - b.double()).abs().max() <= 2.0 ** -21 * float(b.abs().max())
    out = _cabi.proj_gemm_3xtf32(a.to(DEV), hi, lo).cpu()
    want = a.double() @ b.double().t()
    ref32 = (a @ b.t()).double()
    err = (out.double() - want).abs().max()
    err32 = (ref32 - want).abs().max()
    scale = float(want.abs().max())
    assert err <= max(4 * float(err32), 2e-6 * scale), "3xTF32 err %g vs fp32 err %g (scale %g)" % (err, err32, scale)


def test_proj_gemm_strided_a_and_repeatability():
    g = torch.Generator().manual_seed(3)
    big = torch.randn(300, 700, generator=g).to(DEV)
    a = big[:, 100:612]                       # row stride 700, 16-byte aligned start
    b = (torch.randn(384, 512, generator=g) * 0.05).to(DEV)
    hi, lo = _cabi.split_tf32(b)
    o1 = _cabi.proj_gemm_3xtf32(a, hi, lo)
    o2 = _cabi.proj_gemm_3xtf32(a, hi, lo)
    assert torch.equal(o1, o2)
    want = a.double() @ b.double().t()
    assert (o1.double() - want).abs().max() <= 2e-6 * float(want.abs().max())


def test_proj_gemm_timing_report():
    """Not an assertion on speed: prints the kernel time next to cuBLAS fp32 for the cfg2 shape."""
    g = torch.Generator().manual_seed(1)
    a = torch.randn(7680, 512, generator=g).to(DEV)
    b = (torch.randn(2048, 512, generator=g) * 0.05).to(DEV)
    hi, lo = _cabi.split_tf32(b)
    out = torch.empty(7680, 2048, device=DEV)
    for fn, name in ((lambda: _cabi.proj_gemm_3xtf32(a, hi, lo, out=out), "3xtf32"),
                     (lambda: torch.mm(a, b.t(), out=out), "cublas fp32")):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print("%s: %.1f us" % (name, e0.elapsed_time(e1) / 20 * 1e3))

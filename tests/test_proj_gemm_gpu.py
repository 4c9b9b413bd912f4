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


# ---- fp16-split variant (gvqa_proj_gemm_3xf16): same accuracy bar, plus the range guard --------------------
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (256, 512, 128), (7680, 2064, 512), (59, 1200, 300),
                                   (1000, 2048, 512), (130, 260, 36), (7680, 1216, 300), (300, 32, 512)])
def test_proj_gemm_3xf16_matches_fp64(m, n, k):
    g = torch.Generator().manual_seed(m + n + k + 1)
    a = torch.randn(m, k, generator=g) * 3.0
    b = torch.randn(n, k, generator=g) * 0.05
    hi, lo = _cabi.split_f16(b.to(DEV))
    recon = hi.double() + lo.double() / 2048.0
    assert (recon.cpu() - b.double()).abs().max() <= 2.0 ** -21 * float(b.abs().max())
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo, overflow=flag).cpu()
    assert int(flag) == 0
    want = a.double() @ b.double().t()
    err = (out.double() - want).abs().max()
    err32 = ((a @ b.t()).double() - want).abs().max()
    scale = float(want.abs().max())
    assert err <= max(4 * float(err32), 2e-6 * scale), "3xF16 err %g vs fp32 err %g (scale %g)" % (err, err32, scale)


@pytest.mark.parametrize("m,n,k,relu", [(256, 512, 512, True), (7680, 512, 512, False), (59, 1842, 512, True),
                                        (15360, 300, 900, True), (256, 1842, 1536, False)])
def test_linear_epilogue_bias_relu_in_the_gemm(m, n, k, relu):
    """gvqa_linear_3xf16 = nn.Linear (+ ReLU) in one launch, incl. N not a multiple of 4 / 64 (the answer head's 1842
    columns live in a 1844-wide buffer) and ragged M: against float64, and the plain product + eager epilogue."""
    from graphvqa_b200.tc_linear import TensorCoreLinear
    g = torch.Generator().manual_seed(m + n)
    a = torch.randn(m, k, generator=g)
    w = torch.randn(n, k, generator=g) / k ** 0.5
    b = torch.randn(n, generator=g)
    lin = TensorCoreLinear(mode="3xf16")
    lin.flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    got = lin(a.to(DEV), w.to(DEV), b.to(DEV), relu=relu)
    assert got.shape == (m, n) and int(lin.flag) == 0
    want = a.double() @ w.double().t() + b.double()
    if relu:
        want = want.clamp(min=0)
    assert (got.cpu().double() - want).abs().max() <= 2e-5
    plain = lin(a.to(DEV), w.to(DEV)) + b.to(DEV)
    assert torch.equal(got, torch.relu(plain) if relu else plain)       # same accumulators, same single rounding
    ref = TensorCoreLinear(mode="3xtf32")(a.to(DEV), w.to(DEV), b.to(DEV), relu=relu)
    assert (ref - got).abs().max() <= 2e-5


def test_proj_gemm_3xf16_small_and_large_magnitudes():
    """Blocks of A/B in fp16's subnormal range (values ~1e-6) next to blocks of order 1e3-1e4: the error stays at
    fp32 level relative to each output row (tiny operands lose relative, not absolute, precision)."""
    g = torch.Generator().manual_seed(9)
    a = torch.randn(256, 256, generator=g)
    a[:, :64] *= 1e-6
    a[:, 64:128] *= 2.0e3
    b = torch.randn(128, 256, generator=g) * 0.05
    b[:32] *= 1e-5
    hi, lo = _cabi.split_f16(b.to(DEV))
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo, overflow=flag).cpu()
    want = a.double() @ b.double().t()
    # rows of `want` are dominated by the 2e4-scaled block; the bar is relative to each row's magnitude
    rel = ((out.double() - want).abs().max(1).values / want.abs().max(1).values).max()
    assert float(rel) <= 2e-6 and int(flag) == 0


def test_proj_gemm_3xf16_flags_out_of_range_input():
    a = torch.ones(128, 64)
    a[5, 7] = 7.0e4
    b = torch.ones(128, 64) * 0.01
    hi, lo = _cabi.split_f16(b.to(DEV))
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo, overflow=flag)
    assert int(flag) == 1


def test_proj_gemm_3xf16_timing_report():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(7680, 512, generator=g).to(DEV)
    b = (torch.randn(2064, 512, generator=g) * 0.05).to(DEV)
    hi, lo = _cabi.split_f16(b)
    thi, tlo = _cabi.split_tf32(b)
    out = torch.empty(7680, 2064, device=DEV)
    for fn, name in ((lambda: _cabi.proj_gemm_3xf16(a, hi, lo, out=out), "3xf16"),
                     (lambda: _cabi.proj_gemm_3xtf32(a, thi, tlo, out=out), "3xtf32")):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print("%s: %.1f us" % (name, e0.elapsed_time(e1) / 20 * 1e3))


def test_proj_gemm_3xf16_batched_matches_fp64():
    g = torch.Generator().manual_seed(12)
    z, m, n, k = 5, 256, 528, 512
    a = torch.randn(z, m, k, generator=g)
    b = torch.randn(z * n, k, generator=g) * 0.05
    hi, lo = _cabi.split_f16(b.to(DEV))
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = _cabi.proj_gemm_3xf16_batched(a.to(DEV), hi.unflatten(0, (z, n)), lo.unflatten(0, (z, n)), overflow=flag).cpu()
    want = torch.einsum("zmk,znk->zmn", a.double(), b.view(z, n, k).double())
    assert int(flag) == 0
    assert (out.double() - want).abs().max() <= 2e-6 * float(want.abs().max())


@pytest.mark.parametrize("shapes", [
    [(7680, 2064, 512), (5, 256, 528, 512), (15360, 32, 512)],      # gat_seq at cfg2: hop-0 projection + both pre-pass products
    [(120, 1216, 300), (5, 4, 320, 512), (236, 32, 300)],           # reference dims, cfg1-sized batch, three different K
    [(59, 260, 36), (130, 16, 64)],
])
def test_proj_gemm_3xf16_grouped_equals_separate_launches(shapes):
    """One persistent launch over several products returns bit for bit what one launch per product returns
    (same tile decomposition and k order), and both sit at fp32-level error against float64."""
    g = torch.Generator().manual_seed(len(shapes) + shapes[0][0])
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    problems, separate, wants = [], [], []
    for shp in shapes:
        if len(shp) == 4:
            z, m, n, k = shp
            a = (torch.randn(z, m, k, generator=g) * 2.0).to(DEV)
            b = (torch.randn(z * n, k, generator=g) * 0.05).to(DEV)
            hi, lo = (t.unflatten(0, (z, n)) for t in _cabi.split_f16(b))
            separate.append(_cabi.proj_gemm_3xf16_batched(a, hi, lo, overflow=flag))
            wants.append(torch.einsum("zmk,znk->zmn", a.double(), b.view(z, n, k).double()))
        else:
            m, n, k = shp
            a = (torch.randn(m, k, generator=g) * 2.0).to(DEV)
            b = (torch.randn(n, k, generator=g) * 0.05).to(DEV)
            hi, lo = _cabi.split_f16(b)
            separate.append(_cabi.proj_gemm_3xf16(a, hi, lo, overflow=flag))
            wants.append(a.double() @ b.double().t())
        problems.append((a, hi, lo, None))
    outs = _cabi.proj_gemm_3xf16_grouped(problems, overflow=flag)
    assert int(flag) == 0
    for out, sep, want in zip(outs, separate, wants):
        assert torch.equal(out, sep)
        assert (out.double() - want).abs().max() <= 2e-6 * float(want.abs().max())


def test_proj_gemm_3xf16_grouped_rejects_bad_counts_and_skips_empty():
    a = torch.randn(64, 64, device=DEV)
    hi, lo = _cabi.split_f16(torch.randn(32, 64, device=DEV))
    with pytest.raises(ValueError):
        _cabi.proj_gemm_3xf16_grouped([], None)
    with pytest.raises(ValueError):
        _cabi.proj_gemm_3xf16_grouped([(a, hi, lo, None)] * 4, None)
    empty = torch.empty(0, 64, device=DEV)
    outs = _cabi.proj_gemm_3xf16_grouped([(empty, hi, lo, None), (a, hi, lo, None)], None)
    assert outs[0].shape == (0, 32)
    assert torch.equal(outs[1], _cabi.proj_gemm_3xf16(a, hi, lo))


@pytest.mark.parametrize("m,n,k", [(7680, 2064, 512), (1000, 2048, 512), (59, 1200, 300), (300, 32, 64)])
def test_proj_gemm_3xf16_double_buffered_accumulator_variant(m, n, k):
    """Opt-in variant (GVQA_GEMM_DB=1 / debug flag 64): the two hi*hi accumulators hold consecutive tiles instead of
    the two K-halves of one tile.  Same accuracy bar as the default at K <= 512."""
    g = torch.Generator().manual_seed(m + n + k + 7)
    a = torch.randn(m, k, generator=g) * 3.0
    b = torch.randn(n, k, generator=g) * 0.05
    hi, lo = _cabi.split_f16(b.to(DEV))
    want = a.double() @ b.double().t()
    err32 = ((a @ b.t()).double() - want).abs().max()
    scale = float(want.abs().max())
    base = _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo).cpu()
    try:
        _cabi.lib().gvqa_debug_set_gemm_flags(64)
        out = _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo).cpu()
    finally:
        _cabi.lib().gvqa_debug_set_gemm_flags(0)
    assert (out.double() - want).abs().max() <= max(4 * float(err32), 2e-6 * scale)
    assert (out - base).abs().max() <= 4e-6 * scale

"""Per-k-block timeline of CTA 0 of the fp16-split GEMM (clock64 stamps, see GVQA_F16_TRACE in proj_gemm_f16.cu)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200 import _cabi
DEV = 'cuda:0'
g = torch.Generator().manual_seed(1)
a = torch.randn(7680, 512, generator=g).to(DEV); b = (torch.randn(2064, 512, generator=g) * 0.05).to(DEV)
hi, lo = _cabi.split_f16(b); out = torch.empty(7680, 2064, device=DEV)
for _ in range(3): _cabi.proj_gemm_3xf16(a, hi, lo, out=out)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_cabi.lib().gvqa_debug_set_gemm_flags(flags)
tr = torch.zeros(1100 * 8, dtype=torch.int64, device=DEV)
_cabi.lib().gvqa_debug_set_gemm_trace(tr.data_ptr())
_cabi.proj_gemm_3xf16(a, hi, lo, out=out); torch.cuda.synchronize()
_cabi.lib().gvqa_debug_set_gemm_trace(None)
_cabi.lib().gvqa_debug_set_gemm_flags(0)
t = tr.cpu().view(1100, 8)
t0 = int(t[0, 0])
print("flags=%d  cycles since the first producer issue" % flags)
print("it   prod  landed conv_go conv_done | mma_top a_rdy0 a_rdy1 committed")
for it in range(0, 56):
    r = [int(x) - t0 if int(x) > 0 else -1 for x in t[it]]
    print("%2d %6d %6d %6d %6d | %6d %6d %6d %6d" % (it, *r))
print("tile  mma_wait_acc  mma_go | epi_acc_full  tmem_released  stores_issued")
for i in range(7):
    r = [int(x) - t0 if int(x) > 0 else -1 for x in t[1024 + i, :5]]
    print("%2d %8d %8d | %8d %8d %8d" % (i, *r))

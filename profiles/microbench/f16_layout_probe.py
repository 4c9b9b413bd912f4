import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200 import _cabi
DEV = 'cuda:0'
g = torch.Generator().manual_seed(0)
M, N, K = 128, 128, 64
b = (torch.randint(1, 200, (N, K), generator=g).float() / 256.0)   # exactly representable in fp16, distinct-ish
hi, lo = _cabi.split_f16(b.to(DEV))
for k0 in (0, 1, 2, 3, 7, 8, 15, 16, 17, 31, 32, 33, 48, 63):
    a = torch.zeros(M, K); a[:, k0] = 1.0
    out = _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo).cpu()
    row = out[0]                       # should equal b[:, k0]
    match = [k for k in range(K) if torch.allclose(row, b[:, k], atol=1e-6)]
    print("k0=%2d -> output row 0 matches B column(s) %s ; out[0,:4]=%s want=%s" % (k0, match, [round(float(v), 4) for v in row[:4]], [round(float(v), 4) for v in b[:4, k0]]))
# and the converse: which A row index maps where
a = torch.zeros(M, K); a[5, 0] = 1.0
out = _cabi.proj_gemm_3xf16(a.to(DEV), hi, lo).cpu()
print("nonzero output rows for A[5,0]=1:", out.abs().sum(1).nonzero().flatten().tolist()[:10])

"""GINE / GCN / LCGN sequence modules at the BASELINE shapes (cfg3: gine_seq, B=256, F=512; cfg5: lcgn_seq, 128
graphs per GPU, 300 -> 512, 4 iterations): ms per forward, eager launches, non-bug-faithful (conv results used)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200 import gcn_gine, lcgn as lcgn_mod
from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology
dev = torch.device('cuda:0')
def timed(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
g = torch.Generator().manual_seed(0)
B = 256
ei, batch, mx = synthetic_topology(B, 30, 60, seed=1); ei, batch = ei.to(dev), batch.to(dev)
n, e = batch.numel(), ei.size(1)
x = torch.randn(n, 512, generator=g).to(dev); ea = torch.randn(e, 512, generator=g).to(dev); ins = torch.randn(5, B, 512, generator=g).to(dev)
csr = GraphCSR.build(ei, batch, B)
with torch.no_grad():
    m = gcn_gine.gine_seq(512, 512, 512, bug_faithful=False).eval().to(dev)
    ms = timed(lambda: m(x, ei, ea, ins, batch, csr=csr)); print("gine_seq cfg3 (B=256, F=512, 5 hops, conv results used): %.3f ms = %.0f questions/s" % (ms, B / ms * 1e3))
    m = gcn_gine.gcn_seq(512, 512, 512, bug_faithful=False).eval().to(dev)
    ms = timed(lambda: m(x, ei, ins, batch, csr=csr)); print("gcn_seq        (B=256, F=512, 5 hops, conv results used): %.3f ms = %.0f questions/s" % (ms, B / ms * 1e3))
    B5 = 128
    ei5, b5, _ = synthetic_topology(B5, 30, 60, seed=2); ei5, b5 = ei5.to(dev), b5.to(dev)
    n5 = b5.numel()
    m = lcgn_mod.lcgn_seq(300, 512, 1, 5).eval().to(dev)
    x5 = torch.randn(n5, 300, generator=g).to(dev); q = torch.randn(B5, 512, generator=g).to(dev); lo = torch.randn(12, B5, 512, generator=g).to(dev)
    xc = torch.randn(n5, 512, generator=g).to(dev); csr5 = GraphCSR.build(ei5, b5, B5)
    ms = timed(lambda: m(x5, ei5, b5, q, lo, x_ctx_init=xc, csr=csr5)); print("lcgn_seq cfg5 per GPU (B=128, 4 iterations): %.3f ms = %.0f questions/s" % (ms, B5 / ms * 1e3))

"""Every hand-written kernel of SURVEY.md section 8 at the BASELINE.json shapes: duration and algorithmic
GB/s against the measured HBM peak.  Launches are queued behind a busy stream (no CPU launch gap) and L2 is
flushed with CLEAN lines (a 256 MB buffer is read, not written: a zero-fill would leave 126 MB of dirty lines
whose write-back is then billed to the kernel under test) before each timed launch; durations are CUDA-event
brackets, so each includes the ~2.7 us an event pair reads around nothing (event_overhead.py)."""
import json, os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphvqa_b200 import _cabi, gat_skip as eng, gcn_gine, lcgn as lcgn_mod
from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology
from graphvqa_b200.my_graph_layernorm import LayerNorm
dev = torch.device('cuda:0')
peaks = os.path.join('/root/repo', 'MEASURED_PEAKS.json')
PEAK = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)      # 256 MB
def timed(fn, reps=24):
    ts = []
    for _ in range(reps):
        flush.sum(); torch.cuda._sleep(150000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[4:])
    return ts[len(ts) // 2]
def report(label, us, nbytes):
    gbs = nbytes / us / 1e3
    print("%-58s %8.1f us  %7.2f MB  %6.0f GB/s  %4.1f %% of %.0f" % (label, us, nbytes / 1e6, gbs, 100 * gbs / PEAK, PEAK))
def graphs(b, n, e, seed=1234):
    ei, batch, mx = synthetic_topology(b, n, e, seed=seed)
    csr = GraphCSR.build(ei.to(dev), batch.to(dev), b, max_nodes_per_graph=mx, max_in_edges_per_graph=synthetic_topology.last_max_edges)
    return ei.to(dev), batch.to(dev), csr
g = torch.Generator().manual_seed(0)
def rnd(*s): return torch.randn(*s, generator=g).to(dev)

def hop_case(label, b, n_, e_, f, h=4):
    ei, batch, csr = graphs(b, n_, e_)
    n, e = batch.numel(), ei.size(1)
    x_l = rnd(n, h * f + 16); a_edge = rnd(e, 32); hprev = rnd(n, f); out = torch.empty(n, f, device=dev)
    gb = rnd(b, f); ag = rnd(b, h); bias = rnd(f); sc = rnd(f); sh = rnd(f)
    fn = lambda: _cabi.gat_hop(x_l, x_l[:, h * f:h * f + 2 * h], a_edge, csr.as_dict(), h, f, out, lde=32, graph_bias=gb,
                               a_graph=ag, h_prev=hprev, bias=bias, ep_scale=sc, ep_shift=sh, epilogue=_cabi.EPI_AFFINE_RELU,
                               inputs_older_than_predecessor=True, **csr.hints())
    report(label, timed(fn), bench.hop_bytes(n, e, h, f))

hop_case("K1 fused GAT hop, cfg2 (B=256, 30/60, F=512)", 256, 30, 60, 512)
hop_case("K1 fused GAT hop, cfg2 at reference dims F=300", 256, 30, 60, 300)
hop_case("K1 fused GAT hop, cfg4 per GPU (B=128, 200/800, F=512)", 128, 200, 800, 512)
hop_case("K1 fused GAT hop, cfg4 on ONE GPU (B=1024, 200/800)", 1024, 200, 800, 512)

# K2 graph LayerNorm
ei, batch, csr = graphs(256, 30, 60)
n = batch.numel()
ln = LayerNorm(512).to(dev).eval(); x = rnd(n, 512)
with torch.no_grad():
    report("K2 graph LayerNorm, cfg2 (N=7680, F=512)", timed(lambda: ln(x, batch, num_graphs=256, csr=csr)), 8 * n * 512 + 4 * n)
# K3 GINE aggregate (cfg3)
e = ei.size(1)
h_ = rnd(n, 512); ea = rnd(e, 512); ins = rnd(256, 512); z = torch.empty(n, 1024, device=dev)
report("K3 GINE gather-relu-scatter, cfg3 (F=512, D=512)", timed(lambda: _cabi.gine_aggregate(h_, ea, ins, csr.as_dict(), 0.0, out=z)),
       4 * (n * 512 + e * 512 + 256 * 512 + n * 1024) + 4 * (n + 1 + e))
# K4 GCN aggregate
xw = rnd(n, 512); gt = rnd(256, 512); bias = rnd(512); outc = torch.empty(n, 512, device=dev)
dinv = _cabi.gcn_degree(csr.as_dict(), n, dev)
report("K4 GCN normalised aggregate, cfg2 (C=512)", timed(lambda: _cabi.gcn_aggregate(xw, gt, dinv, bias, csr.as_dict(), out=outc)),
       4 * (2 * n * 512 + e) + 4 * (n + 1 + e))
# K5 LCGN hop (cfg5: 128 graphs per GPU, C=512)
ei5, batch5, csr5 = graphs(128, 30, 60)
n5 = batch5.numel(); e5 = ei5.size(1)
proj = rnd(n5, 3 * 512); pc = rnd(128, 512); cc = rnd(128, 512); b5 = rnd(512); o5 = torch.empty(n5, 512, device=dev)
report("K5 LCGN hop, cfg5 per GPU (B=128, C=512)", timed(lambda: _cabi.lcgn_hop(proj[:, :512], proj[:, 512:1024], proj[:, 1024:], pc, cc, b5, csr5.as_dict(), 0.2, out=o5)),
       4 * (4 * n5 * 512 + 2 * 128 * 512) + 4 * (n5 + 1 + e5))
# projection GEMM (tensor-bound, for context)
a = rnd(7680, 512); w = rnd(2064, 512) * 0.05; hi, lo = _cabi.split_tf32(w); c = torch.empty(7680, 2064, device=dev)
us = timed(lambda: _cabi.proj_gemm_3xtf32(a, hi, lo, out=c))
print("%-58s %8.1f us  %.0f TFLOP/s fp32-equivalent (3 tf32 MMAs per product: %.0f TFLOP/s tf32)" % (
    "projection GEMM 3xTF32 [7680x512]x[2064x512]^T", us, 2 * 7680 * 2064 * 512 / us / 1e6, 6 * 7680 * 2064 * 512 / us / 1e6))

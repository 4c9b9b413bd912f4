"""Fused GAT hop: kernel variants side by side at the BASELINE shapes (bracketed CUDA-event timing behind a clean
L2 flush, like kernel_roofline.py) and, with --trace, the per-CTA timeline of one launch.

    python profiles/microbench/hop_variants.py [--variants 3,4] [--npc 0,4,8] [--trace] [--cases cfg2,f300,cfg4]
"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
import bench
from graphvqa_b200 import _cabi
from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="3,5,4")
ap.add_argument("--npc", default="0")
ap.add_argument("--cases", default="cfg2,f300,cfg4,cfg4x8")
ap.add_argument("--trace", action="store_true")
ap.add_argument("--reps", type=int, default=40)
args = ap.parse_args()
dev = torch.device('cuda:0')
peaks = os.path.join('/root/repo', 'MEASURED_PEAKS.json')
PEAK = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)      # 256 MB, read (clean lines)
g = torch.Generator().manual_seed(0)
def rnd(*s): return torch.randn(*s, generator=g).to(dev)

def timed(fn, reps):
    ts = []
    for _ in range(reps):
        flush.sum(); torch.cuda._sleep(150000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[4:])
    return ts[len(ts) // 2], ts[0]

CASES = {"cfg2": (256, 30, 60, 512), "f300": (256, 30, 60, 300), "cfg4": (128, 200, 800, 512), "cfg4x8": (1024, 200, 800, 512)}
for case in args.cases.split(","):
    b, n_, e_, f = CASES[case]
    h = 4
    ei, batch, mx = synthetic_topology(b, n_, e_, seed=1234)
    csr = GraphCSR.build(ei.to(dev), batch.to(dev), b, max_nodes_per_graph=mx, max_in_edges_per_graph=synthetic_topology.last_max_edges)
    n, e = batch.numel(), ei.size(1)
    x_l = rnd(n, h * f + 16); a_edge = rnd(e, 32); hprev = rnd(n, f); out = torch.empty(n, f, device=dev)
    gb = rnd(b, f); ag = rnd(b, h); bias = rnd(f); sc = rnd(f); sh = rnd(f)
    nbytes = bench.hop_bytes(n, e, h, f)
    ref = None
    for variant in [int(v) for v in args.variants.split(",")]:
        for npc in ([int(v) for v in args.npc.split(",")] if variant == 3 else [0]):
            if npc:
                os.environ["GVQA_HOP_NPC"] = str(npc)
            else:
                os.environ.pop("GVQA_HOP_NPC", None)
            slab = (None, None)
            if variant == 5:
                si, sf = _cabi.build_hop_slabs(csr.as_dict(), a_edge, ag.view(1, b, h), 1, h, n)
                slab = (si, sf[0])
            fn = lambda: _cabi.gat_hop(x_l, x_l[:, h * f:h * f + 2 * h], a_edge, csr.as_dict(), h, f, out, lde=32, graph_bias=gb,
                                       a_graph=ag, h_prev=hprev, bias=bias, ep_scale=sc, ep_shift=sh,
                                       epilogue=_cabi.EPI_AFFINE_RELU, variant=variant, slab_idx=slab[0], slab_f=slab[1],
                                       **csr.hints())
            fn(); torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
            err = float((out - ref).abs().max())
            med, best = timed(fn, args.reps)
            print("%-7s variant %d npc %-2d  median %6.1f us (best %6.1f)  %7.2f MB  %5.0f GB/s  %4.1f %% of %.0f   max|d vs first| %.1e"
                  % (case, variant, npc, med, best, nbytes / 1e6, nbytes / med / 1e3, 100 * nbytes / med / 1e3 / PEAK, PEAK, err), flush=True)
            if args.trace:
                grid_cap = n // 4 + 64           # >= the largest grid any variant launches (npc >= 4)
                tr = torch.zeros(grid_cap * 8, dtype=torch.int64, device=dev)
                flush.sum(); torch.cuda.synchronize()
                _cabi.lib().gvqa_debug_set_hop_trace(tr.data_ptr())
                fn(); torch.cuda.synchronize()
                _cabi.lib().gvqa_debug_set_hop_trace(None)
                t = tr.cpu().view(grid_cap, 8).double()
                t = t[t[:, 0] > 0]
                t0 = t[:, 0].min()
                def q(x): return "min %.2f p10 %.2f med %.2f p90 %.2f max %.2f" % tuple(np.percentile(np.asarray(x), [0, 10, 50, 90, 100]))
                start = (t[:, 0] - t0) / 1e3
                fin = (t[:, 4:8] - t0) / 1e3
                fin = torch.where(t[:, 4:8] > 0, fin, torch.zeros_like(fin))
                print("    CTAs %d  start: %s" % (t.shape[0], q(start)))
                if variant == 5:
                    print("    slab + a_node window landed: %s" % q((t[:, 1] - t0) / 1e3))
                    print("    softmax done: %s" % q((t[:, 3] - t0) / 1e3))
                elif variant == 4:
                    fin = fin * 0
                    first = (t[:, 1] - t0) / 1e3
                    print("    first chunk ready: %s" % q(first[t[:, 1] > 0]))
                    chunks = t[:, 2]
                    print("    chunks per CTA: %s  (total %d)" % (q(chunks), int(chunks.sum())))
                    smid = t[:, 3].long()
                    per_sm = torch.zeros(int(smid.max()) + 1).index_add_(0, smid, chunks.float())
                    print("    chunks per SM: %s" % q(per_sm[per_sm > 0]))
                else:
                    print("    softmax done: %s" % q((t[:, 3] - t0) / 1e3))
                print("    CTA finish: %s ; kernel span %.2f us" % (q(fin.max(1).values), float(fin.max())))

"""Fused aggregate-project hop at a BASELINE shape: per-launch duration (CUDA events, L2 flushed between launches), the
replayed gat_seq step in both hop modes, and the small per-hop kernels.  python fused_hop_probe.py [graphs nodes edges feat]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphvqa_b200 import _cabi, gat_skip as eng                     # noqa: E402
from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology   # noqa: E402

DEV = "cuda:0"
graphs, nodes, edges, feat = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 30, 60, 512)))
heads, hops = 4, 5
torch.manual_seed(0)
model = eng.gat_seq(feat, feat, feat, 512, hops, dropout=0.1, gat_heads=heads).eval().to(DEV)
ei, batch, mx = synthetic_topology(graphs, nodes, edges, seed=1234)
n, e = batch.numel(), ei.size(1)
sets = []
for s in range(4):
    g = torch.Generator().manual_seed(s)
    sets.append([torch.randn(n, feat, generator=g).to(DEV), ei.to(DEV), torch.randn(e, feat, generator=g).to(DEV),
                 torch.randn(hops, graphs, 512, generator=g).to(DEV), batch.to(DEV)])
csr = GraphCSR.build(sets[0][1], sets[0][4], graphs, max_nodes_per_graph=mx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, reps=20):
    out = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) * 1e3)
    out.sort()
    return out[len(out) // 2], out[0]


with torch.no_grad():
    for mode in ("split", "fused"):
        model.hop_mode = mode
        outs = [model(*s, csr=csr) for s in sets]
        torch.cuda.synchronize()
        if mode == "split":
            ref_out = outs[0]
        else:
            print("max |fused - split| over the 5-hop output: %.3g  (max |out| %.3g)" % (float((outs[0] - ref_out).abs().max()), float(ref_out.abs().max())))
        gs = []
        for s in sets:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                model(*s, csr=csr)
            gs.append(gr)
        for _ in range(3):
            for gr in gs:
                gr.replay()
        torch.cuda.synchronize()
        K = 40
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            gs[i % 4].replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / K
        print("%s: %.4f ms per step  (%.0f k questions/s)" % (mode, ms, graphs / ms))
    # pieces of a fused hop
    pk = model.packed()
    window = _cabi.fused_window(mx)
    plan = csr.fused_plan(window)
    print("window", window, "tiles", int(plan[1]))
    x = sets[0][0]
    a_node = torch.randn(n, 2 * heads, device=DEV)
    a_edge = torch.randn(e, heads * hops, device=DEV)
    a_graph = torch.randn(graphs, heads, device=DEV)
    gb = torch.randn(graphs, feat, device=DEV)
    alpha = _cabi.gat_alpha(a_node, a_edge, csr.as_dict(), heads, a_graph=a_graph)
    out = torch.empty(n, feat, device=DEV)
    print("skinny a_node   us (median, min):", timed(lambda: _cabi.skinny_matmul(x, pk["v_node"][0], out=a_node)))
    print("alpha kernel    us (median, min):", timed(lambda: _cabi.gat_alpha(a_node, a_edge, csr.as_dict(), heads, a_graph=a_graph, out=alpha)))
    f = lambda: _cabi.gat_fused_hop(x, pk["w_fused"][0], plan, csr.as_dict(), alpha, heads, feat, out, window=window, skip=x,
                                    graph_bias=gb, bias=model.convs[0].bias, ep_scale=pk["scale"][0], ep_shift=pk["shift"][0],
                                    epilogue=2)
    med, mn = timed(f)
    flops = 3 * 2.0 * n * feat * heads * feat
    print("fused hop       us (median, min): (%.1f, %.1f)  -> %.0f TFLOP/s of fp16 tensor math (3 products)" % (med, mn, flops / med / 1e6))
    a_part = torch.empty(_cabi.fused_part_blocks(n, feat), n, 2 * heads, device=DEV)
    terms = _cabi.fused_logit_terms(csr.as_dict(), a_edge, a_graph[None].expand(hops, -1, -1).contiguous(), hops, heads, n)
    f2 = lambda: _cabi.gat_fused_hop(x, pk["w_fused"][0], plan, csr.as_dict(), alpha, heads, feat, out, window=window, skip=x,
                                     graph_bias=gb, bias=model.convs[0].bias, ep_scale=pk["scale"][0], ep_shift=pk["shift"][0],
                                     epilogue=2, v_next=pk["v_node"][1], a_part=a_part, logit_terms=terms[0], a_node=a_node)
    med, mn = timed(f2)
    print("fused hop, softmax inside + next-hop logits   us (median, min): (%.1f, %.1f)" % (med, mn))
    f = f2
    print("plan kernel     us (median, min):", timed(lambda: _cabi.fused_plan(csr.graph_ptr, n, graphs, window)))
    if os.environ.get("FUSED_TRACE"):
        import ctypes
        buf = torch.zeros(1100 * 8, dtype=torch.int64, device=DEV)
        _cabi.lib().gvqa_debug_set_fused_trace(ctypes.c_void_p(buf.data_ptr()))
        f(); torch.cuda.synchronize()
        _cabi.lib().gvqa_debug_set_fused_trace(None)
        t = buf.cpu().view(1100, 8)
        t0 = int(t[0, 0])
        print("stage: producer conv_full conv_agg conv_st mma_a0 mma_a1 - commit   (cycles from first producer stamp)")
        for i in range(0, 36):
            print(i, [int(v) - t0 if v else 0 for v in t[i].tolist()])
        for i in range(1024, 1026):
            print("item", i - 1024, [int(v) - t0 if v else 0 for v in t[i].tolist()])
        print("epilogue passes: start, acc staged, skip landed, stored, next skip issued, logits done")
        for i in range(1040, 1048):
            print("pass", i - 1040, [int(v) - t0 if v else 0 for v in t[i].tolist()][:6])
        spans = t[100:137].reshape(-1, 2)[:148]
        st, en = spans[:, 0], spans[:, 1]
        base = int(st[st > 0].min())
        busy = [(i, int(st[i]) - base, int(en[i]) - base) for i in range(148) if int(st[i]) > 0]
        durs = sorted(e - s_ for _, s_, e in busy)
        print("per-CTA span (ns): start %d..%d, end %d..%d, duration min/median/max %d/%d/%d" % (
            min(s_ for _, s_, _ in busy), max(s_ for _, s_, _ in busy), min(e for _, _, e in busy), max(e for _, _, e in busy),
            durs[0], durs[len(durs) // 2], durs[-1]))
        print("ends (us) of CTAs 0,2,..: " + " ".join("%.1f" % (e / 1e3) for i, _, e in busy if i % 2 == 0))
        print("prologue: start, lists staged, softmax done", [int(v) - t0 if v else 0 for v in t[1060].tolist()][:3])

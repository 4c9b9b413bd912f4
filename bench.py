#!/usr/bin/env python
"""Benchmark of the GraphVQA scene-graph message-passing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic GQA-shaped scene graphs
(BASELINE.json configs[1], "cfg2": 256 graphs/GPU x 30 nodes x 60 edges, F=512, D=512, 4 heads,
5-hop GAT-skip): destination-CSR build + edge-logit pre-pass + 5 x (node projection + fused hop).
Metric: questions/sec (one question = one scene graph).  Rank 0 prints ONE JSON line.

--impl reference times the CPU restatement of the reference's PyG dataflow (oracle/) on the host
cores: the reference itself cannot run here (torch_geometric/torch_scatter are not installable).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of one fused-hop launch at cfg2 from the committed ncu --set full
# capture (cold L2; the 15.7 MB output is still dirty in L2 when the kernel ends)
NCU_TRAFFIC_BYTES = 85176576          # 80.934 MB read + 4.242 MB written (gat_hop_slab_kernel<4,4,0>, first captured launch)
NCU_TRAFFIC_SOURCE = "profiles/r02/hop_slab_ncu_raw.csv"
# the same for one gat_fused_hop_kernel launch at cfg2 (ncu --set full, profiles/r02/fused_hop_ncu_raw.csv)
FUSED_NCU_TRAFFIC_BYTES = 21990144         # 21.938 MB read + 0.052 MB written (the output is still dirty in L2)
FUSED_NCU_TRAFFIC_SOURCE = "profiles/r02/fused_hop_ncu_raw.csv"
CFG2 = dict(name="cfg2", graphs=256, nodes=30, edges=60, feat=512, ins=512, heads=4, hops=5)
METRIC = "questions/sec (batched scene-graph inference, 5-hop GAT-skip stack)"
UNIT = "questions/s"
# the ONE workload string both arms print (config.workload); implementation details go into other config keys
WORKLOAD = ("cfg2: 256 synthetic GQA-shape scene graphs per GPU (30 nodes/60 edges), F=512, D=512, 4 heads, "
            "5-hop GAT-skip (gat_seq.forward)")
DTYPE = "f32 (projection on tcgen05 with 3xf16-split operands, fp32 accumulate; everything else fp32)"
SG_VOCAB = 2577          # rebuilt GQA scene-graph vocabulary


def hop_bytes(n, e, h, c):
    """Algorithmic bytes of one fused GAT hop (BASELINE.md section 3 / SURVEY.md section 8d)."""
    return 4 * (n * h * c + 2 * n * c + 2 * n * h + e * h) + 4 * (n + 1 + e) + 16 * c


def make_inputs(cfg, seed):
    from graphvqa_b200.graph_batch import synthetic_topology
    ei, batch, max_nodes = synthetic_topology(cfg["graphs"], cfg["nodes"], cfg["edges"], seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(batch.numel(), cfg["feat"], generator=g)
    ea = torch.randn(ei.size(1), cfg["feat"], generator=g)
    ins = torch.randn(cfg["hops"], cfg["graphs"], cfg["ins"], generator=g)
    return dict(x=x, edge_index=ei, edge_attr=ea, instr_vectors=ins, batch=batch, max_nodes=max_nodes,
                max_edges=synthetic_topology.last_max_edges)


def randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for bn in model.bns:
            bn.running_mean.normal_(0, 0.1, generator=g)
            bn.running_var.uniform_(0.5, 1.5, generator=g)


def model_kwargs(cfg):
    return dict(in_channels=cfg["feat"], out_channels=cfg["feat"], edge_attr_dim=cfg["feat"],
                ins_dim=cfg["ins"], num_ins=cfg["hops"], dropout=0.1, gat_heads=cfg["heads"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_baseline(cfg, steps, warmup, sample_graphs=None):
    """The reference's PyG dataflow restated in plain PyTorch (oracle/), timed on the host cores."""
    from oracle import graphvqa_oracle as orc
    sub = dict(cfg)
    if sample_graphs:
        sub["graphs"] = sample_graphs
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = orc.gat_seq(**model_kwargs(sub)).eval()
    randomise_bn(model, 7)
    inp = make_inputs(sub, seed=1234)
    args = (inp["x"], inp["edge_index"], inp["edge_attr"], inp["instr_vectors"], inp["batch"])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model(*args)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    mean = sum(times) / len(times)
    return dict(value=sub["graphs"] / mean, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample="%d graphs of %s per step (full 5-hop gat_seq, fp32, torch %d threads), mean of %d steps "
                       "after %d warm-up" % (sub["graphs"], cfg["name"], torch.get_num_threads(), len(times), warmup),
                ms_per_step=mean * 1e3)


def make_token_inputs(cfg, seed, sym_per_graph=4):
    """The same synthetic batch at the reference's own boundary: token ids instead of pre-encoded features
    (x [N,12], edge_attr [E,1] token ids, un-offset added_sym_edge) + what the text side hands over."""
    from graphvqa_b200.graph_batch import synthetic_topology
    ei, batch, max_nodes = synthetic_topology(cfg["graphs"], cfg["nodes"], cfg["edges"], seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    n, e, b = batch.numel(), ei.size(1), cfg["graphs"]
    x = torch.randint(4, SG_VOCAB, (n, 12), generator=g)
    x[torch.rand(n, 12, generator=g) < 0.6] = 1
    x[:, 0] = torch.randint(4, SG_VOCAB, (n,), generator=g)
    ea = torch.randint(4, SG_VOCAB, (e, 1), generator=g)
    sym = torch.randint(0, cfg["edges"], (b * sym_per_graph,), generator=g)          # graph-local, un-offset
    ins = torch.randn(cfg["hops"], b, cfg["ins"], generator=g)
    q0 = torch.randn(b, cfg["ins"], generator=g)
    return dict(x=x, edge_index=ei, edge_attr=ea, added_sym_edge=sym, batch=batch, instr_vectors=ins, q0=q0,
                max_nodes=max_nodes, max_edges=synthetic_topology.last_max_edges)


def per_graph_tensors(tok, cfg):
    """Split a token batch back into the per-graph tuples a dataset's __getitem__ yields (input of the collator)."""
    n_, e_, b = cfg["nodes"], cfg["edges"], cfg["graphs"]
    k = tok["added_sym_edge"].numel() // b
    return [(tok["x"][i * n_:(i + 1) * n_], tok["edge_index"][:, i * e_:(i + 1) * e_] - i * n_,
             tok["edge_attr"][i * e_:(i + 1) * e_], tok["added_sym_edge"][i * k:(i + 1) * k]) for i in range(b)]


def cpu_graph_side(cfg, steps, warmup):
    """The reference's graph side at its own boundary (scene-graph encoder -> gat_seq -> pooling -> logit_fc,
    pipeline_model_gat.py:751, 791-816), restated in oracle/, timed on the host cores."""
    import types
    from oracle import graphvqa_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = orc.GraphSide(SG_VOCAB, sg_emb_dim=cfg["feat"], question_hidden_dim=cfg["ins"]).eval()
    randomise_bn(model.gat_seq, 7)
    tok = make_token_inputs(cfg, seed=1234)
    g = types.SimpleNamespace(**{k: tok[k] for k in ("x", "edge_index", "edge_attr", "added_sym_edge", "batch")})
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model(g, tok["instr_vectors"], tok["q0"])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return dict(value=cfg["graphs"] / mean, unit=UNIT, ms_per_step=mean * 1e3, steps=len(times),
                what="token ids -> scene-graph encoder -> 5-hop gat_seq -> attention pooling -> logit_fc (oracle "
                     "GraphSide, F=%d), %d graphs per step" % (cfg["feat"], cfg["graphs"]))


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port: the reference's own modules need
    torch_geometric / torch_scatter, which cannot be installed here) on all host cores.  Honours --steps / --warmup
    as given; under torchrun rank 0 alone runs it, on the WHOLE job's graphs (weak scaling: 256 per GPU)."""
    if rank != 0:
        return
    cfg = dict(CFG2)
    cfg["graphs"] = CFG2["graphs"] * max(1, args.gpus)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base = cpu_baseline(cfg, steps=steps, warmup=warmup)
    side = cpu_graph_side(CFG2, steps=2, warmup=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "graphs_per_step": cfg["graphs"], "impl_note":
                   "CPU restatement of the reference's PyG dataflow (oracle/graphvqa_oracle.py); the reference's "
                   "own modules need torch_geometric/torch_scatter which are not installable here"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        # the engine arm's e2e crosses the boundary the reference's loop has (token ids in, answer logits out) and so
        # also runs the scene-graph encoder, the pooling and logit_fc; the same work on the host cores, for a
        # like-for-like e2e comparison (this line's value / e2e time gat_seq.forward only, i.e. LESS work)
        "graph_side": side,
        "gpu_launches": 0,
    }
    emit(line)


def _bracketed(fn, flush, reps=24):
    """Median CUDA-event bracket of one launch behind a CLEAN L2 flush (a 256 MB buffer is read, not written) with
    the stream kept busy, like profiles/microbench/kernel_roofline.py; includes the ~2.7 us an event pair reads."""
    ts = []
    for _ in range(reps):
        flush.sum(); torch.cuda._sleep(150000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[4:])
    return ts[len(ts) // 2]


def _back_to_back(fns, reps=20):
    """Average duration of one launch inside a stream of launches: `fns` are the same kernel on DIFFERENT buffer sets
    (together larger than L2), captured round-robin `reps` times each into ONE CUDA graph (no host launch cost, no
    per-launch event pair, no idle gaps) and replayed between two events -- how the kernel runs inside a replayed
    step, where its neighbours keep the GPU busy."""
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph, stream=side):
        for _ in range(reps):
            for fn in fns:
                fn()
    gph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gph.replay()
    gph.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps * len(fns))


def variant_rooflines(dev, peak, variant, tensor_peak=1675.8):
    """The other HBM-bound kernels of SURVEY.md section 8 at their BASELINE shapes (cfg3 GINE, cfg5 LCGN per GPU,
    graph LayerNorm, GCN) and the fused hop at the reference width F=300 and at cfg4 per GPU, measured in this
    process: algorithmic bytes / duration.  Two durations: `avg_launch_us` = event bracket
    around ONE launch behind a clean L2 flush (conservative: includes the ~2.7 us an event pair reads around nothing
    and a cold start, which dominate kernels of a few microseconds), `back_to_back_us` = the launch inside a stream of
    launches over rotating buffer sets larger than L2."""
    from graphvqa_b200 import _cabi
    from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology
    from graphvqa_b200.my_graph_layernorm import LayerNorm
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    g = torch.Generator().manual_seed(0)

    def rnd(*shape):
        return torch.randn(*shape, generator=g).to(dev)

    def graphs(b, n_, e_):
        ei, batch, mx = synthetic_topology(b, n_, e_, seed=1234)
        csr = GraphCSR.build(ei.to(dev), batch.to(dev), b, max_nodes_per_graph=mx,
                             max_in_edges_per_graph=synthetic_topology.last_max_edges)
        return ei.size(1), batch.numel(), batch.to(dev), csr
    out = {}

    def add(name, kernel, make, nbytes, shape):
        """make() -> a closure launching the kernel on a fresh buffer set"""
        sets = max(2, min(12, int(1.5 * (130 << 20) / max(nbytes, 1)) + 1))       # rotating sets > L2 (126 MB)
        fns = [make() for _ in range(sets)]
        us = _bracketed(fns[0], flush)
        b2b = _back_to_back(fns)
        gbs = nbytes / us / 1e3
        out[name] = {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "avg_launch_us": us, "algorithmic_bytes_per_launch": nbytes, "method": "event_bracket, clean L2 flush",
                     "back_to_back_us": b2b, "back_to_back_frac": nbytes / b2b / 1e3 / peak,
                     "back_to_back_method": "%d rotating buffer sets (%.0f MB > L2), 20 rounds captured in one CUDA graph, 2 replays" % (
                         sets, sets * nbytes / 1e6), "shape": shape}

    def hop(name, b, n_, e_, f, h=4):
        e, n, _, csr = graphs(b, n_, e_)

        def make():
            x_l, a_edge, hprev, o = rnd(n, h * f + 16), rnd(e, 32), rnd(n, f), torch.empty(n, f, device=dev)
            gb, ag, bias, sc, sh = rnd(b, f), rnd(b, h), rnd(f), rnd(f), rnd(f)
            si = sf = None
            if variant in (_cabi.VARIANT_AUTO, _cabi.VARIANT_SLAB):      # per-batch slabs of the one-round-trip prologue
                si, sf = _cabi.build_hop_slabs(csr.as_dict(), a_edge, ag.view(1, b, h), 1, h, n)
                sf = sf[0]
            return lambda: _cabi.gat_hop(x_l, x_l[:, h * f:h * f + 2 * h], a_edge, csr.as_dict(), h, f, o, lde=32,
                                         graph_bias=gb, a_graph=ag, h_prev=hprev, bias=bias, ep_scale=sc, ep_shift=sh,
                                         epilogue=_cabi.EPI_AFFINE_RELU, variant=variant, slab_idx=si, slab_f=sf,
                                         **csr.hints())
        add(name, "gvqa_gat_hop_f32", make, hop_bytes(n, e, h, f), "B=%d, %d nodes/%d edges, F=%d, H=%d" % (b, n_, e_, f, h))
    hop("gat_hop_cfg2", 256, 30, 60, 512)
    hop("gat_hop_f300", 256, 30, 60, 300)
    hop("gat_hop_cfg4_per_gpu", 128, 200, 800, 512)

    def fused_hop(name, b, n_, e_, f, h=4):
        """the default one-kernel hop (gvqa_gat_fused_hop_f32) at another shape: tensor-bound, against the bf16 peak"""
        e, n, _, csr = graphs(b, n_, e_)
        window = _cabi.fused_window(csr.max_nodes_per_graph)
        plan = csr.fused_plan(window)
        w = rnd(h * f, f) * (1.0 / f ** 0.5)
        pack = _cabi.fused_pack(w, h, f, f)
        flops = 3 * 2.0 * n * (h * f) * f

        def make():
            x, o = rnd(n, f), torch.empty(n, f, device=dev)
            a_node, a_edge, ag = rnd(n, 2 * h), rnd(e, h), rnd(1, b, h)
            gb, bias, sc, sh, vn = rnd(b, f), rnd(f), rnd(f), rnd(f), rnd(2 * h, f)
            terms = _cabi.fused_logit_terms(csr.as_dict(), a_edge, ag, 1, h, n)
            scratch = torch.empty(e, h, device=dev)
            a_part = torch.empty(_cabi.fused_part_blocks(n, f), n, 2 * h, device=dev)
            return lambda: _cabi.gat_fused_hop(x, pack, plan, csr.as_dict(), scratch, h, f, o, window=window, skip=x,
                                               graph_bias=gb, bias=bias, ep_scale=sc, ep_shift=sh,
                                               epilogue=_cabi.EPI_AFFINE_RELU, v_next=vn, a_part=a_part,
                                               logit_terms=terms[0], a_node=a_node)
        nbytes = 4 * (2 * n * f) + 4 * h * f * f
        sets = max(2, min(12, int(1.5 * (130 << 20) / max(nbytes, 1)) + 1))
        fns = [make() for _ in range(sets)]
        us = _bracketed(fns[0], flush)
        b2b = _back_to_back(fns)
        out[name] = {"kernel": "gvqa_gat_fused_hop_f32", "bound": "tensor", "achieved": flops / b2b / 1e6, "peak": tensor_peak,
                     "unit": "TFLOP/s", "frac": flops / b2b / 1e6 / tensor_peak, "avg_launch_us": b2b,
                     "method": "back to back: %d rotating buffer sets, 20 rounds captured in one CUDA graph, 2 replays" % sets,
                     "bracketed_us": us, "bracketed_frac": flops / us / 1e6 / tensor_peak,
                     "tensor_flops_per_launch": flops, "shape": "B=%d, %d nodes/%d edges, F=%d, H=%d" % (b, n_, e_, f, h)}
    fused_hop("gat_fused_hop_cfg2", 256, 30, 60, 512)
    fused_hop("gat_fused_hop_f300", 256, 30, 60, 300)
    fused_hop("gat_fused_hop_cfg4_per_gpu", 128, 200, 800, 512)
    e, n, batch, csr = graphs(256, 30, 60)
    ln = LayerNorm(512).to(dev).eval()

    def make_ln():
        x = rnd(n, 512)
        return lambda: ln(x, batch, num_graphs=256, csr=csr)
    with torch.no_grad():
        add("graph_layernorm_cfg2", "gvqa_graph_layernorm_f32", make_ln, 8 * n * 512 + 4 * n, "N=%d, F=512" % n)

    def make_gine():
        h_, ea, ins, z = rnd(n, 512), rnd(e, 512), rnd(256, 512), torch.empty(n, 1024, device=dev)
        return lambda: _cabi.gine_aggregate(h_, ea, ins, csr.as_dict(), 0.0, out=z)
    add("gine_cfg3", "gvqa_gine_aggregate_f32", make_gine, 4 * (n * 512 + e * 512 + 256 * 512 + n * 1024) + 4 * (n + 1 + e),
        "cfg3: B=256, 30/60, F=512, D=512")
    dinv = _cabi.gcn_degree(csr.as_dict(), n, dev)

    def make_gcn():
        xw, gt, bias, oc = rnd(n, 512), rnd(256, 512), rnd(512), torch.empty(n, 512, device=dev)
        return lambda: _cabi.gcn_aggregate(xw, gt, dinv, bias, csr.as_dict(), out=oc)
    add("gcn_cfg2", "gvqa_gcn_aggregate_f32", make_gcn, 4 * (2 * n * 512 + e) + 4 * (n + 1 + e), "B=256, 30/60, C=512")
    e5, n5, _, csr5 = graphs(128, 30, 60)

    def make_lcgn():
        proj, pc, cc, b5, o5 = rnd(n5, 1536), rnd(128, 512), rnd(128, 512), rnd(512), torch.empty(n5, 512, device=dev)
        return lambda: _cabi.lcgn_hop(proj[:, :512], proj[:, 512:1024], proj[:, 1024:], pc, cc, b5, csr5.as_dict(), 0.2, out=o5)
    add("lcgn_cfg5_per_gpu", "gvqa_lcgn_hop_f32", make_lcgn, 4 * (4 * n5 * 512 + 2 * 128 * 512) + 4 * (n5 + 1 + e5),
        "cfg5 per GPU: B=128, 30/60, C=512")
    return out


def collate_throughput(cfg, reps=10):
    """Loader side: per-graph tensors -> int32 wire batch with the destination-CSR built on
    the host, one core."""
    from graphvqa_b200.collate import WireCollator
    tok = make_token_inputs(cfg, seed=99)
    graphs = [tuple(t.to(torch.int32) for t in g) for g in per_graph_tensors(tok, cfg)]
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        c = WireCollator(depth=2)
        for _ in range(2):
            c(graphs)
        t0 = time.perf_counter()
        for _ in range(reps):
            w = c(graphs)
        dt = (time.perf_counter() - t0) / reps
    finally:
        torch.set_num_threads(threads)
    return {"graphs_per_s_per_core": cfg["graphs"] / dt, "ms_per_batch": dt * 1e3, "batch_graphs": cfg["graphs"],
            "what": "collate.WireCollator: concatenation + int32 narrowing + gvqa_build_csr_host into a pinned arena "
                    "(per-graph tensors in, as a dataset __getitem__ yields them)",
            "wire_bytes": sum(t.numel() * t.element_size() for t in (w.x, w.edge_index, w.edge_attr, w.edge_sign))
            + sum(t.numel() * t.element_size() for t in w.csr_host.values() if torch.is_tensor(t))}


def cfg4_sharded(args, dev, rank, world, steps):
    """BASELINE cfg4 on `world` GPUs: B=1024 graphs x 200 nodes / 800 edges, F=512, sharded by contiguous graph
    range (dist.shard_scene_graphs), per rank scene-graph encoder -> 5-hop gat_seq -> pooling -> logit_fc, then ONE
    all_gather_into_tensor of the [B/G, 1842] logits -- inside the timed region.  Rank 0 afterwards runs the whole
    batch alone: shard_parity_max_abs = max |gathered - single-GPU| (the sharded == single-GPU invariant)."""
    import torch.distributed as dist
    from graphvqa_b200 import dist as gdist
    from graphvqa_b200.graph_batch import SceneGraphBatch
    from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
    cfg = dict(name="cfg4", graphs=1024, nodes=200, edges=800, feat=512, ins=512, heads=4, hops=5)
    b = cfg["graphs"]
    tok = make_token_inputs(cfg, seed=4321)                       # the same batch on every rank (same seed)
    full = SceneGraphBatch(x=tok["x"], edge_index=tok["edge_index"], edge_attr=tok["edge_attr"], batch=tok["batch"],
                           added_sym_edge=tok["added_sym_edge"], num_graphs=b, max_nodes_per_graph=tok["max_nodes"],
                           max_in_edges_per_graph=tok["max_edges"])
    torch.manual_seed(0)
    model = PipelineModel(VocabSpec(text_vocab_size=64, sg_vocab_size=SG_VOCAB), sg_emb_dim=cfg["feat"]).eval()
    randomise_bn(model.gat_seq, 7)
    model = model.to(dev)
    model.gat_seq.kernel_variant = args.variant
    model.gat_seq.hop_mode = args.hop_mode
    model.strict_range = False
    lo, hi = gdist.graph_range(b, rank, world)
    shard = gdist.shard_scene_graphs(full, rank, world, num_graphs=b).to(device=dev)
    ins = tok["instr_vectors"][:, lo:hi].contiguous().to(dev)
    q_enc = tok["q0"][lo:hi].contiguous().to(dev).unsqueeze(0)

    def step():
        local = model.graph_side(shard, ins, q_enc, hi - lo)
        return gdist.all_gather_logits(local, b)
    with torch.no_grad():
        for _ in range(3):
            gathered = step()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            gathered = step()
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t) / steps
        model.gat_seq.check_overflow()
        parity = None
        if rank == 0:
            single = model.graph_side(full.to(device=dev), tok["instr_vectors"].to(dev), tok["q0"].to(dev).unsqueeze(0), b)
            parity = float((gathered - single).abs().max())
        dist.barrier()
    return {"workload": "cfg4: 1024 synthetic scene graphs (200 nodes/800 edges), F=512, sharded over %d GPUs by "
                        "contiguous graph range; token ids -> scene-graph encoder -> 5-hop gat_seq -> attention pooling "
                        "-> logit_fc per rank" % world,
            "value": b / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "graphs_per_gpu": hi - lo,
            "collective": "all_gather_into_tensor[B/G=%d,1842] fp32 (NCCL), inside the timed region" % (hi - lo),
            "added_sym_edge_entries": int(tok["added_sym_edge"].numel()),
            "shard_parity_max_abs": parity, "shard_parity_bar": 2e-5, "launch": "eager (no CUDA graph)"}


def run_engine(args, rank, local_rank, world):
    import torch.distributed as dist
    from graphvqa_b200 import _cabi
    from graphvqa_b200 import gat_skip as eng
    from graphvqa_b200.graph_batch import GraphCSR

    assert torch.cuda.is_available(), "bench.py (engine arm) needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = CFG2
    b, n, e = cfg["graphs"], cfg["graphs"] * cfg["nodes"], cfg["graphs"] * cfg["edges"]
    heads, c, hops = cfg["heads"], cfg["feat"], cfg["hops"]

    torch.manual_seed(0)
    model = eng.gat_seq(**model_kwargs(cfg)).eval()
    randomise_bn(model, 7)
    model = model.to(dev)
    model.kernel_variant = args.variant
    model.hop_mode = args.hop_mode
    fused = args.hop_mode == "fused" and args.variant == 0 and args.projection == "3xf16"
    if args.gemm_flags:
        _cabi.lib().gvqa_debug_set_gemm_flags(args.gemm_flags)
    model.projection = args.projection
    if args.l2_persist:
        granted = _cabi.l2_persist_limit(args.l2_persist << 20, dev)
        model.l2_persist = granted > 0

    # R input sets (> L2 in total: 4 x ~50 MB) rotated step to step so no step finds its inputs in L2
    R = 4
    host_sets = [make_inputs(cfg, seed=1234 + 10 * (rank * R + r)) for r in range(R)]
    keys = ("x", "edge_index", "edge_attr", "instr_vectors", "batch")
    pinned = [{k: s[k].pin_memory() for k in keys} for s in host_sets]
    dev_sets = [{k: s[k].to(dev) for k in keys} for s in host_sets]
    max_nodes = max(s["max_nodes"] for s in host_sets)
    max_edges = max(s["max_edges"] for s in host_sets)
    in_bytes = sum(host_sets[0][k].numel() * host_sets[0][k].element_size() for k in keys)

    hints = dict(max_nodes_per_graph=max_nodes, max_in_edges_per_graph=max_edges)

    def step(s):   # the CSR build is part of the step: gat_seq.forward runs it beside the pre-pass GEMMs
        return model(s["x"], s["edge_index"], s["edge_attr"], s["instr_vectors"], s["batch"], csr_hints=hints)

    def barrier():
        if world > 1:
            dist.barrier()

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            step(dev_sets[i % R])
        torch.cuda.synchronize()

        graphs = None
        if not args.no_graph:
            graphs, outs = [], []
            side = torch.cuda.Stream()
            for r in range(R):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side):
                    outs.append(step(dev_sets[r]))
                graphs.append(gph)
            for gph in graphs:
                gph.replay()
            torch.cuda.synchronize()

        def run_step(i):
            if graphs is not None:
                graphs[i % R].replay()
            else:
                step(dev_sets[i % R])

        # ---------------- timed region: K steps, device events, max over ranks -----------------
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); torch.cuda.synchronize()
        with ClockSampler(local_rank) as clocks:
            ev0.record()
            for i in range(args.steps):
                run_step(i)
            ev1.record()
            torch.cuda.synchronize()
            # keep the sampler alive for at least ~0.5 s of identical load so it gets samples
            t_end = time.time() + 0.6
            while time.time() < t_end:
                run_step(0)
            torch.cuda.synchronize()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t) / args.steps
        value = b * world / (ms_per_step / 1e3)

        # ---------------- fused-hop kernel time, live, eager launches with events ---------------
        model.hop_events, model.gemm_events = [], (None if fused else [])
        for i in range(args.steps):
            step(dev_sets[i % R])
        torch.cuda.synchronize()
        hop_ms = [a.elapsed_time(z) for a, z in model.hop_events]
        gemm_ms = [a.elapsed_time(z) for a, z in (model.gemm_events or [])]
        model.hop_events = model.gemm_events = None
        gemm_us = 1e3 * sum(gemm_ms) / len(gemm_ms) if gemm_ms else None
        hop_us = 1e3 * sum(hop_ms) / len(hop_ms)
        algo = hop_bytes(n, e, heads, c)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = algo / (hop_us * 1e-6) / 1e9

        # ---------------- fused-hop time inside the replayed graph (differential) ----------------
        # The bracket above puts two event records around every launch; an event pair around an EMPTY stream
        # position already reads ~2.7 us on this hardware (profiles/microbench/event_overhead.py).  Second
        # view: replay the step graphs with and without the hop launches (the GEMMs' time does not depend on
        # the data) and attribute the difference to the hops, launch gaps as they really are in the graph.
        hop_in_graph_us, hop_in_graph_spread = None, None
        if graphs is not None:
            model.skip_hop_launch = True
            for r in range(R):
                step(dev_sets[r])
            torch.cuda.synchronize()
            graphs_nohop = []
            for r in range(R):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side):
                    step(dev_sets[r])
                graphs_nohop.append(gph)
            model.skip_hop_launch = False

            def timed(gs):
                a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(args.steps):
                    gs[i % R].replay()
                z.record(); torch.cuda.synchronize()
                return a.elapsed_time(z) / args.steps
            for gs in (graphs, graphs_nohop):
                timed(gs)
            # 15 interleaved rounds (both graphs see the same clock / power state), median of the per-round differences
            diffs = sorted(timed(graphs) - timed(graphs_nohop) for _ in range(15))
            hop_in_graph_us = 1e3 * diffs[len(diffs) // 2] / hops
            hop_in_graph_spread = [1e3 * diffs[0] / hops, 1e3 * diffs[-1] / hops]
            del graphs_nohop

        # ---------------- third view: the event pairs as nodes of the replayed graph ------------------
        hop_graph_bracket_us = None
        if graphs is not None:
            try:
                model.hop_events = []
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side):
                    step(dev_sets[0])
                pairs, model.hop_events = model.hop_events, None
                acc = []
                for i in range(max(10, args.steps // 2)):
                    graphs[(i + 1) % R].replay()          # different inputs in between: keep L2 honest
                    gph.replay()
                    torch.cuda.synchronize()
                    acc += [a.elapsed_time(z) for a, z in pairs]
                hop_graph_bracket_us = 1e3 * sum(acc) / len(acc)
                del gph
            except Exception:                              # external events unsupported: keep the other two views
                model.hop_events = None

        # ---------------- end to end (1): the operator boundary, pre-encoded fp32 features over PCIe ---------
        # through the public host-buffer API (graphvqa_b200.host_api.GatSeqHostRunner): copies of
        # neighbouring batches overlap the kernels of the current one (3 streams, 3 device slots).
        from graphvqa_b200.host_api import GatSeqHostRunner, GraphSideHostRunner

        def pipelined(runner, submit_args, steps, read):
            for i in range(6):
                runner.submit(*submit_args(i))
            runner.drain()
            torch.cuda.synchronize(); barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(runner.s_h2d)
            checksum = 0.0
            for i in range(steps):
                t_id = runner.submit(*submit_args(i))
                if i >= 2:
                    checksum += read(runner.result(t_id - 2))   # consume results as they arrive (slot i-2 is reused
                                                                # by batch i+1, so it is read before then)
            runner.drain()
            e1.record(runner.s_d2h)
            torch.cuda.synchronize()
            t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            return b * world / (float(t2) / steps / 1e3)

        runner = GatSeqHostRunner(model, dev, depth=3, use_cuda_graph=not args.no_graph,
                                  max_nodes_per_graph=max_nodes, max_in_edges_per_graph=max_edges)
        e2e_operator = pipelined(runner, lambda i: (pinned[i % R],), args.steps, lambda t: float(t[0, 0]))
        del runner

        # ---------------- end to end (2, the headline): the reference's own boundary -----------------------
        # token ids in, answer logits out (mainExplain_gat.py:710-714, 758-765): a wire-format batch from the
        # loader-side collator (int32 tokens + host-built destination-CSR) plus the text side's instruction vectors
        # and question summary go up; scene-graph encoder -> 5-hop gat_seq -> attention pooling -> logit_fc run as
        # ONE CUDA graph per slot; short_answer_logits [B,1842] come back.
        from graphvqa_b200.collate import WireCollator
        from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
        torch.manual_seed(0)
        pm = PipelineModel(VocabSpec(text_vocab_size=64, sg_vocab_size=SG_VOCAB), sg_emb_dim=cfg["feat"]).eval()
        pm.gat_seq.load_state_dict(model.state_dict())           # the SAME hop stack as the resident measurement
        pm = pm.to(dev)
        pm.gat_seq.kernel_variant, pm.gat_seq.projection, pm.gat_seq.hop_mode = args.variant, args.projection, args.hop_mode
        collator = WireCollator(depth=R + 1)
        wires, text = [], []
        for r in range(R):
            tok = make_token_inputs(cfg, seed=1234 + 10 * (rank * R + r))
            wires.append(collator(per_graph_tensors(tok, cfg)))
            text.append((tok["instr_vectors"].pin_memory(), tok["q0"].pin_memory()))
        gs_runner = GraphSideHostRunner(pm, dev, depth=3, use_cuda_graph=not args.no_graph)
        gs_bytes = GraphSideHostRunner.bytes_per_batch(wires[0], *text[0])
        e2e_value = pipelined(gs_runner, lambda i: (wires[i % R], text[i % R][0], text[i % R][1]), args.steps,
                              lambda t: float(t[0, 0]))
        # the same graph side with everything resident (CUDA-graph replay of slot 0), for the compute share of e2e
        gph0 = gs_runner.slots[0].graph
        graph_side_ms = None
        if gph0 is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(args.steps):
                gph0.replay()
            ev1.record(); torch.cuda.synchronize()
            graph_side_ms = ev0.elapsed_time(ev1) / args.steps
        pm.gat_seq.check_overflow()
        del gs_runner

        # ---------------- the other kernels of the path, the loader side, the sharded large-graph config --------
        tensor_peak = (json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}).get("bf16_tflops", 2250.0)
        variants = variant_rooflines(dev, peak, args.variant, tensor_peak) if (rank == 0 and not args.skip_variants) else None
        sharded = cfg4_sharded(args, dev, rank, world, steps=max(3, min(args.steps, 10))) if (world > 1 and not args.skip_cfg4) else None
        collate = collate_throughput(cfg) if rank == 0 else None

    model.check_overflow()      # the fp16-split projection's range flag: the timed path computed valid results
    if rank != 0:
        return
    if hop_in_graph_us:
        primary_us, primary_method = hop_in_graph_us, "in_graph_differential"
    else:
        primary_us, primary_method = hop_us, "event_bracket"
    primary_achieved = algo / (primary_us * 1e-6) / 1e9
    peaks_all = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    base = cpu_baseline(cfg, steps=3, warmup=1) if (not args.skip_cpu and world == 1) else None   # N=1 only
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "step": ("gat_seq.forward = CSR build + row-tile plan + pre-pass (edge / instruction terms, hop 0 node "
                            "logits) + 5 x ONE tcgen05 kernel per hop (aggregate the input rows per head, project the "
                            "aggregate, hop epilogue)") if fused else
                           ("gat_seq.forward = CSR build + edge-logit pre-pass + 5 x (fp32-accurate tcgen05 projection + "
                            "fused hop)"),
                   "hop_mode": "fused" if fused else "split",
                   "hop_kernel_variant": args.variant,
                   "graphs_per_gpu": b, "parallelism": "graph-sharded x%d, no data-path collective" % world,
                   "l2_hygiene": "4 distinct input sets (~200 MB > 126 MB L2) rotated step to step",
                   "cuda_graph": graphs is not None},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": gs_bytes, "d2h_bytes_per_step": b * 1842 * 4,
                "boundary": "graphvqa_b200.host_api.GraphSideHostRunner: int32 token ids + host-built CSR + the text "
                            "side's instr_vectors / question summary in pinned host memory -> scene-graph encoder -> "
                            "5-hop gat_seq -> attention pooling -> logit_fc -> short_answer_logits[B,1842] in pinned "
                            "host memory; one CUDA graph per slot, 3 slots, 3 streams",
                "graph_side_resident_ms": graph_side_ms,
                "note": "does MORE work per question than `value` (encoder, pooling and answer head on top of the hop "
                        "stack); the reference arm's `graph_side` object times the same work on the host cores"},
        "e2e_operator": {"value": e2e_operator, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": n * c * 4,
                         "boundary": "graphvqa_b200.host_api.GatSeqHostRunner: the five gat_seq.forward tensors "
                                     "(pre-encoded fp32 features) up, node states down; PCIe-bound"},
        # per step: 5 CSR kernels + hops x (projection GEMM + fused hop); the edge-logit and instruction pre-pass
        # products ride in hop 0's projection launch (grouped) or cost two launches of their own
        "gpu_launches": args.steps * ((4 + 1 + 1 + 1 + hops) if fused else         # CSR (4 kernels), plan, pre-pass GEMM, logit terms, hops
                                      (4 + (0 if model.group_prepass and model.projection == "3xf16" else 2) + 2 * hops
                                       + (1 if model.use_slabs and args.variant in (0, 5) else 0))),
        "roofline": None,
    }
    hbm_roofline = {"bound": "hbm",
                     "kernel": ("gat_hop_slab_kernel (gvqa_gat_hop_f32; hops 1-4; hop 0 runs gat_hop_block_kernel while the "
                                "slabs are built)" if (model.use_slabs and args.variant in (0, 5)) else
                                "gvqa_gat_hop_f32, variant %d" % args.variant),
                     "achieved": primary_achieved, "peak": peak, "unit": "GB/s", "frac": primary_achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo, "avg_launch_us": primary_us, "method": primary_method,
                     "bracketed_us": hop_us, "bracketed_frac": achieved / peak, "launches_bracketed": len(hop_ms),
                     "in_graph_us": hop_in_graph_us, "in_graph_us_min_max": hop_in_graph_spread,
                     "in_graph_bracketed_us": hop_graph_bracket_us,
                     "note": "in_graph_differential: median over 15 interleaved rounds of (CUDA events around K replays of "
                             "the step graph minus K replays of the same graph captured without the 5 fused-hop "
                             "launches and without the per-batch slab build that serves them), per hop (the launch as it "
                             "runs in the timed region); in_graph_us_min_max = "
                             "smallest and largest round.  bracketed_*: eager launches with an event pair around every hop "
                             "launch of K steps; an event pair around an empty stream position already reads ~2.7 us "
                             "and around a 32-element kernel ~6 us (profiles/r01/event_overhead.txt)"}
    if fused:
        # the dominant kernel is tensor-bound: per launch 3 fp16 products (hi*hi, hi*lo, lo*hi) of
        # Z[N, H*F] x W'[C, H*F]^T on tcgen05 -- against the measured dense bf16/fp16 matmul peak
        tpeak = peaks_all.get("bf16_tflops", 2250.0)
        flops = 3 * 2.0 * n * (heads * cfg["feat"]) * c
        tf = flops / (primary_us * 1e-6) / 1e12
        line["roofline"] = {
            "bound": "tensor", "kernel": "gat_fused_hop_kernel<128,4> (gvqa_gat_fused_hop_f32): aggregate + project + hop epilogue",
            "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
            "traffic": FUSED_NCU_TRAFFIC_BYTES, "traffic_source": FUSED_NCU_TRAFFIC_SOURCE,
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks_all else "nominal dense bf16",
            "peak_sustained": peaks_all.get("bf16_tflops_sustained"),
            "frac_of_sustained": (tf / peaks_all["bf16_tflops_sustained"]) if peaks_all.get("bf16_tflops_sustained") else None,
            "tensor_flops_per_launch": flops, "useful_fp32_flops_per_launch": flops / 3,
            "avg_launch_us": primary_us, "method": primary_method,
            "bracketed_us": hop_us, "bracketed_frac": flops / (hop_us * 1e-6) / 1e12 / tpeak, "launches_bracketed": len(hop_ms),
            "in_graph_us": hop_in_graph_us, "in_graph_us_min_max": hop_in_graph_spread,
            "in_graph_bracketed_us": hop_graph_bracket_us,
            "hbm_bytes_per_launch_algorithmic": 4 * (2 * n * c) + 4 * heads * c * cfg["feat"],
            "note": "algorithmic work of one hop = the three split-precision tensor-core products (16.1 GFLOP of useful fp32 "
                    "work); HBM traffic is h in + h out + the packed weights, x_l[N, H*C] is never materialised.  "
                    "in_graph_differential: median over 15 interleaved rounds of (K replays of the step graph minus K "
                    "replays of the same graph captured without the 5 hop launches), per hop"}
    else:
        line["roofline"] = hbm_roofline
    if gemm_us:
        # the other big kernel of the step, tensor-bound: 3 fp16 products (hi*hi, hi*lo', lo'*hi) of [N x F] x
        # [H*C+16 x F]^T per launch against the measured dense bf16/fp16 matmul peak
        peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
        tpeak = peaks.get("bf16_tflops", 2250.0)
        nproj = heads * c + 16
        flops = 3 * 2.0 * n * nproj * cfg["feat"]
        tf = flops / (gemm_us * 1e-6) / 1e12
        line["roofline_projection"] = {
            "bound": "tensor", "kernel": "proj_gemm_3xf16_kernel (gvqa_proj_gemm_3xf16)" if args.projection == "3xf16"
            else "projection (%s)" % args.projection, "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
            "frac": tf / tpeak, "avg_launch_us": gemm_us, "method": "event_bracket", "launches_bracketed": len(gemm_ms),
            "tensor_flops_per_launch": flops, "useful_fp32_flops_per_launch": flops / 3,
            "peak_sustained": peaks.get("bf16_tflops_sustained"),
            "frac_of_sustained": (tf / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None,
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks else "nominal"}
    if base is not None:
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if variants is not None:
        line["roofline_variants"] = variants
    if collate is not None:
        line["collate"] = collate
    if sharded is not None:
        line["extra"] = {"cfg4_sharded": sharded}
    emit(line)


class _OneJsonLine:
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner from
    C code at communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the line goes to
    the saved descriptor at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.fd, (text.rstrip("\n") + "\n").encode())


_OUT = None


def emit(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global _OUT
    _OUT = _OneJsonLine()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-variants", action="store_true", help="skip the roofline_variants object")
    ap.add_argument("--skip-cfg4", action="store_true", help="N > 1: skip the sharded cfg4 run with the logits all-gather")
    ap.add_argument("--variant", type=int, default=0,
                    help="fused-hop kernel: 0 auto (slab), 1 gather, 2 staged, 3 block, 4 persistent warp-specialised, "
                         "5 block with the one-round-trip slab prologue")
    ap.add_argument("--hop-mode", default="fused", choices=["fused", "split"],
                    help="fused: one tensor-core kernel per hop that aggregates the input rows and projects the aggregate "
                         "(gvqa_gat_fused_hop_f32; needs --variant 0 and --projection 3xf16); split: projection GEMM + hop kernel")
    ap.add_argument("--projection", default="3xf16", choices=["3xf16", "3xtf32", "cublas"])
    ap.add_argument("--gemm-flags", type=int, default=0, help="debug flags of the projection GEMM (experiments)")
    ap.add_argument("--l2-persist", type=int, default=0, help="MiB of L2 set aside to keep x_l resident (0 = off)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner out of stdout: ONE JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_engine(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

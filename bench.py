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
NCU_TRAFFIC_BYTES = 87152384
NCU_TRAFFIC_SOURCE = "profiles/r01/hop_final_ncu_raw.csv"
CFG2 = dict(name="cfg2", graphs=256, nodes=30, edges=60, feat=512, ins=512, heads=4, hops=5)
METRIC = "questions/sec (batched scene-graph inference, 5-hop GAT-skip stack)"
UNIT = "questions/s"


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
                sample="%d graphs of %s per step (full 5-hop gat_seq, fp32, torch %d threads), mean of %d steps"
                       % (sub["graphs"], cfg["name"], torch.get_num_threads(), len(times)),
                ms_per_step=mean * 1e3)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = CFG2
    steps = max(1, min(args.steps, 5))
    base = cpu_baseline(cfg, steps=steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 256 synthetic GQA-shape scene graphs (30 nodes/60 edges), F=512, D=512, "
                               "4 heads, 5-hop GAT-skip (gat_seq.forward)", "impl_note":
                   "CPU restatement of the reference's PyG dataflow (oracle/graphvqa_oracle.py); the reference's "
                   "own modules need torch_geometric/torch_scatter which are not installable here"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
        model.hop_events, model.gemm_events = [], []
        for i in range(args.steps):
            step(dev_sets[i % R])
        torch.cuda.synchronize()
        hop_ms = [a.elapsed_time(z) for a, z in model.hop_events]
        gemm_ms = [a.elapsed_time(z) for a, z in model.gemm_events]
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

        # ---------------- end to end: pinned host buffers -> H2D -> hot path -> D2H --------------
        # through the public host-buffer API (graphvqa_b200.host_api.GatSeqHostRunner): copies of
        # neighbouring batches overlap the kernels of the current one (3 streams, 3 device slots).
        from graphvqa_b200.host_api import GatSeqHostRunner
        runner = GatSeqHostRunner(model, dev, depth=3, use_cuda_graph=not args.no_graph,
                                  max_nodes_per_graph=max_nodes, max_in_edges_per_graph=max_edges)
        for i in range(6):
            runner.submit(pinned[i % R])
        runner.drain()
        torch.cuda.synchronize(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(runner.s_h2d)
        checksum = 0.0
        for i in range(args.steps):
            t_id = runner.submit(pinned[i % R])
            if i >= 2:
                checksum += float(runner.result(t_id - 2)[0, 0])    # consume results as they arrive (slot i-2 is
                                                                    # reused by batch i+1, so it is read before then)
        runner.drain()
        e1.record(runner.s_d2h)
        torch.cuda.synchronize()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_value = b * world / (float(t2) / args.steps / 1e3)

    model.check_overflow()      # the fp16-split projection's range flag: the timed path computed valid results
    if rank != 0:
        return
    if hop_in_graph_us:
        primary_us, primary_method = hop_in_graph_us, "in_graph_differential"
    else:
        primary_us, primary_method = hop_us, "event_bracket"
    primary_achieved = algo / (primary_us * 1e-6) / 1e9
    base = cpu_baseline(cfg, steps=3, warmup=1) if (not args.skip_cpu and world == 1) else None   # N=1 only
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: %d synthetic GQA-shape scene graphs per GPU (30 nodes/60 edges), F=512, "
                               "D=512, 4 heads, 5-hop GAT-skip (gat_seq.forward: CSR build + edge-logit pre-pass + "
                               "5 x (fp32-accurate tcgen05 projection + fused hop))" % b,
                   "graphs_per_gpu": b, "parallelism": "graph-sharded x%d, no data-path collective" % world,
                   "l2_hygiene": "4 distinct input sets (~200 MB > 126 MB L2) rotated step to step",
                   "cuda_graph": graphs is not None},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": n * c * 4},
        # per step: 5 CSR kernels + hops x (projection GEMM + fused hop); the edge-logit and instruction pre-pass
        # products ride in hop 0's projection launch (grouped) or cost two launches of their own
        "gpu_launches": args.steps * (5 + (0 if model.group_prepass and model.projection == "3xf16" else 2) + 2 * hops),
        "roofline": {"bound": "hbm", "kernel": "gat_hop_block_kernel (gvqa_gat_hop_f32)",
                     "achieved": primary_achieved, "peak": peak, "unit": "GB/s", "frac": primary_achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo, "avg_launch_us": primary_us, "method": primary_method,
                     "bracketed_us": hop_us, "bracketed_frac": achieved / peak, "launches_bracketed": len(hop_ms),
                     "in_graph_us": hop_in_graph_us, "in_graph_us_min_max": hop_in_graph_spread,
                     "in_graph_bracketed_us": hop_graph_bracket_us,
                     "note": "in_graph_differential: median over 15 interleaved rounds of (CUDA events around K replays of "
                             "the step graph minus K replays of the same graph captured without the 5 fused-hop "
                             "launches), per hop (the launch as it runs in the timed region); in_graph_us_min_max = "
                             "smallest and largest round.  bracketed_*: eager launches with an event pair around every hop "
                             "launch of K steps; an event pair around an empty stream position already reads ~2.7 us "
                             "and around a 32-element kernel ~6 us (profiles/r01/event_overhead.txt)"},
    }
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
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--variant", type=int, default=0, help="fused-hop kernel: 0 auto, 1 gather, 2 staged, 3 block")
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

"""Pre-pass launches of gat_seq at cfg2, timed back to back behind a busy stream (us per launch)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphvqa_b200 import _cabi, gat_skip as eng
from graphvqa_b200.graph_batch import GraphCSR
dev = torch.device('cuda:0')
cfg = bench.CFG2
torch.manual_seed(0)
model = eng.gat_seq(**bench.model_kwargs(cfg)).eval().to(dev)
inp = bench.make_inputs(cfg, 1234)
d = {k: inp[k].to(dev) for k in ("x", "edge_index", "edge_attr", "instr_vectors", "batch")}
pk = model.packed()
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
def t(fn, label, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda._sleep(100000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[3:])
    print("%-46s median %.1f us (event bracket, L2 flushed)" % (label, ts[len(ts) // 2]))
ins = d["instr_vectors"].contiguous()
t(lambda: GraphCSR.build(d["edge_index"], d["batch"], cfg["graphs"]), "gvqa_build_csr (5 kernels + 2 memsets)")
t(lambda: _cabi.proj_gemm_3xtf32(d["edge_attr"], *pk["edge_split"]), "edge logits: GEMM [15360x512]x[32x512]^T")
t(lambda: _cabi.skinny_matmul(d["edge_attr"], pk["v_edge"]), "edge logits: skinny kernel (old path)")
t(lambda: _cabi.proj_gemm_3xtf32(ins.view(-1, 512), *pk["ins_split"]), "instruction terms: GEMM [1280x512]x[2640x512]^T")

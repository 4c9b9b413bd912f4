"""Fused GAT hop at cfg2: cold (L2 flushed) vs warm (same inputs relaunched) vs right after the projection GEMM."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphvqa_b200 import _cabi, gat_skip as eng
from graphvqa_b200.graph_batch import GraphCSR
dev = torch.device('cuda:0')
cfg = bench.CFG2
torch.manual_seed(0)
model = eng.gat_seq(**bench.model_kwargs(cfg)).eval(); bench.randomise_bn(model, 7); model = model.to(dev)
inp = bench.make_inputs(cfg, 1234)
d = {k: inp[k].to(dev) for k in ("x", "edge_index", "edge_attr", "instr_vectors", "batch")}
b = cfg["graphs"]
csr = GraphCSR.build(d["edge_index"], d["batch"], b, max_nodes_per_graph=inp["max_nodes"], max_in_edges_per_graph=inp["max_edges"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
model.kernel_variant = variant
def run(mode, reps=30):
    ts = []
    for _ in range(reps):
        model.hop_events = []
        if mode == "cold":
            # run the stack but flush L2 before each hop: patch by timing only hop 0 after a flush
            flush.zero_()
        with torch.no_grad():
            model(d["x"], d["edge_index"], d["edge_attr"], d["instr_vectors"], d["batch"], csr=csr)
        torch.cuda.synchronize()
        ts.append([a.elapsed_time(z) * 1e3 for a, z in model.hop_events])
    model.hop_events = None
    t = torch.tensor(ts)[5:]
    return t.mean(0).tolist()
print("variant", variant, "per-hop us in situ (hop 0..4):", ["%.1f" % v for v in run("insitu")])
# isolated relaunches of one hop on fixed buffers: warm L2
n, e, h, c = b * cfg["nodes"], b * cfg["edges"], cfg["heads"], cfg["feat"]
x_l = torch.randn(n, h * c + 16, device=dev); a_edge = torch.randn(e, 5 * h, device=dev)
hprev = torch.randn(n, c, device=dev); out = torch.empty(n, c, device=dev)
gb = torch.randn(b, c, device=dev); ag = torch.randn(b, h, device=dev)
bias = torch.randn(c, device=dev); sc = torch.rand(c, device=dev); sh = torch.randn(c, device=dev)
def hop():
    _cabi.gat_hop(x_l, x_l[:, h * c:h * c + 2 * h], a_edge, csr.as_dict(), h, c, out, lde=a_edge.stride(0), graph_bias=gb,
                  a_graph=ag, h_prev=hprev, bias=bias, ep_scale=sc, ep_shift=sh, epilogue=_cabi.EPI_AFFINE_RELU,
                  variant=variant, **csr.hints())
for mode in ("warm", "cold"):
    ts = []
    for i in range(40):
        if mode == "cold": flush.zero_()
        torch.cuda._sleep(200000)      # let the CPU run ahead so the launches below are queued, not latency-bound
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); hop(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[5:])
    print("isolated %s: median %.1f us  min %.1f us" % (mode, ts[len(ts) // 2], ts[0]))

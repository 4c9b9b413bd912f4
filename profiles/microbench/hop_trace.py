"""Timeline of the fused GAT hop at cfg2 (in situ, hop 2 of the stack): per-CTA globaltimer stamps."""
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
tr = torch.zeros(600 * 8, dtype=torch.int64, device=dev)
def run():
    with torch.no_grad():
        return model(d["x"], d["edge_index"], d["edge_attr"], d["instr_vectors"], d["batch"], csr=csr)
for _ in range(5): run()
torch.cuda.synchronize()
# trace only the LAST hop launch of one forward: every hop overwrites the buffer, the last one stays
_cabi.lib().gvqa_debug_set_hop_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
_cabi.lib().gvqa_debug_set_hop_trace(None)
t = tr.cpu().view(600, 8).double()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
t = (t - t0) / 1e3
import numpy as np
def q(x): return "min %.2f  p10 %.2f  med %.2f  p90 %.2f  max %.2f" % tuple(np.percentile(x.numpy(), [0, 10, 50, 90, 100]))
print("CTAs traced:", t.shape[0], " (us since the first CTA started)")
print("CTA start            ", q(t[:, 0]))
print("indices loaded       ", q(t[:, 1]))
print("logit terms loaded   ", q(t[:, 2]))
print("softmax done         ", q(t[:, 3]))
fin = t[:, 4:8]
print("warp finish (all)    ", q(fin.flatten()))
print("CTA finish (last warp)", q(fin.max(1).values))
print("per-CTA: prologue (start->softmax) ", q(t[:, 3] - t[:, 0]), " stream phase ", q(fin.max(1).values - t[:, 3]))
print("kernel span (first start -> last finish): %.2f us" % float(fin.max()))

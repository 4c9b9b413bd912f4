"""Accuracy of the two hop paths at cfg2 against a float64 evaluation of the oracle (the reference's dataflow restated):
max / rms absolute error of every hop's output."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphvqa_b200 import gat_skip as eng                             # noqa: E402
from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology   # noqa: E402
from oracle import graphvqa_oracle as orc                            # noqa: E402

DEV = "cuda:0"
cfg = dict(in_channels=512, out_channels=512, edge_attr_dim=512, ins_dim=512, num_ins=5, gat_heads=4)
torch.manual_seed(81)
o = orc.gat_seq(**cfg).eval()
g = torch.Generator().manual_seed(82)
for bn in o.bns:
    bn.running_mean.normal_(0, 0.1, generator=g); bn.running_var.uniform_(0.5, 1.5, generator=g)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.normal_(0, 0.1, generator=g)
e = eng.gat_seq(**cfg).eval()
e.load_state_dict(o.state_dict())
e = e.to(DEV)
ei, batch, mx = synthetic_topology(256, 30, 60, seed=1234)
x = torch.randn(batch.numel(), 512, generator=g)
ea = torch.randn(ei.size(1), 512, generator=g)
ins = torch.randn(5, 256, 512, generator=g)
o64 = orc.gat_seq(**cfg).eval().double()
o64.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in o.state_dict().items()})
with torch.no_grad():
    _, want = o64(x.double(), ei, ea.double(), ins.double(), batch, return_hops=True)
    _, ref32 = o(x, ei, ea, ins, batch, return_hops=True)
    dargs = [t.to(DEV) for t in (x, ei, ea, ins, batch)]
    csr = GraphCSR.build(dargs[1], dargs[4], 256, read_hints=True)
    rows = {}
    for mode in ("fused", "split"):
        e.hop_mode = mode
        _, hops = e(*dargs, csr=csr, return_hops=True)
        rows[mode] = [(h.cpu().double() - w) for h, w in zip(hops, want)]
    rows["oracle fp32 (CPU, PyTorch)"] = [(h.double() - w) for h, w in zip(ref32, want)]
print("cfg2 (256 graphs x 30 nodes / 60 edges, F = 512, 4 heads): error against the float64 oracle, per hop output")
print("max |h| per hop: " + " ".join("%.2f" % float(w.abs().max()) for w in want))
for name, errs in rows.items():
    print("%-28s max abs: %s" % (name, " ".join("%.2e" % float(d.abs().max()) for d in errs)))
    print("%-28s rms    : %s" % ("", " ".join("%.2e" % float(d.pow(2).mean().sqrt()) for d in errs)))
# signed view: does the truncating accumulate shrink the outputs by a consistent factor?
for mode in ("fused", "split"):
    outs = []
    for d, w in zip(rows[mode], want):
        big = w.abs() > 0.25 * w.abs().max()
        rel = (d[big] / w[big])
        outs.append("%.2e (+-%.1e)" % (float(rel.mean()), float(rel.std())))
    print("%-8s mean (std) of (got - want) / want over entries with |want| > max/4: %s" % (mode, " ".join(outs)))

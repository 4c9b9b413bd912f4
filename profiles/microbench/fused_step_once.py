"""Two eager gat_seq.forward calls in hop_mode "fused" at cfg2 (for ncu launch lists / --set full captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphvqa_b200 import gat_skip as eng                             # noqa: E402
from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology   # noqa: E402

DEV = "cuda:0"
graphs, nodes, edges, feat = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 30, 60, 512)))
torch.manual_seed(0)
model = eng.gat_seq(feat, feat, feat, 512, 5, dropout=0.1, gat_heads=4).eval().to(DEV)
model.hop_mode = os.environ.get("GVQA_HOP_MODE", "fused")
ei, batch, mx = synthetic_topology(graphs, nodes, edges, seed=1234)
g = torch.Generator().manual_seed(1)
args = [torch.randn(batch.numel(), feat, generator=g).to(DEV), ei.to(DEV), torch.randn(ei.size(1), feat, generator=g).to(DEV),
        torch.randn(5, graphs, 512, generator=g).to(DEV), batch.to(DEV)]
with torch.no_grad():
    for _ in range(2):
        out = model(*args, csr_hints=dict(max_nodes_per_graph=mx))
torch.cuda.synchronize()
print(float(out.abs().max()))

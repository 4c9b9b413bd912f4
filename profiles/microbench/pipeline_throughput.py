"""Whole PipelineModel (reference dims: F=300, 5 hops, 3-layer Transformer text stack) on synthetic GQA-shaped
inputs, B graphs of 30 nodes / 60 edges, 12-token questions: questions/s of `answer_logits` (inference path:
encoder + question encoder + coarse instruction decoder + gat_seq + pooling + answer head), eager launches."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200.graph_batch import SceneGraphBatch, synthetic_topology
from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = PipelineModel(VocabSpec(text_vocab_size=3657, sg_vocab_size=2577)).eval().to(dev)
for B in (256, 1024):
    ei, batch, mx = synthetic_topology(B, 30, 60, seed=3)
    g = torch.Generator().manual_seed(1)
    n, e = batch.numel(), ei.size(1)
    graphs = SceneGraphBatch(x=torch.randint(4, 2577, (n, 12), generator=g), edge_index=ei,
                             edge_attr=torch.randint(4, 2577, (e, 1), generator=g), batch=batch,
                             added_sym_edge=torch.zeros(0, dtype=torch.int64), num_graphs=B,
                             max_nodes_per_graph=mx).to(device=dev)
    q = torch.randint(4, 3657, (12, B), generator=g).to(dev)
    with torch.no_grad():
        for _ in range(5): out = m.answer_logits(q, graphs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 20
        for _ in range(K): out = m.answer_logits(q, graphs)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    m.gat_seq.check_overflow()
    print("PipelineModel.answer_logits B=%d: %.3f ms per batch = %.0f questions/s (eager, logits %s finite=%s)" % (
        B, ms, B / ms * 1e3, tuple(out.shape), bool(torch.isfinite(out).all())))

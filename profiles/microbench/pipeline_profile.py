import sys, torch
sys.path.insert(0, '/root/repo')
from graphvqa_b200.graph_batch import SceneGraphBatch, synthetic_topology
from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = PipelineModel(VocabSpec(text_vocab_size=3657, sg_vocab_size=2577)).eval().to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ei, batch, mx = synthetic_topology(B, 30, 60, seed=3)
g = torch.Generator().manual_seed(1)
n, e = batch.numel(), ei.size(1)
graphs = SceneGraphBatch(x=torch.randint(4, 2577, (n, 12), generator=g), edge_index=ei,
                         edge_attr=torch.randint(4, 2577, (e, 1), generator=g), batch=batch,
                         added_sym_edge=torch.zeros(0, dtype=torch.int64), num_graphs=B, max_nodes_per_graph=mx).to(device=dev)
q = torch.randint(4, 3657, (12, B), generator=g).to(dev)
with torch.no_grad():
    for _ in range(3): m.answer_logits(q, graphs)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5): m.answer_logits(q, graphs)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))

"""Where the graph side's time goes (cfg2 at the token-id boundary: scene-graph encoder -> gat_seq -> pooling ->
logit_fc): torch profiler kernel table of eager steps + CUDA-graph replay time."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphvqa_b200.collate import WireCollator
from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
from graphvqa_b200.host_api import GraphSideHostRunner
dev = torch.device('cuda:0')
cfg = bench.CFG2
torch.manual_seed(0)
pm = PipelineModel(VocabSpec(text_vocab_size=64, sg_vocab_size=bench.SG_VOCAB), sg_emb_dim=cfg["feat"]).eval()
bench.randomise_bn(pm.gat_seq, 7)
pm = pm.to(dev)
tok = bench.make_token_inputs(cfg, seed=1234)
wire = WireCollator(depth=2)(bench.per_graph_tensors(tok, cfg))
ins, q0 = tok["instr_vectors"].pin_memory(), tok["q0"].pin_memory()
runner = GraphSideHostRunner(pm, dev, depth=2, use_cuda_graph=True)
for _ in range(6):
    runner(wire, ins, q0)
g = wire.to(device=dev)
insd, q0d = ins.to(dev), q0.to(dev).unsqueeze(0)
pm.strict_range = False
with torch.no_grad():
    for _ in range(3):
        pm.graph_side(g, insd, q0d, cfg["graphs"])
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            pm.graph_side(g, insd, q0d, cfg["graphs"])
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
gph = runner.slots[0].graph
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    gph.replay()
e1.record(); torch.cuda.synchronize()
print("graph-side CUDA-graph replay: %.4f ms per 256-graph batch" % (e0.elapsed_time(e1) / 50))

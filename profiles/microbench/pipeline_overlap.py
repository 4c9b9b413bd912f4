"""whole-model questions/s of PipelineModel (reference dims F=300, D=512, B=256 scene graphs of
30 nodes / 60 edges, questions of 12 tokens) with the text side on a second CUDA stream beside the scene-graph encoder
and CSR build (overlap_text=True, default) vs everything on one stream; answer_logits (coarse decoder only) and the
reference-shaped forward(SAMPLE_FLAG=True) that validate() calls (mainExplain_gat.py:758-765)."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from graphvqa_b200.graph_batch import SceneGraphBatch
from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
dev = torch.device('cuda:0')
cfg = dict(bench.CFG2, feat=300)
torch.manual_seed(0)
m = PipelineModel(VocabSpec(text_vocab_size=3657, sg_vocab_size=bench.SG_VOCAB)).eval()
bench.randomise_bn(m.gat_seq, 7)
m = m.to(dev)
tok = bench.make_token_inputs(cfg, seed=1234)
g = SceneGraphBatch(x=tok["x"], edge_index=tok["edge_index"], edge_attr=tok["edge_attr"], batch=tok["batch"],
                    added_sym_edge=tok["added_sym_edge"], num_graphs=cfg["graphs"], max_nodes_per_graph=tok["max_nodes"],
                    max_in_edges_per_graph=tok["max_edges"]).to(device=dev)
q = torch.randint(4, 3657, (12, cfg["graphs"]), generator=torch.Generator().manual_seed(1)).to(dev)

def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) / reps * 1e3

with torch.no_grad():
    for strict in (True, False):
        m.strict_range = strict
        for overlap in (False, True):
            m.overlap_text = overlap
            ms, wall = timed(lambda: m.answer_logits(q, g), 30)
            print("answer_logits   strict_range=%-5s overlap_text=%-5s  %.3f ms device, %.3f ms wall per batch -> %.0f questions/s"
                  % (strict, overlap, ms, wall, cfg["graphs"] / wall * 1e3))
    m.strict_range = True
    for overlap in (False, True):
        m.overlap_text = overlap
        ms, wall = timed(lambda: m(q, g, None, None, SAMPLE_FLAG=True), 3)
        print("forward(SAMPLE) strict_range=True  overlap_text=%-5s  %.3f ms device, %.3f ms wall per batch -> %.0f questions/s"
              % (overlap, ms, wall, cfg["graphs"] / wall * 1e3))

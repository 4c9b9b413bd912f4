"""CPU: the oracle restatements of the steps either side of the hop stack --
GroundTruth_SceneGraph_Encoder / MetaLayer, MyConditionalGlobalAttention and the whole graph side -- pinned to
golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py graph_side), and, when the reference
tree is present (build container), compared live with the reference's own classes."""
import types

import pytest
import torch

from oracle import graphvqa_oracle as orc
from oracle import run_reference as rr
from oracle.golden_utils import deterministic_fill

SG_VOCAB = 2577


def _ns(d):
    return types.SimpleNamespace(**{k: v for k, v in d.items() if torch.is_tensor(v)})


def _graph_side(seed):
    torch.manual_seed(0)
    return deterministic_fill(orc.GraphSide(SG_VOCAB).eval(), seed)      # keyed by state_dict name: same values
                                                                         # as the reference PipelineModel's fill


def test_graph_side_matches_reference_golden(golden):
    base, fx = golden("pipeline_gat"), golden("graph_side_gat")
    m = _graph_side(fx["fill_seed"])
    g = _ns(base)
    with torch.no_grad():
        parts = m(g, fx["instr_vectors"], fx["q0"], return_parts=True)
    # same dense torch ops in the same order as the reference over the shim: bit-exact
    assert torch.equal(parts["x_encoded"], base["x_encoded"])
    assert torch.equal(parts["edge_attr_encoded"], base["edge_attr_encoded"])
    assert torch.equal(parts["x_executed"], fx["x_executed"])
    assert torch.equal(parts["pooled"], fx["pooled"])
    assert torch.equal(parts["short_answer_logits"], base["short_answer_logits"])


def test_encoder_edge_cases_match_reference_golden(golden):
    """single-node graphs, nodes without in-edges (scatter_mean clamp), un-offset added_sym_edge of 7 graphs."""
    fx = golden("encoder_pool_edge")
    enc = _graph_side(fx["fill_seed"]).scene_graph_encoder
    e = fx["enc"]
    deg = torch.zeros(e["batch"].numel()).index_add_(0, e["edge_index"][1], torch.ones(e["edge_index"].size(1)))
    assert int((deg == 0).sum()) >= 7                      # the fixture really contains in-degree-0 nodes
    assert int(e["added_sym_edge"].max()) < e["edge_index"].size(1) and e["added_sym_edge"].numel() > 8
    with torch.no_grad():
        x_enc, e_enc, _ = enc(_ns(e))
    assert torch.equal(x_enc, e["x_encoded"])
    assert torch.equal(e_enc, e["edge_attr_encoded"])


def test_pooling_edge_cases_match_reference_golden(golden):
    """an empty graph in the middle of the batch and `size` beyond batch.max()+1: their rows are zero."""
    p = golden("encoder_pool_edge")["pool"]
    pool = orc.MyConditionalGlobalAttention(20, 16).eval()
    pool.load_state_dict(p["state"])
    with torch.no_grad():
        out = pool(p["x"], p["u"], p["batch"], size=p["size"])
        out_default = pool(p["x"], p["u"][:5], p["batch"])
    assert torch.equal(out, p["out"]) and torch.equal(out_default, p["out_default_size"])
    assert float(out[2].abs().max()) == 0.0 and float(out[5].abs().max()) == 0.0
    assert out_default.shape == (5, 16)


def test_interleaved_layernorm_option_is_layernorm_of_the_hop():
    """gat_seq(..., interleaved_ln=ln): hop i < last is LayerNorm(conv + skip) (no BatchNorm, no ReLU)."""
    torch.manual_seed(3)
    m = orc.gat_seq(16, 16, 16, 8, 3, gat_heads=2).eval()
    ln = orc.LayerNorm(16)
    with torch.no_grad():
        ln.weight.fill_(1.2); ln.bias.fill_(-0.1)
    ei = torch.tensor([[0, 1, 2, 3, 4, 1, 0], [0, 1, 2, 3, 4, 0, 1]])
    batch = torch.tensor([0, 0, 1, 1, 1])
    x, ea, ins = torch.randn(5, 16), torch.randn(7, 16), torch.randn(3, 2, 8)
    with torch.no_grad():
        _, hops = m(x, ei, ea, ins, batch, return_hops=True, interleaved_ln=ln)
        x_cat = torch.cat((x, ins[0][batch]), -1)
        e_cat = torch.cat((ea, ins[0][batch[ei[0]]]), -1)
        want0 = ln(m.convs[0](x_cat, ei, e_cat) + x, batch)
    assert torch.allclose(hops[0], want0, atol=1e-6)


@pytest.mark.skipif(not rr.available(), reason="reference tree not present (GPU box)")
def test_oracle_equals_reference_classes_live():
    mod = rr.load("pipeline_model_gat")
    g = torch.Generator().manual_seed(5)
    # MetaLayer at a non-reference width
    torch.manual_seed(1)
    ref_layer = mod.get_gt_scene_graph_encoding_layer(24, 24).eval()
    mine = orc.MetaLayer(24, 24).eval()
    mine.load_state_dict(ref_layer.state_dict())
    n, e = 13, 40
    x, ea = torch.randn(n, 24, generator=g), torch.randn(e, 24, generator=g)
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(2, n, (e,), generator=g)])
    batch = torch.zeros(n, dtype=torch.long)
    with torch.no_grad():
        rx, re, _ = ref_layer(x, ei, ea, None, batch)
        ox, oe, _ = mine(x, ei, ea, None, batch)
    assert torch.equal(rx, ox) and torch.equal(re, oe)
    # pooling
    torch.manual_seed(2)
    ref_pool = mod.MyConditionalGlobalAttention(24, 32).eval()
    mine = orc.MyConditionalGlobalAttention(24, 32).eval()
    mine.load_state_dict(ref_pool.state_dict())
    pb = torch.tensor([0, 0, 0, 1, 2, 2, 2, 2, 4, 4, 4, 4, 4])
    u = torch.randn(5, 32, generator=g)
    with torch.no_grad():
        assert torch.equal(ref_pool(x, u, pb), mine(x, u, pb))
        assert torch.equal(ref_pool(x, u, pb, size=5), mine(x, u, pb, size=5))

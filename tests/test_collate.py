"""Host-side collate (graphvqa_b200/collate.py): hand-built scene graphs with known answers; the golden fixture
tests/golden/collate_debug.pt produced by the reference's OWN loader (unmodified gqa_dataset_entry.py run over
oracle/torchtext_shim + oracle/pyg_shim by ``python -m oracle.make_golden collate``); and -- when the reference
tree is mounted (build container only) -- the same comparison live plus the structural facts SURVEY.md section 4
records."""
import json
import os

import pytest
import torch

from graphvqa_b200.collate import SceneGraphVocab, collate_scene_graphs, convert_scene_graph

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "meta_info")), reason="reference tree not mounted")


def _vocab():
    return SceneGraphVocab.from_tokens(["cat", "dog", "dog", "red", "on", "on", "on", "near", "<self>", "<unk>"])


def test_vocab_order_is_specials_then_frequency_then_alphabet():
    v = _vocab()
    assert v.itos == ["<unk>", "<pad>", "<start>", "<end>", "on", "dog", "<self>", "cat", "near", "red"]
    assert v["zebra"] == 0 and v.pad_id == 1 and v.self_id == 6


def test_convert_self_loops_order_and_synthesised_reverse_edges():
    v = _vocab()
    sg = {"objects": {
        "20": {"name": "dog", "attributes": ["red", "red"], "relations": [{"object": "10", "name": "on"}]},
        "10": {"name": "cat", "attributes": [], "relations": [{"object": "20", "name": "near"},
                                                              {"object": "30", "name": "on"}]},
        "30": {"name": "zebra", "attributes": ["red"], "relations": []},
    }}
    x, ei, ea, sym = convert_scene_graph(sg, v)
    # nodes in sorted-id order: "10" cat, "20" dog, "30" zebra (unknown -> 0); slot 0 = name, then distinct attributes
    assert x.shape == (3, 12)
    assert x[:, 0].tolist() == [v["cat"], v["dog"], 0]
    assert x[1, 1:3].tolist() == [v["red"], v.pad_id] and x[0, 1:].eq(v.pad_id).all()
    # node 0: self, ->1 (reverse exists), ->2 + synthesised 2->0; node 1: self, ->0; node 2: self
    assert ei.t().tolist() == [[0, 0], [0, 1], [0, 2], [2, 0], [1, 1], [1, 0], [2, 2]]
    assert ea.squeeze(1).tolist() == [v.self_id, v["near"], v["on"], v["on"], v.self_id, v["on"], v.self_id]
    assert sym.tolist() == [3]


def test_empty_graph_becomes_the_two_node_dummy():
    x, ei, ea, sym = convert_scene_graph({"objects": {}}, _vocab())
    assert x.shape == (2, 12) and ei.t().tolist() == [[0, 0], [0, 1], [1, 1], [1, 0]] and sym.numel() == 0


def test_collate_offsets_edge_index_only():
    v = _vocab()
    sg = {"objects": {"a": {"name": "dog", "attributes": [], "relations": [{"object": "b", "name": "on"}]},
                      "b": {"name": "cat", "attributes": [], "relations": []}}}
    b = collate_scene_graphs([sg, sg, {"objects": {}}], v, pin_memory=False)
    assert b.num_graphs == 3 and b.batch.tolist() == [0, 0, 1, 1, 2, 2]
    one = convert_scene_graph(sg, v)[1]
    assert torch.equal(b.edge_index[:, :one.size(1)], one)
    assert torch.equal(b.edge_index[:, one.size(1):2 * one.size(1)], one + 2)
    assert b.added_sym_edge.tolist() == [2, 2]           # graph-local positions, NOT offset (Appendix A)
    assert b.x.dtype == torch.int64 and b.edge_attr.shape == (b.edge_index.size(1), 1)
    assert b.max_nodes_per_graph == 2


@needs_ref
def test_reference_vocabulary_facts():
    v = SceneGraphVocab.from_meta_info(os.path.join(REF, "meta_info"))
    assert len(v) == 2577
    assert v.itos[:4] == ["<unk>", "<pad>", "<start>", "<end>"] and v.itos[v.self_id] == "<self>"
    assert v.self_id == 1069


@needs_ref
def test_debug_scene_graphs_node_and_edge_counts():
    v = SceneGraphVocab.from_meta_info(os.path.join(REF, "meta_info"))
    with open(os.path.join(REF, "debug_sceneGraphs.json")) as f:
        sgs = json.load(f)
    got = sorted((x.size(0), ei.size(1)) for x, ei, _, _ in (convert_scene_graph(sg, v) for sg in sgs.values()))
    assert got == sorted([(21, 85), (12, 40), (20, 107), (6, 23)])
    b = collate_scene_graphs(list(sgs.values()), v, pin_memory=False)
    assert b.x.size(0) == 59 and b.edge_index.size(1) == 255 and int(b.x[:, 0].eq(0).sum()) == 0


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate_debug.pt")


def _canonical(x):
    """attribute slots sorted: the reference fills them in ``set`` iteration order (hash-seed dependent)"""
    return torch.cat([x[:, :1], x[:, 1:].sort(dim=1).values], dim=1)


def test_collate_matches_the_reference_loader_golden():
    """Bit-exact (int64) against what GQA_gt_sg_feature_lookup.convert_one_gqa_scene_graph and
    Batch.from_data_list produced for the four debug scene graphs and for the empty-graph dummy."""
    fx = torch.load(GOLDEN, weights_only=False)
    v = SceneGraphVocab(fx["itos"])
    assert len(v) == 2577 and v.self_id == fx["self_id"] == 1069 and v.pad_id == fx["pad_id"] == 1
    keys = [g["key"] for g in fx["graphs"]]
    for g in fx["graphs"]:
        x, ei, ea, sym = convert_scene_graph(fx["scene_graphs"][g["key"]], v)
        assert torch.equal(_canonical(x), g["x"]), g["key"]
        assert torch.equal(ei, g["edge_index"]) and torch.equal(ea, g["edge_attr"]), g["key"]
        assert torch.equal(sym, g["added_sym_edge"]), g["key"]
    x, ei, ea, sym = convert_scene_graph({"objects": {}}, v)
    assert torch.equal(_canonical(x), fx["empty"]["x"]) and torch.equal(ei, fx["empty"]["edge_index"])
    assert torch.equal(ea, fx["empty"]["edge_attr"]) and torch.equal(sym, fx["empty"]["added_sym_edge"])
    b = collate_scene_graphs([fx["scene_graphs"][k] for k in keys], v, pin_memory=False)
    want = fx["batch"]
    assert torch.equal(_canonical(b.x), want["x"]) and torch.equal(b.edge_index, want["edge_index"])
    assert torch.equal(b.edge_attr, want["edge_attr"]) and torch.equal(b.batch, want["batch"])
    assert torch.equal(b.added_sym_edge, want["added_sym_edge"])


@needs_ref
def test_vocabulary_and_graphs_equal_the_reference_loader_live():
    """The unmodified reference loader, run here, against the restatement: identical vocabulary (all 2577 tokens in
    the same order, built from meta_info/ by both) and identical tensors for every debug graph."""
    import hashlib
    from oracle import run_reference as rr
    _, lookup = rr.load_scene_graph_lookup("debug")
    ref_vocab = lookup.SG_ENCODING_TEXT.vocab
    v = SceneGraphVocab.from_meta_info(os.path.join(REF, "meta_info"))
    assert v.itos == list(ref_vocab.itos)
    fx = torch.load(GOLDEN, weights_only=False)
    assert hashlib.sha256("\n".join(v.itos).encode()).hexdigest() == fx["itos_sha256"]
    for key, sg in lookup.sg_json_data.items():
        d = lookup.convert_one_gqa_scene_graph(sg)
        x, ei, ea, sym = convert_scene_graph(sg, v)
        assert torch.equal(_canonical(x), _canonical(d.x)), key
        assert torch.equal(ei, d.edge_index) and torch.equal(ea, d.edge_attr) and torch.equal(sym, d.added_sym_edge), key

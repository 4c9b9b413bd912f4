"""Host-side collate of GQA scene graphs: JSON scene graph -> token ids -> one
``SceneGraphBatch`` in pinned memory, ready for ``.to(device, non_blocking=True)``.

Restates, without torchtext / spaCy / torch_geometric:

* the scene-graph vocabulary of ``GQA_gt_sg_feature_lookup.build_scene_graph_encoding_vocab``
  (gqa_dataset_entry.py:136-163): one pseudo-sentence made of the name / attribute / relation lists of
  ``meta_info/`` plus the ontology tables of Constants.py:96-106 plus ``<self>``, fed to a legacy torchtext
  ``Field(init_token="<start>", eos_token="<end>")`` -- i.e. specials ``<unk> <pad> <start> <end>`` first, then
  every token by descending frequency, ties alphabetical, lower-casing off, unknown tokens -> id 0;
* ``convert_one_gqa_scene_graph`` (gqa_dataset_entry.py:190-372): objects sorted by id are the nodes; per
  node 12 token slots (name + its de-duplicated attributes, ``<pad>`` elsewhere), one ``<self>`` loop edge
  first, then each outgoing relation in JSON order, each followed by a synthesised reverse edge (same
  relation token) when the graph has no edge in the opposite direction; the positions of the synthesised
  edges are recorded graph-locally in ``added_sym_edge``; an empty graph becomes the reference's two-node
  ``<UNK>`` dummy;
* ``torch_geometric.data.Batch.from_data_list`` as the reference's collate uses it (gqa_dataset_entry.py:654;
  SURVEY.md Appendix A): node offsets are added to ``edge_index`` only -- ``added_sym_edge`` stays graph-local,
  which the encoder reproduces (pipeline_model_gat.py:590).

The attribute order inside a node follows Python's ``set`` iteration order in the reference (:287), which is
not deterministic across interpreter runs; here attributes keep their first-occurrence order.  The encoder
sums the 12 slot embeddings (:583-585), so the order does not change any result.

Parity status: pinned against the reference's own loader.  The unmodified ``gqa_dataset_entry.py`` runs in the
build container on minimal ``torchtext`` / ``torch_geometric`` stand-ins (the absent third-party wheels; their
vocabulary-ordering and batching rules are restated from the published behaviour), and its
``GQA_gt_sg_feature_lookup('debug')`` output for the four graphs of debug_sceneGraphs.json -- vocabulary (all
2577 tokens, same order), per-graph tensors, the empty-graph dummy and the collated batch -- is committed as
``tests/golden/collate_debug.pt``; tests/test_collate.py compares bit for bit (attribute slots as sorted sets).
"""
import collections
import json
import os

import torch

from .graph_batch import SceneGraphBatch

MAX_OBJ_TOKEN_LEN = 12          # gqa_dataset_entry.py:265 (1 name + up to 11 attributes)
SPECIALS = ("<unk>", "<pad>", "<start>", "<end>")


class SceneGraphVocab:
    """token -> id with torchtext's unknown-token behaviour (missing -> 0)."""

    def __init__(self, itos):
        self.itos = list(itos)
        self.stoi = {t: i for i, t in enumerate(self.itos)}
        self.pad_id = self.stoi["<pad>"]
        self.self_id = self.stoi.get("<self>", 0)

    def __len__(self):
        return len(self.itos)

    def __getitem__(self, token):
        return self.stoi.get(token, 0)

    @classmethod
    def from_tokens(cls, tokens):
        """Legacy torchtext ``Vocab(counter, specials=...)`` order: specials, then by (-frequency, token)."""
        counter = collections.Counter(tokens)
        for s in SPECIALS:
            counter.pop(s, None)
        words = sorted(counter.items(), key=lambda kv: kv[0])
        words.sort(key=lambda kv: kv[1], reverse=True)
        return cls(list(SPECIALS) + [w for w, _ in words])

    @classmethod
    def from_meta_info(cls, meta_info_dir):
        """The reference's scene-graph vocabulary from its ``meta_info/`` directory."""
        def lines(name):
            with open(os.path.join(meta_info_dir, name)) as f:
                return f.read().splitlines()

        def table(name):
            with open(os.path.join(meta_info_dir, name)) as f:
                return json.load(f)

        tokens = lines("name_gqa.txt") + lines("attr_gqa.txt") + lines("rel_gqa.txt")
        tokens += table("objects.json") + table("predicates.json") + table("attributes.json")
        tokens.append("<self>")
        return cls.from_tokens(tokens)


_EMPTY_GRAPH = {"objects": {
    "0": {"name": "<UNK>", "relations": [{"object": "1", "name": "<UNK>"}], "attributes": ["<UNK>"]},
    "1": {"name": "<UNK>", "relations": [{"object": "0", "name": "<UNK>"}], "attributes": ["<UNK>"]},
}}


def convert_scene_graph(sg, vocab):
    """One GQA scene graph (the JSON object with an ``objects`` dict) ->
    (x [n,12] i64, edge_index [2,e] i64, edge_attr [e,1] i64, added_sym_edge [k] i64)."""
    if len(sg["objects"]) == 0:
        sg = _EMPTY_GRAPH
    objs = sg["objects"]
    ids = sorted(objs.keys())
    index = {o: i for i, o in enumerate(ids)}
    present = {(index[o], index[r["object"]]) for o in ids for r in objs[o]["relations"]}
    x = torch.full((len(ids), MAX_OBJ_TOKEN_LEN), vocab.pad_id, dtype=torch.int64)
    src, dst, rel, sym = [], [], [], []
    for o in ids:
        i, obj = index[o], objs[o]
        x[i, 0] = vocab[obj["name"]]
        for slot, attr in enumerate(dict.fromkeys(obj.get("attributes", []))):
            if slot + 1 >= MAX_OBJ_TOKEN_LEN:
                raise ValueError("object %r has more than %d distinct attributes" % (o, MAX_OBJ_TOKEN_LEN - 1))
            x[i, slot + 1] = vocab[attr]
        src.append(i); dst.append(i); rel.append(vocab.self_id)          # the self-loop comes first
        for r in obj["relations"]:
            j, tok = index[r["object"]], vocab[r["name"]]
            src.append(i); dst.append(j); rel.append(tok)
            if (j, i) not in present:                                     # synthesised reverse edge
                src.append(j); dst.append(i); rel.append(tok)
                sym.append(len(rel) - 1)
    return (x, torch.tensor([src, dst], dtype=torch.int64), torch.tensor(rel, dtype=torch.int64).unsqueeze(1),
            torch.tensor(sym, dtype=torch.int64))


def collate_scene_graphs(scene_graphs, vocab, pin_memory=True):
    """List of GQA scene graphs -> one ``SceneGraphBatch`` of disjoint components (host tensors)."""
    xs, eis, eas, syms, batch = [], [], [], [], []
    offset = max_nodes = max_edges = 0
    for g, sg in enumerate(scene_graphs):
        x, ei, ea, sym = convert_scene_graph(sg, vocab)
        xs.append(x); eis.append(ei + offset); eas.append(ea); syms.append(sym)      # sym stays graph-local
        batch.append(torch.full((x.size(0),), g, dtype=torch.int64))
        offset += x.size(0)
        max_nodes, max_edges = max(max_nodes, x.size(0)), max(max_edges, ei.size(1))
    if not xs:
        raise ValueError("collate_scene_graphs: empty list")
    out = SceneGraphBatch(x=torch.cat(xs), edge_index=torch.cat(eis, dim=1), edge_attr=torch.cat(eas),
                          batch=torch.cat(batch), added_sym_edge=torch.cat(syms), num_graphs=len(xs),
                          max_nodes_per_graph=max_nodes, max_in_edges_per_graph=max_edges)
    if pin_memory and torch.cuda.is_available():
        for name in SceneGraphBatch._TENSOR_FIELDS:
            t = getattr(out, name)
            if t is not None:
                setattr(out, name, t.pin_memory())
    return out


# ------------------------------------------------------------------------------------------------------------
# Wire format: what a loader worker hands to the GPU process.
# ------------------------------------------------------------------------------------------------------------
class _PinnedArena:
    """Reusable pinned host buffers, one per field, grown geometrically (cudaHostAlloc costs ~100 us per call)."""

    def __init__(self, pin):
        self.pin, self.bufs = pin, {}

    def get(self, name, count, dtype):
        t = self.bufs.get(name)
        if t is None or t.numel() < count or t.dtype != dtype:
            t = torch.empty(max(int(count * 1.25), 16), dtype=dtype)
            if self.pin:
                t = t.pin_memory()
            self.bufs[name] = t
        return t[:count]


class WireCollator:
    """Batches per-graph tensors into the engine's wire format: int32 everywhere, ONE pinned arena per in-flight
    batch, the destination-CSR built on the host (``gvqa_build_csr_host``) -- the GPU then runs no CSR kernel and
    receives ~1/3 of the bytes of the reference's int64 COO (torch_geometric ``Batch.from_data_list``,
    gqa_dataset_entry.py:631-675).  Vectorised: concatenations and two ``repeat_interleave`` calls, no per-edge
    Python.  ``depth`` arenas are used round-robin (a batch stays valid until ``depth`` more have been collated).

    ``graphs``: list of ``(x [n,12], edge_index [2,e], edge_attr [e,1], added_sym_edge [k])`` integer tensors, i.e.
    what ``convert_scene_graph`` (or the reference's ``convert_one_gqa_scene_graph``) yields per sample."""

    def __init__(self, depth=4, pin_memory=True):
        pin = bool(pin_memory) and torch.cuda.is_available()
        self.arenas = [_PinnedArena(pin) for _ in range(depth)]
        self.count = 0

    def __call__(self, graphs):
        from . import _cabi
        if not graphs:
            raise ValueError("WireCollator: empty list")
        arena = self.arenas[self.count % len(self.arenas)]
        self.count += 1
        b = len(graphs)
        n_per = torch.tensor([g[0].size(0) for g in graphs], dtype=torch.int64)
        e_per = torch.tensor([g[1].size(1) for g in graphs], dtype=torch.int64)
        n, e = int(n_per.sum()), int(e_per.sum())
        i32 = torch.int32
        def cat_into(name, tensors, shape, dim=0):
            """concatenate (one call) and narrow to int32 (one pass) into the arena"""
            dst = arena.get(name, n * MAX_OBJ_TOKEN_LEN if name == "x" else (2 * e if name == "edge_index" else e), i32)
            dst = dst.view(shape)
            if tensors[0].dtype == i32:
                torch.cat(tensors, dim=dim, out=dst)
            else:
                dst.copy_(torch.cat(tensors, dim=dim))
            return dst
        x = cat_into("x", [g[0] for g in graphs], (n, MAX_OBJ_TOKEN_LEN))
        ei = cat_into("edge_index", [g[1] for g in graphs], (2, e), dim=1)
        node_off = (n_per.cumsum(0) - n_per).to(i32)
        ei += torch.repeat_interleave(node_off, e_per, output_size=e)     # node offsets on edge_index only
        ea = cat_into("edge_attr", [g[2].view(-1) for g in graphs], (e,)).view(e, 1)
        batch = arena.get("batch", n, i32)
        batch.copy_(torch.repeat_interleave(torch.arange(b, dtype=i32), n_per, output_size=n))
        # the reference applies the graph-local, un-offset added_sym_edge indices to the BATCHED edge rows
        sym = torch.cat([g[3].reshape(-1) for g in graphs]).to(torch.int64)
        sign = arena.get("edge_sign", e, torch.float32)
        sign.fill_(1.0)
        if sym.numel():
            sign[sym[(sym >= 0) & (sym < e)]] = -1.0
        parts = dict(rowptr=arena.get("rowptr", n + 1, i32), col_src=arena.get("col_src", max(e, 1), i32),
                     perm=arena.get("perm", max(e, 1), i32), graph_ptr=arena.get("graph_ptr", b + 1, i32),
                     node_graph=arena.get("node_graph", max(n, 1), i32), stats=arena.get("stats", 8, i32))
        _cabi.check(_cabi.lib().gvqa_build_csr_host(
            ei.data_ptr(), 4, e, batch.data_ptr(), 4, n, b, parts["rowptr"].data_ptr(), parts["col_src"].data_ptr(),
            parts["perm"].data_ptr(), parts["graph_ptr"].data_ptr(), parts["node_graph"].data_ptr(),
            parts["stats"].data_ptr()), "gvqa_build_csr_host")
        # the row tiles of the fused hop (gvqa_gat_fused_plan_host): with them the GPU runs no plan kernel either
        window = _cabi.fused_window(int(parts["stats"][0]))
        max_tiles = _cabi.lib().gvqa_gat_fused_max_tiles(n, b)
        parts["tiles"] = arena.get("tiles", 4 * max_tiles, i32).view(max_tiles, 4)
        parts["tile_count"] = arena.get("tile_count", 1, i32)
        _cabi.check(_cabi.lib().gvqa_gat_fused_plan_host(parts["graph_ptr"].data_ptr(), b, window, parts["tiles"].data_ptr(),
                                                          parts["tile_count"].data_ptr(), max_tiles),
                    "gvqa_gat_fused_plan_host")
        parts["tile_window"] = window
        return SceneGraphBatch(x=x, edge_index=ei, edge_attr=ea, batch=batch, added_sym_edge=sym.to(i32), edge_sign=sign,
                               num_graphs=b, max_nodes_per_graph=int(parts["stats"][0]),
                               max_in_edges_per_graph=int(parts["stats"][1]), csr_host=parts)


def collate_wire(scene_graphs, vocab, collator=None):
    """Raw GQA scene graphs -> wire-format ``SceneGraphBatch`` (conversion per graph + ``WireCollator``)."""
    collator = collator or WireCollator(depth=1)
    return collator([convert_scene_graph(sg, vocab) for sg in scene_graphs])

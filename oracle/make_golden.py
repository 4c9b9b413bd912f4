"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (over oracle/pyg_shim).

TEST INFRASTRUCTURE, build-container only:   python -m oracle.make_golden            (all model fixtures)
                                             python -m oracle.make_golden collate    (collate_debug.pt only)

Each fixture holds the inputs, the module ``state_dict`` and the outputs the reference's own code
produced on CPU (fp32).  Topologies come from the reference's ``debug_sceneGraphs.json`` (4 graphs:
21/12/20/6 nodes), converted with the edge rules of gqa_dataset_entry.py:255-332 (self-loop first,
then each relation, each followed by a synthesized reverse edge when the reverse is absent).
Feature values are seeded random numbers (the GloVe vocabulary needs network access).
"""
import hashlib
import json
import os

import torch

from . import run_reference as rr

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def debug_graph_list():
    """Per-graph (edge_index[2,e] local ids, num_nodes, added_sym_edge positions) of the 4 debug graphs."""
    with open(os.path.join(rr.REFERENCE_ROOT, "debug_sceneGraphs.json")) as f:
        sgs = json.load(f)
    out = []
    for key in sgs:
        objs = sgs[key]["objects"]
        ids = sorted(objs.keys())
        idx = {o: i for i, o in enumerate(ids)}
        present = {(idx[o], idx[r["object"]]) for o in ids for r in objs[o]["relations"]}
        src, dst, sym = [], [], []
        for o in ids:
            i = idx[o]
            src.append(i); dst.append(i)
            for r in objs[o]["relations"]:
                j = idx[r["object"]]
                src.append(i); dst.append(j)
                if (j, i) not in present:
                    src.append(j); dst.append(i)
                    sym.append(len(src) - 1)
        out.append((torch.tensor([src, dst], dtype=torch.long), len(ids), torch.tensor(sym, dtype=torch.long)))
    return out


def debug_topology():
    """(edge_index[2,E], batch[N]) of the 4 debug scene graphs batched as disjoint components."""
    eis, batch, off = [], [], 0
    for gi, (ei, n, _) in enumerate(debug_graph_list()):
        eis.append(ei + off); batch += [gi] * n; off += n
    return torch.cat(eis, dim=1), torch.tensor(batch, dtype=torch.long)


def debug_token_batch(seed):
    """The debug graphs as the reference's collate would deliver them (torch_geometric Batch via the
    shim: node offsets on edge_index only, added_sym_edge left graph-local), with seeded random
    token ids in place of the GloVe vocabulary lookups (pad id 1 in unused attribute slots)."""
    import torch_geometric
    g = torch.Generator().manual_seed(seed)
    data = []
    for ei, n, sym in debug_graph_list():
        x = torch.randint(4, rr.SG_VOCAB_SIZE, (n, 12), generator=g)
        x[torch.rand(n, 12, generator=g) < 0.6] = 1
        x[:, 0] = torch.randint(4, rr.SG_VOCAB_SIZE, (n,), generator=g)
        ea = torch.randint(4, rr.SG_VOCAB_SIZE, (ei.size(1), 1), generator=g)
        d = torch_geometric.data.Data(x=x, edge_index=ei, edge_attr=ea)
        d.added_sym_edge = sym
        data.append(d)
    return torch_geometric.data.Batch.from_data_list(data)


def _randomise_bn(module, gen):
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1, generator=gen)
            m.running_var.uniform_(0.5, 1.5, generator=gen)
            with torch.no_grad():
                m.weight.uniform_(0.5, 1.5, generator=gen)
                m.bias.normal_(0, 0.1, generator=gen)


def _state_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode()); h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def canonical_tokens(x):
    """Node token rows with the attribute slots (1..11) sorted: the reference fills them in ``set`` iteration
    order (gqa_dataset_entry.py:287), which depends on the interpreter's string hash seed; the encoder sums the
    slot embeddings, so the order carries no information."""
    return torch.cat([x[:, :1], x[:, 1:].sort(dim=1).values], dim=1)


def make_collate_fixture():
    """tests/golden/collate_debug.pt: what the reference's OWN loader (``GQA_gt_sg_feature_lookup('debug')`` of the
    unmodified gqa_dataset_entry.py, on oracle/torchtext_shim + oracle/pyg_shim) produces for the four graphs of
    debug_sceneGraphs.json: its scene-graph vocabulary, every graph's tensors and the collated ``Batch``."""
    os.makedirs(OUT, exist_ok=True)
    _, lookup = rr.load_scene_graph_lookup("debug")      # puts the shims on sys.path
    import torch_geometric
    vocab = lookup.SG_ENCODING_TEXT.vocab
    graphs, data = [], []
    for key, sg in lookup.sg_json_data.items():
        d = lookup.convert_one_gqa_scene_graph(sg)
        data.append(d)
        graphs.append(dict(key=key, x=canonical_tokens(d.x), edge_index=d.edge_index, edge_attr=d.edge_attr,
                           added_sym_edge=d.added_sym_edge))
    empty = lookup.convert_one_gqa_scene_graph({"objects": {}})
    b = torch_geometric.data.Batch.from_data_list(data)
    fx = dict(meta=dict(generator="oracle/make_golden.py collate", reference=rr.REFERENCE_ROOT, torch=torch.__version__),
              scene_graphs=lookup.sg_json_data, itos=list(vocab.itos),
              itos_sha256=hashlib.sha256("\n".join(vocab.itos).encode()).hexdigest(),
              self_id=int(vocab.stoi["<self>"]), pad_id=int(vocab.stoi[lookup.SG_ENCODING_TEXT.pad_token]),
              graphs=graphs,
              empty=dict(x=canonical_tokens(empty.x), edge_index=empty.edge_index, edge_attr=empty.edge_attr,
                         added_sym_edge=empty.added_sym_edge),
              batch=dict(x=canonical_tokens(b.x), edge_index=b.edge_index, edge_attr=b.edge_attr, batch=b.batch,
                         added_sym_edge=b.added_sym_edge))
    path = os.path.join(OUT, "collate_debug.pt")
    torch.save(fx, path)
    print(os.path.basename(path), os.path.getsize(path))


def make_graph_side_fixtures():
    """tests/golden/graph_side_gat.pt + encoder_pool_edge.pt.

    graph_side_gat: intermediates of the UNMODIFIED reference PipelineModel on the pipeline_gat.pt inputs, captured
    with forward hooks: what the text side hands to the graph side (``instr_vectors``, ``questions_encoded[0]``) and
    what the graph side produces after the hop stack (``x_executed``), after the conditional attention pooling
    (``pooled``) and after ``logit_fc``.
    encoder_pool_edge: the reference's GroundTruth_SceneGraph_Encoder on a 7-graph batch with single-node graphs,
    nodes without in-edges (scatter_mean's clamp, pipeline_model_gat.py:96) and un-offset ``added_sym_edge`` entries
    of more than 4 graphs (:590); and the reference's MyConditionalGlobalAttention with an EMPTY graph in the middle
    of the batch and ``size`` larger than ``batch.max() + 1``."""
    from .golden_utils import deterministic_fill, state_hash
    os.makedirs(OUT, exist_ok=True)
    meta = dict(generator="oracle/make_golden.py graph_side", reference=rr.REFERENCE_ROOT, torch=torch.__version__)
    mod = rr.load("pipeline_model_gat")
    import torch_geometric
    with torch.no_grad():
        base = torch.load(os.path.join(OUT, "pipeline_gat.pt"), weights_only=False)
        torch.manual_seed(0)
        model = deterministic_fill(mod.PipelineModel().eval(), seed=base["fill_seed"])
        assert state_hash(model.state_dict()) == base["state_sha256"]
        batch_obj = debug_token_batch(seed=606)
        assert torch.equal(batch_obj.x, base["x"])
        cap = {}
        hooks = [
            model.gat_seq.register_forward_hook(
                lambda m, a, kw, out: cap.update(instr_vectors=kw["instr_vectors"].clone(), x_executed=out.clone()),
                with_kwargs=True),
            model.graph_global_attention_pooling.register_forward_hook(
                lambda m, a, kw, out: cap.update(q0=kw["u"].clone(), pooled=out.clone()), with_kwargs=True)]
        _, logits = model(base["questions"], batch_obj, base["programs_input"], None, SAMPLE_FLAG=False)
        for hk in hooks:
            hk.remove()
        assert torch.equal(logits, base["short_answer_logits"])
        torch.save(dict(meta=meta, inputs="pipeline_gat.pt", fill_seed=base["fill_seed"],
                        state_sha256=base["state_sha256"], **cap), os.path.join(OUT, "graph_side_gat.pt"))

        # ---- edge cases -------------------------------------------------------------------------
        gen = torch.Generator().manual_seed(808)
        sizes = [1, 5, 1, 9, 3, 7, 2]
        data = []
        for n in sizes:
            src = list(range(1, n))          # node 0 of every graph has NO self-loop and no in-edge
            dst = list(range(1, n))
            m_extra = 2 * n if n > 1 else 0
            s_ = torch.randint(0, n, (m_extra,), generator=gen).tolist()
            d_ = (torch.randint(1, n, (m_extra,), generator=gen).tolist() if n > 1 else [])
            ei = torch.tensor([src + s_, dst + d_], dtype=torch.long).reshape(2, -1)
            x = torch.randint(4, rr.SG_VOCAB_SIZE, (n, 12), generator=gen)
            x[torch.rand(n, 12, generator=gen) < 0.5] = 1
            ea = torch.randint(4, rr.SG_VOCAB_SIZE, (ei.size(1), 1), generator=gen)
            d = torch_geometric.data.Data(x=x, edge_index=ei, edge_attr=ea)
            k = min(3, ei.size(1))
            d.added_sym_edge = torch.randperm(max(ei.size(1), 1), generator=gen)[:k] if ei.size(1) else torch.zeros(0, dtype=torch.long)
            data.append(d)
        b = torch_geometric.data.Batch.from_data_list(data)
        enc = model.scene_graph_encoder
        x_enc, e_enc, _ = enc(b)
        fx = dict(meta=meta, fill_seed=base["fill_seed"], state_sha256=base["state_sha256"],
                  enc=dict(x=b.x, edge_index=b.edge_index, edge_attr=b.edge_attr, batch=b.batch,
                           added_sym_edge=b.added_sym_edge, num_graphs=len(sizes), x_encoded=x_enc,
                           edge_attr_encoded=e_enc))
        # pooling: 6 graph slots, graph 2 empty, graph 5 (the last slot, beyond batch.max()) empty too
        torch.manual_seed(809)
        pool = mod.MyConditionalGlobalAttention(num_node_features=20, num_out_features=16).eval()
        pb = torch.tensor([0] * 4 + [1] * 1 + [3] * 6 + [4] * 2, dtype=torch.long)
        px = torch.randn(pb.numel(), 20, generator=gen)
        pu = torch.randn(6, 16, generator=gen)
        fx["pool"] = dict(state=pool.state_dict(), x=px, u=pu, batch=pb, size=6, out=pool(px, pu, pb, size=6),
                          out_default_size=pool(px, pu[:5], pb))
        torch.save(fx, os.path.join(OUT, "encoder_pool_edge.pt"))
    for fn in ("graph_side_gat.pt", "encoder_pool_edge.pt"):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


def main():
    import sys
    if sys.argv[1:] == ["collate"]:
        make_collate_fixture()
        return
    if sys.argv[1:] == ["graph_side"]:
        make_graph_side_fixtures()
        return
    os.makedirs(OUT, exist_ok=True)
    ei, batch = debug_topology()
    n, e, b = batch.numel(), ei.size(1), int(batch.max()) + 1
    gat_skip = rr.load("gat_skip")
    lnmod = rr.load("graph_utils.my_graph_layernorm")
    lcgn = rr.load("lcgn")
    gine = rr.load("pipeline_model_gine")
    gcn = rr.load("pipeline_model_gcn")
    meta = dict(generator="oracle/make_golden.py", reference=rr.REFERENCE_ROOT, torch=torch.__version__)

    with torch.no_grad():
        # 1. gat_seq, small dims, weights stored
        gen = torch.Generator().manual_seed(101)
        torch.manual_seed(101)
        f, d, hops = 32, 16, 3
        m = gat_skip.gat_seq(f, f, f, d, hops, dropout=0.1, gat_heads=4).eval()
        _randomise_bn(m, gen)
        for c in m.convs:
            c.bias.normal_(0, 0.1, generator=gen)
        x = torch.randn(n, f, generator=gen); ea = torch.randn(e, f, generator=gen)
        ins = torch.randn(hops, b, d, generator=gen)
        out = m(x, ei, ea, ins, batch)
        conv0 = m.convs[0]
        x_cat = torch.cat((x, ins[0][batch]), -1); e_cat = torch.cat((ea, ins[0][batch[ei[0]]]), -1)
        c_out, (_, alpha) = conv0(x_cat, ei, e_cat, return_attention_weights=True)
        torch.save(dict(meta=meta, config=dict(in_channels=f, out_channels=f, edge_attr_dim=f, ins_dim=d,
                                               num_ins=hops, dropout=0.1, gat_heads=4),
                        state=m.state_dict(), x=x, edge_index=ei, edge_attr=ea, instr_vectors=ins, batch=batch,
                        out=out, conv0_out=c_out, conv0_alpha=alpha), os.path.join(OUT, "gat_seq_small.pt"))

        # 2. gat_seq at the reference dims (F=300, D=512, H=4, 5 hops); weights by seed + hash
        gen = torch.Generator().manual_seed(202)
        torch.manual_seed(202)
        m = gat_skip.gat_seq(300, 300, 300, 512, 5, dropout=0.1, gat_heads=4).eval()
        sd_hash = _state_hash(m.state_dict())
        x = torch.randn(n, 300, generator=gen); ea = torch.randn(e, 300, generator=gen)
        ins = torch.randn(5, b, 512, generator=gen)
        out = m(x, ei, ea, ins, batch)
        torch.save(dict(meta=meta, seed=202, state_sha256=sd_hash, x=x, edge_index=ei, edge_attr=ea,
                        instr_vectors=ins, batch=batch, out=out), os.path.join(OUT, "gat_seq_refdims.pt"))

        # 3. graph LayerNorm
        gen = torch.Generator().manual_seed(303)
        ln = lnmod.LayerNorm(300)
        ln.weight.fill_(1.3); ln.bias.fill_(-0.2)
        x = torch.randn(n, 300, generator=gen) * 2 + 0.5
        torch.save(dict(meta=meta, state=ln.state_dict(), x=x, batch=batch, out=ln(x, batch)),
                   os.path.join(OUT, "graph_layernorm.pt"))

        # 4. lcgn_seq small (x_ctx = the torch.randn the reference draws at lcgn.py:306)
        gen = torch.Generator().manual_seed(404)
        torch.manual_seed(404)
        m = lcgn.lcgn_seq(in_channels=24, out_channels=32, edge_attr_dim=24, num_ins=5, gat_cmd_dim=32,
                          question_dim=32, dropout=0.1).eval()
        m.lcgn.bias.normal_(0, 0.1, generator=gen)
        x = torch.randn(n, 24, generator=gen); q = torch.randn(b, 32, generator=gen)
        lo = torch.randn(7, b, 32, generator=gen)
        torch.manual_seed(405); x_ctx = torch.randn(n, 32)
        torch.manual_seed(405); out = m(x, ei, batch, q, lo)
        torch.save(dict(meta=meta, config=dict(in_channels=24, out_channels=32, edge_attr_dim=24, num_ins=5,
                                               gat_cmd_dim=32, question_dim=32, dropout=0.1),
                        state=m.state_dict(), x=x, edge_index=ei, batch=batch, q_encoding=q, lstm_outputs=lo,
                        x_ctx=x_ctx, out=out), os.path.join(OUT, "lcgn_seq_small.pt"))

        # 5/6. gine_seq / gcn_seq small: bug-faithful output + the (discarded) conv results
        for name, mod, cls in (("gine", gine, "gine_seq"), ("gcn", gcn, "gcn_seq")):
            gen = torch.Generator().manual_seed(505)
            torch.manual_seed(505)
            m = getattr(mod, cls)(32, 32, 16, dropout=0.1).eval()
            _randomise_bn(m, gen)
            x = torch.randn(n, 32, generator=gen); ea = torch.randn(e, 32, generator=gen)
            ins = torch.randn(5, b, 16, generator=gen)
            conv_out = []
            hooks = [c.register_forward_hook(lambda _m, _i, o: conv_out.append(o.clone())) for c in m.convs]
            out = m(x, ei, ea, ins, batch) if name == "gine" else m(x, ei, ins, batch)
            for hk in hooks:
                hk.remove()
            torch.save(dict(meta=meta, config=dict(in_channels=32, out_channels=32, ins_dim=16, dropout=0.1),
                            state=m.state_dict(), x=x, edge_index=ei, edge_attr=ea, instr_vectors=ins,
                            batch=batch, out=out, conv_out=conv_out), os.path.join(OUT, "%s_seq_small.pt" % name))
        # 7. the whole PipelineModel of every variant (67M parameters: weights reproduced from a seed by
        #    oracle/golden_utils.deterministic_fill, identified by a hash)
        from .golden_utils import deterministic_fill, state_hash
        batch_obj = debug_token_batch(seed=606)
        gen = torch.Generator().manual_seed(607)
        nb = 4
        questions = torch.randint(4, rr.TEXT_VOCAB_SIZE, (9, nb), generator=gen)
        programs = torch.randint(4, rr.TEXT_VOCAB_SIZE, (6, nb * 5), generator=gen)
        programs[0] = 2
        for variant in ("gat", "gcn", "gine", "lcgn"):
            mod = rr.load("pipeline_model_" + variant)
            torch.manual_seed(0)
            model = deterministic_fill(mod.PipelineModel().eval(), seed=700)
            fx = dict(meta=meta, variant=variant, fill_seed=700, state_sha256=state_hash(model.state_dict()),
                      questions=questions, programs_input=programs, x=batch_obj.x, edge_index=batch_obj.edge_index,
                      edge_attr=batch_obj.edge_attr, added_sym_edge=batch_obj.added_sym_edge, batch=batch_obj.batch)
            if variant == "lcgn":
                torch.manual_seed(708); fx["x_ctx"] = torch.randn(batch_obj.batch.numel(), 512)
                torch.manual_seed(708)
            prog_out, logits = model(questions, batch_obj, programs, None, SAMPLE_FLAG=False)
            fx["short_answer_logits"] = logits
            fx["programs_output_slice"] = prog_out[:, :, :48].clone()
            fx["programs_argmax"] = prog_out.argmax(-1)
            x_enc, e_enc, _ = model.scene_graph_encoder(batch_obj)
            fx["x_encoded"], fx["edge_attr_encoded"] = x_enc, e_enc
            if variant == "gat":
                torch.manual_seed(1)
                sampled, _ = model.program_decoder.sample(model.question_encoder(questions), None)
                fx["sampled_programs"] = sampled
            torch.save(fx, os.path.join(OUT, "pipeline_%s.pt" % variant))
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()

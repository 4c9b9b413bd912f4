"""CPU tests of the multi-GPU host logic with the gloo backend (world_size 2 and 3): graph-range
sharding with index re-basing, and the logits all-gather (equal and ragged shards).  The per-rank
compute is a stand-in per-graph function -- the CUDA engine itself is covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import random_graphs
from graphvqa_b200.dist import all_gather_logits, graph_range, shard_scene_graphs
from graphvqa_b200.graph_batch import SceneGraphBatch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _per_graph_feature(graphs, num_graphs, width=5):
    """Deterministic stand-in for the per-graph model output: depends on node features and topology."""
    deg = torch.zeros(graphs.batch.numel()).index_add_(0, graphs.edge_index[1],
                                                      torch.ones(graphs.edge_index.size(1)))
    node = graphs.x.float().sum(1) * (1 + deg) + graphs.x.float()[graphs.edge_index[0]].sum(1).new_zeros(1)
    # like the encoder (pipeline_model_gat.py:590): rows `added_sym_edge` of the BATCHED edge array change sign
    sign = torch.ones(graphs.edge_index.size(1))
    sym = getattr(graphs, "added_sym_edge", None)
    if sym is not None and sym.numel():
        sign[sym] = -1.0
    src_sum = torch.zeros(graphs.batch.numel()).index_add_(
        0, graphs.edge_index[1], sign * (graphs.x.float().sum(1)[graphs.edge_index[0]] + graphs.edge_attr.float()[:, 0]))
    out = torch.zeros(num_graphs, width)
    for k in range(width):
        out[:, k].index_add_(0, graphs.batch, (node + (k + 1) * src_sum) * (k + 1))
    return out


def _make_batch(num_graphs, seed):
    ei, batch = random_graphs(num_graphs, 1, 9, 1.5, seed=seed, isolated=True)
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 50, (batch.numel(), 3), generator=g)
    ea = torch.randint(0, 50, (ei.size(1), 1), generator=g)
    # graph-local, un-offset positions of synthesized reverse edges, concatenated over the graphs (what
    # Batch.from_data_list yields): small numbers with repeats, all pointing into the first rows of the batch
    sym = torch.randint(0, max(2, ei.size(1) // 3), (2 * num_graphs,), generator=g)
    return SceneGraphBatch(x=x, edge_index=ei, edge_attr=ea, batch=batch, added_sym_edge=sym, num_graphs=num_graphs)


def test_graph_range_partitions_exactly():
    for b in (1, 7, 8, 256, 1000):
        for w in (1, 2, 3, 8):
            spans = [graph_range(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shards_rebase_indices_and_cover_the_batch():
    graphs = _make_batch(11, seed=3)
    full = _per_graph_feature(graphs, 11)
    parts, n_seen, e_seen = [], 0, 0
    for r in range(3):
        s = shard_scene_graphs(graphs, r, 3)
        lo, hi = graph_range(11, r, 3)
        assert s.num_graphs == hi - lo
        assert int(s.batch.min()) == 0 and int(s.batch.max()) == hi - lo - 1
        assert int(s.edge_index.min()) >= 0 and int(s.edge_index.max()) < s.batch.numel()
        assert s.edge_attr.size(0) == s.edge_index.size(1)
        n_seen += s.batch.numel(); e_seen += s.edge_index.size(1)
        parts.append(_per_graph_feature(s, hi - lo))
    assert n_seen == graphs.batch.numel() and e_seen == graphs.edge_index.size(1)
    assert torch.equal(torch.cat(parts), full)        # graphs are independent: bitwise identical


def test_shards_keep_the_sym_edge_rows_of_the_single_batch():
    """`added_sym_edge` indexes rows of the batched edge array (un-offset, pipeline_model_gat.py:590): every shard
    must negate exactly the rows of its own slice that the single-GPU run negates -- not the full list."""
    graphs = _make_batch(9, seed=11)
    assert graphs.added_sym_edge.numel() > 0
    negated = torch.zeros(graphs.edge_index.size(1), dtype=torch.bool)
    negated[graphs.added_sym_edge] = True
    seen = 0
    for r in range(4):
        s = shard_scene_graphs(graphs, r, 4)
        assert s.added_sym_edge.numel() == 0 or int(s.added_sym_edge.max()) < s.edge_index.size(1)
        local = torch.zeros(s.edge_index.size(1), dtype=torch.bool)
        local[s.added_sym_edge] = True
        assert torch.equal(local, negated[s.edge_mask])
        seen += int(local.sum())
    assert seen == int(negated.sum())
    # a stale index beyond the edge array (the reference would fault) is dropped, not propagated
    graphs.added_sym_edge = torch.cat([graphs.added_sym_edge, torch.tensor([10 ** 6])])
    assert int(shard_scene_graphs(graphs, 0, 2).added_sym_edge.max()) < graphs.edge_index.size(1)


def _worker(rank, world, port, num_graphs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        graphs = _make_batch(num_graphs, seed=5)
        lo, hi = graph_range(num_graphs, rank, world)
        local = _per_graph_feature(shard_scene_graphs(graphs, rank, world), hi - lo)
        gathered = all_gather_logits(local, num_graphs)
        want = _per_graph_feature(graphs, num_graphs)
        ret[rank] = bool(torch.equal(gathered, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,num_graphs", [(2, 8), (2, 7), (3, 10)])
def test_all_gather_logits_gloo(world, num_graphs):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, num_graphs, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)

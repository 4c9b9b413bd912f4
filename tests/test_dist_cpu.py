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
    src_sum = torch.zeros(graphs.batch.numel()).index_add_(0, graphs.edge_index[1],
                                                          graphs.x.float().sum(1)[graphs.edge_index[0]])
    out = torch.zeros(num_graphs, width)
    for k in range(width):
        out[:, k].index_add_(0, graphs.batch, (node + (k + 1) * src_sum) * (k + 1))
    return out


def _make_batch(num_graphs, seed):
    ei, batch = random_graphs(num_graphs, 1, 9, 1.5, seed=seed, isolated=True)
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 50, (batch.numel(), 3), generator=g)
    ea = torch.randint(0, 50, (ei.size(1), 1), generator=g)
    return SceneGraphBatch(x=x, edge_index=ei, edge_attr=ea, batch=batch, num_graphs=num_graphs)


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

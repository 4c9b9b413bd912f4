"""Multi-GPU execution: graphs shard, nothing else moves.

The reference scales with DistributedDataParallel + DistributedSampler (mainExplain_gat.py:197-202,
226-229, 259-263): one process per GPU, every rank sees whole scene graphs, and -- at inference --
no collective at all (each rank even writes its own result dump, :938-942).  The engine keeps that
shape: a batch of disjoint graphs is split into contiguous graph ranges (``shard_scene_graphs``),
each rank runs the unchanged single-GPU path on its range, and ONE collective assembles the
answer logits (``all_gather_logits``: NCCL all-gather of [B/G, 1842] fp32, < 1 MB per rank).
Every op on the path is per-graph, so the G-GPU result equals the 1-GPU result row for row (including the
reference's un-offset ``added_sym_edge`` negation, see ``shard_scene_graphs``).
"""
import os

import torch
import torch.distributed as dist

from .graph_batch import SceneGraphBatch


def init_distributed(backend=None):
    """env:// initialisation like util/misc.py:370-392 of the reference (RANK / WORLD_SIZE /
    LOCAL_RANK); returns (rank, local_rank, world_size).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def graph_range(num_graphs, rank, world):
    """Contiguous, balanced graph range of ``rank`` (first ``num_graphs % world`` ranks get one more)."""
    base, rem = divmod(num_graphs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_scene_graphs(graphs, rank, world, num_graphs=None):
    """The sub-batch of graphs [lo, hi) of ``graphs`` (any object with the SceneGraphBatch fields) with
    node ids re-based to start at 0.  Works on CPU or GPU tensors; ``batch`` must be sorted.
    ``added_sym_edge``: the reference applies these indices to rows of the BATCHED edge array although
    they are graph-local (Batch.from_data_list does not offset them; pipeline_model_gat.py:590), so the
    single-GPU run negates the embedding of edge rows ``added_sym_edge`` of the full batch.  To make the
    sharded result equal the single-GPU result row for row, the shard carries exactly those rows that
    fall inside its edge slice, re-based to shard-local edge positions (out-of-range entries, which the
    reference would fault on, are dropped).  This is NOT what per-rank collation under
    DistributedSampler produces (each rank would then negate the low-numbered rows of its own batch);
    ``collate_scene_graphs`` on the rank's own graphs reproduces that behaviour."""
    b = num_graphs if num_graphs is not None else getattr(graphs, "num_graphs", None)
    if b is None:
        b = int(graphs.batch.max()) + 1
    lo, hi = graph_range(b, rank, world)
    batch = graphs.batch
    bounds = torch.searchsorted(batch, torch.tensor([lo, hi], dtype=batch.dtype, device=batch.device))
    n0, n1 = int(bounds[0]), int(bounds[1])
    ei = graphs.edge_index
    keep = (ei[1] >= n0) & (ei[1] < n1)
    out = SceneGraphBatch(num_graphs=hi - lo,
                          max_nodes_per_graph=getattr(graphs, "max_nodes_per_graph", 0),
                          max_in_edges_per_graph=getattr(graphs, "max_in_edges_per_graph", 0))
    out.x = None if getattr(graphs, "x", None) is None else graphs.x[n0:n1]
    out.y = None if getattr(graphs, "y", None) is None else graphs.y[n0:n1]
    out.batch = batch[n0:n1] - lo
    out.edge_index = ei[:, keep] - n0
    out.edge_attr = None if getattr(graphs, "edge_attr", None) is None else graphs.edge_attr[keep]
    sym = getattr(graphs, "added_sym_edge", None)
    if sym is not None:
        mask = torch.zeros(ei.size(1), dtype=torch.bool, device=ei.device)
        sym = sym[(sym >= 0) & (sym < ei.size(1))]
        mask[sym] = True
        sym = mask[keep].nonzero().flatten()
    out.added_sym_edge = sym
    out.node_range, out.edge_mask = (n0, n1), keep
    return out


def all_gather_logits(local_logits, num_graphs_total=None):
    """Assemble [B_total, A] from per-rank [B_r, A] (ranks hold consecutive graph ranges as produced by
    ``graph_range``).  Equal shards use one all_gather_into_tensor; ragged shards are padded to the
    largest shard first.  Single process: returns the input."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    a = local_logits.size(1)
    if num_graphs_total is None:
        t = torch.tensor([local_logits.size(0)], dtype=torch.int64, device=local_logits.device)
        dist.all_reduce(t)
        num_graphs_total = int(t)
    sizes = [graph_range(num_graphs_total, r, world) for r in range(world)]
    sizes = [hi - lo for lo, hi in sizes]
    mx = max(sizes)
    if all(s == mx for s in sizes):
        out = local_logits.new_empty(world * mx, a)
        dist.all_gather_into_tensor(out, local_logits.contiguous())
        return out
    padded = local_logits.new_zeros(mx, a)
    padded[:local_logits.size(0)] = local_logits
    out = local_logits.new_empty(world * mx, a)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * mx:r * mx + s] for r, s in enumerate(sizes)])


@torch.no_grad()
def distributed_answer_logits(model, questions, graphs, rank=None, world=None):
    """Shard ``(questions[L,B], graphs)`` by graph range, run ``model.answer_logits`` on this rank's
    shard and all-gather the [B, 1842] logits."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    b = questions.size(1)
    lo, hi = graph_range(b, rank, world)
    shard = shard_scene_graphs(graphs, rank, world, num_graphs=b)
    local = model.answer_logits(questions[:, lo:hi].contiguous(), shard)
    return all_gather_logits(local, b)

"""Restatement of the torch_geometric / torch_scatter primitives the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference imports these from
un-vendored, un-pinned wheels (README.md:79-90 of the reference shows an install
example only); its call signatures (``softmax(src, index, ptr, size_i)``,
``MessagePassing.message(..., index, ptr, size_i)``, ``OptPairTensor``) place
them at torch_geometric 1.6.2-1.7.x / torch_scatter 2.0.x.  Each function below
restates the published definition of one primitive and names the reference call
site that depends on it.  All arithmetic is done with plain dense torch ops
(``index_add_`` / ``scatter_reduce_``) in the dtype of the input.
"""
import math

import torch


def scatter_sum(src, index, dim_size):
    """torch_scatter.scatter(src, index, dim=0, dim_size=dim_size, reduce='add').

    Used by MessagePassing.aggregate (reference gat_skip.py:155 -> propagate),
    graph_utils/my_graph_layernorm.py:64,69 and pipeline_model_gat.py:179.
    Rows of the output that receive nothing stay 0.
    """
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def scatter_mean(src, index, dim_size):
    """torch_scatter.scatter_mean: sum / count.clamp(min=1) (pipeline_model_gat.py:96)."""
    total = scatter_sum(src, index, dim_size)
    count = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
    count.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    count = count.clamp(min=1)
    return total / count.view((-1,) + (1,) * (src.dim() - 1))


def scatter_max(src, index, dim_size):
    """torch_scatter.scatter_max value output; segments that receive nothing hold 0."""
    out = torch.full((dim_size,) + tuple(src.shape[1:]), float("-inf"),
                     dtype=src.dtype, device=src.device)
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    out.scatter_reduce_(0, idx, src, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)


def degree(index, num_nodes, dtype=torch.float32):
    """torch_geometric.utils.degree (my_graph_layernorm.py:61)."""
    out = torch.zeros(num_nodes, dtype=dtype, device=index.device)
    return out.index_add_(0, index, torch.ones(index.numel(), dtype=dtype, device=index.device))


def segment_softmax(src, index, num_nodes):
    """torch_geometric.utils.softmax(src, index, ptr=None, num_nodes).

    out = exp(src - max_seg[index]) / (sum_seg(exp)[index] + 1e-16), independently per
    trailing column.  Call sites: gat_skip.py:188, lcgn.py:211, pipeline_model_gat.py:178.
    """
    seg_max = scatter_max(src, index, num_nodes)
    out = (src - seg_max.index_select(0, index)).exp()
    seg_sum = scatter_sum(out, index, num_nodes)
    return out / (seg_sum.index_select(0, index) + 1e-16)


def glorot_(tensor):
    """torch_geometric.nn.inits.glorot: U(-a, a), a = sqrt(6 / (size(-2) + size(-1)))."""
    bound = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
    with torch.no_grad():
        tensor.uniform_(-bound, bound)
    return tensor


def gcn_norm(edge_index, num_nodes, dtype=torch.float32):
    """torch_geometric.nn.conv.gcn_conv.gcn_norm with the GCNConv defaults
    (improved=False, add_self_loops=True, edge_weight=None).

    ``add_remaining_self_loops`` drops every existing self-loop edge and appends exactly
    one loop per node (weight 1 here because edge_weight is None); duplicate non-loop
    edges are kept and counted.  deg is accumulated over the *target* index (col);
    norm = deg^-1/2[row] * w * deg^-1/2[col] with inf -> 0.
    Returns (edge_index_with_loops [2, E'], norm [E']).
    Call site: baseline_and_test_models/pipeline_model_gcn.py:660 (GCNConv.forward).
    """
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loops = torch.arange(num_nodes, dtype=row.dtype, device=row.device)
    row = torch.cat([row[keep], loops])
    col = torch.cat([col[keep], loops])
    weight = torch.ones(row.numel(), dtype=dtype, device=row.device)
    deg = scatter_sum(weight, col, num_nodes)
    dinv = deg.pow(-0.5)
    dinv = torch.where(torch.isinf(dinv), torch.zeros_like(dinv), dinv)
    return torch.stack([row, col]), dinv[row] * weight * dinv[col]

"""torch_geometric.utils subset: softmax, degree, self-loop helpers."""
import torch

from oracle.pyg_semantics import degree as _degree
from oracle.pyg_semantics import segment_softmax as _segment_softmax


def softmax(src, index, ptr=None, num_nodes=None):
    if ptr is not None:
        raise NotImplementedError("shim: CSR-ptr softmax is not used by the reference")
    if num_nodes is None:
        num_nodes = int(index.max()) + 1 if index.numel() > 0 else 0
    return _segment_softmax(src, index, num_nodes)


def degree(index, num_nodes=None, dtype=None):
    if num_nodes is None:
        num_nodes = int(index.max()) + 1 if index.numel() > 0 else 0
    return _degree(index, num_nodes, dtype or torch.get_default_dtype())


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    edge_index = torch.cat([edge_index, loops.unsqueeze(0).repeat(2, 1)], dim=1)
    if edge_weight is not None:
        edge_weight = torch.cat([edge_weight, edge_weight.new_full((num_nodes,), fill_value)])
    return edge_index, edge_weight

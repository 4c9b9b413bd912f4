"""torch_geometric.nn.conv subset: MessagePassing (aggr='add', source_to_target), GCNConv, GINEConv."""
import inspect

import torch
from torch.nn import Parameter

from oracle.pyg_semantics import gcn_norm, glorot_, scatter_sum


class MessagePassing(torch.nn.Module):
    """propagate(): for every parameter of self.message named ``foo_j`` / ``foo_i`` gather
    kwargs['foo'] (element 0 / 1 if it is a pair) at edge_index[0] / edge_index[1];
    ``index`` = edge_index[1], ``ptr`` = None, ``size_i`` = number of target nodes; other
    parameters are passed through; messages are summed per target node."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        if aggr != "add" or flow != "source_to_target":
            raise NotImplementedError("shim supports aggr='add', flow='source_to_target' only")
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        sizes = [None, None] if size is None else list(size)

        def note_size(side, tensor):
            if torch.is_tensor(tensor) and sizes[side] is None:
                sizes[side] = tensor.size(self.node_dim)

        names = [n for n in inspect.signature(self.message).parameters]
        for name in names:  # first pass: infer node counts like PyG's __set_size__
            if name.endswith("_j") or name.endswith("_i"):
                data = kwargs.get(name[:-2])
                if isinstance(data, (tuple, list)):
                    note_size(0, data[0]); note_size(1, data[1])
                else:
                    note_size(0, data); note_size(1, data)
        if sizes[0] is None:
            sizes[0] = sizes[1]
        if sizes[1] is None:
            sizes[1] = sizes[0]

        collected = {}
        for name in names:
            if name.endswith("_j") or name.endswith("_i"):
                side = 0 if name.endswith("_j") else 1
                data = kwargs.get(name[:-2])
                if isinstance(data, (tuple, list)):
                    data = data[side]
                if torch.is_tensor(data):
                    data = data.index_select(self.node_dim, src if side == 0 else dst)
                collected[name] = data
            elif name == "index":
                collected[name] = dst
            elif name == "ptr":
                collected[name] = None
            elif name == "size_i":
                collected[name] = sizes[1]
            elif name == "size_j":
                collected[name] = sizes[0]
            elif name == "edge_index":
                collected[name] = edge_index
            else:
                collected[name] = kwargs.get(name)
        msg = self.message(**collected)
        out = scatter_sum(msg, dst, sizes[1])
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


class GCNConv(MessagePassing):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__(aggr="add", node_dim=0)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        self.bias = Parameter(torch.Tensor(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.weight)
        if self.bias is not None:
            self.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight=None):
        edge_index, norm = gcn_norm(edge_index, x.size(0), x.dtype)
        x = torch.matmul(x, self.weight)
        out = self.propagate(edge_index, x=x, edge_weight=norm)
        if self.bias is not None:
            out += self.bias
        return out

    def message(self, x_j, edge_weight):
        return edge_weight.view(-1, 1) * x_j


class GINEConv(MessagePassing):
    def __init__(self, nn, eps=0.0, train_eps=False):
        super().__init__(aggr="add", node_dim=0)
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))
        self.reset_parameters()

    def reset_parameters(self):
        from .inits import reset
        reset(self.nn)
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, edge_index, edge_attr=None):
        if torch.is_tensor(x):
            x = (x, x)
        assert x[0].size(-1) == edge_attr.size(-1)
        out = self.propagate(edge_index, x=x, edge_attr=edge_attr)
        if x[1] is not None:
            out += (1 + self.eps) * x[1]
        return self.nn(out)

    def message(self, x_j, edge_attr):
        return torch.relu(x_j + edge_attr)

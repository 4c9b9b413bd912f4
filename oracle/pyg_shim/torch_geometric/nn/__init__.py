"""torch_geometric.nn subset."""
import torch

from . import inits  # noqa: F401
from . import conv  # noqa: F401
from .conv import GCNConv, GINEConv, MessagePassing  # noqa: F401


class MetaLayer(torch.nn.Module):
    """edge model first (on x[row], x[col]), then node model on the UPDATED edge_attr."""

    def __init__(self, edge_model=None, node_model=None, global_model=None):
        super().__init__()
        self.edge_model, self.node_model, self.global_model = edge_model, node_model, global_model
        self.reset_parameters()

    def reset_parameters(self):
        for item in (self.node_model, self.edge_model, self.global_model):
            if hasattr(item, "reset_parameters"):
                item.reset_parameters()

    def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
        row, col = edge_index[0], edge_index[1]
        if self.edge_model is not None:
            edge_attr = self.edge_model(x[row], x[col], edge_attr, u,
                                        batch if batch is None else batch[row])
        if self.node_model is not None:
            x = self.node_model(x, edge_index, edge_attr, u, batch)
        if self.global_model is not None:
            u = self.global_model(x, edge_index, edge_attr, u, batch)
        return x, edge_attr, u

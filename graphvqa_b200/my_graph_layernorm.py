"""Per-graph LayerNorm with the reference surface (graph_utils/my_graph_layernorm.py:11-81).

``LayerNorm(in_channels, eps=1e-5, affine=True).forward(x, batch=None)``: statistics over all
nodes x channels of each graph, two-pass variance, eps added to the standard deviation, and --
faithfully -- SCALAR affine parameters: the reference builds them with ``torch.Tensor([in_channels])``
(my_graph_layernorm.py:40-41), i.e. shape [1].  Runs ``gvqa_graph_layernorm_f32`` (one CTA per
graph, single HBM read).  The reference's ``int(batch.max()) + 1`` host sync (:59) is avoided
when the caller passes ``num_graphs`` / a GraphCSR.
"""
import torch
from torch import nn

from . import _cabi


class LayerNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        if affine:
            self.weight = nn.Parameter(torch.ones(1))
            self.bias = nn.Parameter(torch.zeros(1))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def reset_parameters(self):
        if self.weight is not None:
            with torch.no_grad():
                self.weight.fill_(1.0)
                self.bias.fill_(0.0)

    def forward(self, x, batch=None, num_graphs=None, csr=None):
        _cabi.require_cuda(x, batch)
        if torch.is_grad_enabled() and (x.requires_grad or (self.weight is not None and self.training)):
            raise NotImplementedError("LayerNorm: the B200 engine is inference-only; use torch.no_grad()/eval()")
        x = x.contiguous().float()
        n = x.size(0)
        if csr is not None:
            graph_ptr, b, hint = csr.graph_ptr, csr.num_graphs, csr.max_nodes_per_graph
        elif batch is None:
            graph_ptr, b, hint = torch.tensor([0, n], dtype=torch.int32, device=x.device), 1, n
        else:
            if num_graphs is None:
                num_graphs = int(batch.max()) + 1      # same host sync as the reference (:59)
            b, hint = num_graphs, 0
            counts = torch.bincount(batch, minlength=b)
            graph_ptr = torch.zeros(b + 1, dtype=torch.int32, device=x.device)
            graph_ptr[1:] = counts.cumsum(0).to(torch.int32)
        return _cabi.graph_layernorm(x, graph_ptr, b, None if self.weight is None else self.weight.detach(),
                                     None if self.bias is None else self.bias.detach(), self.eps,
                                     max_nodes_per_graph=hint)

    def __repr__(self):
        return f"{type(self).__name__}({self.in_channels})"

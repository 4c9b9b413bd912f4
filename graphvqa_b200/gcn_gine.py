"""GCN / GINE variants behind the reference surface
(baseline_and_test_models/pipeline_model_gcn.py:622-669, pipeline_model_gine.py:622-674).

``GCNConv`` / ``GINEConv`` restate the torch_geometric operators the reference instantiates
 with the same ``state_dict`` keys (``weight [in,out]`` / ``bias``;
``nn.0.*``, ``nn.2.*``, ``eps``) and run their message passing through ``gvqa_gcn_aggregate_f32`` /
``gvqa_gine_aggregate_f32``; the dense halves (x @ W, the GINE MLP) are library GEMMs.

``gcn_seq`` / ``gine_seq`` reproduce the reference's behaviour bit for bit in structure: the conv
result is computed and DISCARDED (the reference never assigns ``conv_res`` to ``h``,
pipeline_model_gcn.py:660-668, pipeline_model_gine.py:665-673), so the output is ``x`` pushed
through 4 x (BatchNorm1d -> ReLU -> Dropout).  With ``bug_faithful=True`` (default) the dead
convolutions are skipped unless ``return_conv=True`` asks for them; ``bug_faithful=False`` feeds
``h = conv_res`` forward (what the authors presumably intended).
"""
import torch
from torch import nn

from . import _cabi
from .gat_skip import _glorot_, _require_inference, _strict_fp32_matmul
from .graph_batch import GraphCSR
from .tc_linear import TensorCoreLinear


class GCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self._lin = TensorCoreLinear()
        self.reset_parameters()

    def reset_parameters(self):
        _glorot_(self.weight)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x, edge_index, csr=None, ins=None, dinv=None):
        """x [N, in] (or x [N, F] with ``ins`` [B, in-F]: the split form of cat([x, ins[batch]]))."""
        _require_inference(self, x)
        _cabi.require_cuda(x, edge_index)
        n = x.size(0)
        if csr is None:
            csr = GraphCSR.build(edge_index, torch.zeros(n, dtype=torch.int64, device=x.device), 1)
        d = csr.as_dict()
        if dinv is None:
            dinv = _cabi.gcn_degree(d, n, x.device)
        x = x.contiguous().float()
        f = x.size(1)
        xw = self._lin(x, self.weight[:f].t())                          # x @ W[:F]   (weight is [in, out])
        graph_term = None if ins is None else self._lin(ins.contiguous().float(), self.weight[f:].t())
        return _cabi.gcn_aggregate(xw, graph_term, dinv, self.bias, d)


class GINEConv(nn.Module):
    def __init__(self, nn_module, eps=0.0, train_eps=False):
        super().__init__()
        self.nn = nn_module
        self._lin = TensorCoreLinear()
        self.initial_eps = eps
        if train_eps:
            self.eps = nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))
        self._eps_host = float(eps)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._eps_host = float(self.eps.detach().cpu())

    def forward(self, x, edge_index, edge_attr, csr=None, ins=None):
        """x [N, in], edge_attr [E, in]  (or the split form: x [N,F], edge_attr [E,F], ins [B,D])."""
        _require_inference(self, x, edge_attr)
        _cabi.require_cuda(x, edge_index, edge_attr)
        n = x.size(0)
        if csr is None:
            csr = GraphCSR.build(edge_index, torch.zeros(n, dtype=torch.int64, device=x.device), 1)
        if ins is None:
            assert x.size(-1) == edge_attr.size(-1)
        z = _cabi.gine_aggregate(x.contiguous().float(), edge_attr.contiguous().float(),
                                 None if ins is None else ins.contiguous().float(), csr.as_dict(), self._eps_host)
        return self._apply_nn(z)

    def _apply_nn(self, z):
        """self.nn(z) with its Linear layers on the tensor-core GEMM (any other layer runs as is)."""
        for layer in self.nn:
            if isinstance(layer, nn.Linear):
                z = self._lin(z, layer.weight, layer.bias)
            else:
                with _strict_fp32_matmul():
                    z = layer(z)
        return z


class _seq_base(nn.Module):
    _folded = None

    def _bn_relu(self, h, i):
        """eval-mode BatchNorm1d + ReLU (+ Dropout = identity) as ONE kernel: the statistics fold into a per-channel
        scale / shift (cached per parameter version), ``gvqa_affine_relu_f32`` applies them."""
        bn = self.bns[i]
        key = tuple((t.data_ptr(), t._version) for t in (bn.running_mean, bn.running_var, bn.weight, bn.bias)
                    if t is not None)
        if self._folded is None:
            self._folded = {}
        hit = self._folded.get(i)
        if hit is None or hit[0] != key:
            inv = torch.rsqrt(bn.running_var.detach().double() + bn.eps)
            g = bn.weight.detach().double() if bn.affine else torch.ones_like(inv)
            b = bn.bias.detach().double() if bn.affine else torch.zeros_like(inv)
            hit = self._folded[i] = (key, (g * inv).float().contiguous(),
                                     (b - bn.running_mean.detach().double() * g * inv).float().contiguous())
        return _cabi.affine_relu(h.contiguous(), hit[1], hit[2], relu=True)


class gcn_seq(_seq_base):
    def __init__(self, in_channels, out_channels, ins_dim, dropout=0.0, bug_faithful=True):
        super().__init__()
        self.convs = nn.ModuleList([GCNConv(in_channels + ins_dim, out_channels) for _ in range(5)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(4)])
        self.dropout, self.bug_faithful = dropout, bug_faithful

    def forward(self, x, edge_index, instr_vectors, batch, csr=None, return_conv=False):
        _require_inference(self, x, instr_vectors)
        _cabi.require_cuda(x, edge_index, instr_vectors, batch)
        need_conv = return_conv or not self.bug_faithful
        dinv = None
        if need_conv:
            if csr is None:
                csr = GraphCSR.build(edge_index, batch, instr_vectors.size(1))
            dinv = _cabi.gcn_degree(csr.as_dict(), x.size(0), x.device)
        h, conv_out = x.contiguous().float(), []
        for i, conv in enumerate(self.convs):
            if need_conv:
                conv_res = conv(h, edge_index, csr=csr, ins=instr_vectors[i], dinv=dinv)
                conv_out.append(conv_res)
                if not self.bug_faithful:
                    h = conv_res
            if i != 4:
                h = self._bn_relu(h, i)
        return (h, conv_out) if return_conv else h


class gine_seq(_seq_base):
    def __init__(self, in_channels, out_channels, ins_dim, dropout=0.0, bug_faithful=True):
        super().__init__()
        self.convs = nn.ModuleList([
            GINEConv(nn.Sequential(nn.Linear(in_channels + ins_dim, out_channels), nn.ReLU(),
                                   nn.Linear(out_channels, out_channels))) for _ in range(5)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(4)])
        self.dropout, self.bug_faithful = dropout, bug_faithful

    def forward(self, x, edge_index, edge_attr, instr_vectors, batch, csr=None, return_conv=False):
        _require_inference(self, x, edge_attr, instr_vectors)
        _cabi.require_cuda(x, edge_index, edge_attr, instr_vectors, batch)
        need_conv = return_conv or not self.bug_faithful
        if need_conv and csr is None:
            csr = GraphCSR.build(edge_index, batch, instr_vectors.size(1))
        h, conv_out = x.contiguous().float(), []
        for i, conv in enumerate(self.convs):
            if need_conv:
                conv_res = conv(h, edge_index, edge_attr, csr=csr, ins=instr_vectors[i])
                conv_out.append(conv_res)
                if not self.bug_faithful:
                    h = conv_res
            if i != 4:
                h = self._bn_relu(h, i)
        return (h, conv_out) if return_conv else h

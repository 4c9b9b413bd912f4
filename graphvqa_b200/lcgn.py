"""LCGN variant behind the reference surface (baseline_and_test_models/lcgn.py).

``gat_lcgn`` (lcgn.py:17-244) and ``lcgn_seq`` (:251-323) keep the reference's constructor
signatures, attribute names and ``state_dict`` keys.  Per iteration the three node projections
``lin_l``, ``lin_r`` and the hoisted ``cal_x`` run as ONE library GEMM ([N,3C_in] x [3C_in,3C]);
the reference applies ``cal_x`` per EDGE inside ``message()`` (:230) and gathers the command with a
dense one-hot matmul (:150-153) -- both become index gathers inside ``gvqa_lcgn_hop_f32``, which
fuses logits, LeakyReLU, per-destination softmax, aggregation, ``* cal_cmd`` and bias.

``lcgn_seq.forward`` draws ``x_ctx = torch.randn(...)`` on the CPU at every call exactly like the
reference (:306, even in eval mode); pass ``x_ctx_init`` to inject it (parity tests do).
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _cabi
from .gat_skip import _glorot_, _require_inference, _strict_fp32_matmul
from .graph_batch import GraphCSR
from .tc_linear import TensorCoreLinear


class gat_lcgn(nn.Module):
    def __init__(self, in_channels, out_channels, edge_in_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, cmd_dim=512, add_self_loops=True, bias=True, **kwargs):
        super().__init__()
        if concat or heads != 1:
            raise NotImplementedError("GraphVQA instantiates gat_lcgn with heads=1, concat=False (lcgn.py:272-273)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops = add_self_loops
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.lin_r = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.cal_x = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.proj_cmd = nn.Linear(cmd_dim, heads * out_channels, bias=False)
        self.cal_cmd = nn.Linear(cmd_dim, heads * out_channels, bias=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self._packed = None
        self._lin = TensorCoreLinear()
        self.reset_parameters()

    def reset_parameters(self):      # same order as lcgn.py:106-117
        _glorot_(self.lin_l.weight)
        _glorot_(self.lin_r.weight)
        _glorot_(self.proj_cmd.weight)
        _glorot_(self.cal_cmd.weight)
        _glorot_(self.cal_x.weight)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def _weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._packed[0] != key:
            w_node = torch.cat([self.lin_l.weight, self.lin_r.weight, self.cal_x.weight]).detach().contiguous()   # [3C, in]
            w_cmd = torch.cat([self.proj_cmd.weight, self.cal_cmd.weight]).detach().contiguous()                 # [2C, cmd]
            self._packed = (key, w_node, w_cmd)
        return self._packed[1], self._packed[2]

    def forward(self, x, edge_index, cmd, batch, edge_attr=None, size=None, return_attention_weights=None,
                csr=None):
        _require_inference(self, x, cmd)
        _cabi.require_cuda(x, edge_index, cmd, batch)
        if return_attention_weights is not None:
            raise NotImplementedError("gat_lcgn: attention weights are not exported by the fused kernel")
        c = self.out_channels
        if csr is None:
            csr = GraphCSR.build(edge_index, batch, cmd.size(0))
        w_node, w_cmd = self._weights()
        proj = self._lin(x.contiguous().float(), w_node)             # [N, 3C] = lin_l | lin_r | cal_x  (tensor cores)
        cmds = self._lin(cmd.contiguous().float(), w_cmd)            # [B, 2C] = proj_cmd | cal_cmd
        return _cabi.lcgn_hop(proj[:, :c], proj[:, c:2 * c], proj[:, 2 * c:], cmds[:, :c].contiguous(),
                              cmds[:, c:].contiguous(), self.bias, csr.as_dict(), self.negative_slope)


class lcgn_seq(nn.Module):
    def __init__(self, in_channels, out_channels, edge_attr_dim, num_ins, gat_cmd_dim=512, question_dim=512,
                 MAX_ITER_NUM=4, dropout=0.0, gat_heads=1, gat_negative_slope=0.2, gat_bias=True):
        super().__init__()
        self.init_sg_emb_input = nn.Sequential(nn.Linear(in_channels, out_channels), nn.Dropout(dropout))
        self.MAX_ITER_NUM = MAX_ITER_NUM
        self.qInput1 = nn.Linear(question_dim, out_channels)
        for t in range(MAX_ITER_NUM):
            setattr(self, "qInput2_%d" % t, nn.Linear(out_channels, out_channels))
        self.cmd_inter2logits = nn.Linear(out_channels, 1)
        self.dropout = dropout
        self.proj_x_loc = nn.Sequential(nn.Dropout(dropout), nn.Linear(out_channels, out_channels))
        self.proj_x_ctx = nn.Sequential(nn.Dropout(dropout), nn.Linear(out_channels, out_channels))
        self.output_layer = nn.Linear(2 * out_channels, out_channels)
        self.fin_layer = nn.Linear(2 * out_channels, out_channels)
        self.lcgn = gat_lcgn(in_channels=3 * out_channels, out_channels=out_channels, edge_in_channels=1,
                             heads=gat_heads, concat=False, negative_slope=gat_negative_slope, dropout=dropout,
                             bias=gat_bias, cmd_dim=gat_cmd_dim)
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(num_ins - 1)])   # unused (:283)

    def extract_textual_command(self, q_emb, lstm_outputs, t):
        """lcgn.py:292-300: attention over the question tokens (no padding mask, like the reference)."""
        seq = lstm_outputs.transpose(1, 0)
        q_cmd = getattr(self, "qInput2_%d" % t)(q_emb)
        att = F.softmax(self.cmd_inter2logits(q_cmd[:, None, :] * seq).squeeze(-1), dim=-1)
        return torch.bmm(att[:, None, :], seq).squeeze(1)

    def forward(self, x, edge_index, batch, q_encoding, lstm_outputs, edge_attr=None, instr_vectors=None,
                x_ctx_init=None, csr=None):
        _require_inference(self, x, q_encoding, lstm_outputs)
        _cabi.require_cuda(x, edge_index, batch, q_encoding, lstm_outputs)
        if csr is None:
            csr = GraphCSR.build(edge_index, batch, q_encoding.size(0))
        lin = self.lcgn._lin       # node-level Linear layers on the tensor-core GEMM (dropout is identity in eval)

        def L(layer, v):
            return lin(v, layer.weight, layer.bias)
        with _strict_fp32_matmul():
            x_loc = L(self.init_sg_emb_input[0], x.contiguous().float())
            x_ctx = (torch.randn(x_loc.size()).to(x_loc.device) if x_ctx_init is None     # lcgn.py:306
                     else x_ctx_init.to(x_loc.device))
            q_emb = F.relu(self.qInput1(q_encoding))
            proj_x_loc = L(self.proj_x_loc[1], x_loc)
            for t in range(self.MAX_ITER_NUM):
                cmd = self.extract_textual_command(q_emb, lstm_outputs, t)
                x_joint = torch.cat([x_loc, x_ctx, L(self.proj_x_ctx[1], x_ctx) * proj_x_loc], dim=-1)
                msg = self.lcgn(x_joint, edge_index, cmd, batch, csr=csr)
                x_ctx = L(self.output_layer, torch.cat([x_ctx, msg], dim=-1))
            return L(self.fin_layer, torch.cat([x_loc, x_ctx], dim=-1))

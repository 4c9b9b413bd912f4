"""CPU restatement (PyTorch fp32, dense ops only) of the GraphVQA message-passing hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py; parity status: third-party primitives unpinned,
reference-own code pinned through tests/golden/).  It keeps the reference's *materialising*
dataflow -- cat -> Linear -> index_select -> segment softmax -> index_add -- so that it is
also an honest CPU baseline for the PyG path, and it keeps the reference's module / parameter
names so one ``state_dict`` loads into the reference, this oracle and the CUDA engine alike.

Every class cites the reference lines it follows (paths relative to the reference root).
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import pyg_semantics as pyg


# --------------------------------------------------------------------------------------
# GAT with edge features and skip connection            (reference: gat_skip.py)
# --------------------------------------------------------------------------------------
def gat_conv(x, edge_index, edge_attr, w_l, w_e, att_l, att_r, att_e, bias, heads,
             negative_slope=0.2):
    """gat.forward + gat.message (gat_skip.py:111-208) for a single tensor ``x`` and
    concat=False.  Returns (out [N,C], alpha [E,H]) with alpha in the caller's edge order.

    x [N,Fin], edge_attr [E,Fe], w_l [H*C,Fin] (lin_l is lin_r, gat_skip.py:76-77),
    w_e [H*C,Fe], att_* [1,H,C], bias [C] or None.
    """
    n = x.size(0)
    h = heads
    c = w_l.size(0) // h
    src, dst = edge_index[0], edge_index[1]
    x_l = F.linear(x, w_l).view(n, h, c)                                   # :133
    a_l = (x_l * att_l).sum(dim=-1)                                        # :134
    a_r = (x_l * att_r).sum(dim=-1)                                        # :135
    a_e = (F.linear(edge_attr, w_e).view(-1, h, c) * att_e).sum(dim=-1)    # :150-151
    # MessagePassing.__collect__: *_j gathers at the source, *_i at the target (Appendix A)
    logits = a_l.index_select(0, src) + a_r.index_select(0, dst)           # :183
    logits = logits + a_e                                                  # :185
    logits = F.leaky_relu(logits, negative_slope)                          # :187
    alpha = pyg.segment_softmax(logits, dst, n)                            # :188
    msg = x_l.index_select(0, src) * alpha.unsqueeze(-1)                   # :208
    out = pyg.scatter_sum(msg, dst, n)                                     # aggregate, aggr='add'
    out = out.mean(dim=1)                                                  # :165-166 (concat=False)
    if bias is not None:
        out = out + bias                                                   # :168
    return out, alpha


class gat(nn.Module):
    """Parameter container + forward with the reference's constructor (gat_skip.py:60-109)."""

    def __init__(self, in_channels, out_channels, edge_in_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0.0, add_self_loops=True, bias=True):
        super().__init__()
        if concat:
            raise NotImplementedError("the reference only instantiates concat=False (gat_skip.py:231)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.dropout = negative_slope, dropout
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.lin_r = self.lin_l
        self.lin_e = nn.Linear(edge_in_channels, heads * out_channels, bias=False)
        self.att_e = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_l = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_r = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # gat_skip.py:101-108 (same call order -> same RNG stream)
        pyg.glorot_(self.lin_l.weight)
        pyg.glorot_(self.lin_r.weight)
        pyg.glorot_(self.lin_e.weight)
        pyg.glorot_(self.att_l)
        pyg.glorot_(self.att_r)
        pyg.glorot_(self.att_e)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, edge_index, edge_attr, size=None, return_attention_weights=None):
        if self.training and self.dropout > 0:
            raise NotImplementedError("oracle restates the eval-mode path only")
        out, alpha = gat_conv(x, edge_index, edge_attr, self.lin_l.weight, self.lin_e.weight,
                              self.att_l, self.att_r, self.att_e, self.bias, self.heads,
                              self.negative_slope)
        if isinstance(return_attention_weights, bool):
            return out, (edge_index, alpha)
        return out


class gat_seq(nn.Module):
    """gat_skip.py:220-279: num_ins hops of [cat ins -> gat -> skip -> BN -> ReLU -> dropout]."""

    def __init__(self, in_channels, out_channels, edge_attr_dim, ins_dim, num_ins,
                 dropout=0.0, gat_heads=4, gat_negative_slope=0.2, gat_bias=True):
        super().__init__()
        self.convs = nn.ModuleList([
            gat(in_channels + ins_dim, out_channels, edge_attr_dim + ins_dim, heads=gat_heads,
                concat=False, negative_slope=gat_negative_slope, dropout=dropout, bias=gat_bias)
            for _ in range(num_ins)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(num_ins - 1)])
        self.dropout = dropout

    def forward(self, x, edge_index, edge_attr, instr_vectors, batch, return_hops=False, interleaved_ln=None):
        """``interleaved_ln`` (a LayerNorm of this file, default None = the reference's behaviour): apply that per-graph
        LayerNorm between the hops INSTEAD of BatchNorm1d+ReLU -- the north star's "interleaved with
        my_graph_layernorm" option, which the reference itself does not use."""
        h = x
        hops = []
        last = len(self.convs) - 1
        for i, conv in enumerate(self.convs):
            ins = instr_vectors[i]                                          # :256
            edge_cat = torch.cat((edge_attr, ins[batch[edge_index[0]]]), dim=-1)   # :257-260
            x_cat = torch.cat((h, ins[batch]), dim=-1)                      # :263-264
            h = conv(x_cat, edge_index, edge_cat) + h                       # :269-270
            if i != last and interleaved_ln is not None:
                h = graph_layernorm(h, batch, instr_vectors.size(1), interleaved_ln.weight, interleaved_ln.bias,
                                    interleaved_ln.eps)
            elif i != last:                                                 # :273-276
                h = F.dropout(F.relu(self.bns[i](h)), p=self.dropout, training=self.training)
            hops.append(h)
        return (h, hops) if return_hops else h


# --------------------------------------------------------------------------------------
# Per-graph LayerNorm                       (reference: graph_utils/my_graph_layernorm.py)
# --------------------------------------------------------------------------------------
def graph_layernorm(x, batch, num_graphs, weight=None, bias=None, eps=1e-5):
    """my_graph_layernorm.py:52-78.  Statistics over all nodes x channels of each graph,
    two-pass variance, eps added to the *std*; weight/bias are shape-[1] scalars (:40-41)."""
    if batch is None:
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + eps)
    else:
        norm = pyg.degree(batch, num_graphs, x.dtype).clamp_(min=1)
        norm = norm.mul_(x.size(-1)).view(-1, 1)
        mean = pyg.scatter_sum(x, batch, num_graphs).sum(dim=-1, keepdim=True) / norm
        x = x - mean[batch]
        var = pyg.scatter_sum(x * x, batch, num_graphs).sum(dim=-1, keepdim=True) / norm
        out = x / (var.sqrt()[batch] + eps)
    if weight is not None and bias is not None:
        out = out * weight + bias
    return out


class LayerNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        if affine:  # the reference builds them from the *list* [in_channels] -> shape [1]
            self.weight = nn.Parameter(torch.ones(1))
            self.bias = nn.Parameter(torch.zeros(1))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x, batch=None):
        num_graphs = None if batch is None else int(batch.max()) + 1
        return graph_layernorm(x, batch, num_graphs, self.weight, self.bias, self.eps)


# --------------------------------------------------------------------------------------
# GCN / GINE variants     (reference: baseline_and_test_models/pipeline_model_{gcn,gine}.py)
# --------------------------------------------------------------------------------------
def gcn_conv(x, edge_index, weight, bias):
    """PyG GCNConv defaults (Appendix A of SURVEY.md); weight is [in, out]."""
    ei, norm = pyg.gcn_norm(edge_index, x.size(0), x.dtype)
    xw = torch.matmul(x, weight)
    out = pyg.scatter_sum(norm.view(-1, 1) * xw.index_select(0, ei[0]), ei[1], x.size(0))
    return out if bias is None else out + bias


def gine_aggregate(x, edge_index, edge_attr, eps=0.0):
    """The message-passing half of PyG GINEConv: (1+eps)*x_i + sum_k relu(x_src(k) + e_k)."""
    msg = torch.relu(x.index_select(0, edge_index[0]) + edge_attr)
    return pyg.scatter_sum(msg, edge_index[1], x.size(0)) + (1.0 + eps) * x


class GCNConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        pyg.glorot_(self.weight)

    def forward(self, x, edge_index):
        return gcn_conv(x, edge_index, self.weight, self.bias)


class GINEConv(nn.Module):
    def __init__(self, mlp, eps=0.0):
        super().__init__()
        self.nn = mlp
        self.register_buffer("eps", torch.Tensor([eps]))

    def forward(self, x, edge_index, edge_attr):
        assert x.size(-1) == edge_attr.size(-1)
        return self.nn(gine_aggregate(x, edge_index, edge_attr, float(self.eps)))


class gcn_seq(nn.Module):
    """pipeline_model_gcn.py:622-669.  NOTE the reference never assigns conv_res to h (:660-668):
    with ``bug_faithful=True`` (default) the output is x pushed through 4x(BN,ReLU,dropout)."""

    def __init__(self, in_channels, out_channels, ins_dim, dropout=0.0, bug_faithful=True):
        super().__init__()
        self.convs = nn.ModuleList([GCNConv(in_channels + ins_dim, out_channels) for _ in range(5)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(4)])
        self.dropout, self.bug_faithful = dropout, bug_faithful

    def forward(self, x, edge_index, instr_vectors, batch, return_conv=False):
        h, conv_out = x, []
        for i, conv in enumerate(self.convs):
            x_cat = torch.cat((h, instr_vectors[i][batch]), dim=-1)
            conv_res = conv(x_cat, edge_index)
            conv_out.append(conv_res)
            if not self.bug_faithful:
                h = conv_res
            if i != 4:
                h = F.dropout(F.relu(self.bns[i](h)), p=self.dropout, training=self.training)
        return (h, conv_out) if return_conv else h


class gine_seq(nn.Module):
    """pipeline_model_gine.py:622-674; same dead-conv behaviour as gcn_seq (:665-673)."""

    def __init__(self, in_channels, out_channels, ins_dim, dropout=0.0, bug_faithful=True):
        super().__init__()
        self.convs = nn.ModuleList([
            GINEConv(nn.Sequential(nn.Linear(in_channels + ins_dim, out_channels), nn.ReLU(),
                                   nn.Linear(out_channels, out_channels))) for _ in range(5)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(4)])
        self.dropout, self.bug_faithful = dropout, bug_faithful

    def forward(self, x, edge_index, edge_attr, instr_vectors, batch, return_conv=False):
        h, conv_out = x, []
        for i, conv in enumerate(self.convs):
            ins = instr_vectors[i]
            edge_cat = torch.cat((edge_attr, ins[batch[edge_index[0]]]), dim=-1)
            x_cat = torch.cat((h, ins[batch]), dim=-1)
            conv_res = conv(x_cat, edge_index, edge_cat)
            conv_out.append(conv_res)
            if not self.bug_faithful:
                h = conv_res
            if i != 4:
                h = F.dropout(F.relu(self.bns[i](h)), p=self.dropout, training=self.training)
        return (h, conv_out) if return_conv else h


# --------------------------------------------------------------------------------------
# LCGN variant                               (reference: baseline_and_test_models/lcgn.py)
# --------------------------------------------------------------------------------------
def lcgn_conv(x, edge_index, cmd, batch, w_l, w_r, w_x, w_pc, w_cc, bias, heads=1,
              negative_slope=0.2):
    """gat_lcgn.forward + message (lcgn.py:120-238), concat=False.

    logits_k = sum_c lin_l(x)[src] * (proj_cmd(cmd)[g] * lin_r(x))[dst]  (:154-158, :209)
    msg_k    = cal_x(x[src]) * cal_cmd(cmd)[g(src)] * alpha_k             (:229-238)
    """
    n, h = x.size(0), heads
    c = w_l.size(0) // h
    src, dst = edge_index[0], edge_index[1]
    x_l = F.linear(x, w_l).view(n, h, c)
    x_r = F.linear(x, w_r).view(n, h, c)
    onehot = F.one_hot(batch).to(x.dtype)                                   # :150
    proj_cmd = onehot.matmul(F.linear(cmd, w_pc)).view(n, h, c)             # :152
    cal_cmd = onehot.matmul(F.linear(cmd, w_cc)).view(n, h, c)              # :153
    x_mul = proj_cmd * x_r                                                  # :154
    logits = (x_l.index_select(0, src) * x_mul.index_select(0, dst)).sum(dim=-1)   # :209
    logits = F.leaky_relu(logits, negative_slope)
    alpha = pyg.segment_softmax(logits, dst, n)
    x_val = F.linear(x.index_select(0, src), w_x).view(-1, h, c)            # :230 (per-edge GEMM)
    msg = x_val * cal_cmd.index_select(0, src) * alpha.unsqueeze(-1)        # :231, :238
    out = pyg.scatter_sum(msg, dst, n).mean(dim=1)
    return out if bias is None else out + bias


class gat_lcgn(nn.Module):
    def __init__(self, in_channels, out_channels, edge_in_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0.0, cmd_dim=512, add_self_loops=True, bias=True):
        super().__init__()
        if concat:
            raise NotImplementedError("the reference only instantiates concat=False (lcgn.py:273)")
        self.heads, self.out_channels, self.negative_slope = heads, out_channels, negative_slope
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.lin_r = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.cal_x = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.proj_cmd = nn.Linear(cmd_dim, heads * out_channels, bias=False)
        self.cal_cmd = nn.Linear(cmd_dim, heads * out_channels, bias=False)
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # lcgn.py:106-117
        pyg.glorot_(self.lin_l.weight)
        pyg.glorot_(self.lin_r.weight)
        pyg.glorot_(self.proj_cmd.weight)
        pyg.glorot_(self.cal_cmd.weight)
        pyg.glorot_(self.cal_x.weight)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, edge_index, cmd, batch, edge_attr=None):
        return lcgn_conv(x, edge_index, cmd, batch, self.lin_l.weight, self.lin_r.weight,
                         self.cal_x.weight, self.proj_cmd.weight, self.cal_cmd.weight, self.bias,
                         self.heads, self.negative_slope)


class lcgn_seq(nn.Module):
    """lcgn.py:251-323.  ``x_ctx_init`` injects the tensor the reference draws with
    torch.randn on every forward (:306); None reproduces the draw from the global CPU RNG."""

    def __init__(self, in_channels, out_channels, edge_attr_dim, num_ins, gat_cmd_dim=512,
                 question_dim=512, MAX_ITER_NUM=4, dropout=0.0, gat_heads=1,
                 gat_negative_slope=0.2, gat_bias=True):
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
        self.lcgn = gat_lcgn(3 * out_channels, out_channels, edge_in_channels=1, heads=gat_heads,
                             concat=False, negative_slope=gat_negative_slope, dropout=dropout,
                             bias=gat_bias, cmd_dim=gat_cmd_dim)
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(num_ins - 1)])

    def extract_textual_command(self, q_emb, lstm_outputs, t):   # :292-300 (no padding mask)
        seq = lstm_outputs.transpose(1, 0)
        q_cmd = getattr(self, "qInput2_%d" % t)(q_emb)
        raw_att = self.cmd_inter2logits(q_cmd[:, None, :] * seq).squeeze(-1)
        att = F.softmax(raw_att, dim=-1)
        return torch.bmm(att[:, None, :], seq).squeeze(1)

    def forward(self, x, edge_index, batch, q_encoding, lstm_outputs, edge_attr=None,
                instr_vectors=None, x_ctx_init=None):
        x_loc = self.init_sg_emb_input(x)
        x_ctx = torch.randn(x_loc.size()).to(x_loc.device) if x_ctx_init is None else x_ctx_init
        q_emb = F.relu(self.qInput1(q_encoding))
        proj_x_loc = self.proj_x_loc(x_loc)
        for t in range(self.MAX_ITER_NUM):
            cmd = self.extract_textual_command(q_emb, lstm_outputs, t)
            x_joint = torch.cat([x_loc, x_ctx, self.proj_x_ctx(x_ctx) * proj_x_loc], dim=-1)
            msg = self.lcgn(x_joint, edge_index=edge_index, cmd=cmd, batch=batch)
            x_ctx = self.output_layer(torch.cat([x_ctx, msg], dim=-1))
        return self.fin_layer(torch.cat([x_loc, x_ctx], dim=-1))


# --------------------------------------------------------------------------------------
# Scene-graph encoder (the step before the hop stack)      (reference: pipeline_model_gat.py)
# --------------------------------------------------------------------------------------
class _EdgeModel(nn.Module):
    """get_gt_scene_graph_encoding_layer.EdgeModel (pipeline_model_gat.py:65-77)."""

    def __init__(self, nf, ef):
        super().__init__()
        self.edge_mlp = nn.Sequential(nn.Linear(2 * nf + ef, ef), nn.ReLU(), nn.Linear(ef, ef))

    def forward(self, src, dest, edge_attr, u=None, batch=None):
        return self.edge_mlp(torch.cat([src, dest, edge_attr], 1))                 # :76-77


class _NodeModel(nn.Module):
    """get_gt_scene_graph_encoding_layer.NodeModel (pipeline_model_gat.py:79-98)."""

    def __init__(self, nf, ef):
        super().__init__()
        self.node_mlp_1 = nn.Sequential(nn.Linear(nf + ef, nf), nn.ReLU(), nn.Linear(nf, nf))
        self.node_mlp_2 = nn.Sequential(nn.Linear(2 * nf, nf), nn.ReLU(), nn.Linear(nf, nf))

    def forward(self, x, edge_index, edge_attr, u=None, batch=None):
        row, col = edge_index[0], edge_index[1]                                    # :93
        out = torch.cat([x[row], edge_attr], dim=1)                                # :94
        out = self.node_mlp_1(out)                                                 # :95
        out = pyg.scatter_mean(out, col, x.size(0))                                # :96 (count clamped to >= 1)
        out = torch.cat([x, out], dim=1)                                           # :97
        return self.node_mlp_2(out)                                                # :98


class MetaLayer(nn.Module):
    """torch_geometric.nn.MetaLayer(EdgeModel(), NodeModel()) (pipeline_model_gat.py:100):
    edges first, then nodes on the UPDATED edge features."""

    def __init__(self, nf, ef):
        super().__init__()
        self.edge_model = _EdgeModel(nf, ef)
        self.node_model = _NodeModel(nf, ef)

    def forward(self, x, edge_index, edge_attr, u=None, batch=None):
        row, col = edge_index[0], edge_index[1]
        edge_attr = self.edge_model(x[row], x[col], edge_attr, u, None if batch is None else batch[row])
        x = self.node_model(x, edge_index, edge_attr, u, batch)
        return x, edge_attr, u


class GroundTruth_SceneGraph_Encoder(nn.Module):
    """pipeline_model_gat.py:553-610 with the vocabulary size / pad id passed in (the reference reads them from
    gqa_dataset_entry class attributes, :556-562).  ``sg_emb_dim`` is 300 in the reference (:560)."""

    def __init__(self, sg_vocab_size, sg_pad_idx=1, sg_emb_dim=300):
        super().__init__()
        self.sg_emb_dim = sg_emb_dim
        self.sg_vocab_embedding = nn.Embedding(sg_vocab_size, sg_emb_dim, padding_idx=sg_pad_idx)
        self.scene_graph_encoding_layer = MetaLayer(sg_emb_dim, sg_emb_dim)
        self.graph_layer_norm = LayerNorm(sg_emb_dim)

    def forward(self, gt_scene_graphs):
        g = gt_scene_graphs
        x_embed_sum = self.sg_vocab_embedding(g.x).sum(dim=-2)                     # :583-585
        edge_attr_embed = self.sg_vocab_embedding(g.edge_attr)                     # :587
        # :590 -- rows `added_sym_edge` of the BATCHED edge array (the indices are graph-local and un-offset)
        sym = getattr(g, "added_sym_edge", None)
        if sym is not None and sym.numel() > 0:
            edge_attr_embed = edge_attr_embed.clone()
            edge_attr_embed[sym, :, :] *= -1
        edge_attr_embed_sum = edge_attr_embed.sum(dim=-2)                          # :594
        x_encoded, edge_attr_encoded, _ = self.scene_graph_encoding_layer(
            x=x_embed_sum, edge_index=g.edge_index, edge_attr=edge_attr_embed_sum, u=None, batch=g.batch)   # :600-606
        x_encoded = self.graph_layer_norm(x_encoded, g.batch)                      # :608
        return x_encoded, edge_attr_encoded, None


# --------------------------------------------------------------------------------------
# Question-conditioned attention pooling (the step after the hop stack)
# --------------------------------------------------------------------------------------
class MyConditionalGlobalAttention(nn.Module):
    """pipeline_model_gat.py:108-185."""

    def __init__(self, num_node_features, num_out_features):
        super().__init__()
        c = num_out_features
        self.gate_nn = nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, 1))                 # :126
        self.node_nn = nn.Sequential(nn.Linear(num_node_features, c), nn.ReLU(), nn.Linear(c, c))  # :127
        self.ques_nn = nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, c))                 # :128

    def forward(self, x, u, batch, size=None):
        x = x.unsqueeze(-1) if x.dim() == 1 else x                                 # :150
        size = int(batch[-1]) + 1 if size is None else size                        # :152
        x = self.node_nn(x)                                                        # :161
        gate = self.gate_nn(self.ques_nn(u)[batch] * x)                            # :171
        assert gate.dim() == x.dim() and gate.size(0) == x.size(0)                 # :172
        gate = pyg.segment_softmax(gate, batch, size)                              # :177
        return pyg.scatter_sum(gate * x, batch, size)                              # :178


class GraphSide(nn.Module):
    """The graph side of PipelineModel.forward (pipeline_model_gat.py:751, 791-816) with the reference's
    sub-module names: scene_graph_encoder -> gat_seq -> graph_global_attention_pooling -> logit_fc.  The text
    side's outputs (``instr_vectors`` [5,B,D], ``questions_encoded[0]`` [B,D]) are inputs."""

    def __init__(self, sg_vocab_size, sg_pad_idx=1, sg_emb_dim=300, question_hidden_dim=512, num_answers=1842):
        super().__init__()
        f, d = sg_emb_dim, question_hidden_dim
        self.scene_graph_encoder = GroundTruth_SceneGraph_Encoder(sg_vocab_size, sg_pad_idx, f)
        self.gat_seq = gat_seq(in_channels=f, out_channels=f, edge_attr_dim=f, ins_dim=d, num_ins=5, dropout=0.1,
                               gat_heads=4, gat_negative_slope=0.2, gat_bias=True)            # :683-687
        self.graph_global_attention_pooling = MyConditionalGlobalAttention(num_node_features=f, num_out_features=d)
        self.logit_fc = nn.Sequential(nn.Dropout(p=0.2), nn.Linear(3 * d, d), nn.ELU(), nn.Dropout(p=0.2),
                                      nn.Linear(d, num_answers))                               # :722-728

    def forward(self, gt_scene_graphs, instr_vectors, q0, return_parts=False):
        g = gt_scene_graphs
        x_encoded, edge_attr_encoded, _ = self.scene_graph_encoder(g)              # :751
        x_executed = self.gat_seq(x=x_encoded, edge_index=g.edge_index, edge_attr=edge_attr_encoded,
                                  instr_vectors=instr_vectors, batch=g.batch)      # :791
        pooled = self.graph_global_attention_pooling(x=x_executed, u=q0, batch=g.batch, size=None)   # :800-805
        logits = self.logit_fc(torch.cat((pooled, q0, pooled * q0), dim=-1))       # :814-816
        if return_parts:
            return dict(x_encoded=x_encoded, edge_attr_encoded=edge_attr_encoded, x_executed=x_executed,
                        pooled=pooled, short_answer_logits=logits)
        return logits

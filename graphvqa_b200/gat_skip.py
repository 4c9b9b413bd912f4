"""B200 engine behind the reference's ``gat_skip`` operator surface.

Mirrors the constructor / ``forward`` signatures, attribute names and ``state_dict`` keys of the
reference's ``gat`` (gat_skip.py:16-213) and ``gat_seq`` (gat_skip.py:220-279), so a checkpoint
trained with the reference loads unchanged, but executes the message passing through the C ABI in
``include/gvqa_b200.h``:

  per batch   gvqa_build_csr (cached on the SceneGraphBatch), edge-logit skinny projection for all
              hops at once (edge_attr is hop-invariant), per-graph instruction terms (2 small bmm);
  per hop     node projection  h @ W_h^T  (gvqa_proj_gemm_3xtf32: tcgen05 split-TF32, fp32-level accuracy),
              gvqa_skinny_matmul_f32 for the collapsed a_l/a_r node logits,
              gvqa_gat_hop_f32: gather + logits + segment softmax + aggregate + head mean + bias +
              skip + BatchNorm(eval) + ReLU in ONE kernel.

The reference's  cat([h, ins[batch]]) @ W^T  is evaluated as  h @ W[:, :F]^T + (ins @ W[:, F:]^T)[batch];
the second, per-graph term is identical for every source row of a graph and the softmax weights
of a destination sum to one, so it is added once per destination as ``graph_bias`` (head mean of
ins @ W[:, F:]^T).  <W x, att_h> is evaluated as x . (W_h^T att_h)  (SURVEY.md section 8a).  Same
mathematics, fp32 rounding differences of order 1e-6.  Inference only: ``forward`` raises in training mode (attention dropout
and BatchNorm batch statistics, gat_skip.py:190/274, are not part of the engine).  CUDA only: there
is no CPU fallback.
"""
import contextlib
import os

import torch
from torch import nn

from . import _cabi
from .graph_batch import GraphCSR


def _glorot_(t):
    """torch_geometric.nn.inits.glorot: U(-a,a), a = sqrt(6/(size(-2)+size(-1)))."""
    bound = (6.0 / (t.size(-2) + t.size(-1))) ** 0.5
    with torch.no_grad():
        t.uniform_(-bound, bound)


@contextlib.contextmanager
def _strict_fp32_matmul():
    """The 1e-4 parity bar rules out TF32 for the projections (hard part 1)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def _collapse(att, weight, heads):
    """V[h,:] = sum_c att[0,h,c] * W[h*C+c,:]  in float64 -> float32  ([H, in])."""
    h = heads
    c = weight.size(0) // h
    v = torch.einsum("hc,hcf->hf", att.detach().double().view(h, c), weight.detach().double().view(h, c, -1))
    return v.float()


def _param_key(module):
    return tuple((p.data_ptr(), p._version) for p in list(module.parameters()) + list(module.buffers()))


def _require_inference(module, *tensors):
    if module.training:
        raise NotImplementedError(
            "%s: the B200 engine implements the eval-mode path only (call .eval()); training through "
            "the fused kernels (autograd, attention dropout, BatchNorm batch statistics) is not built"
            % type(module).__name__)
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError("%s: inputs require grad; run under torch.no_grad()" % type(module).__name__)


class gat(nn.Module):
    """Edge-featured multi-head GAT convolution with the reference signature (gat_skip.py:60-63)."""

    def __init__(self, in_channels, out_channels, edge_in_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0.0, add_self_loops=True, bias=True, **kwargs):
        super().__init__()
        if not isinstance(in_channels, int):
            raise NotImplementedError("bipartite (tuple) in_channels are not used by GraphVQA")
        if concat:
            raise NotImplementedError("concat=True is never instantiated by GraphVQA (gat_skip.py:231)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops = add_self_loops   # stored, never used -- like the reference (:73)
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.lin_r = self.lin_l                 # shared module: both keys appear in the state_dict
        self.lin_e = nn.Linear(edge_in_channels, heads * out_channels, bias=False)
        self.att_e = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_l = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_r = nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self._packed = None
        self._lin = None
        self.reset_parameters()

    def reset_parameters(self):
        # same order as gat_skip.py:101-108 so a seeded construction matches the reference
        _glorot_(self.lin_l.weight)
        _glorot_(self.lin_r.weight)
        _glorot_(self.lin_e.weight)
        _glorot_(self.att_l)
        _glorot_(self.att_r)
        _glorot_(self.att_e)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    # ---- weight prepack (cached until a parameter changes) ---------------------------------
    def packed(self):
        key = _param_key(self)
        if self._packed is None or self._packed["key"] != key:
            h = self.heads
            v_l = _collapse(self.att_l, self.lin_l.weight, h)
            v_r = _collapse(self.att_r, self.lin_l.weight, h)
            v_e = _collapse(self.att_e, self.lin_e.weight, h)
            self._packed = dict(key=key, v_node=torch.cat([v_l, v_r]).contiguous(), v_l=v_l, v_r=v_r,
                                v_edge=v_e.contiguous())
        return self._packed

    def forward(self, x, edge_index, edge_attr, size=None, return_attention_weights=None, csr=None):
        """x [N, in], edge_index [2,E] i64, edge_attr [E, edge_in] -> [N, out]
        (+ (edge_index, alpha[E,H]) when return_attention_weights is a bool, gat_skip.py:170-175)."""
        _require_inference(self, x, edge_attr)
        _cabi.require_cuda(x, edge_attr, edge_index)
        assert x.dim() == 2, "Static graphs not supported in `GATConv`."
        n, e, h, c = x.size(0), edge_index.size(1), self.heads, self.out_channels
        if csr is None:  # stand-alone call: the whole input is one graph
            csr = GraphCSR.build(edge_index, torch.zeros(n, dtype=torch.int64, device=x.device), 1)
        pk = self.packed()
        x = x.contiguous().float()
        edge_attr = edge_attr.contiguous().float()
        if x.size(1) % 4 == 0:      # fp32-accurate tcgen05 GEMM (tf32 split: the full fp32 range, no flag to watch)
            if self._lin is None:
                from .tc_linear import TensorCoreLinear
                self._lin = TensorCoreLinear()
            x_l = self._lin(x, self.lin_l.weight)
        else:                       # widths the TMA descriptors cannot address: strict-fp32 library GEMM
            with _strict_fp32_matmul():
                x_l = torch.mm(x, self.lin_l.weight.t())
        a_node = _cabi.skinny_matmul(x, pk["v_node"])
        a_edge = _cabi.skinny_matmul(edge_attr, pk["v_edge"]) if e > 0 else x.new_zeros(1, h)
        out = torch.empty(n, c, dtype=torch.float32, device=x.device)
        want_alpha = isinstance(return_attention_weights, bool)
        alpha = torch.zeros(e, h, dtype=torch.float32, device=x.device) if want_alpha else None
        _cabi.gat_hop(x_l, a_node, a_edge, csr.as_dict(), h, c, out, bias=self.bias, alpha_out=alpha,
                      negative_slope=self.negative_slope, **csr.hints())
        if want_alpha:
            return out, (edge_index, alpha)
        return out

    def __repr__(self):
        return "{}({}, {}, heads={})".format(type(self).__name__, self.in_channels, self.out_channels, self.heads)


class gat_seq(nn.Module):
    """num_ins hops of [cat instruction -> gat -> skip -> BatchNorm1d -> ReLU -> Dropout]
    (no BN/ReLU after the last hop), reference signature gat_skip.py:224-225, 249."""

    def __init__(self, in_channels, out_channels, edge_attr_dim, ins_dim, num_ins,
                 dropout=0.0, gat_heads=4, gat_negative_slope=0.2, gat_bias=True):
        super().__init__()
        if in_channels != out_channels:
            raise ValueError("the skip connection (gat_skip.py:270) needs in_channels == out_channels")
        self.convs = nn.ModuleList([
            gat(in_channels=in_channels + ins_dim, out_channels=out_channels,
                edge_in_channels=edge_attr_dim + ins_dim, heads=gat_heads, concat=False,
                negative_slope=gat_negative_slope, dropout=dropout, bias=gat_bias)
            for _ in range(num_ins)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(out_channels) for _ in range(num_ins - 1)])
        self.dropout = dropout
        self.in_channels, self.edge_attr_dim, self.ins_dim = in_channels, edge_attr_dim, ins_dim
        self.kernel_variant = _cabi.VARIANT_AUTO
        # per-batch slabs for the hop kernel's one-round-trip prologue (hop variant 5): built once per forward beside
        # the pre-pass (one small launch), consumed by all hops when kernel_variant is AUTO or SLAB
        self.use_slabs = os.environ.get("GVQA_HOP_SLABS", "1") != "0"
        # node projection h @ W_h^T, all with fp32-level accuracy:
        #   "3xf16"  hand-written tcgen05 GEMM on fp16-split operands (default; inputs must fit fp16's range,
        #            guarded by a device flag, see check_overflow)
        #   "3xtf32" the same kernel structure on tf32-split operands (full fp32 range, ~1.4x slower)
        #   "cublas" torch.mm with TF32 off (fp32 SIMT, ~6x slower)
        self.projection = "3xf16"
        # "3xf16" only: hop 0's projection and the two pre-pass products share ONE persistent launch (their short
        # tiles fill the projection's last, partially empty wave); False = three launches
        self.group_prepass = os.environ.get("GVQA_GROUP_PREPASS", "1") != "0"
        # "fused": one tensor-core kernel per hop that aggregates the INPUT rows per head and projects the aggregate
        # (gvqa_gat_fused_hop_f32; x_l [N, H*C] is never materialised); "split": projection GEMM + hop kernel.
        # Shapes the fused kernel does not take (heads 8, ...) and the other projection kinds use "split".
        # An explicit kernel_variant (a hop kernel of the split path) also selects "split".
        self.hop_mode = os.environ.get("GVQA_HOP_MODE", "fused")
        self._overflow, self._overflow_pending = None, []
        self.overflow_external = False   # True: the caller reads / clears the fp16 range flag itself (host runner)
        self._side = None
        # keep x_l (written by the GEMM, read once by the hop kernel) resident in L2 between the two
        self.l2_persist = False
        self.skip_hop_launch = False  # measurement only (bench.py): omit the fused-hop launches, results are garbage
        self.gemm_events = None     # same for the projection GEMM launches
        self.hop_events = None      # set to a list to collect (start, end) CUDA events per fused-hop launch
        self._packed = None
        self.__dict__["_interleaved_ln"] = None
        self._slab_l2_on = None

    def set_interleaved_layernorm(self, layer_norm):
        """Use ``layer_norm`` (a ``my_graph_layernorm.LayerNorm``; None restores the default) instead of
        BatchNorm1d+ReLU between the hops: the per-graph LayerNorm then runs as the fused hop kernel's epilogue
        (GVQA_EPI_GRAPH_LN, one CTA per graph, no extra HBM pass).  The reference interleaves BatchNorm
        (gat_skip.py:272-276) and applies its graph LayerNorm once before hop 0 (pipeline_model_gat.py:608); this is
        the option BASELINE.json's north star asks the kernel to offer.  The module is not registered as a
        sub-module, so the ``state_dict`` stays the reference's."""
        self.__dict__["_interleaved_ln"] = layer_norm

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()

    # ---- weight prepack ---------------------------------------------------------------------
    def packed(self):
        # one pack per projection kind, all kept alive: captured CUDA graphs hold raw pointers into the pack they
        # were recorded with, and the host runner's full-range rerun switches the projection temporarily
        key = _param_key(self)
        if self._packed is None or self._packed.get("key") != key:
            self._packed = {"key": key}
        if self.projection in self._packed:
            return self._packed[self.projection]
        f, fe = self.in_channels, self.edge_attr_dim
        w_h, w_ins, v_node, v_graph, v_edge, scale, shift = [], [], [], [], [], [], []
        for i, conv in enumerate(self.convs):
            pk = conv.packed()
            w = conv.lin_l.weight.detach()
            w_h.append(w[:, :f].contiguous())                     # [HC, F]
            heads, c = conv.heads, conv.out_channels
            w_ins.append(w[:, f:].double().view(heads, c, -1).mean(0).float().t().contiguous())   # [D, C]
            v_node.append(torch.cat([pk["v_l"][:, :f], pk["v_r"][:, :f]]).contiguous())        # [2H, F]
            v_graph.append((pk["v_l"][:, f:].double() + pk["v_r"][:, f:].double()
                            + pk["v_edge"][:, fe:].double()).float().t().contiguous())          # [D, H]
            v_edge.append(pk["v_edge"][:, :fe])                   # [H, Fe]
            if i < len(self.bns):
                bn = self.bns[i]
                inv = torch.rsqrt(bn.running_var.detach().double() + bn.eps)
                g = bn.weight.detach().double() if bn.affine else torch.ones_like(inv)
                b = bn.bias.detach().double() if bn.affine else torch.zeros_like(inv)
                scale.append((g * inv).float().contiguous())
                shift.append((b - bn.running_mean.detach().double() * g * inv).float().contiguous())
        # tcgen05 projection: the collapsed a_l / a_r logit vectors ride along as 16 extra output columns
        # (8 used at H=4) so one GEMM yields x_l and a_node; split into tf32 hi/lo once
        w_split = edge_split = ins_split = None
        v_edge_all = torch.cat(v_edge).contiguous()                      # [hops*H, Fe]
        if w_h[0].is_cuda:
            def pad16(t):
                r = (-t.size(0)) % 16
                return torch.cat([t, t.new_zeros(r, t.size(1))]) if r else t
            split = _cabi.split_f16 if self.projection == "3xf16" else _cabi.split_tf32
            w_split = [split(pad16(torch.cat([w, vn]))) for w, vn in zip(w_h, v_node)]
            # pre-pass operands for the same GEMM: all hops' edge-logit vectors as one [hops*H (+pad), Fe] matrix,
            # and per hop the instruction weights [C + H (+pad), D] = [W_ins_mean^T ; V_graph^T] stacked over hops
            edge_split = split(pad16(v_edge_all))
            ins_rows = [pad16(torch.cat([wi.t(), vg.t()])) for wi, vg in zip(w_ins, v_graph)]
            ins_split = split(torch.cat(ins_rows).contiguous())
            ins_ld = ins_rows[0].size(0)
        w_fused = v0_split = None
        if w_h[0].is_cuda and self.projection == "3xf16" and _cabi.fused_supported(heads, f, c):
            w_fused = [_cabi.fused_pack(w, heads, c, f) for w in w_h]
            v0_split = _cabi.split_f16(pad16(v_node[0]))      # hop 0's node logits ride in the pre-pass GEMM launch
        pack = dict(w_h=w_h, w_split=w_split, w_fused=w_fused, v0_split=v0_split, w_ins=torch.stack(w_ins), v_node=v_node,
                    v_graph=torch.stack(v_graph), v_edge=v_edge_all, edge_split=edge_split,
                    ins_split=ins_split, ins_ld=ins_ld if ins_split is not None else 0,
                    scale=scale, shift=shift)
        self._packed[self.projection] = pack
        return pack

    def fused_path_ready(self):
        """True when forward() will run the one-kernel hops (and therefore accepts ``a_edge_all`` in place of
        ``edge_attr``)."""
        return (self.hop_mode == "fused" and self.projection == "3xf16" and self.kernel_variant == _cabi.VARIANT_AUTO
                and self._interleaved_ln is None and self.gemm_events is None and not self.training
                and self.convs[0].lin_l.weight.is_cuda and self.packed().get("w_fused") is not None)

    def forward(self, x, edge_index, edge_attr, instr_vectors, batch, csr=None, return_hops=False, csr_hints=None,
                a_edge_all=None):
        """x [N,F], edge_index [2,E] i64, edge_attr [E,Fe], instr_vectors [num_ins,B,D], batch [N] i64
        -> h [N,F].  ``csr`` (a GraphCSR) may be passed to reuse the per-batch pre-pass; otherwise it is built
        here, on a side stream, concurrently with the pre-pass GEMMs and the first projection (none of which
        needs the topology).  ``csr_hints``: optional max_nodes_per_graph / max_in_edges_per_graph loader hints."""
        _require_inference(self, x, edge_attr, instr_vectors)
        _cabi.require_cuda(x, edge_index, edge_attr, instr_vectors, batch)
        num_hops = len(self.convs)
        n, e = x.size(0), edge_index.size(1)
        b = instr_vectors.size(1)
        heads, c = self.convs[0].heads, self.convs[0].out_channels
        side = None
        if csr is None:
            cur = torch.cuda.current_stream(x.device)
            side = self._side_stream(x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                csr = GraphCSR.build(edge_index, batch, b, **(csr_hints or {}))
                csr_ready = torch.cuda.Event()
                csr_ready.record(side)
            for t in (csr.rowptr, csr.col_src, csr.perm, csr.graph_ptr, csr.node_graph, csr.stats):
                t.record_stream(cur)
        csr_d = csr.as_dict()
        pk = self.packed()
        x = x.contiguous().float()
        if edge_attr is not None:
            edge_attr = edge_attr.contiguous().float()
        ins = instr_vectors[:num_hops].contiguous().float()

        # hop-invariant pre-pass: all hops' edge logits in one sweep over edge_attr, and the
        # per-graph instruction terms of x_l and of the logits
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing and not self.overflow_external:
            self._poll_overflow()
        tensor_core = self.projection in ("3xtf32", "3xf16")
        if self.projection == "3xf16":
            if self._overflow is None or self._overflow.device != x.device:
                self._overflow = torch.zeros(1, dtype=torch.int32, device=x.device)
            flag = self._overflow

            def gemm(a, split, out=None):
                return _cabi.proj_gemm_3xf16(a, split[0], split[1], out=out, overflow=flag)
        else:
            def gemm(a, split, out=None):
                return _cabi.proj_gemm_3xtf32(a, split[0], split[1], out=out)
        if self.hop_mode == "fused" and pk.get("w_fused") is not None and self._interleaved_ln is None and n > 0 \
                and self.gemm_events is None and self.kernel_variant == _cabi.VARIANT_AUTO:
            return self._forward_fused(x, edge_attr, ins, csr, side, csr_ready if side is not None else None, pk, flag,
                                       return_hops, batch=batch, a_edge_all=a_edge_all)
        hc = heads * c
        fused_logits = tensor_core
        ldx = hc + (-(-2 * heads // 16) * 16 if fused_logits else 0)
        x_l = torch.empty(n, ldx, dtype=torch.float32, device=x.device)
        grouped = self.projection == "3xf16" and self.group_prepass and n > 0 and self.gemm_events is None
        if grouped:
            ld = pk["ins_ld"]
            bh, bl = (t.unflatten(0, (num_hops, ld)) for t in pk["ins_split"])
            problems = [(x, pk["w_split"][0][0], pk["w_split"][0][1], x_l), (ins, bh, bl, None)]
            if e > 0:
                problems.append((edge_attr, pk["edge_split"][0], pk["edge_split"][1], None))
            outs = _cabi.proj_gemm_3xf16_grouped(problems, overflow=flag)
            g_all = outs[1]
            a_edge_all = outs[2] if e > 0 else x.new_zeros(1, num_hops * heads)
            graph_bias_all = [g_all[i, :, :c] for i in range(num_hops)]
            a_graph_all = [g_all[i, :, c:c + heads] for i in range(num_hops)]
        elif e == 0:
            a_edge_all = x.new_zeros(1, num_hops * heads)
        elif tensor_core:
            a_edge_all = gemm(edge_attr, pk["edge_split"])                      # [E, hops*H (+pad)], one sweep
        else:
            a_edge_all = _cabi.skinny_matmul(edge_attr, pk["v_edge"])           # [E, hops*H]
        if grouped:
            pass
        elif tensor_core:
            # one GEMM for every hop's per-graph terms: rows = (hop, graph), columns = (hop', C + H); only the
            # hop == hop' blocks are used (the cross blocks are wasted flops, ~5 us, cheaper than 5 launches)
            ld = pk["ins_ld"]
            if self.projection == "3xf16":      # one batched launch: [hops, B, D] x [hops, C+H(+pad), D]^T
                bh, bl = (t.unflatten(0, (num_hops, ld)) for t in pk["ins_split"])
                g_all = _cabi.proj_gemm_3xf16_batched(ins, bh, bl, overflow=flag)            # [hops, B, ld]
                graph_bias_all = [g_all[i, :, :c] for i in range(num_hops)]      # [B, C] views, row stride ld
                a_graph_all = [g_all[i, :, c:c + heads] for i in range(num_hops)]
            else:                               # one plain GEMM incl. the unused cross blocks (hop != hop')
                g_all = gemm(ins.view(num_hops * b, -1), pk["ins_split"]).view(num_hops, b, num_hops, ld)
                graph_bias_all = [g_all[i, :, i, :c] for i in range(num_hops)]  # row stride hops*ld
                a_graph_all = [g_all[i, :, i, c:c + heads] for i in range(num_hops)]
        else:
            with _strict_fp32_matmul():
                graph_bias_all = torch.bmm(ins, pk["w_ins"])                    # [hops, B, C]
                a_graph_all = torch.bmm(ins, pk["v_graph"])                     # [hops, B, H]

        # per-batch slabs of the one-round-trip hop prologue: built on the side stream, concurrently with hop 0 (which
        # therefore still runs the block kernel) and the next projection; hops >= 1 consume them
        slab_idx = slab_f = slab_ready = None
        if self.use_slabs and n > 0 and num_hops > 1 and self.kernel_variant in (_cabi.VARIANT_AUTO, _cabi.VARIANT_SLAB) \
                and not self.skip_hop_launch:
            if torch.is_tensor(a_graph_all):                    # [hops, B, H] from the bmm
                ag3 = a_graph_all
            elif g_all.dim() == 3:                              # [hops, B, ld]: columns [c, c+H)
                ag3 = g_all[:, :, c:c + heads]
            else:                                               # [hops, B, hops, ld]: the (i, :, i, c:c+H) blocks
                ld4 = g_all.size(3)
                ag3 = torch.as_strided(g_all, (num_hops, b, heads), (b * num_hops * ld4 + ld4, num_hops * ld4, 1),
                                       g_all.storage_offset() + c)
            cur = torch.cuda.current_stream(x.device)
            sstream = self._side_stream(x.device)
            sstream.wait_stream(cur)            # the pre-pass products (and, without a side CSR build, the topology)
            with torch.cuda.stream(sstream):
                slab_idx, slab_f = _cabi.build_hop_slabs(csr_d, a_edge_all, ag3, num_hops, heads, n,
                                                         l2_persist=self._slab_l2(x.device))
                slab_ready = torch.cuda.Event()
                slab_ready.record(sstream)
            for t in (slab_idx, slab_f):
                t.record_stream(cur)
        h = x
        hops = []
        a_node = x_l[:, hc:hc + 2 * heads] if fused_logits else \
            torch.empty(n, 2 * heads, dtype=torch.float32, device=x.device)
        if self.l2_persist:
            _cabi.l2_window(x_l, x.device, 1.0)
        for i in range(num_hops):
            if fused_logits:
                if self.gemm_events is not None:
                    gev = (torch.cuda.Event(enable_timing=True, external=capturing),
                           torch.cuda.Event(enable_timing=True, external=capturing))
                    gev[0].record()
                if not (grouped and i == 0):
                    gemm(h, pk["w_split"][i], out=x_l)
                if self.gemm_events is not None:
                    gev[1].record()
                    self.gemm_events.append(gev)
            else:
                with _strict_fp32_matmul():
                    torch.mm(h, pk["w_h"][i].t(), out=x_l)
                _cabi.skinny_matmul(h, pk["v_node"][i], out=a_node)
            if side is not None:                # the hop is the first consumer of the topology (the event, not the
                torch.cuda.current_stream(x.device).wait_event(csr_ready)   # stream: the slab build queued behind
                side = None                                                 # the CSR build must not hold hop 0 up)
            if i == 1 and slab_ready is not None:
                torch.cuda.current_stream(x.device).wait_event(slab_ready)
            use_slab = slab_idx is not None and i >= 1
            last = i == num_hops - 1
            ln = None if last else self._interleaved_ln
            if last:
                epi = dict(epilogue=_cabi.EPI_NONE)
            elif ln is not None:
                epi = dict(epilogue=_cabi.EPI_GRAPH_LN, ln_eps=ln.eps,
                           ln_weight=None if ln.weight is None else ln.weight.detach(),
                           ln_bias=None if ln.bias is None else ln.bias.detach())
            else:
                epi = dict(epilogue=_cabi.EPI_AFFINE_RELU, ep_scale=pk["scale"][i], ep_shift=pk["shift"][i])
            h_out = torch.empty(n, c, dtype=torch.float32, device=x.device)
            if self.hop_events is not None:
                ext = capturing       # inside a CUDA-graph capture the pair becomes two event-record nodes
                ev = (torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext))
                ev[0].record()
            if not self.skip_hop_launch:
                _cabi.gat_hop(x_l, a_node, a_edge_all[:, i * heads:], csr_d, heads, c, h_out,
                              lde=a_edge_all.stride(0), graph_bias=graph_bias_all[i], a_graph=a_graph_all[i],
                              h_prev=h, bias=self.convs[i].bias, negative_slope=self.convs[i].negative_slope,
                              variant=(self.kernel_variant if use_slab or self.kernel_variant != _cabi.VARIANT_SLAB
                                       else _cabi.VARIANT_BLOCK),
                              slab_idx=slab_idx if use_slab else None, slab_f=slab_f[i] if use_slab else None, **epi,
                              # topology and pre-pass outputs are older than the projection launched just above
                              # (except hop 0 of a grouped launch, whose predecessor also wrote the pre-pass outputs)
                              inputs_older_than_predecessor=fused_logits and not (grouped and i == 0), **csr.hints())
            if self.hop_events is not None:
                ev[1].record()
                self.hop_events.append(ev)
            h = h_out
            if return_hops:
                hops.append(h)
        if self.l2_persist:
            _cabi.l2_window(None, x.device)
        if self.projection == "3xf16" and not capturing and not self.overflow_external:
            self._queue_overflow_check()
        return (h, hops) if return_hops else h

    def _forward_fused(self, x, edge_attr, ins, csr, side, csr_ready, pk, flag, return_hops, batch=None, a_edge_all=None):
        """Hops as ONE kernel each (hop_mode "fused"): per hop the collapsed node logits (skinny matvec), the softmax
        weights of all in-edges (gvqa_gat_alpha_f32) and gvqa_gat_fused_hop_f32."""
        num_hops = len(self.convs)
        n, e, b = x.size(0), csr.num_edges, ins.size(1)
        heads, c = self.convs[0].heads, self.convs[0].out_channels
        cur = torch.cuda.current_stream(x.device)
        # per-batch: row tiles (beside the CSR build when that runs on the side stream), pre-pass products
        window = _cabi.fused_window(csr.max_nodes_per_graph or 256)
        if side is not None and batch is not None and batch.dtype == torch.int64 and window not in csr._fused_plans:
            # the CSR is being built on the side stream: plan straight from `batch` on a second side stream, beside it
            side2 = self._side_stream2(x.device)
            side2.wait_stream(cur)
            with torch.cuda.stream(side2):
                plan = _cabi.fused_plan_from_batch(batch.contiguous(), b, window)
                plan_ready = torch.cuda.Event()
                plan_ready.record(side2)
            csr._fused_plans[window] = plan
            for t in plan:
                t.record_stream(cur)
        elif side is not None:
            with torch.cuda.stream(side):
                plan = csr.fused_plan(window)
                plan_ready = torch.cuda.Event()
                plan_ready.record(side)
            for t in plan:
                t.record_stream(cur)
        else:
            plan, plan_ready = csr.fused_plan(window), None
        ld = pk["ins_ld"]
        bh, bl = (t.unflatten(0, (num_hops, ld)) for t in pk["ins_split"])
        # ONE pre-pass launch: hop 0's node logits x @ [V_l ; V_r]^T, the per-graph instruction terms of all hops and
        # (with edges) all hops' edge logits in one sweep over edge_attr
        problems = [(x, pk["v0_split"][0], pk["v0_split"][1], None), (ins, bh, bl, None)]
        if e > 0 and a_edge_all is None:
            problems.append((edge_attr, pk["edge_split"][0], pk["edge_split"][1], None))
        outs = _cabi.proj_gemm_3xf16_grouped(problems, overflow=flag)
        a_node, g_all = outs[0], outs[1]
        if a_edge_all is None:
            a_edge_all = outs[2] if e > 0 else x.new_zeros(1, num_hops * heads)
        elif a_edge_all.dtype != torch.float32 or a_edge_all.dim() != 2 or a_edge_all.stride(1) != 1 \
                or a_edge_all.size(0) != e or a_edge_all.size(1) < num_hops * heads:
            raise ValueError("gat_seq: a_edge_all must be float32 [E, >= hops*heads] with unit column stride")
        csr_d = csr.as_dict()
        alpha = torch.empty(max(e, 1), heads, dtype=torch.float32, device=x.device)    # scratch of the kernels' generic path
        # hops >= 1 get their node logits from the previous hop's epilogue (partial sums per 128-column block)
        # (two buffers: a hop reads the previous hop's block while it writes its own)
        a_part = torch.empty(2, _cabi.fused_part_blocks(n, c), n, 2 * heads, dtype=torch.float32, device=x.device)
        h, hops = x, []
        terms = None
        for i in range(num_hops):
            if i == 0:
                if csr_ready is not None:
                    cur.wait_event(csr_ready)
                if plan_ready is not None:
                    cur.wait_event(plan_ready)
                # hop-invariant logit terms of all hops, CSR order (the kernels run the softmax in their tile prologues)
                terms = _cabi.fused_logit_terms(csr_d, a_edge_all, g_all[:, :, c:c + heads], num_hops, heads, n)
            last = i == num_hops - 1
            epi = dict(epilogue=_cabi.EPI_NONE) if last else \
                dict(epilogue=_cabi.EPI_AFFINE_RELU, ep_scale=pk["scale"][i], ep_shift=pk["shift"][i])
            h_out = torch.empty(n, c, dtype=torch.float32, device=x.device)
            if self.hop_events is not None:
                ext = torch.cuda.is_current_stream_capturing()
                ev = (torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext))
                ev[0].record()
            if not self.skip_hop_launch:
                _cabi.gat_fused_hop(h, pk["w_fused"][i], plan, csr_d, alpha, heads, c, h_out, window=window, skip=h,
                                    graph_bias=g_all[i, :, :c], bias=self.convs[i].bias, overflow=flag,
                                    v_next=None if last else pk["v_node"][i + 1], a_part=None if last else a_part[i & 1],
                                    logit_terms=terms[i], a_node=a_node if i == 0 else a_part[(i - 1) & 1],
                                    negative_slope=self.convs[i].negative_slope,
                                    # (the predecessor of hop i >= 1 is hop i - 1: it writes h and the node logits only)
                                    inputs_older_than_predecessor=i > 0, **epi)
            if self.hop_events is not None:
                ev[1].record()
                self.hop_events.append(ev)
            h = h_out
            if return_hops:
                hops.append(h)
        if not torch.cuda.is_current_stream_capturing() and not self.overflow_external:
            self._queue_overflow_check()
        return (h, hops) if return_hops else h

    def _slab_l2(self, device):
        """GVQA_SLAB_L2=1 (experiment): the slab build writes its ~6 MB through a persisting L2 access-policy window,
        so the hop kernels' slab fetch (one per CTA, on the critical path of the prologue) hits L2 instead of DRAM."""
        if self._slab_l2_on is None:
            self._slab_l2_on = os.environ.get("GVQA_SLAB_L2", "0") != "0" and _cabi.l2_persist_limit(16 << 20, device) > 0
        return self._slab_l2_on

    def _side_stream(self, device):
        if self._side is None or self._side.device != device:
            self._side = torch.cuda.Stream(device)
        return self._side

    def _side_stream2(self, device):
        side2 = self.__dict__.get("_side2")
        if side2 is None or side2.device != device:
            side2 = self.__dict__["_side2"] = torch.cuda.Stream(device)
        return side2

    # ---- fp16 range guard of the "3xf16" projection -----------------------------------------------
    # The kernels OR a device flag when an input element does not fit fp16 (|x| >= 65504 or not finite).
    # Reading it must not stall the stream, so every eager forward queues an asynchronous copy of the flag and
    # the NEXT forward (or an explicit check_overflow()) looks at the copies that have landed.
    def _queue_overflow_check(self):
        host = torch.empty(1, dtype=torch.int32).pin_memory()
        host.copy_(self._overflow, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._overflow_pending.append((host, ev))
        del self._overflow_pending[:-8]

    def _poll_overflow(self, wait=False):
        keep = []
        for host, ev in self._overflow_pending:
            if wait:
                ev.synchronize()
            if ev.query():
                if int(host) != 0:
                    self._overflow_pending = []
                    self._overflow.zero_()
                    raise FloatingPointError(
                        "gat_seq: an input of the fp16-split projection was outside fp16's range (|x| >= 65504 or "
                        "not finite); the results of that forward are invalid. Set gat_seq.projection = '3xtf32' "
                        "(full fp32 range, ~1.4x slower projection) and rerun.")
            else:
                keep.append((host, ev))
        self._overflow_pending = keep

    def check_overflow(self):
        """Synchronising check of the fp16 range flag (after CUDA-graph replays or before trusting results)."""
        if self.projection != "3xf16" or self._overflow is None:
            return
        self._poll_overflow(wait=True)
        if int(self._overflow) != 0:
            self._overflow.zero_()
            raise FloatingPointError("gat_seq: input outside fp16's range in the fp16-split projection; use "
                                     "gat_seq.projection = '3xtf32'")

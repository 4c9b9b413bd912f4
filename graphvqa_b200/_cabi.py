"""ctypes binding of the C ABI in include/gvqa_b200.h (graphvqa_b200/lib/libgvqa_b200.so).

This is the only place the package talks to native code.  There is NO fallback: if the shared
library is missing or a call fails, a RuntimeError is raised.  Device pointers are taken from
torch tensors (``data_ptr()``) and work is enqueued on torch's current CUDA stream, so calls
compose with torch ops and are CUDA-graph capturable.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgvqa_b200.so")
ABI_VERSION = 5

EPI_NONE, EPI_AFFINE, EPI_AFFINE_RELU, EPI_GRAPH_LN = 0, 1, 2, 3
VARIANT_AUTO, VARIANT_GATHER, VARIANT_STAGED, VARIANT_BLOCK, VARIANT_WS, VARIANT_SLAB = 0, 1, 2, 3, 4, 5
HOP_INPUTS_OLDER_THAN_PREDECESSOR = 1

_c_i32, _c_i64, _c_f32, _c_vp, _c_sz = (ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                                        ctypes.c_void_p, ctypes.c_size_t)


class GatHopArgs(ctypes.Structure):
    """Mirror of ``struct gvqa_gat_hop_args`` (field order and types must match the header)."""
    _fields_ = [
        ("x_l", _c_vp), ("ldx", _c_i64), ("graph_bias", _c_vp), ("a_node", _c_vp), ("ld_a_node", _c_i64),
        ("a_graph", _c_vp),
        ("a_edge", _c_vp), ("lde", _c_i64), ("rowptr", _c_vp), ("col_src", _c_vp), ("perm", _c_vp),
        ("graph_ptr", _c_vp), ("node_graph", _c_vp), ("h_prev", _c_vp), ("bias", _c_vp),
        ("ep_scale", _c_vp), ("ep_shift", _c_vp), ("h_out", _c_vp), ("alpha_out", _c_vp),
        ("num_nodes", _c_i64), ("num_edges", _c_i64), ("num_graphs", _c_i64),
        ("heads", _c_i32), ("channels", _c_i32), ("negative_slope", _c_f32), ("epilogue", _c_i32),
        ("max_nodes_per_graph", _c_i32), ("max_in_edges_per_graph", _c_i32), ("variant", _c_i32),
        ("ld_graph_bias", _c_i64), ("ld_a_graph", _c_i64), ("flags", _c_i32), ("ln_eps", _c_f32),
        ("ln_weight", _c_vp), ("ln_bias", _c_vp), ("slab_idx", _c_vp), ("slab_f", _c_vp), ("sched", _c_vp),
    ]


class GatFusedArgs(ctypes.Structure):
    """Mirror of ``struct gvqa_gat_fused_args``."""
    _fields_ = [
        ("h_in", _c_vp), ("ld_h", _c_i64), ("w_pack", _c_vp), ("tiles", _c_vp), ("tile_count", _c_vp),
        ("rowptr", _c_vp), ("col_src", _c_vp), ("node_graph", _c_vp), ("alpha", _c_vp),
        ("logit_terms", _c_vp), ("a_node", _c_vp), ("a_node_part_stride", _c_i64), ("a_node_parts", _c_i32),
        ("negative_slope", _c_f32), ("ld_a_node", _c_i64), ("flags", _c_i32),
        ("skip", _c_vp), ("ld_skip", _c_i64), ("graph_bias", _c_vp), ("ld_graph_bias", _c_i64),
        ("bias", _c_vp), ("ep_scale", _c_vp), ("ep_shift", _c_vp), ("h_out", _c_vp), ("overflow", _c_vp),
        ("num_nodes", _c_i64), ("in_channels", _c_i32), ("channels", _c_i32), ("heads", _c_i32),
        ("epilogue", _c_i32), ("window", _c_i32), ("w_scale", _c_f32),
        ("v_next", _c_vp), ("a_part", _c_vp), ("a_part_blocks", _c_i32),
    ]


class GatSlabPlan(ctypes.Structure):
    """Mirror of ``struct gvqa_gat_slab_plan``."""
    _fields_ = [("nodes_per_cta", _c_i32), ("edge_capacity", _c_i32), ("num_ctas", _c_i32), ("idx_words", _c_i64),
                ("f_words_per_hop", _c_i64)]


class GemmProblem(ctypes.Structure):
    """Mirror of ``struct gvqa_gemm_problem`` (one product of a grouped launch)."""
    _fields_ = [
        ("a", _c_vp), ("lda", _c_i64), ("stride_a", _c_i64), ("b_hi", _c_vp), ("b_lo", _c_vp), ("ldb", _c_i64),
        ("stride_b", _c_i64), ("c", _c_vp), ("ldc", _c_i64), ("stride_c", _c_i64), ("m", _c_i64),
        ("n", _c_i32), ("k", _c_i32), ("batch", _c_i32), ("relu", _c_i32), ("bias", _c_vp),
    ]


MAX_GROUPED_PROBLEMS = 3

# name -> (restype, argtypes); must list every symbol declared in include/gvqa_b200.h
SIGNATURES = {
    "gvqa_abi_version": (ctypes.c_int, []),
    "gvqa_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "gvqa_device_set_l2_persist_limit": (ctypes.c_int, [_c_sz]),
    "gvqa_stream_set_l2_window": (ctypes.c_int, [_c_vp, _c_sz, _c_f32, _c_vp]),
    "gvqa_csr_workspace_bytes": (_c_sz, [_c_i64, _c_i64]),
    "gvqa_build_csr": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp,
                                      _c_vp, _c_vp, _c_vp, _c_sz, _c_vp]),
    "gvqa_skinny_matmul_f32": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, ctypes.c_int,
                                              ctypes.c_int, _c_vp]),
    "gvqa_gat_hop_f32": (ctypes.c_int, [ctypes.POINTER(GatHopArgs), _c_vp]),
    "gvqa_gat_hop_slab_plan": (ctypes.c_int, [_c_i64, _c_i64, _c_i32, ctypes.POINTER(GatSlabPlan)]),
    "gvqa_gat_hop_build_slabs_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i32,
                                                    _c_i64, _c_i64, _c_i32, _c_vp, _c_vp, _c_vp]),
    "gvqa_gat_fused_supported": (ctypes.c_int, [_c_i32, _c_i32, _c_i32]),
    "gvqa_gat_fused_pack_halves": (_c_i64, [_c_i32, _c_i32, _c_i32]),
    "gvqa_gat_fused_pack_f16": (ctypes.c_int, [_c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_f32, _c_vp, _c_vp]),
    "gvqa_gat_fused_max_tiles": (_c_i64, [_c_i64, _c_i64]),
    "gvqa_gat_fused_part_blocks": (_c_i32, [_c_i64, _c_i32]),
    "gvqa_gat_fused_window": (_c_i32, [_c_i32]),
    "gvqa_gat_fused_plan": (ctypes.c_int, [_c_vp, _c_i64, _c_i32, _c_vp, _c_vp, _c_i64, _c_vp]),
    "gvqa_gat_fused_plan_from_batch": (ctypes.c_int, [_c_vp, _c_i64, _c_i64, _c_i32, _c_vp, _c_vp, _c_i64, _c_vp]),
    "gvqa_gat_fused_plan_host": (ctypes.c_int, [_c_vp, _c_i64, _c_i32, _c_vp, _c_vp, _c_i64]),
    "gvqa_gat_alpha_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64,
                                          _c_f32, _c_i64, _c_i32, _c_vp, _c_vp, _c_vp]),
    "gvqa_gat_fused_hop_f32": (ctypes.c_int, [ctypes.POINTER(GatFusedArgs), _c_vp]),
    "gvqa_gat_fused_logit_terms_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i32,
                                                      _c_i64, _c_i64, _c_i32, _c_vp, _c_i64, _c_vp]),
    "gvqa_debug_set_fused_trace": (None, [_c_vp]),
    "gvqa_graph_layernorm_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i32,
                                                _c_f32, _c_i32, _c_vp]),
    "gvqa_debug_set_gemm_trace": (None, [_c_vp]),
    "gvqa_debug_set_gemm_flags": (None, [ctypes.c_int]),
    "gvqa_debug_set_hop_trace": (None, [_c_vp]),
    "gvqa_split_tf32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_vp]),
    "gvqa_proj_gemm_3xtf32": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i32,
                                             _c_i32, _c_vp]),
    "gvqa_split_f16": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_vp]),
    "gvqa_proj_gemm_3xf16": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i32,
                                            _c_i32, _c_vp, _c_vp]),
    "gvqa_linear_3xf16": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_vp, _c_i32, _c_vp, _c_i64, _c_i64, _c_i32,
                                         _c_i32, _c_vp, _c_vp]),
    "gvqa_proj_gemm_3xf16_batched": (ctypes.c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_i64,
                                                    _c_i64, _c_i64, _c_i32, _c_i32, _c_i32, _c_vp, _c_vp]),
    "gvqa_proj_gemm_3xf16_grouped": (ctypes.c_int, [ctypes.POINTER(GemmProblem), _c_i32, _c_vp, _c_vp]),
    "gvqa_gine_aggregate_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                               _c_i32, _c_i32, _c_f32, _c_vp]),
    "gvqa_gcn_degree_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_vp]),
    "gvqa_gcn_aggregate_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                              _c_i32, _c_vp]),
    "gvqa_gather_add_relu_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_vp]),
    "gvqa_gather_add_relu_i32_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_vp]),
    "gvqa_embedding_sum_f32": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_i32, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_vp]),
    "gvqa_affine_relu_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_vp]),
    "gvqa_gather_add_relu_strided_f32": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp, _c_i32, _c_vp,
                                                        _c_i64, _c_i32, _c_i32, _c_vp]),
    "gvqa_graph_scale_rows_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_vp]),
    "gvqa_attention_pool_gate_f32": (ctypes.c_int, [_c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64,
                                                    _c_i32, _c_vp]),
    "gvqa_build_csr_host": (ctypes.c_int, [_c_vp, _c_i32, _c_i64, _c_vp, _c_i32, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp,
                                           _c_vp, _c_vp, _c_vp]),
    "gvqa_segment_mean_rows_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_vp]),
    "gvqa_attention_pool_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_vp]),
    "gvqa_lcgn_hop_f32": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                         _c_vp, _c_i64, _c_i32, _c_f32, _c_vp]),
}

_lib = None


def lib():
    """Load the shared library once; raise loudly when it is absent or stale."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "graphvqa_b200: CUDA library %s not found. Build it with "
                "`python -m graphvqa_b200.build` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        handle.gvqa_abi_version.restype, handle.gvqa_abi_version.argtypes = ctypes.c_int, []
        if handle.gvqa_abi_version() != ABI_VERSION:      # checked first: a stale library may lack newer symbols
            raise RuntimeError("graphvqa_b200: ABI version mismatch (library %d, binding %d); rebuild with "
                               "`python -m graphvqa_b200.build`" % (handle.gvqa_abi_version(), ABI_VERSION))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        raise RuntimeError("graphvqa_b200: %s failed: %s (status %d)"
                           % (what, lib().gvqa_error_string(status).decode(), status))


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_handle(device):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("graphvqa_b200 runs on CUDA tensors only (got a %s tensor); "
                               "there is no CPU fallback" % t.device)


def require_f32c(**tensors):
    for name, t in tensors.items():
        if t is None:
            continue
        if t.dtype != torch.float32:
            raise TypeError("%s must be float32, got %s" % (name, t.dtype))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % name)


# ------------------------------------------------------------------------------------------
# thin tensor-level wrappers
# ------------------------------------------------------------------------------------------
def build_csr(edge_index, batch, num_graphs):
    """Returns dict(rowptr, col_src, perm, graph_ptr, node_graph, stats) of int32 tensors."""
    require_cuda(edge_index, batch)
    if edge_index.dtype != torch.int64 or batch.dtype != torch.int64:
        raise TypeError("edge_index and batch must be int64 (the reference's layout)")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must be [2, E]")
    edge_index = edge_index.contiguous()
    batch = batch.contiguous()
    dev = edge_index.device
    n, e = batch.numel(), edge_index.size(1)
    i32 = dict(dtype=torch.int32, device=dev)
    out = dict(rowptr=torch.empty(n + 1, **i32), col_src=torch.empty(max(e, 1), **i32),
               perm=torch.empty(max(e, 1), **i32), graph_ptr=torch.empty(num_graphs + 1, **i32),
               node_graph=torch.empty(max(n, 1), **i32), stats=torch.empty(8, **i32))
    ws_bytes = lib().gvqa_csr_workspace_bytes(n, e)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().gvqa_build_csr(ptr(edge_index), e, ptr(batch), n, num_graphs, ptr(out["rowptr"]),
                                   ptr(out["col_src"]), ptr(out["perm"]), ptr(out["graph_ptr"]),
                                   ptr(out["node_graph"]), ptr(out["stats"]), ptr(ws), ws_bytes,
                                   stream_handle(dev)), "gvqa_build_csr")
    return out


def skinny_matmul(x, v, out=None):
    """out[M,K] = x[M,F] @ v[K,F]^T ; x may be a row-strided view (stride(1) == 1)."""
    require_cuda(x, v)
    if x.dtype != torch.float32 or v.dtype != torch.float32:
        raise TypeError("skinny_matmul expects float32")
    if x.dim() != 2 or v.dim() != 2 or x.size(1) != v.size(1) or x.stride(1) != 1 or not v.is_contiguous():
        raise ValueError("skinny_matmul: x [M,F] (unit column stride), v [K,F] contiguous")
    m, f = x.shape
    k = v.size(0)
    if out is None:
        out = torch.empty(m, k, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().gvqa_skinny_matmul_f32(ptr(x), x.stride(0) if m > 1 else f, ptr(v), ptr(out), m, f, k,
                                           stream_handle(x.device)), "gvqa_skinny_matmul_f32")
    return out


def gat_hop(x_l, a_node, a_edge, csr, heads, channels, h_out, *, ldx=None, lde=None, graph_bias=None,
            a_graph=None, h_prev=None, bias=None, ep_scale=None, ep_shift=None, alpha_out=None,
            negative_slope=0.2, epilogue=EPI_NONE, num_graphs=None, max_nodes_per_graph=0,
            max_in_edges_per_graph=0, variant=VARIANT_AUTO, inputs_older_than_predecessor=False,
            ln_weight=None, ln_bias=None, ln_eps=1e-5, sched=None, slab_idx=None, slab_f=None):
    require_cuda(x_l, a_node, a_edge, h_out, graph_bias, a_graph, h_prev, bias, ep_scale, ep_shift, alpha_out,
                 ln_weight, ln_bias, sched)
    if variant == VARIANT_WS and sched is None:
        sched = hop_sched(h_out.device)
    require_f32c(h_out=h_out, h_prev=h_prev, bias=bias, ep_scale=ep_scale, ep_shift=ep_shift, alpha_out=alpha_out)
    for name, t in (("graph_bias", graph_bias), ("a_graph", a_graph)):    # [B, .] row-strided views are fine
        if t is not None and (t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1):
            raise ValueError("gat_hop: %s must be float32 [B, .] with unit column stride" % name)
    n = h_out.size(0)
    e = csr["num_edges"]
    a = GatHopArgs()
    a.x_l, a.ldx = ptr(x_l), (x_l.stride(0) if ldx is None else ldx)
    a.graph_bias, a.a_node, a.a_graph = ptr(graph_bias), ptr(a_node), ptr(a_graph)
    a.ld_a_node = a_node.stride(0) if a_node.dim() == 2 and a_node.size(0) > 1 else 0
    a.a_edge, a.lde = ptr(a_edge), (a_edge.stride(0) if lde is None else lde)
    a.rowptr, a.col_src, a.perm = ptr(csr["rowptr"]), ptr(csr["col_src"]), ptr(csr["perm"])
    a.graph_ptr, a.node_graph = ptr(csr["graph_ptr"]), ptr(csr["node_graph"])
    a.h_prev, a.bias, a.ep_scale, a.ep_shift = ptr(h_prev), ptr(bias), ptr(ep_scale), ptr(ep_shift)
    a.h_out, a.alpha_out = ptr(h_out), ptr(alpha_out)
    a.num_nodes, a.num_edges = n, e
    a.num_graphs = csr["num_graphs"] if num_graphs is None else num_graphs
    a.heads, a.channels, a.negative_slope, a.epilogue = heads, channels, negative_slope, epilogue
    a.max_nodes_per_graph, a.max_in_edges_per_graph = max_nodes_per_graph, max_in_edges_per_graph
    a.variant = variant
    a.ld_graph_bias = graph_bias.stride(0) if graph_bias is not None and graph_bias.size(0) > 1 else 0
    a.ld_a_graph = a_graph.stride(0) if a_graph is not None and a_graph.size(0) > 1 else 0
    a.flags = HOP_INPUTS_OLDER_THAN_PREDECESSOR if inputs_older_than_predecessor else 0
    a.ln_eps, a.ln_weight, a.ln_bias, a.sched = ln_eps, ptr(ln_weight), ptr(ln_bias), ptr(sched)
    a.slab_idx, a.slab_f = ptr(slab_idx), ptr(slab_f)
    with torch.cuda.device(h_out.device):
        check(lib().gvqa_gat_hop_f32(ctypes.byref(a), stream_handle(h_out.device)), "gvqa_gat_hop_f32")
    return h_out


# ---- fused aggregate-project hop (gvqa_gat_fused_*) -------------------------------------------------------------
def fused_supported(heads, in_channels, channels):
    return bool(lib().gvqa_gat_fused_supported(heads, in_channels, channels))


def fused_pack(w, heads, channels, in_channels, scale=None):
    """lin_l.weight ([H*C, >= F] fp32, only the first ``in_channels`` columns are used) -> (packed fp16 operand of
    gvqa_gat_fused_hop_f32, scale).  ``scale``: power of two applied before the fp16 split (default: max|W| lands in
    [2^8, 2^9), so the unscaled low parts stay in fp16's normal range); the hop divides it out."""
    require_cuda(w)
    if w.dtype != torch.float32 or w.dim() != 2 or w.stride(1) != 1 or w.size(0) != heads * channels:
        raise ValueError("fused_pack: w must be float32 [heads*channels, >= in_channels] with unit column stride")
    if scale is None:
        import math
        mx = float(w[:, :in_channels].abs().max()) if w.numel() else 0.0
        scale = 2.0 ** max(-14, min(24, 8 - math.floor(math.log2(mx)))) if mx > 0 and math.isfinite(mx) else 1.0
    out = torch.empty(lib().gvqa_gat_fused_pack_halves(heads, channels, in_channels), dtype=torch.float16, device=w.device)
    with torch.cuda.device(w.device):
        check(lib().gvqa_gat_fused_pack_f16(ptr(w), w.stride(0), heads, channels, in_channels, scale, ptr(out),
                                            stream_handle(w.device)), "gvqa_gat_fused_pack_f16")
    return out, scale


def fused_window(max_nodes_per_graph):
    return lib().gvqa_gat_fused_window(int(max_nodes_per_graph))


def fused_plan(graph_ptr, num_nodes, num_graphs, window):
    """Per-batch row tiles of the fused hop: (tiles int32 [max_tiles, 4], count int32 [1]) on graph_ptr's device."""
    require_cuda(graph_ptr)
    max_tiles = lib().gvqa_gat_fused_max_tiles(num_nodes, num_graphs)
    tiles = torch.empty(max_tiles, 4, dtype=torch.int32, device=graph_ptr.device)
    count = torch.empty(1, dtype=torch.int32, device=graph_ptr.device)
    with torch.cuda.device(graph_ptr.device):
        check(lib().gvqa_gat_fused_plan(ptr(graph_ptr), num_graphs, window, ptr(tiles), ptr(count), max_tiles,
                                        stream_handle(graph_ptr.device)), "gvqa_gat_fused_plan")
    return tiles, count


def fused_plan_from_batch(batch, num_graphs, window):
    """The same plan from the int64 ``batch`` vector (device): independent of the CSR build."""
    require_cuda(batch)
    if batch.dtype != torch.int64 or not batch.is_contiguous():
        raise ValueError("fused_plan_from_batch: batch must be a contiguous int64 tensor")
    n = batch.numel()
    max_tiles = lib().gvqa_gat_fused_max_tiles(n, num_graphs)
    tiles = torch.empty(max_tiles, 4, dtype=torch.int32, device=batch.device)
    count = torch.empty(1, dtype=torch.int32, device=batch.device)
    with torch.cuda.device(batch.device):
        check(lib().gvqa_gat_fused_plan_from_batch(ptr(batch), n, num_graphs, window, ptr(tiles), ptr(count), max_tiles,
                                                   stream_handle(batch.device)), "gvqa_gat_fused_plan_from_batch")
    return tiles, count


def fused_plan_host(graph_ptr, num_nodes, num_graphs, window, pin=False):
    """The same plan from a host int32 graph_ptr (loader side, wire format)."""
    if graph_ptr.is_cuda or graph_ptr.dtype != torch.int32 or not graph_ptr.is_contiguous():
        raise ValueError("fused_plan_host: graph_ptr must be a contiguous host int32 tensor")
    max_tiles = lib().gvqa_gat_fused_max_tiles(num_nodes, num_graphs)
    tiles = torch.zeros(max_tiles, 4, dtype=torch.int32, pin_memory=pin)
    count = torch.zeros(1, dtype=torch.int32, pin_memory=pin)
    check(lib().gvqa_gat_fused_plan_host(graph_ptr.data_ptr(), num_graphs, window, tiles.data_ptr(), count.data_ptr(),
                                         max_tiles), "gvqa_gat_fused_plan_host")
    return tiles, count


def fused_part_blocks(num_nodes, channels):
    return lib().gvqa_gat_fused_part_blocks(num_nodes, channels)


def gat_alpha(a_node, a_edge, csr, heads, *, lde=None, a_graph=None, negative_slope=0.2, out=None, alpha_out=None):
    """Softmax weights of all in-edges, CSR order: alpha [E, heads].  ``a_node``: [N, >= 2H] node logits, or
    [parts, N, 2H] partial sums (the fused hop's ``a_part``)."""
    require_cuda(a_node, a_edge, a_graph, out, alpha_out)
    n = csr["rowptr"].numel() - 1
    e = csr["num_edges"]
    if out is None:
        out = torch.empty(max(e, 1), heads, dtype=torch.float32, device=a_node.device)
    parts, part_stride = 1, 0
    if a_node.dim() == 3:
        parts, part_stride = a_node.size(0), a_node.stride(0)
        a_node = a_node[0]
    with torch.cuda.device(a_node.device):
        check(lib().gvqa_gat_alpha_f32(ptr(csr["rowptr"]), ptr(csr["col_src"]), ptr(csr["perm"]), ptr(csr["node_graph"]),
                                       ptr(a_node), a_node.stride(0) if a_node.size(0) > 1 else a_node.size(1),
                                       parts, part_stride, ptr(a_edge), a_edge.stride(0) if lde is None else lde,
                                       ptr(a_graph), (a_graph.stride(0) if a_graph is not None and a_graph.size(0) > 1 else heads),
                                       negative_slope, n, heads, ptr(out), ptr(alpha_out), stream_handle(a_node.device)),
              "gvqa_gat_alpha_f32")
    return out


def fused_logit_terms(csr, a_edge_all, a_graph_all, hops, heads, num_nodes):
    """Hop-invariant logit terms of all hops in CSR order: [hops, E, heads] (gvqa_gat_fused_logit_terms_f32).
    ``a_edge_all`` [E, >= hops*heads]; ``a_graph_all`` [hops, B, heads] (any strides with unit last stride) or None."""
    require_cuda(a_edge_all, a_graph_all)
    e = csr["num_edges"]
    out = torch.empty(hops, (max(e, 1) + 3) // 4 * 4, heads, dtype=torch.float32, device=a_edge_all.device)
    if a_graph_all is not None and (a_graph_all.dim() != 3 or a_graph_all.stride(2) != 1):
        raise ValueError("fused_logit_terms: a_graph_all must be [hops, B, heads] with unit last stride")
    with torch.cuda.device(out.device):
        check(lib().gvqa_gat_fused_logit_terms_f32(
            ptr(csr["rowptr"]), ptr(csr["perm"]), ptr(csr["node_graph"]), ptr(a_edge_all), a_edge_all.stride(0),
            ptr(a_graph_all), a_graph_all.stride(1) if a_graph_all is not None else 0,
            a_graph_all.stride(0) if a_graph_all is not None else 0, hops, num_nodes, e, heads, ptr(out), out.stride(0),
            stream_handle(out.device)), "gvqa_gat_fused_logit_terms_f32")
    return out


def gat_fused_hop(h_in, w_pack, plan, csr, alpha, heads, channels, h_out, *, window, skip=None, graph_bias=None,
                  bias=None, ep_scale=None, ep_shift=None, epilogue=EPI_NONE, overflow=None, v_next=None, a_part=None,
                  logit_terms=None, a_node=None, negative_slope=0.2, inputs_older_than_predecessor=False):
    """``alpha`` [E, heads]: the softmax weights (gat_alpha) -- or, with ``logit_terms`` ([E, heads], this hop's block of
    fused_logit_terms) and ``a_node`` ([N, 2*heads] or [parts, N, 2*heads] partial sums), scratch: the kernel computes
    the weights in its tile prologues.  ``v_next`` [2*heads, channels] + ``a_part`` [fused_part_blocks, N, 2*heads]: the
    epilogue also emits the next hop's node logits as partial sums."""
    require_cuda(h_in, w_pack[0], alpha, h_out, skip, graph_bias, bias, ep_scale, ep_shift, overflow, v_next, a_part,
                 logit_terms, a_node)
    require_f32c(v_next=v_next, a_part=a_part, logit_terms=logit_terms, alpha=alpha)
    require_f32c(h_out=h_out, bias=bias, ep_scale=ep_scale, ep_shift=ep_shift)
    for name, t in (("h_in", h_in), ("skip", skip), ("graph_bias", graph_bias)):
        if t is not None and (t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1):
            raise ValueError("gat_fused_hop: %s must be float32 2-D with unit column stride" % name)
    a = GatFusedArgs()
    a.h_in, a.ld_h = ptr(h_in), (h_in.stride(0) if h_in.size(0) > 1 else h_in.size(1))
    a.w_pack, a.w_scale, a.tiles, a.tile_count = ptr(w_pack[0]), w_pack[1], ptr(plan[0]), ptr(plan[1])
    a.rowptr, a.col_src, a.node_graph, a.alpha = ptr(csr["rowptr"]), ptr(csr["col_src"]), ptr(csr["node_graph"]), ptr(alpha)
    a.skip = ptr(skip)
    a.ld_skip = skip.stride(0) if skip is not None and skip.size(0) > 1 else 0
    a.graph_bias = ptr(graph_bias)
    a.ld_graph_bias = graph_bias.stride(0) if graph_bias is not None and graph_bias.size(0) > 1 else 0
    a.bias, a.ep_scale, a.ep_shift, a.h_out, a.overflow = ptr(bias), ptr(ep_scale), ptr(ep_shift), ptr(h_out), ptr(overflow)
    a.num_nodes, a.in_channels, a.channels, a.heads = h_out.size(0), h_in.size(1), channels, heads
    a.epilogue, a.window = epilogue, window
    a.v_next, a.a_part, a.a_part_blocks = ptr(v_next), ptr(a_part), (a_part.size(0) if a_part is not None else 0)
    a.logit_terms, a.a_node, a.negative_slope = ptr(logit_terms), ptr(a_node), negative_slope
    a.a_node_parts, a.a_node_part_stride = 1, 0
    a.ld_a_node = 0
    a.flags = HOP_INPUTS_OLDER_THAN_PREDECESSOR if inputs_older_than_predecessor else 0
    if a_node is not None:
        if a_node.size(-1) < 2 * heads or a_node.stride(-1) != 1:
            raise ValueError("gat_fused_hop: a_node must be [N, >= 2*heads] or [parts, N, 2*heads]")
        a.ld_a_node = a_node.stride(-2)
        if a_node.dim() == 3:
            a.a_node_parts, a.a_node_part_stride = a_node.size(0), a_node.stride(0)
    with torch.cuda.device(h_out.device):
        check(lib().gvqa_gat_fused_hop_f32(ctypes.byref(a), stream_handle(h_out.device)), "gvqa_gat_fused_hop_f32")
    return h_out


def build_hop_slabs(csr, a_edge_all, a_graph_all, hops, heads, num_nodes, l2_persist=False):
    """Per-batch slabs of the one-round-trip hop prologue (hop variant 5).  ``a_edge_all`` [E, >= hops*H] (hop j at
    columns j*H..), ``a_graph_all`` None or a [hops, B, H] view with unit column stride (row / hop strides free).
    Returns (slab_idx int32, slab_f float32 [hops, f_words_per_hop])."""
    require_cuda(a_edge_all, a_graph_all)
    e = csr["num_edges"]
    plan = GatSlabPlan()
    check(lib().gvqa_gat_hop_slab_plan(num_nodes, e, heads, ctypes.byref(plan)), "gvqa_gat_hop_slab_plan")
    dev = a_edge_all.device
    # one allocation (index slabs first, then the per-hop term slabs): a single L2 access-policy window can cover it
    buf = torch.empty(plan.idx_words + hops * plan.f_words_per_hop, dtype=torch.float32, device=dev)
    slab_idx = buf[:plan.idx_words].view(torch.int32)
    slab_f = buf[plan.idx_words:].view(hops, plan.f_words_per_hop)
    if l2_persist:
        l2_window(buf, dev, 1.0)
    if a_graph_all is not None and (a_graph_all.dim() != 3 or a_graph_all.stride(2) != 1 or a_graph_all.dtype != torch.float32):
        raise ValueError("build_hop_slabs: a_graph_all must be a float32 [hops, B, H] view with unit column stride")
    with torch.cuda.device(dev):
        check(lib().gvqa_gat_hop_build_slabs_f32(
            ptr(csr["rowptr"]), ptr(csr["col_src"]), ptr(csr["perm"]), ptr(csr["node_graph"]), ptr(a_edge_all),
            a_edge_all.stride(0), ptr(a_graph_all), 0 if a_graph_all is None else a_graph_all.stride(1),
            0 if a_graph_all is None else a_graph_all.stride(0), hops, num_nodes, e, heads, ptr(slab_idx), ptr(slab_f),
            stream_handle(dev)), "gvqa_gat_hop_build_slabs_f32")
    if l2_persist:
        l2_window(None, dev)
    return slab_idx, slab_f


_hop_sched = {}


def hop_sched(device):
    """Scheduler words of the warp-specialised hop kernel (variant 4): one zeroed int32[2] per device, shared by
    all launches of a stream-ordered sequence (the kernel leaves it zero)."""
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    t = _hop_sched.get(key)
    if t is None:
        t = _hop_sched[key] = torch.zeros(2, dtype=torch.int32, device=device)
    return t


def graph_layernorm(x, graph_ptr, num_graphs, weight, bias, eps, out=None, max_nodes_per_graph=0):
    require_cuda(x, graph_ptr, weight, bias)
    require_f32c(x=x, weight=weight, bias=bias, out=out)
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().gvqa_graph_layernorm_f32(ptr(x), ptr(graph_ptr), ptr(weight), ptr(bias), ptr(out),
                                             x.size(0), num_graphs, x.size(1), eps, max_nodes_per_graph,
                                             stream_handle(x.device)), "gvqa_graph_layernorm_f32")
    return out


def gine_aggregate(h, edge_attr, ins, csr, eps, out=None):
    """z[N, F+D] of GINEConv's propagate + self term on the split inputs (see the header)."""
    require_cuda(h, edge_attr, ins)
    require_f32c(h=h, edge_attr=edge_attr, ins=ins, out=out)
    n, f = h.shape
    d = 0 if ins is None else ins.size(1)
    if out is None:
        out = torch.empty(n, f + d, dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        check(lib().gvqa_gine_aggregate_f32(ptr(h), ptr(edge_attr), ptr(ins), ptr(csr["rowptr"]),
                                            ptr(csr["col_src"]), ptr(csr["perm"]), ptr(csr["node_graph"]),
                                            ptr(out), n, f, d, eps, stream_handle(h.device)),
              "gvqa_gine_aggregate_f32")
    return out


def gcn_degree(csr, num_nodes, device):
    dinv = torch.empty(max(num_nodes, 1), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib().gvqa_gcn_degree_f32(ptr(csr["rowptr"]), ptr(csr["col_src"]), ptr(dinv), num_nodes,
                                        stream_handle(device)), "gvqa_gcn_degree_f32")
    return dinv


def gcn_aggregate(xw, graph_term, dinv, bias, csr, out=None):
    require_cuda(xw, graph_term, dinv, bias)
    require_f32c(xw=xw, graph_term=graph_term, dinv=dinv, bias=bias, out=out)
    n, c = xw.shape
    if out is None:
        out = torch.empty(n, c, dtype=torch.float32, device=xw.device)
    with torch.cuda.device(xw.device):
        check(lib().gvqa_gcn_aggregate_f32(ptr(xw), ptr(graph_term), ptr(dinv), ptr(bias), ptr(csr["rowptr"]),
                                           ptr(csr["col_src"]), ptr(csr["node_graph"]), ptr(out), n, c,
                                           stream_handle(xw.device)), "gvqa_gcn_aggregate_f32")
    return out


def lcgn_hop(xl, xr, xv, proj_cmd, cal_cmd, bias, csr, negative_slope, out=None):
    """xl/xr/xv: [N,C] views sharing one row stride (unit column stride)."""
    require_cuda(xl, xr, xv, proj_cmd, cal_cmd, bias)
    require_f32c(proj_cmd=proj_cmd, cal_cmd=cal_cmd, bias=bias, out=out)
    n, c = xl.shape
    ld = xl.stride(0) if n > 1 else c
    for t in (xl, xr, xv):
        if t.dtype != torch.float32 or t.stride(1) != 1 or (n > 1 and t.stride(0) != ld):
            raise ValueError("lcgn_hop: xl/xr/xv must be float32 [N,C] views with a common row stride")
    if out is None:
        out = torch.empty(n, c, dtype=torch.float32, device=xl.device)
    with torch.cuda.device(xl.device):
        check(lib().gvqa_lcgn_hop_f32(ptr(xl), ptr(xr), ptr(xv), ld, ptr(proj_cmd), ptr(cal_cmd), ptr(bias),
                                      ptr(csr["rowptr"]), ptr(csr["col_src"]), ptr(csr["node_graph"]), ptr(out),
                                      n, c, negative_slope, stream_handle(xl.device)), "gvqa_lcgn_hop_f32")
    return out


def split_tf32(w):
    """(hi, lo) with hi = tf32(w), lo = tf32(w - hi): the weight half of the 3xTF32 GEMM."""
    require_cuda(w)
    w = w.contiguous().float()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    with torch.cuda.device(w.device):
        check(lib().gvqa_split_tf32(ptr(w), ptr(hi), ptr(lo), w.numel(), stream_handle(w.device)), "gvqa_split_tf32")
    return hi, lo


def proj_gemm_3xtf32(a, b_hi, b_lo, out=None):
    """out[M,N] = a[M,K] @ b[N,K]^T with fp32-level accuracy on the tcgen05 tensor cores."""
    require_cuda(a, b_hi, b_lo)
    require_f32c(b_hi=b_hi, b_lo=b_lo)
    if out is not None and (out.dtype != torch.float32 or out.dim() != 2 or out.stride(1) != 1):
        raise ValueError("proj_gemm_3xtf32: out must be float32 [M,N] with unit column stride")
    if a.dtype != torch.float32 or a.dim() != 2 or a.stride(1) != 1:
        raise ValueError("proj_gemm_3xtf32: a must be float32 [M,K] with unit column stride")
    m, k = a.shape
    n = b_hi.size(0)
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().gvqa_proj_gemm_3xtf32(ptr(a), a.stride(0) if m > 1 else k, ptr(b_hi), ptr(b_lo), b_hi.stride(0),
                                          ptr(out), out.stride(0), m, n, k, stream_handle(a.device)),
              "gvqa_proj_gemm_3xtf32")
    return out


def split_f16(w):
    """w [rows, cols] float32 -> (hi, lo') float16 [rows, ld] with ld = cols rounded up to 8 (zero padded)."""
    require_cuda(w)
    if w.dtype != torch.float32 or w.dim() != 2 or w.stride(1) != 1:
        raise ValueError("split_f16: w must be float32 [rows, cols] with unit column stride")
    rows, cols = w.shape
    ld = (cols + 7) // 8 * 8
    hi = torch.empty(rows, ld, dtype=torch.float16, device=w.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(w.device):
        check(lib().gvqa_split_f16(ptr(w), w.stride(0) if rows > 1 else cols, ptr(hi), ptr(lo), ld, rows, cols,
                                   stream_handle(w.device)), "gvqa_split_f16")
    return hi[:, :cols], lo[:, :cols]


def proj_gemm_3xf16(a, b_hi, b_lo, out=None, overflow=None, bias=None, relu=False):
    """out[M,N] = act(a[M,K] @ b[N,K]^T + bias), fp32-level accuracy from fp16 tensor-core operands (|a| < 65504).
    ``overflow``: optional int32[1] device tensor, set to 1 when an element of ``a`` does not fit fp16.
    ``bias`` [N] / ``relu``: the nn.Linear epilogue fused into the GEMM's store."""
    require_cuda(a, b_hi, b_lo, overflow, bias)
    require_f32c(bias=bias)
    if a.dtype != torch.float32 or a.dim() != 2 or a.stride(1) != 1:
        raise ValueError("proj_gemm_3xf16: a must be float32 [M,K] with unit column stride")
    for t in (b_hi, b_lo):
        if t.dtype != torch.float16 or t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 8:
            raise ValueError("proj_gemm_3xf16: b_hi / b_lo must come from split_f16")
    m, k = a.shape
    n = b_hi.size(0)
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        if bias is None and not relu:
            check(lib().gvqa_proj_gemm_3xf16(ptr(a), a.stride(0) if m > 1 else k, ptr(b_hi), ptr(b_lo), b_hi.stride(0),
                                             ptr(out), out.stride(0), m, n, k, ptr(overflow), stream_handle(a.device)),
                  "gvqa_proj_gemm_3xf16")
        else:
            check(lib().gvqa_linear_3xf16(ptr(a), a.stride(0) if m > 1 else k, ptr(b_hi), ptr(b_lo), b_hi.stride(0),
                                          ptr(bias), 1 if relu else 0, ptr(out), out.stride(0), m, n, k, ptr(overflow),
                                          stream_handle(a.device)), "gvqa_linear_3xf16")
    return out


def proj_gemm_3xf16_batched(a, b_hi, b_lo, overflow=None):
    """out[z] = a[z] @ b[z]^T for a [Z,M,K] float32 (contiguous) and b_hi/b_lo [Z,N,K] float16 views of split_f16
    output reshaped to [Z, N, ld] (one launch for all z)."""
    require_cuda(a, b_hi, b_lo, overflow)
    if a.dtype != torch.float32 or a.dim() != 3 or not a.is_contiguous():
        raise ValueError("proj_gemm_3xf16_batched: a must be contiguous float32 [Z,M,K]")
    for t in (b_hi, b_lo):
        if t.dtype != torch.float16 or t.dim() != 3 or t.stride(2) != 1 or t.stride(1) % 8 or t.stride(0) % 8:
            raise ValueError("proj_gemm_3xf16_batched: b_hi / b_lo must be [Z,N,K] views of split_f16 output")
    z, m, k = a.shape
    n = b_hi.size(1)
    out = torch.empty(z, m, n, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().gvqa_proj_gemm_3xf16_batched(ptr(a), k, m * k, ptr(b_hi), ptr(b_lo), b_hi.stride(1), b_hi.stride(0),
                                                 ptr(out), n, m * n, m, n, k, z, ptr(overflow),
                                                 stream_handle(a.device)), "gvqa_proj_gemm_3xf16_batched")
    return out


def proj_gemm_3xf16_grouped(problems, overflow=None):
    """Up to three independent products in ONE persistent launch.  ``problems``: (a, b_hi, b_lo, out) tuples, each
    either 2-D (a [M,K], b [N,K], out [M,N] or None) or 3-D batched (a [Z,M,K] contiguous, b [Z,N,K] views of
    split_f16 output, out None or contiguous [Z,M,N]).  Returns the list of outputs."""
    if not 1 <= len(problems) <= MAX_GROUPED_PROBLEMS:
        raise ValueError("proj_gemm_3xf16_grouped: 1..%d problems" % MAX_GROUPED_PROBLEMS)
    arr = (GemmProblem * len(problems))()
    outs = []
    device = problems[0][0].device
    for q, prob in zip(arr, problems):
        a, b_hi, b_lo, out = prob[:4]
        bias, relu = (prob[4], prob[5]) if len(prob) > 4 else (None, False)      # optional Linear epilogue
        require_cuda(a, b_hi, b_lo, out, overflow, bias)
        require_f32c(bias=bias)
        q.bias, q.relu = ptr(bias), 1 if relu else 0
        if a.dtype != torch.float32 or a.dim() not in (2, 3) or a.stride(-1) != 1 or a.device != device:
            raise ValueError("proj_gemm_3xf16_grouped: a must be float32 [M,K] or [Z,M,K] with unit column stride")
        for t in (b_hi, b_lo):
            if t.dtype != torch.float16 or t.dim() != a.dim() or t.stride(-1) != 1 or any(s % 8 for s in t.stride()[:-1]):
                raise ValueError("proj_gemm_3xf16_grouped: b_hi / b_lo must come from split_f16")
        if a.dim() == 3:
            if not a.is_contiguous():
                raise ValueError("proj_gemm_3xf16_grouped: batched a must be contiguous")
            z, m, k = a.shape
            n = b_hi.size(1)
            if out is None:
                out = torch.empty(z, m, n, dtype=torch.float32, device=device)
            elif not out.is_contiguous() or out.shape != (z, m, n):
                raise ValueError("proj_gemm_3xf16_grouped: batched out must be contiguous [Z,M,N]")
            q.lda, q.stride_a, q.ldb, q.stride_b, q.ldc, q.stride_c = k, m * k, b_hi.stride(1), b_hi.stride(0), n, m * n
        else:
            (m, k), z = a.shape, 1
            n = b_hi.size(0)
            if out is None:
                out = torch.empty(m, (n + 3) // 4 * 4, dtype=torch.float32, device=device)[:, :n]
            q.lda, q.stride_a, q.ldb, q.stride_b, q.ldc, q.stride_c = (a.stride(0) if m > 1 else k), 0, b_hi.stride(0), 0, \
                out.stride(0), 0
        q.a, q.b_hi, q.b_lo, q.c, q.m, q.n, q.k, q.batch = ptr(a), ptr(b_hi), ptr(b_lo), ptr(out), m, n, k, z
        outs.append(out)
    with torch.cuda.device(device):
        check(lib().gvqa_proj_gemm_3xf16_grouped(arr, len(problems), ptr(overflow), stream_handle(device)),
              "gvqa_proj_gemm_3xf16_grouped")
    return outs


def l2_persist_limit(nbytes, device):
    with torch.cuda.device(device):
        r = lib().gvqa_device_set_l2_persist_limit(nbytes)
    if r < 0:
        check(r, "gvqa_device_set_l2_persist_limit")
    return r


def l2_window(tensor, device, hit_ratio=1.0):
    """Mark ``tensor`` (or clear with None) as L2-persisting for work enqueued on the current stream."""
    with torch.cuda.device(device):
        check(lib().gvqa_stream_set_l2_window(ptr(tensor), 0 if tensor is None else tensor.numel() * tensor.element_size(),
                                              hit_ratio, stream_handle(device)), "gvqa_stream_set_l2_window")


def gather_add_relu(a, b, c, bias, edge_index, relu=True):
    """out[k] = act(a[src_k] + b[dst_k] + c[k] + bias); edge_index [2,E] int64 (reference layout) or int32 (the
    loader-side wire format); a / b / c float32 [., F] with unit column stride (row-strided views are fine).
    ``edge_index=None``: no gather, out[k] = act(a[k] + b[k] + c[k] + bias) over the rows of ``a``."""
    require_cuda(a, b, c, bias, edge_index)
    require_f32c(bias=bias)
    for name, t in (("a", a), ("b", b), ("c", c)):
        if t is not None and (t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1):
            raise ValueError("gather_add_relu: %s must be float32 [., F] with unit column stride" % name)
    if edge_index is not None and (edge_index.dtype not in (torch.int64, torch.int32) or edge_index.dim() != 2
                                   or edge_index.size(0) != 2):
        raise TypeError("gather_add_relu: edge_index must be int64 or int32 [2, E]")
    e, f = (edge_index.size(1) if edge_index is not None else a.size(0)), a.size(1)
    out = torch.empty(e, f, dtype=torch.float32, device=a.device)
    ld = lambda t: 0 if t is None else (t.stride(0) if t.size(0) > 1 else f)
    with torch.cuda.device(a.device):
        check(lib().gvqa_gather_add_relu_strided_f32(ptr(a), ld(a), ptr(b), ld(b), ptr(c), ld(c), ptr(bias),
                                                     ptr(None if edge_index is None else edge_index.contiguous()),
                                                     8 if edge_index is None else edge_index.element_size(), ptr(out), e, f,
                                                     1 if relu else 0, stream_handle(a.device)),
              "gvqa_gather_add_relu_strided_f32")
    return out


def embedding_sum(table, tokens, sign=None):
    """out[n] = sign[n] * sum_t table[tokens[n, t]] for tokens [N, T] (or [N]) int32 / int64; table [V, F] float32."""
    require_cuda(table, tokens, sign)
    require_f32c(table=table, sign=sign)
    if tokens.dtype not in (torch.int32, torch.int64):
        raise TypeError("embedding_sum: tokens must be int32 or int64")
    tokens = tokens.contiguous()
    n = tokens.size(0)
    t = tokens.numel() // max(n, 1) if n else 1
    out = torch.empty(n, table.size(1), dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        check(lib().gvqa_embedding_sum_f32(ptr(table), table.size(0), ptr(tokens), tokens.element_size(), ptr(sign), ptr(out),
                                           n, max(t, 1), table.size(1), stream_handle(table.device)), "gvqa_embedding_sum_f32")
    return out


def affine_relu(x, scale, shift, relu=True, out=None):
    """out = act(x * scale[c] + shift[c]) for x [N,C] float32 contiguous (BatchNorm1d eval folded + ReLU)."""
    require_cuda(x, scale, shift)
    require_f32c(x=x, scale=scale, shift=shift, out=out)
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().gvqa_affine_relu_f32(ptr(x), ptr(scale), ptr(shift), ptr(out), x.size(0), x.size(1),
                                         1 if relu else 0, stream_handle(x.device)), "gvqa_affine_relu_f32")
    return out


def graph_scale_rows(x, q, node_graph, out=None):
    """out[n] = x[n] * q[node_graph[n]]  (x [N,C], q [B,C] float32 contiguous, node_graph int32 [N])."""
    require_cuda(x, q, node_graph)
    require_f32c(x=x, q=q, out=out)
    if node_graph.dtype != torch.int32:
        raise TypeError("graph_scale_rows: node_graph must be int32 (GraphCSR.node_graph)")
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().gvqa_graph_scale_rows_f32(ptr(x), ptr(q), ptr(node_graph), ptr(out), x.size(0), x.size(1),
                                              stream_handle(x.device)), "gvqa_graph_scale_rows_f32")
    return out


def attention_pool_gate(hid, w_gate, b_gate, x, graph_ptr, num_graphs):
    """out[g] = sum_n softmax_g(<hid[n], w_gate> + b_gate)[n] * x[n]: gate Linear(C,1) + per-graph softmax + pooling
    in one kernel.  Returns (out [B,C], gate [N])."""
    require_cuda(hid, w_gate, b_gate, x, graph_ptr)
    w_gate = w_gate.reshape(-1).contiguous().float()
    require_f32c(hid=hid, x=x)
    gate = torch.empty(max(x.size(0), 1), dtype=torch.float32, device=x.device)
    out = torch.empty(num_graphs, x.size(1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().gvqa_attention_pool_gate_f32(ptr(hid), hid.size(1), ptr(w_gate), ptr(b_gate), ptr(gate), ptr(x),
                                                 ptr(graph_ptr), ptr(out), num_graphs, x.size(1),
                                                 stream_handle(x.device)), "gvqa_attention_pool_gate_f32")
    return out, gate[:x.size(0)]


def build_csr_host(edge_index, batch, num_graphs, pin=True):
    """Loader-side CSR build on HOST tensors (int32 or int64 inputs): dict(rowptr, col_src, perm, graph_ptr,
    node_graph, stats) of int32 CPU tensors (pinned when a CUDA runtime is present and ``pin``), identical to what
    ``build_csr`` produces on the device.  Needs no GPU."""
    for t in (edge_index, batch):
        if t.is_cuda or t.dtype not in (torch.int32, torch.int64):
            raise TypeError("build_csr_host: CPU int32 / int64 tensors expected")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must be [2, E]")
    edge_index, batch = edge_index.contiguous(), batch.contiguous()
    n, e = batch.numel(), edge_index.size(1)
    pin = bool(pin) and torch.cuda.is_available()

    def buf(count):
        t = torch.empty(count, dtype=torch.int32)
        return t.pin_memory() if pin else t
    out = dict(rowptr=buf(n + 1), col_src=buf(max(e, 1)), perm=buf(max(e, 1)), graph_ptr=buf(num_graphs + 1),
               node_graph=buf(max(n, 1)), stats=buf(8))
    check(lib().gvqa_build_csr_host(ptr(edge_index), edge_index.element_size(), e, ptr(batch), batch.element_size(), n,
                                    num_graphs, ptr(out["rowptr"]), ptr(out["col_src"]), ptr(out["perm"]),
                                    ptr(out["graph_ptr"]), ptr(out["node_graph"]), ptr(out["stats"])),
          "gvqa_build_csr_host")
    return out


def segment_mean_rows(values, csr, mean=True):
    """scatter_mean (or sum) of per-edge rows by target node over the destination-CSR."""
    require_cuda(values)
    require_f32c(values=values)
    n, f = csr["rowptr"].numel() - 1, values.size(1)
    out = torch.empty(n, f, dtype=torch.float32, device=values.device)
    with torch.cuda.device(values.device):
        check(lib().gvqa_segment_mean_rows_f32(ptr(values), ptr(csr["perm"]), ptr(csr["rowptr"]), ptr(out), n, f,
                                               1 if mean else 0, stream_handle(values.device)),
              "gvqa_segment_mean_rows_f32")
    return out


def attention_pool(gate, x, graph_ptr, num_graphs):
    """out[g] = sum_n softmax_g(gate)[n] * x[n]  (PyG softmax semantics)."""
    require_cuda(gate, x, graph_ptr)
    gate = gate.reshape(-1).contiguous().float()
    require_f32c(x=x)
    out = torch.empty(num_graphs, x.size(1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().gvqa_attention_pool_f32(ptr(gate), ptr(x), ptr(graph_ptr), ptr(out), num_graphs, x.size(1),
                                            stream_handle(x.device)), "gvqa_attention_pool_f32")
    return out

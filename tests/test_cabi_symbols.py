"""CPU: the C-ABI library builds/loads and exports every symbol include/gvqa_b200.h declares,
and the ctypes binding lists exactly those symbols (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from graphvqa_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gvqa_b200.h")).read()
    return sorted(set(re.findall(r"GVQA_API\s+[\w\s\*]+?\b(gvqa_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def handle():
    if not os.path.exists(_cabi.LIB_PATH):
        from graphvqa_b200 import build
        build.build()
    return ctypes.CDLL(_cabi.LIB_PATH)


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "gvqa_gat_hop_f32" in syms and "gvqa_build_csr" in syms and len(syms) >= 7


def test_library_exports_every_declared_symbol(handle):
    for name in declared_symbols():
        assert hasattr(handle, name), "libgvqa_b200.so does not export %s" % name


def test_binding_covers_header_exactly():
    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_abi_version_and_error_strings(handle):
    lib = _cabi.lib()
    assert lib.gvqa_abi_version() == _cabi.ABI_VERSION
    assert lib.gvqa_error_string(0) == b"ok"
    assert b"NULL" in lib.gvqa_error_string(-1)
    assert lib.gvqa_csr_workspace_bytes(10, 20) >= 10 * 4 + 20 * 4
    assert lib.gvqa_csr_workspace_bytes(-1, 0) == 0


def test_argument_errors_are_reported_before_any_launch():
    # host-side validation only: no device is touched for these calls
    lib = _cabi.lib()
    assert lib.gvqa_gat_hop_f32(None, None) == -1
    a = _cabi.GatHopArgs()
    a.heads, a.channels, a.num_nodes, a.ldx, a.lde = 4, 30, 8, 120, 4        # channels % 4 != 0
    a.x_l = a.a_node = a.rowptr = a.node_graph = a.h_out = 256
    assert lib.gvqa_gat_hop_f32(ctypes.byref(a), None) == -3
    a.channels, a.ldx = 32, 64                                                # ldx < H*C
    assert lib.gvqa_gat_hop_f32(ctypes.byref(a), None) == -2
    a.ldx, a.heads = 96, 3                                                    # heads not in {1,2,4,8}
    assert lib.gvqa_gat_hop_f32(ctypes.byref(a), None) == -3
    assert lib.gvqa_skinny_matmul_f32(256, 8, 256, 256, 4, 8, 64, None) == -3  # k > 32
    assert lib.gvqa_graph_layernorm_f32(None, None, None, None, None, 4, 1, 8, 1e-5, 0, None) == -1


def test_product_fails_loudly_on_cpu_tensors():
    import torch
    from graphvqa_b200.gat_skip import gat_seq
    m = gat_seq(8, 8, 8, 4, 2).eval()
    ei = torch.tensor([[0, 1], [1, 0]])
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 8), ei, torch.randn(2, 8), torch.randn(2, 1, 4), torch.zeros(2, dtype=torch.long))
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.randn(2, 8), ei, torch.randn(2, 8), torch.randn(2, 1, 4), torch.zeros(2, dtype=torch.long))

"""CPU: the C-ABI library builds/loads and exports every symbol include/gvqa_b200.h declares,
and the ctypes binding lists exactly those symbols (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from graphvqa_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="gvqa_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"GVQA_API\s+[\w\s\*]+?\b(gvqa_\w+)\s*\(", text)))


def all_declared_symbols():
    return sorted(set(sum((declared_symbols(h) for h in os.listdir(os.path.join(ROOT, "include")) if h.endswith(".h")), [])))


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
    for name in all_declared_symbols():
        assert hasattr(handle, name), "libgvqa_b200.so does not export %s" % name


def test_binding_covers_header_exactly():
    assert sorted(_cabi.SIGNATURES) == all_declared_symbols()
    # the production header carries no debug hooks; they live in gvqa_b200_debug.h
    assert not [n for n in declared_symbols() if "debug" in n]
    assert sorted(n for n in _cabi.SIGNATURES if "debug" in n) == declared_symbols("gvqa_b200_debug.h")


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


def test_header_is_plain_c_and_struct_layouts_match_the_binding(tmp_path):
    """include/gvqa_b200.h compiles as C (gcc, no CUDA headers) and the two argument structs have the size and
    field offsets the ctypes mirrors assume."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    classes = (("gvqa_gat_hop_args", _cabi.GatHopArgs), ("gvqa_gemm_problem", _cabi.GemmProblem),
               ("gvqa_gat_fused_args", _cabi.GatFusedArgs))
    fields = {struct: [n for n, _ in cls._fields_] for struct, cls in classes}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "gvqa_b200.h"', 'int main(void) {']
    for struct, names in fields.items():
        src.append('  printf("%s %%zu", sizeof(struct %s));' % (struct, struct))
        for n in names:
            src.append('  printf(" %%zu", offsetof(struct %s, %s));' % (struct, n))
        src.append('  printf("\\n");')
    src += ['  return 0;', '}']
    c_file = tmp_path / "layout.c"
    c_file.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c_file), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line, (struct, cls) in zip(out, classes):
        got = line.split()
        assert got[0] == struct
        want = [ctypes.sizeof(cls)] + [getattr(cls, n).offset for n, _ in cls._fields_]
        assert [int(v) for v in got[1:]] == want, struct


def test_fused_hop_argument_errors_are_reported_before_any_launch():
    lib = _cabi.lib()
    assert lib.gvqa_gat_fused_hop_f32(None, None) == -1
    assert lib.gvqa_gat_fused_supported(4, 512, 512) == 1 and lib.gvqa_gat_fused_supported(8, 512, 512) == 0
    assert lib.gvqa_gat_fused_supported(4, 30, 32) == 0                        # widths must be multiples of 4
    assert lib.gvqa_gat_fused_window(30) == 128 and lib.gvqa_gat_fused_window(200) == 256
    assert lib.gvqa_gat_fused_part_blocks(7680, 512) == 6 and lib.gvqa_gat_fused_pack_halves(4, 512, 300) == 512 * 4 * 304 * 2
    a = _cabi.GatFusedArgs()
    a.num_nodes, a.in_channels, a.channels, a.heads, a.window, a.w_scale = 8, 64, 64, 8, 128, 1.0
    assert lib.gvqa_gat_fused_hop_f32(ctypes.byref(a), None) == -3           # heads 8: the split path's job
    a.heads, a.window = 4, 100
    assert lib.gvqa_gat_fused_hop_f32(ctypes.byref(a), None) == -3           # window must be 128 or 256
    a.window = 128
    assert lib.gvqa_gat_fused_hop_f32(ctypes.byref(a), None) == -1           # NULL operands
    a.h_in = a.w_pack = a.tiles = a.tile_count = a.rowptr = a.col_src = a.node_graph = a.alpha = a.h_out = 256
    a.epilogue = 2
    assert lib.gvqa_gat_fused_hop_f32(ctypes.byref(a), None) == -1           # affine epilogue without scale / shift
    a.epilogue, a.h_out = 0, 260
    assert lib.gvqa_gat_fused_hop_f32(ctypes.byref(a), None) == -4           # not 16-byte aligned
    a.h_out, a.num_nodes = 256, 0
    assert lib.gvqa_gat_fused_hop_f32(ctypes.byref(a), None) == 0            # nothing to do
    assert lib.gvqa_gat_fused_plan(None, 4, 128, None, None, 8, None) == -1
    assert lib.gvqa_gat_fused_plan(256, 4, 100, 256, 256, 8, None) == -2


def test_fused_plan_on_the_host():
    import torch
    gp = torch.tensor([0, 30, 60, 90, 120, 150, 400, 401], dtype=torch.int32)
    tiles, count = _cabi.fused_plan_host(gp, 401, 7, 256)
    assert int(count) == 5          # four graphs of 30 | one of 30 | a 250-node graph in two chunks (window = the graph) | 1
    assert tiles[:5].tolist() == [[0, 120, 0, 0], [120, 30, 120, 0], [150, 128, 150, 0], [278, 122, 150, 0], [400, 1, 400, 0]]


def test_grouped_gemm_argument_errors_are_reported_before_any_launch():
    lib = _cabi.lib()
    q = (_cabi.GemmProblem * 4)()
    assert lib.gvqa_proj_gemm_3xf16_grouped(None, 1, None, None) == -2
    assert lib.gvqa_proj_gemm_3xf16_grouped(q, 0, None, None) == -2
    assert lib.gvqa_proj_gemm_3xf16_grouped(q, 4, None, None) == -2
    q[0].m, q[0].n, q[0].k, q[0].batch, q[0].lda, q[0].ldb, q[0].ldc = 128, 128, 64, 1, 64, 64, 128
    assert lib.gvqa_proj_gemm_3xf16_grouped(q, 1, None, None) == -1          # NULL operands
    q[0].a = q[0].b_hi = q[0].b_lo = q[0].c = 256
    q[0].k, q[0].lda, q[0].ldb = 62, 62, 62                                  # k not a multiple of 4
    assert lib.gvqa_proj_gemm_3xf16_grouped(q, 1, None, None) == -3
    q[0].k, q[0].lda, q[0].ldb = 64, 64, 64
    q[0].a = 260                                                             # not 16-byte aligned
    assert lib.gvqa_proj_gemm_3xf16_grouped(q, 1, None, None) == -4
    q[0].a, q[0].ldc = 256, 64                                               # ldc < n
    assert lib.gvqa_proj_gemm_3xf16_grouped(q, 1, None, None) == -2
    q[0].ldc, q[0].m = 128, 0                                                # nothing to do
    assert lib.gvqa_proj_gemm_3xf16_grouped(q, 1, None, None) == 0

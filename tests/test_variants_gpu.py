"""GPU parity tests (through the C ABI) of graph LayerNorm and the GCN / GINE / LCGN variants
against the CPU oracle and the golden fixtures generated from the reference's own code."""
import pytest
import torch

from conftest import random_graphs
from graphvqa_b200 import gcn_gine, lcgn
from graphvqa_b200.my_graph_layernorm import LayerNorm
from oracle import graphvqa_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


def _load(cls, fx):
    m = cls(**fx["config"]).eval()
    m.load_state_dict(fx["state"])
    return m.to(DEV)


# ---------------------------------------------------------------- graph LayerNorm -------------
def test_layernorm_golden(golden):
    fx = golden("graph_layernorm")
    m = LayerNorm(300)
    m.load_state_dict(fx["state"])
    m = m.to(DEV).eval()
    with torch.no_grad():
        out = m(fx["x"].to(DEV), fx["batch"].to(DEV)).cpu()
    assert (out - fx["out"]).abs().max() <= TOL


@pytest.mark.parametrize("f,graphs,n_hi", [(300, 7, 40), (512, 33, 30), (512, 3, 260), (8, 5, 3)])
def test_layernorm_matches_oracle(f, graphs, n_hi):
    _, batch = random_graphs(graphs, 1, n_hi, 0.0, seed=f + graphs)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch.numel(), f, generator=g) * 3 + 1
    m = LayerNorm(f)
    with torch.no_grad():
        m.weight.fill_(0.7); m.bias.fill_(0.3)
        want = orc.graph_layernorm(x, batch, graphs, m.weight, m.bias, 1e-5)
        m = m.to(DEV).eval()
        got = m(x.to(DEV), batch.to(DEV), num_graphs=graphs).cpu()
        # single-node graph: variance 0 -> 0 / (0 + eps) * w + b
        one = m(torch.full((1, f), 2.5, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV), num_graphs=1)
    assert (want - got).abs().max() <= TOL
    assert torch.allclose(one.cpu(), torch.full((1, f), 0.3))


def test_layernorm_no_batch_and_no_affine():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(37, 64, generator=g)
    m = LayerNorm(64, affine=False).to(DEV).eval()
    with torch.no_grad():
        got = m(x.to(DEV)).cpu()
    want = orc.graph_layernorm(x, None, None)
    assert (want - got).abs().max() <= TOL


# ---------------------------------------------------------------- GINE / GCN -------------------
@pytest.mark.parametrize("name", ["gine", "gcn"])
def test_gine_gcn_seq_golden(golden, name):
    fx = golden(name + "_seq_small")
    m = _load(getattr(gcn_gine, name + "_seq"), fx)
    keys = ("x", "edge_index", "edge_attr", "instr_vectors", "batch") if name == "gine" else \
        ("x", "edge_index", "instr_vectors", "batch")
    with torch.no_grad():
        out, conv = m(*[fx[k].to(DEV) for k in keys], return_conv=True)
        out_fast = m(*[fx[k].to(DEV) for k in keys])
    assert (out.cpu() - fx["out"]).abs().max() <= TOL
    assert torch.equal(out, out_fast)                      # dead convs skipped: same result
    for a, b in zip(conv, fx["conv_out"]):
        assert (a.cpu() - b).abs().max() <= TOL


@pytest.mark.parametrize("f,d", [(300, 512), (64, 32)])
def test_gine_gcn_conv_matches_oracle_and_fixed_mode(f, d):
    b = 5
    ei, batch = random_graphs(b, 1, 20, 2.0, seed=3, isolated=True)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(batch.numel(), f, generator=g); ea = torch.randn(ei.size(1), f, generator=g)
    ins = torch.randn(5, b, d, generator=g)
    for name in ("gine", "gcn"):
        torch.manual_seed(5)
        o = getattr(orc, name + "_seq")(f, f, d, bug_faithful=False).eval()
        e = getattr(gcn_gine, name + "_seq")(f, f, d, bug_faithful=False).eval()
        e.load_state_dict(o.state_dict())
        e = e.to(DEV)
        with torch.no_grad():
            if name == "gine":
                want = o(x, ei, ea, ins, batch)
                got = e(x.to(DEV), ei.to(DEV), ea.to(DEV), ins.to(DEV), batch.to(DEV)).cpu()
            else:
                want = o(x, ei, ins, batch)
                got = e(x.to(DEV), ei.to(DEV), ins.to(DEV), batch.to(DEV)).cpu()
        scale = max(1.0, float(want.abs().max()))
        assert (want - got).abs().max() <= TOL * scale, name


def test_gcn_conv_standalone_self_loops_and_duplicates():
    # existing self loops are replaced by exactly one loop per node, duplicate edges are counted
    ei = torch.tensor([[0, 0, 1, 2, 2, 3], [1, 1, 1, 0, 2, 0]])
    torch.manual_seed(0)
    o = orc.GCNConv(8, 12).eval(); e = gcn_gine.GCNConv(8, 12).eval()
    x = torch.randn(5, 8)
    with torch.no_grad():
        o.bias.normal_(); e.load_state_dict(o.state_dict())
        want = o(x, ei)
        got = e.to(DEV)(x.to(DEV), ei.to(DEV)).cpu()
    assert (want - got).abs().max() <= 1e-5


# ---------------------------------------------------------------- LCGN ------------------------
def test_lcgn_seq_golden(golden):
    fx = golden("lcgn_seq_small")
    m = _load(lcgn.lcgn_seq, fx)
    with torch.no_grad():
        out = m(fx["x"].to(DEV), fx["edge_index"].to(DEV), fx["batch"].to(DEV), fx["q_encoding"].to(DEV),
                fx["lstm_outputs"].to(DEV), x_ctx_init=fx["x_ctx"]).cpu()
    assert (out - fx["out"]).abs().max() <= TOL


def test_lcgn_seq_reference_dims_and_rng_draw():
    b = 6
    ei, batch = random_graphs(b, 1, 30, 2.0, seed=9, isolated=True)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(batch.numel(), 300, generator=g)
    q = torch.randn(b, 512, generator=g); lo = torch.randn(12, b, 512, generator=g)
    torch.manual_seed(11)
    o = orc.lcgn_seq(300, 512, 300, 5).eval()
    e = lcgn.lcgn_seq(300, 512, 300, 5).eval()
    e.load_state_dict(o.state_dict())
    e = e.to(DEV)
    with torch.no_grad():
        torch.manual_seed(12); want = o(x, ei, batch, q, lo)                  # reference-style CPU randn draw
        torch.manual_seed(12); got = e(x.to(DEV), ei.to(DEV), batch.to(DEV), q.to(DEV), lo.to(DEV)).cpu()
    assert (want - got).abs().max() <= TOL * max(1.0, float(want.abs().max()))


# ---------------------------------------------------------------- BASELINE-size parity (cfg3, cfg5) ------------------
def test_gine_and_gcn_conv_at_cfg3_size():
    """BASELINE cfg3: the cfg2 batch (256 graphs x 30 nodes / 60 edges, F=512, D=512) through GINEConv / GCNConv --
    the conv results the reference computes (and discards), against the CPU oracle at full size."""
    from graphvqa_b200.graph_batch import synthetic_topology
    b, f, d = 256, 512, 512
    ei, batch, _ = synthetic_topology(b, 30, 60, seed=1234)
    g = torch.Generator().manual_seed(31)
    x = torch.randn(batch.numel(), f, generator=g); ea = torch.randn(ei.size(1), f, generator=g)
    ins = torch.randn(5, b, d, generator=g)
    for name in ("gine", "gcn"):
        torch.manual_seed(32)
        o = getattr(orc, name + "_seq")(f, f, d).eval()
        e = getattr(gcn_gine, name + "_seq")(f, f, d).eval()
        e.load_state_dict(o.state_dict())
        e = e.to(DEV)
        with torch.no_grad():
            if name == "gine":
                want, want_conv = o(x, ei, ea, ins, batch, return_conv=True)
                got, got_conv = e(x.to(DEV), ei.to(DEV), ea.to(DEV), ins.to(DEV), batch.to(DEV), return_conv=True)
            else:
                want, want_conv = o(x, ei, ins, batch, return_conv=True)
                got, got_conv = e(x.to(DEV), ei.to(DEV), ins.to(DEV), batch.to(DEV), return_conv=True)
        assert (want - got.cpu()).abs().max() <= TOL, name                       # bug-faithful sequence output
        for i, (a, c) in enumerate(zip(want_conv, got_conv)):
            scale = max(1.0, float(a.abs().max()))
            assert (a - c.cpu()).abs().max() <= TOL * scale, "%s conv %d" % (name, i)


def test_lcgn_seq_at_cfg5_per_gpu_size():
    """BASELINE cfg5 per GPU: lcgn_seq(300 -> 512), 128 graphs x 30 nodes / 60 edges, L = 12 question tokens,
    4 iterations, x_ctx injected; against the CPU oracle at full size."""
    from graphvqa_b200.graph_batch import synthetic_topology
    b = 128
    ei, batch, _ = synthetic_topology(b, 30, 60, seed=77)
    g = torch.Generator().manual_seed(41)
    x = torch.randn(batch.numel(), 300, generator=g)
    q = torch.randn(b, 512, generator=g); lo = torch.randn(12, b, 512, generator=g)
    x_ctx = torch.randn(batch.numel(), 512, generator=g)
    torch.manual_seed(42)
    o = orc.lcgn_seq(300, 512, 300, 5).eval()
    e = lcgn.lcgn_seq(300, 512, 300, 5).eval()
    e.load_state_dict(o.state_dict())
    e = e.to(DEV)
    with torch.no_grad():
        want = o(x, ei, batch, q, lo, x_ctx_init=x_ctx)
        got = e(x.to(DEV), ei.to(DEV), batch.to(DEV), q.to(DEV), lo.to(DEV), x_ctx_init=x_ctx.to(DEV)).cpu()
    assert (want - got).abs().max() <= TOL * max(1.0, float(want.abs().max()))


def test_affine_relu_kernel_is_batchnorm_eval_plus_relu():
    g = torch.Generator().manual_seed(5)
    bn = torch.nn.BatchNorm1d(300).eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.3, generator=g); bn.running_var.uniform_(0.5, 1.5, generator=g)
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.normal_(0, 0.1, generator=g)
    x = torch.randn(777, 300, generator=g)
    want = torch.relu(bn(x))
    inv = torch.rsqrt(bn.running_var.double() + bn.eps)
    scale = (bn.weight.double() * inv).float(); shift = (bn.bias.double() - bn.running_mean.double() * bn.weight.double() * inv).float()
    from graphvqa_b200 import _cabi
    got = _cabi.affine_relu(x.to(DEV), scale.to(DEV), shift.to(DEV)).cpu()
    assert (got - want.detach()).abs().max() <= 2e-6

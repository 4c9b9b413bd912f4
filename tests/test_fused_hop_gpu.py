"""GPU parity (through the C ABI) of the fused aggregate-project hop -- gvqa_gat_fused_* in include/gvqa_b200.h:
weight prepack layout, row-tile plan (device == host == a Python restatement), softmax weights, one hop against a
float64 restatement of gat_skip.py:133-168 + :270-275, and gat_seq in hop_mode "fused" against the CPU oracle at the
BASELINE shapes.  Tolerance: 1e-4 absolute fp32 (BASELINE.json north_star); plan and packing bit-exact."""
import pytest
import torch

from conftest import random_graphs
from graphvqa_b200 import _cabi
from graphvqa_b200 import gat_skip as eng
from graphvqa_b200.graph_batch import GraphCSR, synthetic_topology
from oracle import graphvqa_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


def _plan_reference(graph_ptr, win):
    """Greedy packing of whole graphs into <= 128 rows; larger graphs cut into 128-row chunks."""
    gp, tiles, g, b = graph_ptr, [], 0, len(graph_ptr) - 1
    while g < b:
        r0, n_g = gp[g], gp[g + 1] - gp[g]
        if n_g <= 0:
            g += 1
            continue
        if n_g > 128:
            for s in range(r0, r0 + n_g, 128):
                w0 = r0 if n_g <= win else max(r0, min(s - (win - 128) // 2, r0 + n_g - win))
                tiles.append([s, min(128, r0 + n_g - s), w0, 0])
            g += 1
            continue
        end = g + 1
        while end < b and gp[end + 1] - r0 <= 128:
            end += 1
        tiles.append([r0, gp[end] - r0, r0, 0])
        g = end
    return tiles


@pytest.mark.parametrize("sizes,win", [([30] * 256, 128), ([200] * 8, 256), ([1, 0, 127, 1, 1, 300, 5, 128, 129, 0, 3], 128),
                                       ([700, 2, 2], 256), ([], 128)])
def test_plan_device_host_and_reference_agree(sizes, win):
    gp = torch.zeros(len(sizes) + 1, dtype=torch.int32)
    if sizes:
        gp[1:] = torch.tensor(sizes).cumsum(0)
    n, b = int(gp[-1]), len(sizes)
    want = _plan_reference(gp.tolist(), win)
    tiles_h, count_h = _cabi.fused_plan_host(gp, n, b, win)
    tiles_d, count_d = _cabi.fused_plan(gp.to(DEV), n, b, win)
    batch = torch.repeat_interleave(torch.arange(b), torch.tensor(sizes, dtype=torch.long)) if sizes else torch.zeros(0, dtype=torch.long)
    tiles_b, count_b = _cabi.fused_plan_from_batch(batch.to(DEV), b, win)       # straight from the batch vector
    assert int(count_h) == int(count_d) == int(count_b) == len(want)
    assert tiles_h[:len(want)].tolist() == want
    assert tiles_d.cpu()[:len(want)].tolist() == want
    assert tiles_b.cpu()[:len(want)].tolist() == want
    covered = sorted((t[0], t[0] + t[1]) for t in want)       # the tiles partition [0, N)
    assert [c[0] for c in covered] == [0] + [c[1] for c in covered[:-1]] if covered else n == 0
    assert not covered or covered[-1][1] == n


@pytest.mark.parametrize("heads,c,f", [(4, 64, 64), (2, 36, 20), (4, 128, 300)])
def test_pack_layout(heads, c, f):
    g = torch.Generator().manual_seed(c)
    w = torch.randn(heads * c, f + 12, generator=g) * 0.03
    packed, scale = _cabi.fused_pack(w.to(DEV), heads, c, f)
    assert 256.0 <= float(w[:, :f].abs().max()) * scale < 512.0
    fp = (f + 15) // 16 * 16
    wp = torch.zeros(heads, c, fp)
    wp[:, :, :f] = w[:, :f].view(heads, c, -1) * scale
    hi = wp.half()
    lo = (wp - hi.float()).half()
    # [C, Fp/16, H, (16 hi | 16 lo)]
    want = torch.stack([hi.view(heads, c, fp // 16, 16), lo.view(heads, c, fp // 16, 16)], 3).permute(1, 2, 0, 3, 4)
    assert torch.equal(packed.cpu().view(c, fp // 16, heads, 2, 16), want.contiguous())


def _alpha_reference(a_node, a_edge, a_graph, ei, batch, heads, slope):
    src, dst = ei
    logit = a_node[src, :heads] + a_node[dst, heads:2 * heads] + a_edge[:, :heads] + a_graph[batch[dst]]
    logit = torch.where(logit > 0, logit, logit * slope).double()
    n = a_node.size(0)
    mx = torch.full((n, heads), -float("inf"), dtype=torch.float64).scatter_reduce(
        0, dst[:, None].expand(-1, heads), logit, "amax", include_self=True)
    ex = (logit - mx[dst]).exp()
    den = torch.zeros(n, heads, dtype=torch.float64).index_add_(0, dst, ex)
    return ex / (den[dst] + 1e-16)


def _hop_case(graphs, n_lo, n_hi, extra, heads, f, c, seed):
    ei, batch = random_graphs(graphs, n_lo, n_hi, extra, seed=seed, isolated=True)
    g = torch.Generator().manual_seed(seed + 100)
    n, e = batch.numel(), ei.size(1)
    t = dict(ei=ei, batch=batch, h=torch.randn(n, f, generator=g), w=torch.randn(heads * c, f, generator=g) / f ** 0.5,
             a_node=torch.randn(n, 2 * heads, generator=g), a_edge=torch.randn(e, heads, generator=g),
             a_graph=torch.randn(graphs, heads, generator=g), gb=torch.randn(graphs, c, generator=g),
             bias=torch.randn(c, generator=g), skip=torch.randn(n, c, generator=g),
             scale=torch.rand(c, generator=g) + 0.5, shift=torch.randn(c, generator=g))
    return t


def _hop_reference(t, heads, c, epilogue):
    src, dst = t["ei"]
    n = t["h"].size(0)
    alpha = _alpha_reference(t["a_node"], t["a_edge"], t["a_graph"], t["ei"], t["batch"], heads, 0.2)
    x_l = (t["h"].double() @ t["w"].double().t()).view(n, heads, c)
    agg = torch.zeros(n, heads, c, dtype=torch.float64).index_add_(0, dst, alpha[:, :, None] * x_l[src])
    out = agg.mean(1)
    has_in = torch.zeros(n, dtype=torch.bool)
    has_in[dst] = True
    out = out + has_in[:, None] * t["gb"].double()[t["batch"]] + t["bias"].double() + t["skip"].double()
    if epilogue != _cabi.EPI_NONE:
        out = out * t["scale"].double() + t["shift"].double()
        if epilogue == _cabi.EPI_AFFINE_RELU:
            out = out.clamp_min(0)
    return alpha, out


@pytest.mark.parametrize("graphs,n_lo,n_hi,extra,heads,f,c,win,epilogue", [
    (6, 1, 24, 2.0, 4, 64, 64, 128, 2),          # several graphs per tile, in-degree-0 nodes, multi-edges
    (40, 5, 40, 2.0, 4, 300, 300, 128, 2),       # reference dims: ragged k-slice (300 = 9 x 32 + 12) and column tile
    (3, 150, 220, 3.0, 4, 128, 128, 256, 0),     # graphs larger than a tile: chunks with the whole graph as window
    (2, 290, 300, 2.0, 2, 64, 96, 256, 1),       # graphs larger than the window: sources read from global memory
    (3, 150, 220, 3.0, 2, 36, 20, 128, 2),       # window hint too small for the graphs (generic path), tiny widths
    (70, 20, 40, 2.0, 4, 512, 512, 128, 2),      # two column blocks of 256, several items per pair
    (300, 1, 3, 1.0, 4, 32, 32, 128, 2),         # many tiny graphs
    (190, 90, 128, 1.5, 4, 64, 64, 128, 2),      # ~95 pair tiles on 74 pairs: several items per pair (accumulator
                                                 # hand-over between items, helper warps on the last item only)
    (100, 129, 200, 2.0, 2, 32, 160, 256, 1),    # two rounds of two-tile graphs, window 256, ragged column block
])
def test_fused_hop_matches_float64_reference(graphs, n_lo, n_hi, extra, heads, f, c, win, epilogue):
    t = _hop_case(graphs, n_lo, n_hi, extra, heads, f, c, seed=graphs)
    d = {k: v.to(DEV) for k, v in t.items()}
    n = t["h"].size(0)
    csr = GraphCSR.build(d["ei"], d["batch"], graphs)
    alpha = _cabi.gat_alpha(d["a_node"], d["a_edge"], csr.as_dict(), heads, a_graph=d["a_graph"])
    want_alpha, want = _hop_reference(t, heads, c, epilogue)
    perm = csr.perm.cpu().long()[:t["ei"].size(1)]
    assert (alpha.cpu()[:perm.numel()].double() - want_alpha[perm]).abs().max() <= 1e-5
    out = torch.full((n, c), float("nan"), device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    v_next = torch.randn(2 * heads, c, generator=torch.Generator().manual_seed(1)) / c ** 0.5
    a_part = torch.full((_cabi.fused_part_blocks(n, c), n, 2 * heads), float("nan"), device=DEV)
    _cabi.gat_fused_hop(d["h"], _cabi.fused_pack(d["w"], heads, c, f), csr.fused_plan(win), csr.as_dict(), alpha, heads, c,
                        out, window=win, skip=d["skip"], graph_bias=d["gb"], bias=d["bias"], ep_scale=d["scale"],
                        ep_shift=d["shift"], epilogue=epilogue, overflow=flag, v_next=v_next.to(DEV), a_part=a_part)
    err = (out.cpu().double() - want).abs().max()
    assert err <= 2e-5 * max(1.0, float(want.abs().max())), "max|d| = %g" % err
    assert int(flag) == 0
    # the next hop's node logits, emitted by the epilogue as partial sums per 128-column block
    a_next = a_part.cpu().double().sum(0)
    assert (a_next - want @ v_next.double().t()).abs().max() <= 1e-4
    # ... and consumed in that form by the softmax-weight kernel
    alpha2 = _cabi.gat_alpha(a_part, d["a_edge"], csr.as_dict(), heads, a_graph=d["a_graph"])
    alpha2_ref = _cabi.gat_alpha(a_part.sum(0), d["a_edge"], csr.as_dict(), heads, a_graph=d["a_graph"])
    assert (alpha2 - alpha2_ref).abs().max() <= 1e-6
    # the same hop with the softmax inside the kernel (logit terms + node logits instead of precomputed weights)
    terms = _cabi.fused_logit_terms(csr.as_dict(), d["a_edge"], d["a_graph"][None], 1, heads, n)
    out2 = torch.full((n, c), float("nan"), device=DEV)
    scratch = torch.zeros_like(alpha)
    _cabi.gat_fused_hop(d["h"], _cabi.fused_pack(d["w"], heads, c, f), csr.fused_plan(win), csr.as_dict(), scratch, heads, c,
                        out2, window=win, skip=d["skip"], graph_bias=d["gb"], bias=d["bias"], ep_scale=d["scale"],
                        ep_shift=d["shift"], epilogue=epilogue, logit_terms=terms[0], a_node=d["a_node"])
    err2 = (out2.cpu().double() - want).abs().max()
    assert err2 <= 2e-5 * max(1.0, float(want.abs().max())), "in-kernel softmax: max|d| = %g" % err2


def test_fused_hop_hub_exceeds_the_staged_edge_list():
    """A tile with more in-edges than the kernel stages in shared memory takes the generic path."""
    n = 120
    src = list(range(n)) + [i % n for i in range(3000)]
    dst = list(range(n)) + [i % 7 for i in range(3000)]
    ei = torch.tensor([src, dst])
    t = _hop_case(1, n, n, 0.0, 4, 64, 64, seed=77)
    g = torch.Generator().manual_seed(5)
    t["ei"], t["a_edge"] = ei, torch.randn(ei.size(1), 4, generator=g)
    d = {k: v.to(DEV) for k, v in t.items()}
    csr = GraphCSR.build(d["ei"], d["batch"], 1)
    alpha = _cabi.gat_alpha(d["a_node"], d["a_edge"], csr.as_dict(), 4, a_graph=d["a_graph"])
    _, want = _hop_reference(t, 4, 64, 2)
    terms = _cabi.fused_logit_terms(csr.as_dict(), d["a_edge"], d["a_graph"][None], 1, 4, n)
    for kw in (dict(), dict(logit_terms=terms[0], a_node=d["a_node"])):      # weights given / softmax inside the kernel
        out = torch.empty(n, 64, device=DEV)
        _cabi.gat_fused_hop(d["h"], _cabi.fused_pack(d["w"], 4, 64, 64), csr.fused_plan(128), csr.as_dict(),
                            alpha if not kw else torch.zeros_like(alpha), 4, 64, out, window=128, skip=d["skip"],
                            graph_bias=d["gb"], bias=d["bias"], ep_scale=d["scale"], ep_shift=d["shift"], epilogue=2, **kw)
        assert (out.cpu().double() - want).abs().max() <= 2e-5 * max(1.0, float(want.abs().max()))


def test_fused_hop_flags_inputs_outside_fp16_range():
    t = _hop_case(4, 10, 20, 2.0, 4, 64, 64, seed=3)
    t["h"][5, 7] = 1e6
    d = {k: v.to(DEV) for k, v in t.items()}
    csr = GraphCSR.build(d["ei"], d["batch"], 4)
    alpha = _cabi.gat_alpha(d["a_node"], d["a_edge"], csr.as_dict(), 4, a_graph=d["a_graph"])
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = torch.empty(t["h"].size(0), 64, device=DEV)
    _cabi.gat_fused_hop(d["h"], _cabi.fused_pack(d["w"], 4, 64, 64), csr.fused_plan(128), csr.as_dict(), alpha, 4, 64, out,
                        window=128, overflow=flag)
    assert int(flag) == 1


def _pair(cfg, seed):
    torch.manual_seed(seed)
    o = orc.gat_seq(**cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for bn in o.bns:
        bn.running_mean.normal_(0, 0.1, generator=g); bn.running_var.uniform_(0.5, 1.5, generator=g)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.normal_(0, 0.1, generator=g)
    with torch.no_grad():
        for c in o.convs:
            c.bias.normal_(0, 0.1, generator=g)
    e = eng.gat_seq(**cfg).eval()
    e.load_state_dict(o.state_dict())
    e = e.to(DEV)
    e.hop_mode = "fused"
    return o, e


def _inputs(ei, batch, b, f, fe, d, hops, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(batch.numel(), f, generator=g), ei, torch.randn(ei.size(1), fe, generator=g),
            torch.randn(hops, b, d, generator=g), batch)


def _check(o, e, args, hints=True, own_csr=True):
    with torch.no_grad():
        want, want_hops = o(*args, return_hops=True)
        dargs = [a.to(DEV) for a in args]
        csr = GraphCSR.build(dargs[1], dargs[4], args[3].size(1), read_hints=hints) if own_csr else None
        got, got_hops = e(*dargs, csr=csr, return_hops=True)
    for i, (a, c) in enumerate(zip(want_hops, got_hops)):
        assert (a - c.cpu()).abs().max() <= TOL, "hop %d: max|d|=%g" % (i, (a - c.cpu()).abs().max())
    assert (want - got.cpu()).abs().max() <= TOL
    e.check_overflow()
    return got


@pytest.mark.parametrize("own_csr", [True, False])
@pytest.mark.parametrize("f,d,heads,hops", [(300, 512, 4, 5), (512, 512, 4, 5), (64, 32, 2, 2), (36, 20, 2, 2)])
def test_gat_seq_fused_matches_oracle_random_graphs(f, d, heads, hops, own_csr):
    cfg = dict(in_channels=f, out_channels=f, edge_attr_dim=f, ins_dim=d, num_ins=hops, dropout=0.1, gat_heads=heads)
    o, e = _pair(cfg, seed=11)
    ei, batch = random_graphs(6, 1, 24, 2.0, seed=5, isolated=True)
    _check(o, e, _inputs(ei, batch, 6, f, f, d, hops, seed=6), own_csr=own_csr)


def test_gat_seq_fused_cfg2_full_batch_matches_oracle_and_is_deterministic():
    """BASELINE cfg2 at its FULL size (256 graphs x 30 nodes / 60 edges, F=512, 4 heads, 5 hops), every hop."""
    cfg = dict(in_channels=512, out_channels=512, edge_attr_dim=512, ins_dim=512, num_ins=5, gat_heads=4)
    o, e = _pair(cfg, seed=81)
    ei, batch, _ = synthetic_topology(256, 30, 60, seed=1234)
    args = _inputs(ei, batch, 256, 512, 512, 512, 5, seed=82)
    got = _check(o, e, args)
    with torch.no_grad():
        dargs = [a.to(DEV) for a in args]
        again = e(*dargs)
        assert torch.equal(got, again)
        e.hop_mode = "split"
        split = e(*dargs)
    assert (split - got).abs().max() <= TOL


def test_gat_seq_fused_cfg4_shape_large_graphs():
    """BASELINE cfg4 shape (200 nodes / 800 edges per graph, F=512, 5 hops) on a per-GPU slice of 8 graphs:
    every graph is cut into two row tiles that stage the whole graph as their window."""
    cfg = dict(in_channels=512, out_channels=512, edge_attr_dim=512, ins_dim=512, num_ins=5, gat_heads=4)
    o, e = _pair(cfg, seed=41)
    ei, batch, _ = synthetic_topology(8, 200, 800, seed=4321)
    _check(o, e, _inputs(ei, batch, 8, 512, 512, 512, 5, seed=42))


def test_gat_seq_fused_refdims_jittered_graphs_without_hints():
    """Reference dims (F = 300) on graphs of 8-60 nodes, no loader hints (window = 256)."""
    cfg = dict(in_channels=300, out_channels=300, edge_attr_dim=300, ins_dim=512, num_ins=5, gat_heads=4)
    o, e = _pair(cfg, seed=51)
    ei, batch = random_graphs(40, 8, 60, 2.5, seed=9, isolated=True)
    _check(o, e, _inputs(ei, batch, 40, 300, 300, 512, 5, seed=52), hints=False)


@pytest.mark.parametrize("heads", [1, 8])
def test_gat_seq_fused_other_head_counts_use_the_split_path(heads):
    cfg = dict(in_channels=128, out_channels=128, edge_attr_dim=128, ins_dim=64, num_ins=3, gat_heads=heads)
    o, e = _pair(cfg, seed=61)
    ei, batch = random_graphs(6, 1, 24, 2.0, seed=5, isolated=True)
    _check(o, e, _inputs(ei, batch, 6, 128, 128, 64, 3, seed=62))


def test_gat_seq_fused_golden_refdims(golden):
    from oracle.make_golden import _state_hash
    fx = golden("gat_seq_refdims")
    torch.manual_seed(fx["seed"])
    e = eng.gat_seq(300, 300, 300, 512, 5, dropout=0.1, gat_heads=4).eval()
    if _state_hash(e.state_dict()) != fx["state_sha256"]:
        pytest.skip("seeded init differs on this torch build")
    e = e.to(DEV)
    e.hop_mode = "fused"
    with torch.no_grad():
        out = e(*[fx[k].to(DEV) for k in ("x", "edge_index", "edge_attr", "instr_vectors", "batch")]).cpu()
    assert (out - fx["out"]).abs().max() <= TOL


@pytest.mark.parametrize("seed", range(12))
def test_gat_seq_fused_equals_split_on_random_shapes(seed):
    """Fuzz: random widths (multiples of 4), head counts, graph-size mixes (empty, single-node, > 128 and > 256 nodes),
    edge densities and hop counts -- the one-kernel hop against the projection GEMM + hop kernel pair."""
    g = torch.Generator().manual_seed(1000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    f = 4 * ri(3, 80)
    d = 4 * ri(2, 40)
    heads, hops = (2, 4)[ri(0, 1)], ri(1, 4)
    cfg = dict(in_channels=f, out_channels=f, edge_attr_dim=f, ins_dim=d, num_ins=hops, gat_heads=heads)
    _, e = _pair(cfg, seed=seed)
    sizes = [ri(0, 3) for _ in range(ri(0, 6))] + [ri(1, 60) for _ in range(ri(1, 30))]
    if seed % 3 == 0:
        sizes += [ri(129, 300)]
    if seed % 4 == 1:
        sizes += [ri(257, 400)]
    order = torch.randperm(len(sizes), generator=g).tolist()
    sizes = [sizes[i] for i in order]
    src, dst, batch, off = [], [], [], 0
    for b, n in enumerate(sizes):
        m = int(n * (0.5 + 3 * float(torch.rand(1, generator=g))))
        if n:
            src += (torch.randint(0, n, (m,), generator=g) + off).tolist() + list(range(off, off + n))[: n // 2]
            dst += (torch.randint(0, n, (m,), generator=g) + off).tolist() + list(range(off, off + n))[: n // 2]
        batch += [b] * n
        off += n
    ei = torch.tensor([src, dst], dtype=torch.long).reshape(2, -1)
    batch = torch.tensor(batch, dtype=torch.long)
    args = [a.to(DEV) for a in _inputs(ei, batch, len(sizes), f, f, d, hops, seed=seed + 7)]
    with torch.no_grad():
        e.hop_mode = "fused"
        got = e(*args)
        e.hop_mode = "split"
        want = e(*args)
    e.check_overflow()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max() <= TOL, "seed %d: max|d| = %g" % (seed, (got - want).abs().max())

"""Where does the end-to-end (host-buffer) step go?  H2D only / H2D+D2H / full runner, per batch of cfg2."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from graphvqa_b200 import gat_skip as eng
from graphvqa_b200.host_api import GatSeqHostRunner
dev = torch.device('cuda:0')
cfg = bench.CFG2
keys = ("x", "edge_index", "edge_attr", "instr_vectors", "batch")
sets = [bench.make_inputs(cfg, 1234 + 10 * r) for r in range(4)]
pinned = [{k: s[k].pin_memory() for k in keys} for s in sets]
devbuf = [{k: torch.empty_like(pinned[0][k], device=dev) for k in keys} for _ in range(2)]
out_dev = torch.empty(cfg["graphs"] * cfg["nodes"], cfg["feat"], device=dev)
out_host = [torch.empty(out_dev.shape).pin_memory() for _ in range(2)]
nbytes = sum(pinned[0][k].numel() * pinned[0][k].element_size() for k in keys)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, n=40):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d(i):
    with torch.cuda.stream(s1):
        for k in keys: devbuf[i % 2][k].copy_(pinned[i % 4][k], non_blocking=True)
def h2d_d2h(i):
    h2d(i)
    with torch.cuda.stream(s2): out_host[i % 2].copy_(out_dev, non_blocking=True)
ms = timeit(h2d); print("H2D only, 5 tensors %.1f MB: %.3f ms/batch = %.1f GB/s" % (nbytes / 1e6, ms, nbytes / ms / 1e6))
ms = timeit(h2d_d2h); print("H2D + D2H (15.7 MB) concurrently: %.3f ms/batch" % ms)
# one flat pinned staging buffer, one copy
flat = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); dflat = torch.empty(nbytes, dtype=torch.uint8, device=dev)
def h2d_flat(i):
    with torch.cuda.stream(s1): dflat.copy_(flat, non_blocking=True)
ms = timeit(h2d_flat); print("H2D one flat buffer: %.3f ms = %.1f GB/s" % (ms, nbytes / ms / 1e6))
def h2d_flat_d2h(i):
    h2d_flat(i)
    with torch.cuda.stream(s2): out_host[i % 2].copy_(out_dev, non_blocking=True)
ms = timeit(h2d_flat_d2h); print("H2D flat + D2H: %.3f ms" % ms)
torch.manual_seed(0)
model = eng.gat_seq(**bench.model_kwargs(cfg)).eval(); bench.randomise_bn(model, 7); model = model.to(dev)
runner = GatSeqHostRunner(model, dev, depth=int(sys.argv[1]) if len(sys.argv) > 1 else 2, use_cuda_graph=True,
                          max_nodes_per_graph=max(s["max_nodes"] for s in sets), max_in_edges_per_graph=max(s["max_edges"] for s in sets))
with torch.no_grad():
    for i in range(8): runner.submit(pinned[i % 4])
    runner.drain(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 40
    for i in range(n):
        t = runner.submit(pinned[i % 4])
        if i >= 2: runner.result(t - 2)
    runner.drain(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
print("runner depth=%d: %.3f ms/batch = %.0f q/s" % (runner.depth, ms, cfg["graphs"] / ms * 1e3))

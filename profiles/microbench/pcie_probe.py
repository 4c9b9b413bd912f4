"""PCIe copy rates of the box (pinned host <-> device), the ceiling of bench.py's e2e number."""
import torch
dev = "cuda:0"
for mb in (1, 16, 50, 200):
    h = torch.empty(mb * 1024 * 1024 // 4).pin_memory(); d = torch.empty_like(h, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%s %4d MB: %.3f ms  %.1f GB/s" % (name, mb, ms, mb / 1024 / (ms / 1e3)))
# both directions at once on two streams
h1 = torch.empty(50 * 1024 * 1024 // 4).pin_memory(); d1 = torch.empty_like(h1, device=dev)
h2 = torch.empty(16 * 1024 * 1024 // 4).pin_memory(); d2 = torch.empty_like(h2, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
print("duplex 50 MB H2D + 16 MB D2H: %.3f ms per pair" % (e0.elapsed_time(e1) / 10))
# does splitting one H2D transfer over several streams (copy engines) raise the rate?
for nstream in (1, 2, 4):
    n = 48 * 1024 * 1024 // 4
    h = torch.randn(n).pin_memory(); d = torch.empty_like(h, device=dev)
    streams = [torch.cuda.Stream() for _ in range(nstream)]
    chunk = n // nstream
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
    for _ in range(3): go()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(10): go()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print("H2D 48 MB over %d stream(s): %.3f ms  %.1f GB/s" % (nstream, ms, 48 / 1024 / (ms / 1e3)))
# repeated 48 MB H2D from the SAME pinned buffer vs rotating over 4 buffers (host cache / IOMMU effects)
bufs = [torch.randn(48 * 1024 * 1024 // 4).pin_memory() for _ in range(4)]
d = torch.empty_like(bufs[0], device=dev)
for label, seq in (("same buffer", [0] * 12), ("4 buffers rotated", [0, 1, 2, 3] * 3)):
    for i in seq[:4]: d.copy_(bufs[i], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in seq: d.copy_(bufs[i], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / len(seq)
    print("H2D 48 MB %s: %.3f ms  %.1f GB/s" % (label, ms, 48 / 1024 / (ms / 1e3)))

"""What does an event pair around ONE launch cost when the stream is busy (the bench's in-situ protocol)?"""
import torch
dev = torch.device('cuda:0')
a = torch.randn(4096, 4096, device=dev); t = torch.zeros(32, device=dev)
big = torch.randn(64 << 20, device=dev) ; big2 = torch.empty_like(big)
def measure(fn, label, pre):
    ts = []
    for i in range(30):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[5:])
    print("%-44s median %.2f us  min %.2f us" % (label, ts[len(ts) // 2], ts[0]))
busy = lambda: torch.mm(a, a)
measure(lambda: None, "no launch between the events (stream busy)", busy)
measure(lambda: t.add_(1.0), "32-element add (stream busy)", busy)
measure(lambda: big2.copy_(big), "256 MB copy r+w = 512 MB (stream busy)", busy)
measure(lambda: t.add_(1.0), "32-element add (stream idle, CPU-bound)", lambda: torch.cuda.synchronize())

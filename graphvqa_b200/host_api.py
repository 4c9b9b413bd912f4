"""Host-buffer entry point of the hot path: ``gat_seq`` fed from (pinned) host memory.

``GatSeqHostRunner`` is the call a data-loader-side user makes: it takes the five host tensors of one
batch (``x, edge_index, edge_attr, instr_vectors, batch`` -- the arguments of the reference's
``gat_seq.forward``, gat_skip.py:249), moves them to the GPU, builds the CSR, runs the hop stack and
returns the node states in pinned host memory.  Consecutive batches are software-pipelined over three
CUDA streams and ``depth`` device slots: the H2D copy of batch i+1 and the D2H copy of batch i-1
overlap the kernels of batch i (PCIe is full duplex).  ``depth`` = 3 keeps the H2D engine saturated: with 2
slots the host has to wait for batch i-2's result before it may enqueue batch i's copy, which left the link
idle ~45 % of the time (profiles/microbench/e2e_probe.py: 1.75 ms -> 0.98 ms per cfg2 batch, the PCIe bound).  With ``use_cuda_graph`` the per-slot compute
(CSR build + pre-pass + hops) is captured once per input shape and replayed.
"""
import torch

from .graph_batch import GraphCSR

_KEYS = ("x", "edge_index", "edge_attr", "instr_vectors", "batch")


class _Slot:
    def __init__(self):
        self.dev = None            # static device input tensors
        self.out_dev = None        # device output (static when graph-captured)
        self.out_host = None       # pinned host output
        self.flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()   # fp16 range flag of this batch
        self.flag_dev = None       # device snapshot of the flag, taken on the compute stream after the batch
        self.host_in = None        # the caller's host tensors (for the full-range rerun)
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.graph = None
        self.shapes = None
        self.busy = False


class GatSeqHostRunner:
    def __init__(self, model, device, depth=3, use_cuda_graph=True, max_nodes_per_graph=0,
                 max_in_edges_per_graph=0):
        self.model, self.device, self.depth = model, torch.device(device), depth
        self.use_cuda_graph = use_cuda_graph
        self.hints = dict(max_nodes_per_graph=max_nodes_per_graph, max_in_edges_per_graph=max_in_edges_per_graph)
        self.s_h2d, self.s_compute, self.s_d2h = (torch.cuda.Stream(self.device) for _ in range(3))
        self.slots = [_Slot() for _ in range(depth)]
        if hasattr(model, "overflow_external"):
            model.overflow_external = True     # the fp16 range flag travels with each result (see result())
        self.count = 0

    def _forward(self, d):
        return self.model(d["x"], d["edge_index"], d["edge_attr"], d["instr_vectors"], d["batch"], csr_hints=self.hints)

    @torch.no_grad()
    def submit(self, host):
        """Enqueue one batch (dict or 5-tuple of CPU tensors, ideally pinned).  Returns a ticket."""
        if not isinstance(host, dict):
            host = dict(zip(_KEYS, host))
        slot = self.slots[self.count % self.depth]
        if slot.busy:
            slot.d2h_done.synchronize()     # the slot's previous result must have left the device
        shapes = tuple((tuple(host[k].shape), host[k].dtype) for k in _KEYS)
        if slot.shapes != shapes:
            slot.dev = {k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.device) for k in _KEYS}
            slot.shapes, slot.graph, slot.out_dev, slot.out_host = shapes, None, None, None
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(slot.compute_done)          # previous compute on this slot has read its inputs
            for k in _KEYS:
                slot.dev[k].copy_(host[k], non_blocking=True)
            slot.h2d_done.record(self.s_h2d)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(slot.h2d_done)
            self.s_compute.wait_event(slot.d2h_done)          # static output buffer is free again
            if self.use_cuda_graph and slot.graph is None and slot.out_dev is not None:
                # second use of this shape: capture (the first use ran eagerly and warmed everything up)
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(self.device)
                with torch.cuda.graph(g, stream=self.s_compute):
                    slot.out_dev = self._forward(slot.dev)
                slot.graph = g
            flag = getattr(self.model, "_overflow", None)
            if flag is not None:
                flag.zero_()                                  # this batch's fp16 range flag starts clear ...
            if slot.graph is not None:
                slot.graph.replay()
            else:
                slot.out_dev = self._forward(slot.dev)
            flag = getattr(self.model, "_overflow", None)     # (created by the first forward)
            if flag is not None:
                if slot.flag_dev is None:
                    slot.flag_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
                slot.flag_dev.copy_(flag)                     # ... and is snapshotted in stream order, per slot
            slot.compute_done.record(self.s_compute)
        if slot.out_host is None or slot.out_host.shape != slot.out_dev.shape:
            slot.out_host = torch.empty(slot.out_dev.shape, dtype=slot.out_dev.dtype).pin_memory()
        slot.host_in = host
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(slot.compute_done)
            slot.out_host.copy_(slot.out_dev, non_blocking=True)
            if slot.flag_dev is not None:   # fp16-split projection: the batch's range flag travels with the result
                slot.flag_host.copy_(slot.flag_dev, non_blocking=True)
            if slot.graph is None:
                slot.out_dev.record_stream(self.s_d2h)
            slot.d2h_done.record(self.s_d2h)
        slot.busy = True
        self.count += 1
        return self.count - 1

    def result(self, ticket):
        """Pinned host tensor with the node states of batch ``ticket`` (valid until the slot is reused,
        i.e. until ``depth`` more batches have been submitted)."""
        slot = self.slots[ticket % self.depth]
        slot.d2h_done.synchronize()
        if int(slot.flag_host) != 0:
            # an input did not fit fp16: redo this batch with the tf32-split projection (full fp32 range)
            slot.flag_host.zero_()
            prev = self.model.projection
            self.model.projection = "3xtf32"
            try:
                with torch.no_grad(), torch.cuda.stream(self.s_compute):
                    dev = {k: slot.host_in[k].to(self.device, non_blocking=True) for k in _KEYS}
                    slot.out_host.copy_(self._forward(dev))
                self.s_compute.synchronize()
            finally:
                self.model.projection = prev
        return slot.out_host

    def drain(self):
        for s in self.slots:
            if s.busy:
                s.d2h_done.synchronize()

    def __call__(self, *host):
        return self.result(self.submit(host if len(host) != 1 else host[0]))

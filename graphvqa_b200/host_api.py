"""Host-buffer entry points of the hot path.

``GatSeqHostRunner``   the operator boundary: the five host tensors of one ``gat_seq.forward`` call
                       (``x, edge_index, edge_attr, instr_vectors, batch`` -- gat_skip.py:249 of the reference) in,
                       node states out.  Ships pre-encoded fp32 features: ~50 MB per 256-graph batch, PCIe-bound.
``GraphSideHostRunner`` the boundary the reference's own loop has (mainExplain_gat.py:710-714, 758-765): TOKEN IDS in,
                       answer logits out.  One wire-format batch (``collate.WireCollator``: int32 tokens + the
                       loader-built destination-CSR) plus the text side's ``instr_vectors`` / question summary go up
                       (~4 MB per 256-graph batch), ``short_answer_logits`` [B, 1842] come back (1.9 MB); scene-graph
                       encoder -> hop stack -> conditional attention pooling -> logit_fc run as ONE CUDA graph.

Both software-pipeline consecutive batches over three CUDA streams and ``depth`` device slots: the H2D copy of batch
i+1 and the D2H copy of batch i-1 overlap the kernels of batch i (PCIe is full duplex).  ``depth`` = 3 keeps the H2D
engine saturated: with 2 slots the host has to wait for batch i-2's result before it may enqueue batch i's copy, which
left the link idle ~45 % of the time (profiles/microbench/e2e_probe.py).  With ``use_cuda_graph`` the per-slot compute
is captured once per input shape and replayed.

The default fp16-split projection flags inputs outside fp16's range on the device; the flag travels with every
result and ``result()`` redoes such a batch with the tf32 split (full fp32 range), so the caller always receives
valid numbers.
"""
import torch

from .graph_batch import GraphCSR, SceneGraphBatch


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


class _PipelinedRunner:
    """Slot / stream machinery shared by the two runners.  Subclasses define ``_keys`` (names of the host tensors of
    one batch), ``_forward(dev_dict) -> device tensor`` and ``_range_seq()`` (the gat_seq whose fp16 flag guards the
    batch, or None)."""

    _keys = ()

    def __init__(self, device, depth=3, use_cuda_graph=True):
        self.device, self.depth = torch.device(device), depth
        self.use_cuda_graph = use_cuda_graph
        self.s_h2d, self.s_compute, self.s_d2h = (torch.cuda.Stream(self.device) for _ in range(3))
        self.slots = [_Slot() for _ in range(depth)]
        self.count = 0
        seq = self._range_seq()
        if seq is not None:
            seq.overflow_external = True     # the fp16 range flag travels with each result (see result())

    def _range_seq(self):
        return None

    def _forward(self, dev):
        raise NotImplementedError

    def _flag(self):
        seq = self._range_seq()
        return None if seq is None else getattr(seq, "_overflow", None)

    def _forward_full_range(self, dev):
        """``_forward`` with the tf32-split projection (full fp32 range)."""
        seq = self._range_seq()
        prev = seq.projection
        seq.projection = "3xtf32"
        try:
            return self._forward(dev)
        finally:
            seq.projection = prev

    @torch.no_grad()
    def submit(self, host):
        """Enqueue one batch (dict of CPU tensors, ideally pinned).  Returns a ticket."""
        slot = self.slots[self.count % self.depth]
        if slot.busy:
            slot.d2h_done.synchronize()     # the slot's previous result must have left the device
        shapes = tuple((tuple(host[k].shape), host[k].dtype) for k in self._keys)
        if slot.shapes != shapes:
            slot.dev = {k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.device) for k in self._keys}
            slot.shapes, slot.graph, slot.out_dev, slot.out_host = shapes, None, None, None
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(slot.compute_done)          # previous compute on this slot has read its inputs
            for k in self._keys:
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
            flag = self._flag()
            if flag is not None:
                flag.zero_()                                  # this batch's fp16 range flag starts clear ...
            if slot.graph is not None:
                slot.graph.replay()
            else:
                slot.out_dev = self._forward(slot.dev)
            flag = self._flag()                               # (created by the first forward)
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
        """Pinned host tensor with the result of batch ``ticket`` (valid until the slot is reused, i.e. until
        ``depth`` more batches have been submitted)."""
        slot = self.slots[ticket % self.depth]
        slot.d2h_done.synchronize()
        if int(slot.flag_host) != 0:
            # an input did not fit fp16: redo this batch with the tf32-split projection (full fp32 range).  The fp16
            # weight pack stays alive (gat_seq keeps one pack per projection kind), so the slots' captured graphs,
            # which hold raw pointers into it, remain valid.
            slot.flag_host.zero_()
            with torch.no_grad(), torch.cuda.stream(self.s_compute):
                dev = {k: slot.host_in[k].to(self.device, non_blocking=True) for k in self._keys}
                slot.out_host.copy_(self._forward_full_range(dev))
            self.s_compute.synchronize()
        return slot.out_host

    def drain(self):
        for s in self.slots:
            if s.busy:
                s.d2h_done.synchronize()


_GAT_SEQ_KEYS = ("x", "edge_index", "edge_attr", "instr_vectors", "batch")


class GatSeqHostRunner(_PipelinedRunner):
    _keys = _GAT_SEQ_KEYS

    def __init__(self, model, device, depth=3, use_cuda_graph=True, max_nodes_per_graph=0,
                 max_in_edges_per_graph=0):
        self.model = model
        self.hints = dict(max_nodes_per_graph=max_nodes_per_graph, max_in_edges_per_graph=max_in_edges_per_graph)
        super().__init__(device, depth, use_cuda_graph)

    def _range_seq(self):
        return self.model if hasattr(self.model, "overflow_external") else None

    def _forward(self, d):
        return self.model(d["x"], d["edge_index"], d["edge_attr"], d["instr_vectors"], d["batch"], csr_hints=self.hints)

    def submit(self, host):
        if not isinstance(host, dict):
            host = dict(zip(_GAT_SEQ_KEYS, host))
        return super().submit(host)

    def __call__(self, *host):
        return self.result(self.submit(host if len(host) != 1 else host[0]))


class GraphSideHostRunner(_PipelinedRunner):
    """``model``: a ``PipelineModel`` (GAT / GCN / GINE variant).  ``submit(graphs, instr_vectors, q0)`` takes a
    wire-format ``SceneGraphBatch`` from ``collate.WireCollator`` (host tensors), the instruction vectors [5,B,D] and
    the question summary ``questions_encoded[0]`` [B,D] (host, ideally pinned); ``result`` returns
    ``short_answer_logits`` [B,1842] in pinned host memory.  The device never runs a CSR kernel: the topology arrives
    in its final int32 form."""

    _csr_keys = ("rowptr", "col_src", "perm", "graph_ptr", "node_graph", "stats")
    _plan_keys = ("tiles", "tile_count")         # row tiles of the fused hop, built by the loader as well
    _keys = ("x", "edge_index", "edge_attr", "edge_sign") + _csr_keys + _plan_keys + ("instr_vectors", "q0")

    def __init__(self, model, device, depth=3, use_cuda_graph=True):
        self.model = model
        super().__init__(device, depth, use_cuda_graph)

    def _range_seq(self):
        return None          # the model owns one flag for gat_seq and the dense layers (see _flag)

    def _flag(self):
        return self.model._range_flag

    def _forward_full_range(self, d):
        return self._forward(d, full_range=True)

    def _forward(self, d, full_range=False):
        n, b = d["rowptr"].numel() - 1, d["graph_ptr"].numel() - 1
        e = d["edge_index"].size(1)
        hints = self._hints
        csr = GraphCSR({k: d[k] for k in self._csr_keys}, n, e, b, hints[0], hints[1])
        csr._fused_plans[self._tile_window] = (d["tiles"], d["tile_count"])
        g = SceneGraphBatch(x=d["x"], edge_index=d["edge_index"], edge_attr=d["edge_attr"], batch=d["node_graph"],
                            edge_sign=d["edge_sign"], num_graphs=b, max_nodes_per_graph=hints[0],
                            max_in_edges_per_graph=hints[1])
        g._csr = csr
        model = self.model
        strict, model.strict_range = model.strict_range, False      # the flag travels with the result instead
        try:
            return model.graph_side(g, d["instr_vectors"], d["q0"].unsqueeze(0), b, csr=csr, full_range=full_range)
        finally:
            model.strict_range = strict

    def submit(self, graphs, instr_vectors=None, q0=None):
        if isinstance(graphs, dict):
            host = graphs
        else:
            if graphs.csr_host is None or graphs.edge_sign is None:
                raise ValueError("GraphSideHostRunner: pass a wire-format batch (collate.WireCollator)")
            host = dict(x=graphs.x, edge_index=graphs.edge_index, edge_attr=graphs.edge_attr,
                        edge_sign=graphs.edge_sign, instr_vectors=instr_vectors, q0=q0, **graphs.csr_host)
            self._hints = (graphs.max_nodes_per_graph, graphs.max_in_edges_per_graph)
            self._tile_window = int(graphs.csr_host["tile_window"])
        return super().submit(host)

    _hints = (0, 0)
    _tile_window = 128

    def __call__(self, graphs, instr_vectors, q0):
        return self.result(self.submit(graphs, instr_vectors, q0))

    @staticmethod
    def bytes_per_batch(graphs, instr_vectors, q0):
        host = [graphs.x, graphs.edge_index, graphs.edge_attr, graphs.edge_sign, instr_vectors, q0] + \
               [graphs.csr_host[k] for k in GraphSideHostRunner._csr_keys + GraphSideHostRunner._plan_keys]
        return sum(t.numel() * t.element_size() for t in host)

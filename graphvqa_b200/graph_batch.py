"""Batched disjoint scene graphs: the input object of the engine.

``SceneGraphBatch`` is a duck-typed stand-in for ``torch_geometric.data.Batch`` with exactly the
attributes the reference reads (produced by
``gqa_dataset_entry.GQATorchDataset_collate_fn``, gqa_dataset_entry.py:631-675 of the reference):
``x[N,12]`` i64 token ids, ``edge_index[2,E]`` i64 ([0]=source, [1]=target), ``edge_attr[E,1]`` i64,
``added_sym_edge[K]`` i64, ``batch[N]`` i64 (non-decreasing), ``y[N,5]`` f32, ``num_graphs``, and
``.to(device=, non_blocking=)`` as called by mainExplain_gat.py:424-428.

``GraphCSR`` is the engine-side int32 destination-CSR view (built once per batch on the GPU by
``gvqa_build_csr``) that every hop kernel consumes.
"""
import torch

from . import _cabi


class GraphCSR:
    """int32 destination-CSR of a batch (device tensors) + loader hints."""

    def __init__(self, parts, num_nodes, num_edges, num_graphs, max_nodes_per_graph=0,
                 max_in_edges_per_graph=0):
        self.rowptr, self.col_src, self.perm = parts["rowptr"], parts["col_src"], parts["perm"]
        self.graph_ptr, self.node_graph, self.stats = parts["graph_ptr"], parts["node_graph"], parts["stats"]
        self.num_nodes, self.num_edges, self.num_graphs = num_nodes, num_edges, num_graphs
        # loader hints (0 = unknown): they size the shared-memory staged kernels, never correctness
        self.max_nodes_per_graph = max_nodes_per_graph
        self.max_in_edges_per_graph = max_in_edges_per_graph
        self._fused_plans = {}

    def fused_plan(self, window):
        """Row tiles of the fused aggregate-project hop (gvqa_gat_fused_plan), built once per batch and window."""
        if window not in self._fused_plans:
            self._fused_plans[window] = _cabi.fused_plan(self.graph_ptr, self.num_nodes, self.num_graphs, window)
        return self._fused_plans[window]

    def hints(self):
        return dict(max_nodes_per_graph=self.max_nodes_per_graph,
                    max_in_edges_per_graph=self.max_in_edges_per_graph)

    def as_dict(self):
        return dict(rowptr=self.rowptr, col_src=self.col_src, perm=self.perm, graph_ptr=self.graph_ptr,
                    node_graph=self.node_graph, num_edges=self.num_edges, num_graphs=self.num_graphs)

    def read_stats(self):
        """Host copy (synchronises!) of {max_nodes, max_in_edges, max_in_degree, bad_edges}."""
        s = self.stats.cpu().tolist()
        return dict(max_nodes=s[0], max_in_edges=s[1], max_in_degree=s[2], bad_edges=s[3])

    @staticmethod
    def from_host(parts, device, non_blocking=True):
        """Device CSR from the loader-side build (``_cabi.build_csr_host`` / ``collate.collate_wire``): the int32
        arrays are copied as they are, no CSR kernel runs on the GPU.  The hints come from the host stats."""
        st = parts["stats"]
        dev = {k: parts[k].to(device=device, non_blocking=non_blocking)
               for k in ("rowptr", "col_src", "perm", "graph_ptr", "node_graph", "stats")}
        n, b = parts["rowptr"].numel() - 1, parts["graph_ptr"].numel() - 1
        e = int(parts["rowptr"][n]) if n >= 0 else 0
        csr = GraphCSR(dev, n, e, b, int(st[0]), int(st[1]))
        if "tiles" in parts:            # the loader also built the fused hop's row tiles (collate.WireCollator)
            csr._fused_plans[int(parts["tile_window"])] = tuple(
                parts[k].to(device=device, non_blocking=non_blocking) for k in ("tiles", "tile_count"))
        return csr

    def check(self):
        """Synchronising sanity check (debug / first batch of a new data source): raises when the build counted
        edges whose endpoints are out of range or lie in different graphs, or batch ids that are out of range or
        not sorted -- inputs for which the per-graph folding of the instruction terms is not valid."""
        st = self.read_stats()
        if st["bad_edges"]:
            raise ValueError("GraphCSR: %d invalid edges / batch ids (endpoints out of range or in different "
                             "graphs, graph ids outside [0, num_graphs) or unsorted)" % st["bad_edges"])
        return st

    @staticmethod
    def build(edge_index, batch, num_graphs, max_nodes_per_graph=0, max_in_edges_per_graph=0,
              read_hints=False):
        """read_hints=True fills missing hints from the device stats (one host sync)."""
        parts = _cabi.build_csr(edge_index, batch, num_graphs)
        csr = GraphCSR(parts, batch.numel(), edge_index.size(1), num_graphs, max_nodes_per_graph,
                       max_in_edges_per_graph)
        if read_hints and (not max_nodes_per_graph or not max_in_edges_per_graph):
            st = csr.read_stats()
            csr.max_nodes_per_graph, csr.max_in_edges_per_graph = st["max_nodes"], st["max_in_edges"]
        return csr


class SceneGraphBatch:
    """Attribute bag with the reference Batch surface.  Extra: ``max_nodes_per_graph`` hint and a
    lazily built, cached ``csr`` (dropped by ``.to`` to another device)."""

    _TENSOR_FIELDS = ("x", "edge_index", "edge_attr", "added_sym_edge", "batch", "y", "edge_sign")

    def __init__(self, x=None, edge_index=None, edge_attr=None, batch=None, added_sym_edge=None, y=None,
                 num_graphs=None, max_nodes_per_graph=0, max_in_edges_per_graph=0, edge_sign=None, csr_host=None):
        self.x, self.edge_index, self.edge_attr, self.batch = x, edge_index, edge_attr, batch
        self.added_sym_edge, self.y = added_sym_edge, y
        # wire format (collate.WireCollator): `edge_sign` [E] float32 = -1 on the rows the reference negates through
        # `added_sym_edge` (pipeline_model_gat.py:590), +1 elsewhere -- a fixed-shape stand-in for the ragged index
        # list; `csr_host` = the loader-built int32 destination-CSR (dict of host tensors), shipped by .to()
        self.edge_sign, self.csr_host = edge_sign, csr_host
        self.num_graphs = num_graphs
        self.max_nodes_per_graph = max_nodes_per_graph
        self.max_in_edges_per_graph = max_in_edges_per_graph
        self._csr = None

    @property
    def num_nodes(self):
        return self.batch.numel()

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def to(self, device=None, non_blocking=False):
        out = SceneGraphBatch(num_graphs=self.num_graphs, max_nodes_per_graph=self.max_nodes_per_graph,
                              max_in_edges_per_graph=self.max_in_edges_per_graph, csr_host=self.csr_host)
        for name in self._TENSOR_FIELDS:
            t = getattr(self, name)
            setattr(out, name, None if t is None else t.to(device=device, non_blocking=non_blocking))
        if self.csr_host is not None and device is not None and torch.device(device).type == "cuda":
            out._csr = GraphCSR.from_host(self.csr_host, device, non_blocking=non_blocking)   # no CSR kernel on the GPU
        return out

    def pin_memory(self):
        out = SceneGraphBatch(num_graphs=self.num_graphs, max_nodes_per_graph=self.max_nodes_per_graph,
                              max_in_edges_per_graph=self.max_in_edges_per_graph, csr_host=self.csr_host)
        for name in self._TENSOR_FIELDS:
            t = getattr(self, name)
            setattr(out, name, None if t is None else t.pin_memory())
        return out

    def csr(self):
        if self._csr is None:
            if self.num_graphs is None:
                raise ValueError("SceneGraphBatch.num_graphs must be set (avoids a device sync on batch.max())")
            self._csr = GraphCSR.build(self.edge_index, self.batch, self.num_graphs, self.max_nodes_per_graph,
                                       self.max_in_edges_per_graph)
        return self._csr


def get_csr(graph_or_none, edge_index, batch, num_graphs):
    """CSR for (edge_index, batch): reuse the batch object's cache when the tensors are its own."""
    if isinstance(graph_or_none, SceneGraphBatch) and graph_or_none.edge_index is edge_index:
        return graph_or_none.csr()
    return GraphCSR.build(edge_index, batch, num_graphs)


def synthetic_topology(num_graphs, nodes_per_graph, edges_per_graph, seed=1234, jitter=0):
    """GQA-shaped synthetic topology: per graph, one self-loop per
    node listed first (the dataset emits an explicit <self> edge per node,
    gqa_dataset_entry.py:292-297) followed by ``edges_per_graph - nodes`` random intra-graph
    directed edges (duplicates allowed).  ``jitter`` > 0 varies the node count per graph by
    +-jitter.  Returns CPU tensors (edge_index[2,E] i64, batch[N] i64, max_nodes); the largest
    per-graph edge count is ``synthetic_topology.last_max_edges``."""
    g = torch.Generator().manual_seed(seed)
    srcs, dsts, batch, off, max_nodes, max_edges = [], [], [], 0, 0, 0
    for b in range(num_graphs):
        n = nodes_per_graph
        if jitter:
            n = max(1, n + int(torch.randint(-jitter, jitter + 1, (1,), generator=g)))
        extra = max(0, int(round(edges_per_graph * n / nodes_per_graph)) - n)
        loops = torch.arange(n)
        s = torch.cat([loops, torch.randint(0, n, (extra,), generator=g)]) + off
        d = torch.cat([loops, torch.randint(0, n, (extra,), generator=g)]) + off
        srcs.append(s); dsts.append(d)
        batch.append(torch.full((n,), b, dtype=torch.long))
        off += n
        max_nodes = max(max_nodes, n)
        max_edges = max(max_edges, n + extra)
    synthetic_topology.last_max_edges = max_edges
    return torch.stack([torch.cat(srcs), torch.cat(dsts)]), torch.cat(batch), max_nodes

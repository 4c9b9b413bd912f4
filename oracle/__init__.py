"""CPU oracle for the GraphVQA scene-graph message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and there only as the checker
(or as the timed CPU baseline), never as the thing shipped.  The product
package ``graphvqa_b200`` never imports this package and fails loudly when its
CUDA library is missing.

PARITY STATUS: *parity unpinned for the third-party primitives*.  The
reference (codexxxl/GraphVQA) has no tests, no golden vectors and pins no
version of torch_geometric / torch_scatter, whose kernels own the arithmetic
of the path.  What IS pinned:

* the reference's own code for the path (``gat_skip.py``,
  ``graph_utils/my_graph_layernorm.py``, ``baseline_and_test_models/lcgn.py``
  and the ``*_seq`` classes of the pipeline files) is executed UNMODIFIED in
  the build container on top of ``oracle/pyg_shim`` (a minimal restatement of
  the PyG 1.6/1.7 API it calls), by ``oracle/make_golden.py``; the resulting
  input/output vectors are committed under ``tests/golden/`` and both this
  restatement and the CUDA path are checked against them;
* the PyG primitive semantics themselves (segment softmax with +1e-16,
  scatter-add/mean, degree, GCN normalisation, GINE aggregation, glorot) are
  restated from their published definitions in ``oracle/pyg_semantics.py``
  and checked against hand-computed tiny cases in ``tests/``.
"""

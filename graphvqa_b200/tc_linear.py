"""Dense layers of the graph side on the hand-written tensor-core GEMMs.

``TensorCoreLinear()(x, weight, bias)`` evaluates ``x @ weight^T (+ bias)`` with fp32-level accuracy on tcgen05 instead
of cuBLAS' fp32 SIMT kernels, which the 1e-4 parity bar would otherwise force (TF32 off):

* ``mode = "3xtf32"`` (default): ``gvqa_proj_gemm_3xtf32`` -- tf32-split operands, the full fp32 range;
* ``mode = "3xf16"``: ``gvqa_proj_gemm_3xf16`` -- fp16-split operands, ~1.4x faster, inputs must fit fp16's range
  (|x| < 65504); the kernel ORs ``flag`` (a device int32[1]) when one does not, and whoever owns the flag redoes the
  work with "3xtf32" (``PipelineModel.graph_side``, the host runners).  Only callers that own such a flag switch
  this mode on.

The split weights are cached per parameter version and mode.  Used by the scene-graph encoder, the attention pooling,
the answer head and the GCN / GINE / LCGN variants; the Transformer text stack stays plain PyTorch.
"""
import torch

from . import _cabi


class TensorCoreLinear:
    def __init__(self, mode="3xtf32"):
        self._cache = {}
        self.mode = mode
        self.flag = None

    def _split(self, weight, f16):
        key = (weight.data_ptr(), weight._version, tuple(weight.shape), tuple(weight.stride()), f16)
        split = self._cache.get(key)
        if split is None:
            if len(self._cache) > 128:
                self._cache.clear()
            w = weight.detach().contiguous().float()
            split = self._cache[key] = _cabi.split_f16(w) if f16 else _cabi.split_tf32(w)
        return split

    def stacked(self, weights):
        """Row-wise concatenation of several [N_i, K] weights (cached per parameter version): Linear layers that read
        the SAME input become one GEMM whose output columns are sliced afterwards."""
        key = ("stack",) + tuple((w.data_ptr(), w._version, tuple(w.shape), tuple(w.stride())) for w in weights)
        hit = self._cache.get(key)
        if hit is None:
            hit = self._cache[key] = torch.cat([w.detach() for w in weights]).contiguous().float()
        return hit

    def stacked_bias(self, parts):
        """Bias of a stacked Linear: ``parts`` = tensors or ints (that many zeros), cached per parameter version."""
        key = ("bias",) + tuple(p if isinstance(p, int) else (p.data_ptr(), p._version, p.numel()) for p in parts)
        hit = self._cache.get(key)
        if hit is None:
            ref = next(p for p in parts if not isinstance(p, int))
            hit = self._cache[key] = torch.cat([ref.new_zeros(p) if isinstance(p, int) else p.detach().float()
                                                for p in parts]).contiguous()
        return hit

    def group(self, problems):
        """Up to three INDEPENDENT Linear layers in one persistent launch (fp16-split mode; their tiles share the
        waves).  ``problems``: (x, weight, bias, relu) tuples.  Returns the outputs in order."""
        f16 = self.mode == "3xf16" and self.flag is not None
        if not f16 or len(problems) == 1 or len(problems) > _cabi.MAX_GROUPED_PROBLEMS or \
                any(p[0].size(0) == 0 or p[1].size(1) % 4 for p in problems):
            return [self(*p) for p in problems]
        packed = []
        for x, weight, bias, relu in problems:
            _cabi.require_cuda(x, weight, bias)
            hi, lo = self._split(weight, True)
            packed.append((x.contiguous().float(), hi, lo, None,
                           None if bias is None else bias.detach().contiguous().float(), relu))
        return _cabi.proj_gemm_3xf16_grouped(packed, overflow=self.flag)

    def __call__(self, x, weight, bias=None, relu=False):
        """x [M, K] float32 CUDA, weight [N, K] (any strides), bias [N] or None -> act(x @ weight^T + bias) [M, N]."""
        n, k = weight.shape
        _cabi.require_cuda(x, weight, bias)              # raises: there is no CPU fallback
        if k % 4:
            raise ValueError("TensorCoreLinear: the input width must be a multiple of 4 (got %d); pad the weight "
                             "columns or call torch.nn.functional.linear explicitly" % k)
        if x.size(0) == 0:                               # nothing to compute, nothing to launch
            return x.new_zeros(0, n)
        if bias is not None:
            bias = bias.detach().contiguous().float()
        f16 = self.mode == "3xf16" and self.flag is not None
        split = self._split(weight, f16)
        x = x.contiguous().float()
        # the kernels need a leading dimension that is a multiple of 4: pad the row stride, hand back a view
        out = None
        if n % 4:
            out = torch.empty(x.size(0), (n + 3) // 4 * 4, dtype=torch.float32, device=x.device)[:, :n]
        if f16:          # bias and ReLU ride in the GEMM's epilogue
            return _cabi.proj_gemm_3xf16(x, split[0], split[1], out=out, overflow=self.flag, bias=bias, relu=relu)
        y = _cabi.proj_gemm_3xtf32(x, split[0], split[1], out=out)
        if bias is not None:
            y.add_(bias)
        return torch.relu_(y) if relu else y

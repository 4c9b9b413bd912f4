"""Shim of torch_scatter.{scatter, scatter_add, scatter_mean} for dim=0 (see ../README.md)."""
from oracle.pyg_semantics import scatter_mean as _scatter_mean
from oracle.pyg_semantics import scatter_sum as _scatter_sum


def _size(index, dim_size):
    return int(index.max()) + 1 if dim_size is None else dim_size


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    assert dim == 0 and out is None
    return _scatter_sum(src, index, _size(index, dim_size))


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    assert dim == 0 and out is None
    return _scatter_mean(src, index, _size(index, dim_size))


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None
    if reduce in ("add", "sum"):
        return _scatter_sum(src, index, _size(index, dim_size))
    if reduce == "mean":
        return _scatter_mean(src, index, _size(index, dim_size))
    raise NotImplementedError(reduce)

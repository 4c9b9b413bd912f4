"""Shim of the torch_geometric API surface used by the GraphVQA reference (see ../README.md)."""
from . import typing, utils, nn, data  # noqa: F401

__version__ = "1.6.3+shim"

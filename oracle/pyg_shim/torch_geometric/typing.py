"""Type aliases only (torch_geometric.typing)."""
from typing import Optional, Tuple, Union

from torch import Tensor

Adj = Union[Tensor, "SparseTensor"]
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
PairOptTensor = Tuple[Optional[Tensor], Optional[Tensor]]
Size = Optional[Tuple[int, int]]
NoneType = Optional[Tensor]

"""Op-for-op CPU restatement of PyG's `LGConv` (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference imports the operator at `src/lightgcn.py:9`, builds K of them at `:82` and calls
them at `:96`. PyG is absent from this image and un-pinned upstream (`requirements.txt:10`), so
the algorithm is restated from the published PyG 2.2/2.3 sources:

* `torch_geometric/nn/conv/lg_conv.py::LGConv.forward` (normalize=True, aggr='add'),
* `torch_geometric/nn/conv/gcn_conv.py::gcn_norm(edge_index, edge_weight, num_nodes,
  add_self_loops=False, flow='source_to_target')`,
* `MessagePassing.propagate` -> `message` (`edge_weight.view(-1,1) * x_j`) -> `aggregate`
  (`torch_scatter.scatter_sum` == `Tensor.scatter_add_`).

Every step bottoms out in a torch-native op, so on CPU this is the same arithmetic in the same
order as the reference's PyG path for a dense `edge_index`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor


def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int,
             dtype: torch.dtype = torch.float32) -> Tuple[Tensor, Tensor, Tensor]:
    """Symmetric normalisation without self-loops. Returns (w_hat, deg, deg_inv_sqrt)."""
    row, col = edge_index[0], edge_index[1]
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype, device=edge_weight.device)
    deg.scatter_add_(0, col, edge_weight)            # weighted in-degree of the TARGET node
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)          # isolated nodes contribute nothing
    w_hat = dis[row] * edge_weight * dis[col]         # evaluated left to right
    return w_hat, deg, dis


class LGConv(torch.nn.Module):
    """`LGConv(normalize=True)`: x'_i = sum_{j->i} w_hat_ji * x_j; no parameters, no bias."""

    def __init__(self, normalize: bool = True, **kwargs):
        super().__init__()
        self.normalize = normalize

    def reset_parameters(self):
        pass

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        row, col = edge_index[0], edge_index[1]
        if self.normalize:
            edge_weight, _, _ = gcn_norm(edge_index, edge_weight, x.size(0), x.dtype)
        x_j = x.index_select(0, row)                                   # [nnz, d] source rows
        msg = x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j
        out = torch.zeros_like(x)
        out.scatter_add_(0, col.view(-1, 1).expand_as(msg), msg)       # sum at the targets
        return out


def dense_normalised_adjacency(edge_index: Tensor, edge_weight: Tensor, num_nodes: int) -> Tensor:
    """Independent fp64 cross-check: the dense matrix D^-1/2 A_w D^-1/2 (tiny graphs only)."""
    a = torch.zeros(num_nodes, num_nodes, dtype=torch.float64)
    a.index_put_((edge_index[1], edge_index[0]), edge_weight.double(), accumulate=True)
    deg = a.sum(dim=1)
    dis = torch.where(deg > 0, deg.rsqrt(), torch.zeros_like(deg))
    return dis.view(-1, 1) * a * dis.view(1, -1)

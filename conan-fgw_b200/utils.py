"""``torch_geometric.utils.to_dense_batch`` / ``to_dense_adj`` on the library's kernels.

ConAN calls them to prepare the FGW barycenter inputs (``schnet_no_sum.py:242-253``, ``visnet.py:168-177``): node
features and adjacency of every conformer as dense, zero-padded ``[G, n_max, ...]`` blocks.  Same signatures as PyG;
the barycenter solver itself stays on the reference path.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops


def _segments(batch, batch_size, num_nodes, device):
    if batch is None:
        return torch.tensor([0, num_nodes], dtype=torch.int32, device=device), 1
    return ops.segments_from_batch(batch, batch_size)


def _max_nodes(seg_ptr, max_num_nodes):
    if max_num_nodes is not None:
        return int(max_num_nodes)
    if seg_ptr.numel() < 2:
        return 0
    return int((seg_ptr[1:] - seg_ptr[:-1]).max().item())      # host sync, as in PyG


def to_dense_batch(x: torch.Tensor, batch: Optional[torch.Tensor] = None, fill_value: float = 0.0,
                   max_num_nodes: Optional[int] = None, batch_size: Optional[int] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``[N, C]`` node features -> (``[B, n_max, C]`` dense features, ``[B, n_max]`` bool mask); ``batch`` sorted.
    No gradient (the reference feeds the result to the POT barycenter solver, which is not differentiated through)."""
    if not x.is_cuda:
        raise _lib.ConanMPError("to_dense_batch: x must be a CUDA tensor (no CPU path)")
    if x.dim() != 2:
        raise ValueError("to_dense_batch: x must be [N, C]")
    x = x.detach().to(torch.float32).contiguous()
    seg_ptr, B = _segments(batch, batch_size, x.size(0), x.device)
    n_max = _max_nodes(seg_ptr, max_num_nodes)
    out = torch.empty(B, n_max, x.size(1), dtype=torch.float32, device=x.device)
    mask = torch.empty(B, n_max, dtype=torch.uint8, device=x.device)
    _lib.call("cmp_dense_batch", _lib.ptr(x), _lib.ptr(seg_ptr), B, n_max, x.size(1), float(fill_value), _lib.ptr(out),
              _lib.ptr(mask))
    return out, mask.bool()


def to_dense_adj(edge_index: torch.Tensor, batch: Optional[torch.Tensor] = None, edge_attr=None,
                 max_num_nodes: Optional[int] = None, batch_size: Optional[int] = None) -> torch.Tensor:
    """``edge_index [2, E]`` -> ``[B, n_max, n_max]`` adjacency (``adj[b, src, dst]`` counts the edges)."""
    if edge_attr is not None:
        raise NotImplementedError("to_dense_adj: edge_attr is not used by ConAN and not provided")
    if not edge_index.is_cuda:
        raise _lib.ConanMPError("to_dense_adj: edge_index must be a CUDA tensor (no CPU path)")
    ei = edge_index.to(torch.int64).contiguous()
    E = ei.size(1)
    if batch is None:
        num_nodes = int(ei.max().item()) + 1 if E else 0
        bt = None
    else:
        num_nodes = batch.numel()
        bt = batch.to(torch.int64).contiguous()
    seg_ptr, B = _segments(bt, batch_size, num_nodes, ei.device)
    n_max = _max_nodes(seg_ptr, max_num_nodes)
    adj = torch.empty(B, n_max, n_max, dtype=torch.float32, device=ei.device)
    _lib.call("cmp_dense_adj", _lib.ptr(ei), E, _lib.ptr(bt), _lib.ptr(seg_ptr), B, n_max, _lib.ptr(adj))
    return adj

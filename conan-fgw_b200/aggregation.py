"""Conformer-level glue around the backbone, on the library's kernels.

Mirrors, with the reference's attribute names (checkpoints load unchanged):

* ``create_aggregation_index`` - ``conan_fgw/src/model/common.py:414-423``: a Python loop over the batch plus a
  host->device copy every step there; here the index is one ``arange // K`` on the device.
* ``MeanAggregation`` - ``torch_geometric.nn.aggr.MeanAggregation`` as ConAN calls it
  (``common.py:404,410``; ``schnet_based_models.py:61,79,171,242``): sorted index -> segment sums
  (``cmp_batch_to_segments`` + ``cmp_segment_sum_fwd/bwd``) divided by the segment sizes.
* ``ConformerAggregationHead`` - the tail of ``EmbeddingsWithGATAggregation[BaryCenter][Classification].forward``
  (``schnet_based_models.py:164-172,236-244,296-303,358-367``): ``W3d x_3d + Wcov x_cov (+ agg_weight Wbary x_bary)``
  -> mean over the K conformers of a molecule -> ``build_mlp`` / ``build_mlp_class`` head (``:17-45``).
  The 2-D GAT branch and the FGW barycenter stay on the reference path: their outputs come in as tensors.
"""

from __future__ import annotations

import torch
from torch import nn

from . import ops
from .nn import Linear


def create_aggregation_index(num_conformer_graphs: int, num_conformers: int, device) -> torch.Tensor:
    """``index[i] = i // K`` for every conformer graph of the batch (molecules own K consecutive graphs).
    Same values and length as the reference loop (which always appends whole groups of K)."""
    K = int(num_conformers)
    n = -(-int(num_conformer_graphs) // K) * K
    return torch.arange(n, dtype=torch.int64, device=device) // K


class MeanAggregation(nn.Module):
    """``aggr.MeanAggregation()(x, index)`` over a SORTED index (conformers of a molecule are consecutive)."""

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2):
        if dim not in (0, -2) or x.dim() != 2:
            raise ValueError("MeanAggregation: only [N, C] inputs reduced over dim 0 are provided")
        if index is None:
            return ops.segment_sum(x, torch.tensor([0, x.size(0)], dtype=torch.int32, device=x.device), 1) \
                * (1.0 / max(x.size(0), 1))
        if index.numel() != x.size(0):
            raise ValueError(f"MeanAggregation: index has {index.numel()} entries for {x.size(0)} rows")
        seg_ptr, G = ops.segments_from_batch(index, dim_size)
        counts = (seg_ptr[1:] - seg_ptr[:-1]).clamp(min=1).to(torch.float32)
        return ops.segment_sum(x, seg_ptr, G) / counts.unsqueeze(1)


def build_mlp(out_channels: int, is_complex: bool = False):
    """``schnet_based_models.py:17-29`` on the library's Linear (ReLU / Dropout act on [molecules, C] tensors)."""
    if is_complex:
        return nn.Sequential(Linear(out_channels, out_channels // 2), nn.Dropout(0.02), nn.ReLU(),
                             Linear(out_channels // 2, 1), nn.Dropout(0.02))
    return Linear(out_channels, 1)


def build_mlp_class(out_channels: int, is_complex: bool = False):
    """``schnet_based_models.py:32-45``."""
    if is_complex:
        return nn.Sequential(Linear(out_channels, out_channels), nn.ReLU(), Linear(out_channels, out_channels // 2),
                             nn.ReLU(), Linear(out_channels // 2, 1))
    return Linear(out_channels, 1)


class ConformerAggregationHead(nn.Module):
    def __init__(self, out_channels: int, use_covalent: bool = True, use_barycenter: bool = False,
                 agg_weight: float = 0.2, classification: bool = False, is_complex: bool = False):
        super().__init__()
        self.transformation_matrix_3d = Linear(out_channels, out_channels)
        if use_covalent:
            self.transformation_matrix_cov = Linear(out_channels, out_channels)
        if use_barycenter:
            self.transformation_matrix_bary = Linear(out_channels, out_channels)
        self.molecular_regression_lin = (build_mlp_class if classification else build_mlp)(out_channels, is_complex)
        self.conformers_mean_aggr = MeanAggregation()
        self.use_covalent, self.use_barycenter, self.agg_weight = use_covalent, use_barycenter, float(agg_weight)

    def forward(self, x_3d, conformers_index, x_covalent=None, x_bary=None, num_molecules=None):
        # the sum of the branches rides in the residual input of the GEMM kernel: one launch per branch, no adds
        x = self.transformation_matrix_3d(x_3d)
        if self.use_covalent:
            if x_covalent is None:
                raise ValueError("this head was built with use_covalent=True: pass the GAT embedding")
            x = self.transformation_matrix_cov(x_covalent, residual=x)
        if self.use_barycenter:
            if x_bary is None:
                raise ValueError("this head was built with use_barycenter=True: pass the barycenter embedding")
            x = x + self.agg_weight * self.transformation_matrix_bary(x_bary)
        x = self.conformers_mean_aggr(x, conformers_index, dim_size=num_molecules)
        return self.molecular_regression_lin(x)

"""ORACLE (test infrastructure only) - CPU restatement of ConAN's conformer-level glue around the backbone.

Follows ``conan_fgw/src/model/common.py:414-423`` (``create_aggregation_index``),
``torch_geometric.nn.aggr.MeanAggregation`` as used at ``common.py:404,410`` and
``schnet_based_models.py:61,79,171,242`` (scatter-mean of consecutive conformers), and the tail of
``EmbeddingsWithGATAggregation[BaryCenter].forward`` (``schnet_based_models.py:164-172,236-244``):
``x = W3d x_3d + Wcov x_cov (+ agg_weight * Wbary x_bary)`` -> mean over the K conformers -> head
(``build_mlp`` / ``build_mlp_class``, ``schnet_based_models.py:17-45``).

Parity unpinned against PyG (un-vendored): anchored on the reference call sites and closed-form cases.
"""
import torch
from torch.nn import Dropout, Linear, ReLU, Sequential


def create_aggregation_index(num_conformer_graphs: int, num_conformers: int) -> torch.Tensor:
    """The reference's Python loop, literally (common.py:414-423)."""
    index, mol_idx, i = [], -1, 0
    while i < num_conformer_graphs:
        mol_idx += 1
        for _ in range(num_conformers):
            index.append(mol_idx)
            i += 1
    return torch.tensor(index, dtype=torch.long)


def mean_aggregation(x: torch.Tensor, index: torch.Tensor, dim_size=None) -> torch.Tensor:
    """PyG MeanAggregation(x, index): scatter-sum / count (empty segments -> 0)."""
    n = int(index.max()) + 1 if dim_size is None else int(dim_size)
    out = torch.zeros(n, x.size(1), dtype=x.dtype)
    out.index_add_(0, index, x)
    cnt = torch.zeros(n, dtype=x.dtype).index_add_(0, index, torch.ones_like(index, dtype=x.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


def build_mlp(out_channels: int, is_complex: bool = False):
    if is_complex:
        return Sequential(Linear(out_channels, out_channels // 2), Dropout(0.02), ReLU(),
                          Linear(out_channels // 2, 1), Dropout(0.02))
    return Linear(out_channels, 1)


def build_mlp_class(out_channels: int, is_complex: bool = False):
    if is_complex:
        return Sequential(Linear(out_channels, out_channels), ReLU(), Linear(out_channels, out_channels // 2), ReLU(),
                          Linear(out_channels // 2, 1))
    return Linear(out_channels, 1)


class ConformerAggregationHead(torch.nn.Module):
    """Tail of EmbeddingsWithGATAggregation[BaryCenter][Classification].forward, attribute names as there."""

    def __init__(self, out_channels, use_covalent=True, use_barycenter=False, agg_weight=0.2, classification=False,
                 is_complex=False):
        super().__init__()
        self.transformation_matrix_3d = Linear(out_channels, out_channels)
        if use_covalent:
            self.transformation_matrix_cov = Linear(out_channels, out_channels)
        if use_barycenter:
            self.transformation_matrix_bary = Linear(out_channels, out_channels)
        self.molecular_regression_lin = (build_mlp_class if classification else build_mlp)(out_channels, is_complex)
        self.use_covalent, self.use_barycenter, self.agg_weight = use_covalent, use_barycenter, agg_weight

    def forward(self, x_3d, conformers_index, x_covalent=None, x_bary=None):
        x = self.transformation_matrix_3d(x_3d)
        if self.use_covalent:
            x = x + self.transformation_matrix_cov(x_covalent)
        if self.use_barycenter:
            x = x + self.agg_weight * self.transformation_matrix_bary(x_bary)
        x = mean_aggregation(x, conformers_index)
        return self.molecular_regression_lin(x)

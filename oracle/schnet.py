"""Oracle (test infrastructure): SchNet stack as the reference executes it.

Restates torch-geometric 2.3.0 ``torch_geometric/nn/models/schnet.py``
(``SchNet, RadiusInteractionGraph, InteractionBlock, CFConv, GaussianSmearing,
ShiftedSoftplus``; un-vendored dependency pinned at ``environment.yml:163``) in
the unfused op order PyG runs it - materialised ``rbf[E, Ng]``, E-row Linears,
``index_select`` gather, ``index_add_`` scatter - plus the ConAN deltas of
``conan_fgw/src/model/graph_embeddings/schnet_no_sum.py``:

* ``:109-130``  constructor order and the extra ``lin1_bary / lin2_bary / lin2``;
* ``:159-164``  trunk loop ``h = h + interaction(h, edge_index, edge_weight, edge_attr)``;
* ``:177-186``  head order ``lin1 -> lin2 -> act`` then sum readout;
* ``:225-231``  the two heads of ``forward_3d_bary``.

Parameter / buffer names equal PyG's, including the aliased
``interactions.{t}.conv.nn.*`` keys, so one ``state_dict`` loads into both this
oracle and the CUDA modules.  Parity status: unpinned (see ``oracle/__init__``).
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .radius import radius_graph_ref


class ShiftedSoftplus(nn.Module):
    # PyG: softplus(x) - log(2)   
    def __init__(self):
        super().__init__()
        self.shift = math.log(2.0)

    def forward(self, x):
        return F.softplus(x) - self.shift


class GaussianSmearing(nn.Module):
    # PyG: offset = linspace(start, stop, Ng); coeff = -0.5 / delta**2   
    def __init__(self, start: float = 0.0, stop: float = 5.0, num_gaussians: int = 50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)

    def forward(self, dist):
        d = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * d.pow(2))


class RadiusInteractionGraph(nn.Module):
    def __init__(self, cutoff: float = 10.0, max_num_neighbors: int = 32):
        super().__init__()
        self.cutoff = cutoff
        self.max_num_neighbors = max_num_neighbors

    def forward(self, pos, batch):
        ei = radius_graph_ref(pos, self.cutoff, batch, max_num_neighbors=self.max_num_neighbors)
        ei = ei.to(pos.device)
        row, col = ei[0], ei[1]
        ew = (pos[row] - pos[col]).norm(dim=-1)
        return ei, ew


class CFConv(nn.Module):
    """Continuous-filter convolution; message ``x_j * W`` summed at the target."""

    def __init__(self, in_channels, out_channels, num_filters, net, cutoff):
        super().__init__()
        self.lin1 = nn.Linear(in_channels, num_filters, bias=False)
        self.lin2 = nn.Linear(num_filters, out_channels)
        self.nn = net
        self.cutoff = cutoff
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.lin1.weight)
        nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight, edge_attr):
        C = 0.5 * (torch.cos(edge_weight * math.pi / self.cutoff) + 1.0)
        W = self.nn(edge_attr) * C.view(-1, 1)
        x = self.lin1(x)
        msg = x.index_select(0, edge_index[0]) * W            # x_j = x[source]
        out = torch.zeros_like(x).index_add_(0, edge_index[1], msg)
        return self.lin2(out)


class InteractionBlock(nn.Module):
    def __init__(self, hidden_channels, num_gaussians, num_filters, cutoff):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(num_gaussians, num_filters),
            ShiftedSoftplus(),
            nn.Linear(num_filters, num_filters),
        )
        self.conv = CFConv(hidden_channels, hidden_channels, num_filters, self.mlp, cutoff)
        self.act = ShiftedSoftplus()
        self.lin = nn.Linear(hidden_channels, hidden_channels)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.mlp[0].weight)
        self.mlp[0].bias.data.fill_(0)
        nn.init.xavier_uniform_(self.mlp[2].weight)
        self.mlp[2].bias.data.fill_(0)
        self.conv.reset_parameters()
        nn.init.xavier_uniform_(self.lin.weight)
        self.lin.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight, edge_attr):
        x = self.conv(x, edge_index, edge_weight, edge_attr)
        x = self.act(x)
        return self.lin(x)


def segment_sum(x, index, num_segments=None):
    if index is None:
        return x.sum(dim=0, keepdim=True)
    if num_segments is None:
        num_segments = int(index.max()) + 1 if index.numel() else 0
    out = x.new_zeros((num_segments,) + tuple(x.shape[1:]))
    return out.index_add_(0, index, x)


class SumReadout(nn.Module):
    """PyG ``SumAggregation``: ``readout(x, index, dim=0)``."""

    def forward(self, x, index=None, dim=0, dim_size=None):
        assert dim == 0
        return segment_sum(x, index, dim_size)


class SchNet(nn.Module):
    """PyG 2.3.0 ``SchNet`` (regression trunk; dipole / atomref variants not used by ConAN)."""

    def __init__(self, hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50,
                 cutoff=10.0, interaction_graph=None, max_num_neighbors=32, readout="add",
                 dipole=False, mean=None, std=None, atomref=None):
        super().__init__()
        assert not dipole and atomref is None, "oracle covers the configurations ConAN instantiates"
        self.hidden_channels = hidden_channels
        self.num_filters = num_filters
        self.num_interactions = num_interactions
        self.num_gaussians = num_gaussians
        self.cutoff = cutoff
        self.dipole = dipole
        self.mean, self.std, self.scale = mean, std, None
        self.readout = SumReadout()
        self.embedding = nn.Embedding(100, hidden_channels, padding_idx=0)
        self.interaction_graph = interaction_graph or RadiusInteractionGraph(cutoff, max_num_neighbors)
        self.distance_expansion = GaussianSmearing(0.0, cutoff, num_gaussians)
        self.interactions = nn.ModuleList(
            InteractionBlock(hidden_channels, num_gaussians, num_filters, cutoff)
            for _ in range(num_interactions)
        )
        self.lin1 = nn.Linear(hidden_channels, hidden_channels // 2)
        self.act = ShiftedSoftplus()
        self.lin2 = nn.Linear(hidden_channels // 2, 1)
        self.reset_parameters()

    def reset_parameters(self):
        self.embedding.reset_parameters()
        for blk in self.interactions:
            blk.reset_parameters()
        nn.init.xavier_uniform_(self.lin1.weight)
        self.lin1.bias.data.fill_(0)
        nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)

    def trunk(self, z, pos, batch):
        h = self.embedding(z)
        edge_index, edge_weight = self.interaction_graph(pos, batch)
        edge_attr = self.distance_expansion(edge_weight)
        for blk in self.interactions:
            h = h + blk(h, edge_index, edge_weight, edge_attr)
        return h

    def forward(self, z, pos, batch=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h = self.trunk(z, pos, batch)
        h = self.lin2(self.act(self.lin1(h)))          # PyG head order
        if self.mean is not None and self.std is not None:
            h = h * self.std + self.mean
        out = self.readout(h, batch, dim=0)
        if self.scale is not None:
            out = self.scale * out
        return out


COVALENT_BONDS_ATTRS_DIM = 3      # schnet_no_sum.py:23


class SchNetNoSum(SchNet):
    """ConAN's backbone, ``schnet_no_sum.py:90-232`` (both the plain and the ``use_covalent`` trunk)."""

    def __init__(self, device=None, hidden_channels=128, num_filters=128, num_interactions=6,
                 num_gaussians=50, cutoff=10.0, interaction_graph=None, max_num_neighbors=32,
                 readout="add", dipole=False, mean=None, std=None, atomref=None,
                 use_covalent=False, use_readout=True):
        super().__init__(hidden_channels, num_filters, num_interactions, num_gaussians, cutoff,
                         interaction_graph, max_num_neighbors, readout, dipole, mean, std, atomref)
        self.device = device
        self.use_readout = use_readout
        self.use_covalent = use_covalent
        half = hidden_channels // 2
        self.lin1_bary = nn.Linear(hidden_channels, half)
        self.lin2_bary = nn.Linear(half, half)
        self.lin2 = nn.Linear(half, half)
        if use_covalent:                                   # schnet_no_sum.py:131-142
            self.interactions_cov = nn.ModuleList(
                InteractionBlock(hidden_channels, COVALENT_BONDS_ATTRS_DIM, num_filters, cutoff)
                for _ in range(num_interactions))
            self.lin1 = nn.Linear(hidden_channels * 2, half)
            self.lin1_bary = nn.Linear(hidden_channels * 2, half)

    def full_trunk(self, z, pos, batch, data_batch):
        h = self.trunk(z, pos, batch)
        if self.use_covalent:                              # schnet_no_sum.py:166-175
            h_cov = self.embedding(z)
            ei = data_batch.edge_index
            ew = torch.ones(ei.shape[1], dtype=torch.float32)
            ea = data_batch.edge_attr.float()
            for blk in self.interactions_cov:
                h_cov = h_cov + blk(h_cov, ei, ew, ea)
            h = torch.cat([h, h_cov], dim=1)
        return h

    def forward(self, z, pos, batch=None, data_batch=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h = self.full_trunk(z, pos, batch, data_batch)
        h = self.act(self.lin2(self.lin1(h)))          # ConAN head order: lin1 -> lin2 -> ssp
        return self.readout(h, batch, dim=0) if self.use_readout else h

    def forward_3d_bary(self, z, pos, batch=None, data_batch=None):
        batch = torch.zeros_like(z) if batch is None else batch
        hs = self.full_trunk(z, pos, batch, data_batch)
        h = self.act(self.lin2(self.lin1(hs)))
        hb = self.act(self.lin2_bary(self.lin1_bary(hs)))
        return h, hb


class SchNetWithMultipleReturns(SchNet):
    """``schnet_no_sum.py:357-450``: per-atom ``ssp(lin1(h))`` plus the radius graph and its Gaussian expansion."""

    def __init__(self, hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=10.0,
                 interaction_graph=None, max_num_neighbors=32, readout="add", dipole=False, mean=None, std=None,
                 atomref=None, use_covalent=False, use_readout=True):
        super().__init__(hidden_channels, num_filters, num_interactions, num_gaussians, cutoff,
                         interaction_graph, max_num_neighbors, readout, dipole, mean, std, atomref)
        self.use_readout = use_readout
        self.use_covalent = use_covalent
        if use_covalent:
            self.interactions_cov = nn.ModuleList(
                InteractionBlock(hidden_channels, COVALENT_BONDS_ATTRS_DIM, num_filters, cutoff)
                for _ in range(num_interactions))
            self.lin1 = nn.Linear(hidden_channels * 2, hidden_channels // 2)

    def forward(self, z, pos, batch=None, data_batch=None, conformers_index=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h = self.embedding(z)
        edge_index, edge_weight = self.interaction_graph(pos, batch)
        edge_attr = self.distance_expansion(edge_weight)
        for blk in self.interactions:
            h = h + blk(h, edge_index, edge_weight, edge_attr)
        if self.use_covalent:
            h_cov = self.embedding(z)
            ei = data_batch.edge_index
            ew = torch.ones(ei.shape[1], dtype=torch.float32)
            ea = data_batch.edge_attr.float()
            for blk in self.interactions_cov:
                h_cov = h_cov + blk(h_cov, ei, ew, ea)
            h = torch.cat([h, h_cov], dim=1)
        h = self.act(self.lin1(h))
        return h, edge_index, edge_attr


def to_double(module: nn.Module) -> nn.Module:
    """fp64 twin used as ground truth when judging fp32 error budgets."""
    import copy

    m = copy.deepcopy(module).double()
    return m

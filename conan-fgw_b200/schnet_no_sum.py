"""``SchNetNoSum`` - ConAN's SchNet backbone on the sm_100a kernels.

Same constructor, methods and ``state_dict`` as the reference class
(``conan_fgw/src/model/graph_embeddings/schnet_no_sum.py:90-354``):

* ``forward(z, pos, batch)``            -> ``[G, H/2]`` (or ``[N, H/2]`` with ``use_readout=False``), ``:144-188``
* ``forward_3d_bary(z, pos, batch)``    -> two per-atom heads ``([N, H/2], [N, H/2])``, ``:190-232``
* ``forward_w_barycenter(...)``         -> trunk + heads here; the FGW barycenter itself stays on
  the reference path (out of scope, BASELINE.json north_star) and is reached through the
  ``barycenter_fn`` hook, which receives exactly what ``_compute_barycenter`` receives (``:344-350``).
  The reference runs a second, redundant radius search at ``:342``; here the CSR built for the
  trunk is reused.

The head order is ConAN's ``lin1 -> lin2 -> ssp`` (``:177-179``), not PyG's.
"""

from __future__ import annotations

from typing import Callable, Optional

import torch

from torch import nn

from . import _lib
from .nn import InteractionBlock, Linear, SchNet  # noqa: F401

COVALENT_BONDS_ATTRS_DIM = 3      # schnet_no_sum.py:23


def _covalent_trunk(model, z, data_batch, status=None):
    """The 2-D (covalent bond) interaction stack of ``schnet_no_sum.py:166-175``: the same InteractionBlocks on the
    bond graph with unit edge weights and the 3 bond attributes as ``edge_attr`` (generic CFConv entry: arbitrary
    ``edge_index`` / ``edge_attr``, filter MLP on the given attributes, CSR message kernels)."""
    if data_batch is None:
        raise ValueError("use_covalent=True needs data_batch with .edge_index and .edge_attr")
    h_cov = model.embed(z, status)
    ei = data_batch.edge_index
    ew = torch.ones(ei.shape[1], dtype=torch.float32, device=z.device)
    ea = data_batch.edge_attr.float().contiguous()
    for blk in model.interactions_cov:
        h_cov = blk(h_cov, ei, ew, ea, residual=h_cov)
    return h_cov


class SchNetNoSum(SchNet):
    def __init__(self, device=None, hidden_channels: int = 128, num_filters: int = 128, num_interactions: int = 6,
                 num_gaussians: int = 50, cutoff: float = 10.0, interaction_graph: Optional[Callable] = None,
                 max_num_neighbors: int = 32, readout: str = "add", dipole: bool = False,
                 mean: Optional[float] = None, std: Optional[float] = None, atomref=None,
                 use_covalent: bool = False, use_readout: bool = True):
        super().__init__(hidden_channels, num_filters, num_interactions, num_gaussians, cutoff, interaction_graph,
                         max_num_neighbors, readout, dipole, mean, std, atomref)
        self.device = device
        self.use_readout = use_readout
        self.use_covalent = use_covalent
        half = hidden_channels // 2
        # created after reset_parameters(): torch default init, exactly as at schnet_no_sum.py:126-130
        self.lin1_bary = Linear(hidden_channels, half)
        self.lin2_bary = Linear(half, half)
        self.lin2 = Linear(half, half)
        if use_covalent:      # schnet_no_sum.py:131-142 (model name "schnet_covalent", common.py:530-531)
            self.interactions_cov = nn.ModuleList(
                InteractionBlock(hidden_channels, COVALENT_BONDS_ATTRS_DIM, num_filters, cutoff)
                for _ in range(num_interactions))
            self.lin1 = Linear(hidden_channels * 2, half)
            self.lin1_bary = Linear(hidden_channels * 2, half)
        self.barycenter_fn: Optional[Callable] = None

    def full_trunk(self, z, pos, batch, data_batch=None, num_graphs=None):
        h, graph = self.trunk(z, pos, batch, num_graphs)
        if self.use_covalent:
            h = torch.cat([h, _covalent_trunk(self, z, data_batch, graph.status)], dim=1)
        return h, graph

    def _head(self, h, lin1, lin2):
        return lin2(lin1(h), act=_lib.ACT_SSP)

    def forward(self, z, pos, batch=None, data_batch=None, num_graphs=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h, graph = self.full_trunk(z, pos, batch, data_batch, num_graphs)
        h = self._head(h, self.lin1, self.lin2)
        if self.use_readout:
            return self.readout(h, batch, dim=0, seg_ptr=graph.seg_ptr if graph.G else None)
        return h

    def forward_3d_bary(self, z, pos, batch=None, data_batch=None, num_graphs=None, return_graph=False):
        batch = torch.zeros_like(z) if batch is None else batch
        hs, graph = self.full_trunk(z, pos, batch, data_batch, num_graphs)
        h = self._head(hs, self.lin1, self.lin2)
        hb = self._head(hs, self.lin1_bary, self.lin2_bary)
        if return_graph:
            return h, hb, graph
        return h, hb

    def forward_w_barycenter(self, z, pos, num_conformers: int, batch=None, data_batch=None, max_iter: int = 100,
                             epsilon: float = 0.1, num_graphs=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h_3d, h_bary, graph = self.forward_3d_bary(z, pos, batch, data_batch=data_batch, num_graphs=num_graphs,
                                                   return_graph=True)
        if self.barycenter_fn is None:
            raise RuntimeError(
                "forward_w_barycenter: the FGW barycenter stays on the reference path; set "
                "`model.barycenter_fn = <reference SchNetNoSum._compute_barycenter bound to its solver>` "
                "(see INTEGRATION.md)")
        batch_size = int(graph.G / num_conformers)
        _, h_bary = self.barycenter_fn(node_feature=h_bary, edge_index=graph.edge_index(), batch=batch,
                                       batch_size=batch_size, num_conformers=num_conformers)
        h_3d = self.readout(h_3d, batch, dim=0, seg_ptr=graph.seg_ptr)
        return h_3d, h_bary


class SchNetWithMultipleReturns(SchNet):
    """``schnet_no_sum.py:357-450``: per-atom ``ssp(lin1(h))`` together with the radius graph (PyG ``edge_index``) and
    its Gaussian expansion ``edge_attr`` - the inputs ESAN-style consumers expect.  Returning ``edge_index`` /
    ``edge_attr`` materialises them (one host sync for the edge count), so this class runs the interaction blocks through
    the module call sequence of the reference; in bf16 mode the tensor tags still route each block to the fused kernel."""

    def __init__(self, hidden_channels: int = 128, num_filters: int = 128, num_interactions: int = 6,
                 num_gaussians: int = 50, cutoff: float = 10.0, interaction_graph: Optional[Callable] = None,
                 max_num_neighbors: int = 32, readout: str = "add", dipole: bool = False,
                 mean: Optional[float] = None, std: Optional[float] = None, atomref=None,
                 use_covalent: bool = False, use_readout: bool = True):
        super().__init__(hidden_channels, num_filters, num_interactions, num_gaussians, cutoff, interaction_graph,
                         max_num_neighbors, readout, dipole, mean, std, atomref)
        self.use_readout = use_readout
        self.use_covalent = use_covalent
        if use_covalent:
            self.interactions_cov = nn.ModuleList(
                InteractionBlock(hidden_channels, COVALENT_BONDS_ATTRS_DIM, num_filters, cutoff)
                for _ in range(num_interactions))
            self.lin1 = Linear(hidden_channels * 2, hidden_channels // 2)

    def forward(self, z, pos, batch=None, data_batch=None, conformers_index=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h = self.embedding(z)
        edge_index, edge_weight = self.interaction_graph(pos, batch)
        edge_attr = self.distance_expansion(edge_weight)
        for interaction in self.interactions:
            h = interaction(h, edge_index, edge_weight, edge_attr, residual=h)
        if self.use_covalent:
            h = torch.cat([h, _covalent_trunk(self, z, data_batch)], dim=1)
        h = self.lin1(h, act=_lib.ACT_SSP)
        return h, edge_index, edge_attr

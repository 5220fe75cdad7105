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

from . import _lib
from .nn import InteractionBlock, Linear, SchNet  # noqa: F401


class SchNetNoSum(SchNet):
    def __init__(self, device=None, hidden_channels: int = 128, num_filters: int = 128, num_interactions: int = 6,
                 num_gaussians: int = 50, cutoff: float = 10.0, interaction_graph: Optional[Callable] = None,
                 max_num_neighbors: int = 32, readout: str = "add", dipole: bool = False,
                 mean: Optional[float] = None, std: Optional[float] = None, atomref=None,
                 use_covalent: bool = False, use_readout: bool = True):
        super().__init__(hidden_channels, num_filters, num_interactions, num_gaussians, cutoff, interaction_graph,
                         max_num_neighbors, readout, dipole, mean, std, atomref)
        if use_covalent:
            # model name "schnet_covalent" is not selectable from the ConAN CLI (config_parser.py:111-117)
            raise NotImplementedError("use_covalent=True is unreachable from the ConAN CLI and is not provided")
        self.device = device
        self.use_readout = use_readout
        self.use_covalent = use_covalent
        half = hidden_channels // 2
        # created after reset_parameters(): torch default init, exactly as at schnet_no_sum.py:126-130
        self.lin1_bary = Linear(hidden_channels, half)
        self.lin2_bary = Linear(half, half)
        self.lin2 = Linear(half, half)
        self.barycenter_fn: Optional[Callable] = None

    def _head(self, h, lin1, lin2):
        return lin2(lin1(h), act=_lib.ACT_SSP)

    def forward(self, z, pos, batch=None, data_batch=None, num_graphs=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h, graph = self.trunk(z, pos, batch, num_graphs)
        h = self._head(h, self.lin1, self.lin2)
        if self.use_readout:
            return self.readout(h, batch, dim=0, seg_ptr=graph.seg_ptr if graph.G else None)
        return h

    def forward_3d_bary(self, z, pos, batch=None, data_batch=None, num_graphs=None, return_graph=False):
        batch = torch.zeros_like(z) if batch is None else batch
        hs, graph = self.trunk(z, pos, batch, num_graphs)
        h = self._head(hs, self.lin1, self.lin2)
        hb = self._head(hs, self.lin1_bary, self.lin2_bary)
        if return_graph:
            return h, hb, graph
        return h, hb

    def forward_w_barycenter(self, z, pos, num_conformers: int, batch=None, data_batch=None, max_iter: int = 100,
                             epsilon: float = 0.1, num_graphs=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h_3d, h_bary, graph = self.forward_3d_bary(z, pos, batch, num_graphs=num_graphs, return_graph=True)
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

"""PyG-signature modules of the SchNet stack, backed by the sm_100a kernels.

Drop-in for the names ConAN imports from torch-geometric 2.3.0
(``conan_fgw/src/model/graph_embeddings/schnet_no_sum.py:6-9``, ``dimenet.py:9``,
``visnet.py:11``): ``SchNet, InteractionBlock, CFConv, GaussianSmearing,
ShiftedSoftplus, RadiusInteractionGraph, radius_graph``.  Constructor orders,
attribute names and ``state_dict`` keys equal PyG's -
including the aliased ``interactions.{t}.conv.nn.*`` entries - so existing ConAN
checkpoints load with ``strict=True``.
"""

from __future__ import annotations

import math
import os
from typing import Callable, Optional

import torch
from torch import nn

from . import _lib, ops
from .graph import NeighborList, build_neighbor_list, graph_from_edge_index, radius_graph  # noqa: F401


# fused (bf16-mode) trunk: chain the node linears of a block tail in one kernel (cmp_node_chain_fwd); False = one launch
# per Linear (kept for cross-checking)
CHAIN_NODE_LINEARS = True

# "fp32" precision: fused fp32-grade (three-pass) CFConv kernels whenever the graph qualifies (radius-built, promised
# max_atoms <= 128); False = always the exact kernels on materialised [E, *] tensors (same as precision "exact")
FUSED_FP32 = os.environ.get("CMP_FUSED_FP32", "1") != "0"
# "fp32" precision, fused CFConv: node linears on the split-bf16 tcgen05 kernels (chained block tails) instead of the
# exact SIMT GEMM.  Off by default: bf16 hi + lo pairs carry 16 significant bits, which over the 18 node GEMMs of a
# 6-block trunk costs 1e-5 .. 2e-5 on the embeddings and 4e-5 on the gradients (measured, tools/x3_errors.py) - outside
# the 1e-5 bar of this mode.  1.6 x faster (cfg 2: 234 K vs 146 K conformers/s); CMP_FP32_NODE_TC=1 or nn.FP32_NODE_TC.
FP32_NODE_TC = os.environ.get("CMP_FP32_NODE_TC", "0") != "0"


class Linear(nn.Linear):
    """``torch.nn.Linear`` parameters and init, forward through ``cmp_gemm_f32``."""

    tc = False   # True (set by SchNet.set_precision("bf16")): split-bf16 tcgen05 node GEMMs

    def forward(self, x, act=_lib.ACT_NONE, residual=None):
        return ops.linear(x, self.weight, self.bias, act, residual, tc=self.tc)


class Embedding(nn.Embedding):
    """``torch.nn.Embedding`` parameters and init, lookup / weight gradient through the C ABI."""

    def forward(self, z, status=None):
        return ops.embedding(z, self.weight, self.padding_idx, status)


class ShiftedSoftplus(nn.Module):
    def __init__(self):
        super().__init__()
        self.shift = math.log(2.0)

    def forward(self, x):
        return ops.shifted_softplus(x)


class GaussianSmearing(nn.Module):
    def __init__(self, start: float = 0.0, stop: float = 5.0, num_gaussians: int = 50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)

    def forward(self, dist):
        out = ops.gaussian_rbf(dist, self.offset, self.coeff)
        # remember what this expansion was computed from, so CFConv can run the fused geometric kernel
        out._cmp_rbf_of = (dist, self)
        return out


class RadiusInteractionGraph(nn.Module):
    def __init__(self, cutoff: float = 10.0, max_num_neighbors: int = 32):
        super().__init__()
        self.cutoff = cutoff
        self.max_num_neighbors = max_num_neighbors

    def neighbor_list(self, pos, batch, num_graphs=None, max_atoms=None, status=None) -> NeighborList:
        return build_neighbor_list(pos, batch, self.cutoff, self.max_num_neighbors, loop=False,
                                   num_graphs=num_graphs, max_atoms=max_atoms, status=status)

    def forward(self, pos, batch):
        nl = self.neighbor_list(pos, batch)
        return nl.edge_index(), nl.edge_weight()


class SumAggregation(nn.Module):
    """``aggr_resolver('add')``: ``readout(x, index, dim=0)`` over a sorted index."""

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=0, seg_ptr=None):
        if dim not in (0, -2):
            raise ValueError("SumAggregation: only dim=0 is provided")
        if index is None and seg_ptr is None:
            seg_ptr = torch.tensor([0, x.size(0)], dtype=torch.int32, device=x.device)
            return ops.segment_sum(x, seg_ptr, 1)
        if seg_ptr is None:
            seg_ptr, G = ops.segments_from_batch(index, dim_size)
        else:
            G = seg_ptr.numel() - 1
        return ops.segment_sum(x, seg_ptr, G)


class CFConv(nn.Module):
    """Continuous-filter convolution (PyG ``CFConv``, aggr='add')."""

    def __init__(self, in_channels: int, out_channels: int, num_filters: int, nn: nn.Sequential, cutoff: float):
        super().__init__()
        self.lin1 = Linear(in_channels, num_filters, bias=False)
        self.lin2 = Linear(num_filters, out_channels)
        self.nn = nn
        self.cutoff = cutoff
        # "exact": exact-fp32 kernels on materialised tensors;  "fp32": fp32-grade fused tcgen05 kernels where the graph
        # qualifies, else the exact kernels;  "bf16": fused tcgen05 kernels with an f16 filter MLP
        self.precision = "fp32"
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.lin1.weight)
        torch.nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)

    def _standard_mlp(self):
        net = self.nn
        return (isinstance(net, torch.nn.Sequential) and len(net) == 3 and isinstance(net[0], torch.nn.Linear)
                and isinstance(net[1], ShiftedSoftplus) and isinstance(net[2], torch.nn.Linear)
                and net[0].bias is not None and net[2].bias is not None)

    def fused_mode(self, graph, smearing):
        """Which fused geometric kernel applies to this layer on this graph: ``"f16"`` (bf16 mode: f16 filter MLP),
        ``"x3"`` (fp32 mode: fp32-grade three-pass kernels; needs a promised ``max_atoms`` <= 128) or ``None`` (the
        exact kernels).  Fused kernels need a radius-built graph, the Gaussian expansion of its distances, the standard
        filter MLP, supported (num_filters, num_gaussians) and an sm_100 device."""
        if self.precision not in ("bf16", "fp32") or (self.precision == "fp32" and not FUSED_FP32):
            return None
        if not (graph is not None and graph.G > 0 and graph.cutoff is not None
                and not graph.loop and isinstance(smearing, GaussianSmearing) and self._standard_mlp()
                and ops.fused_supported(self.lin1.out_features, smearing.offset.numel())
                and self.nn[0].in_features == smearing.offset.numel()):
            return None
        if self.precision == "bf16":
            return "f16"
        return "x3" if ops.x3_graph_ok(graph) else None

    def fused_ok(self, graph, smearing) -> bool:
        return self.fused_mode(graph, smearing) is not None

    def forward_fused(self, x, graph, smearing, act=_lib.ACT_NONE):
        xp = self.lin1(x)
        net = self.nn
        agg = ops.cfconv_fused(xp, net[0].weight, net[0].bias, net[2].weight, net[2].bias, graph, smearing.offset,
                               smearing.coeff, self.cutoff, x3=self.fused_mode(graph, smearing) == "x3")
        return self.lin2(agg, act=act)

    def filter(self, edge_attr):
        """``self.nn(edge_attr)``: Linear -> ShiftedSoftplus -> Linear with the activation fused."""
        net = self.nn
        if (isinstance(net, torch.nn.Sequential) and len(net) == 3 and isinstance(net[0], Linear)
                and isinstance(net[1], ShiftedSoftplus) and isinstance(net[2], Linear)):
            return net[2](net[0](edge_attr, act=_lib.ACT_SSP))
        return net(edge_attr)

    def forward(self, x, edge_index, edge_weight, edge_attr, graph: Optional[NeighborList] = None,
                act=_lib.ACT_NONE):
        if graph is None:
            graph, perm = graph_from_edge_index(edge_index, edge_weight, x.size(0))
            if perm is not None:
                edge_attr = edge_attr[perm]
        tag = getattr(edge_attr, "_cmp_rbf_of", None)
        if tag is not None and getattr(tag[0], "_cmp_graph", None) is graph and self.fused_ok(graph, tag[1]):
            # edge_attr is the Gaussian expansion of this graph's own distances: run the fused kernel
            return self.forward_fused(x, graph, tag[1], act=act)
        W = self.filter(edge_attr)
        xp = self.lin1(x)
        agg = ops.cfconv_message(xp, W, graph, self.cutoff)
        return self.lin2(agg, act=act)


class InteractionBlock(nn.Module):
    def __init__(self, hidden_channels: int, num_gaussians: int, num_filters: int, cutoff: float):
        super().__init__()
        self.mlp = nn.Sequential(
            Linear(num_gaussians, num_filters),
            ShiftedSoftplus(),
            Linear(num_filters, num_filters),
        )
        self.conv = CFConv(hidden_channels, hidden_channels, num_filters, self.mlp, cutoff)
        self.act = ShiftedSoftplus()
        self.lin = Linear(hidden_channels, hidden_channels)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.mlp[0].weight)
        self.mlp[0].bias.data.fill_(0)
        torch.nn.init.xavier_uniform_(self.mlp[2].weight)
        self.mlp[2].bias.data.fill_(0)
        self.conv.reset_parameters()
        torch.nn.init.xavier_uniform_(self.lin.weight)
        self.lin.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight, edge_attr, graph: Optional[NeighborList] = None,
                residual=None):
        # conv -> ssp (fused into conv.lin2's epilogue) -> lin (+ residual fused when the caller passes it)
        y = self.conv(x, edge_index, edge_weight, edge_attr, graph=graph, act=_lib.ACT_SSP)
        return self.lin(y, residual=residual)

    def forward_fused(self, x, graph, smearing, residual=None):
        y = self.conv.forward_fused(x, graph, smearing, act=_lib.ACT_SSP)
        return self.lin(y, residual=residual)


class SchNet(nn.Module):
    """PyG 2.3.0 ``SchNet`` (positional constructor order as used at ``schnet_no_sum.py:109-122``)."""

    def __init__(self, hidden_channels: int = 128, num_filters: int = 128, num_interactions: int = 6,
                 num_gaussians: int = 50, cutoff: float = 10.0, interaction_graph: Optional[Callable] = None,
                 max_num_neighbors: int = 32, readout: str = "add", dipole: bool = False,
                 mean: Optional[float] = None, std: Optional[float] = None, atomref=None):
        super().__init__()
        if dipole or atomref is not None:
            raise NotImplementedError("dipole / atomref variants are not instantiated by ConAN (SURVEY.md 2.1)")
        if readout not in ("add", "sum"):
            raise NotImplementedError("only the sum readout ConAN uses is provided")
        self.hidden_channels = hidden_channels
        self.num_filters = num_filters
        self.num_interactions = num_interactions
        self.num_gaussians = num_gaussians
        self.cutoff = cutoff
        self.dipole = dipole
        self.mean, self.std, self.scale = mean, std, None
        self.sum_aggr = SumAggregation()
        self.readout = SumAggregation()
        self.embedding = Embedding(100, hidden_channels, padding_idx=0)
        self.interaction_graph = interaction_graph if interaction_graph is not None else \
            RadiusInteractionGraph(cutoff, max_num_neighbors)
        self.distance_expansion = GaussianSmearing(0.0, cutoff, num_gaussians)
        self.interactions = nn.ModuleList(
            InteractionBlock(hidden_channels, num_gaussians, num_filters, cutoff) for _ in range(num_interactions))
        self.lin1 = Linear(hidden_channels, hidden_channels // 2)
        self.act = ShiftedSoftplus()
        self.lin2 = Linear(hidden_channels // 2, 1)
        self.register_buffer("initial_atomref", None)
        self.atomref = None
        self.precision = "fp32"
        # caller's bound on the atoms of any conformer it will feed (None: unknown).  With a bound <= 128 the fused
        # path launches the dense-block kernel only; a conformer that breaks the promise raises through the status word.
        self.max_atoms_hint = None
        # persistent device-side error word shared by every neighbour list this model builds (not part of the state_dict)
        self.register_buffer("status", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.reset_parameters()

    def check_status(self):
        """Raise for device-detected input errors of the forward passes since the last call (unsorted batch, atomic
        number out of range, capacity overflow, broken ``max_atoms_hint`` promise).  One 4-byte device -> host read."""
        from .graph import raise_for_status

        raise_for_status(self.status, reset=True)

    def set_precision(self, precision: str):
        """"exact": exact-fp32 kernels on materialised rbf / filter tensors (the reference's own op sequence).
        "fp32" (default): the 1e-5 parity mode - fp32-grade fused tcgen05 CFConv (hi + lo operand images, three MMA
        passes, fp32 epilogues) and split-bf16 node linears when the model was given ``max_atoms_hint`` <= 128 and the
        shapes are supported, otherwise the exact kernels.  "bf16": fused tcgen05 CFConv with an f16 filter MLP (fp32
        accumulation, fp32 node features); tolerance stated in DESIGN.md."""
        if precision not in ("exact", "fp32", "bf16"):
            raise ValueError("precision must be 'exact', 'fp32' or 'bf16'")
        self.precision = precision
        for m in self.modules():
            if isinstance(m, CFConv):
                m.precision = precision
            if isinstance(m, Linear):
                m.tc = precision == "bf16"
        return self

    def _node_tc(self, on: bool):
        """fp32 mode: the node linears follow the CFConv kernels (tcgen05 with the fused path, exact otherwise)."""
        if self.precision != "fp32":
            return
        mods = [m for blk in self.interactions for m in (blk.conv.lin1, blk.conv.lin2, blk.lin)]
        mods += [m for m in self.children() if isinstance(m, Linear)]       # heads
        for m in mods:
            m.tc = bool(on)

    def reset_parameters(self):
        self.embedding.reset_parameters()
        for blk in self.interactions:
            blk.reset_parameters()
        torch.nn.init.xavier_uniform_(self.lin1.weight)
        self.lin1.bias.data.fill_(0)
        torch.nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)

    # -- trunk shared by every forward variant ------------------------------------------------
    def embed(self, z, status=None):
        return self.embedding(z, status)

    def trunk(self, z, pos, batch, num_graphs=None):
        """Embedding + T interaction blocks.  Returns ``(h[N,H], graph)``."""
        ig = self.interaction_graph
        if isinstance(ig, RadiusInteractionGraph):
            graph = ig.neighbor_list(pos, batch, num_graphs, max_atoms=self.max_atoms_hint, status=self.status)
            h = self.embed(z, graph.status)
            modes = {blk.conv.fused_mode(graph, self.distance_expansion) for blk in self.interactions}
            x3 = modes == {"x3"}
            self._node_tc(x3 and FP32_NODE_TC)
            if len(modes) == 1 and None not in modes:
                # fused path: no edge_index / rbf[E, Ng] / filter[E, F] is ever materialised, no host sync
                blocks = list(self.interactions)
                chained = CHAIN_NODE_LINEARS and all(
                    ops.block_tail_supported(b.conv.lin2, b.lin, n.conv.lin1 if n is not None else None)
                    for b, n in zip(blocks, blocks[1:] + [None]))
                if not chained:
                    for blk in blocks:
                        h = blk.forward_fused(h, graph, self.distance_expansion, residual=h)
                    return h, graph
                # one kernel per block for the node linears: lin2 -> ssp -> lin (+ h) -> the NEXT block's lin1
                sm = self.distance_expansion
                x = blocks[0].conv.lin1(h)
                for blk, nxt in zip(blocks, blocks[1:] + [None]):
                    net = blk.conv.nn
                    agg = ops.cfconv_fused(x, net[0].weight, net[0].bias, net[2].weight, net[2].bias, graph, sm.offset,
                                           sm.coeff, blk.conv.cutoff, x3=x3)
                    if nxt is not None:
                        h, x = ops.block_tail(agg, h, blk.conv.lin2, blk.lin, nxt.conv.lin1)
                    else:
                        h = ops.block_tail(agg, h, blk.conv.lin2, blk.lin)
                return h, graph
            edge_index = None
            edge_weight = graph.edge_weight()      # first host sync: also surfaces device-side input errors
        else:  # user supplied interaction graph: generic path
            edge_index, edge_weight = ig(pos, batch)
            graph, perm = graph_from_edge_index(edge_index, edge_weight, z.numel())
            if perm is not None:
                edge_weight = edge_weight[perm]
            h = self.embed(z, graph.status)
        edge_attr = self.distance_expansion(edge_weight)
        for blk in self.interactions:
            h = blk(h, edge_index, edge_weight, edge_attr, graph=graph, residual=h)
        return h, graph

    def forward(self, z, pos, batch=None, num_graphs=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h, graph = self.trunk(z, pos, batch, num_graphs)
        h = self.lin2(self.lin1(h, act=_lib.ACT_SSP))
        if self.mean is not None and self.std is not None:
            h = h * self.std + self.mean
        seg = graph.seg_ptr if graph.G else None
        out = self.readout(h, batch, dim=0, seg_ptr=seg)
        if self.scale is not None:
            out = self.scale * out
        return out

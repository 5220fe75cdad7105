"""ViSNet on the sm_100a kernels - drop-in for the reference's vendored PyG ViSNet.

Same classes, constructor signatures, attribute tree and ``state_dict`` names as
``conan_fgw/src/model/graph_embeddings/torch_geometric_visnet.py`` (``tgv.py`` below) and the ConAN wrapper
``conan_fgw/src/model/graph_embeddings/visnet.py:82-158``:

  ``CosineCutoff`` tgv:13 - ``ExpNormalSmearing`` tgv:49 - ``Sphere`` tgv:114 - ``VecLayerNorm`` tgv:192 -
  ``Distance`` tgv:285 - ``NeighborEmbedding`` tgv:350 - ``EdgeEmbedding`` tgv:426 - ``ViS_MP`` tgv:468 -
  ``ViSNetBlock`` tgv:741 - ``GatedEquivariantBlock`` tgv:889 - ``EquivariantScalar`` tgv:963 - ``Atomref`` tgv:1017 -
  ``TorchGeometricViSNet`` = tgv ``ViSNet`` :1061 - ``ViSNet`` = ConAN's wrapper ``visnet.py:82``.

Provided configuration = what ConAN instantiates (``visnet.py:84-86``): ``lmax=1``, ``vecnorm_type=None``,
``vertex=False``, non-trainable RBFs; other settings raise ``NotImplementedError``.

Every edge-sized operation (geometry, neighbour embedding, edge embedding, attention message, vector message +
aggregation, edge update, all their backward passes) and every Linear / LayerNorm / SiLU runs through the C ABI
(``csrc/visnet.cu``, ``dense.cu``).  Node-level tensor algebra of a few element-wise ops (splits, ``vec1 * vec2``
sums, gating) is plain tensor arithmetic on CUDA tensors.  Exact-fp32 numerics (1e-5 parity with the oracle).
"""

from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import nn
from torch.autograd import Function

from . import _lib, ops
from ._lib import call, ptr
from .graph import NeighborList, build_neighbor_list
from .nn import Embedding, Linear, SumAggregation


def _c(t):
    return ops._f32c(t)


def _graph_of(edge_index) -> NeighborList:
    g = getattr(edge_index, "_cmp_graph", None)
    if g is None:
        raise _lib.ConanMPError(
            "ViSNet modules need the edge_index produced by this package's Distance / radius_graph (it carries the "
            "CSR neighbour list the kernels run on)")
    return g


# ------------------------------------------------------------------------------------------------
# autograd wrappers
# ------------------------------------------------------------------------------------------------

def _segsum(x, rowptr, perm, N):
    x = _c(x)
    out = torch.empty(N, x.shape[1], dtype=torch.float32, device=x.device)
    call("cmp_csr_segment_sum", ptr(x), ptr(rowptr), ptr(perm), N, x.shape[1], ptr(out))
    return out


def _gather(x, idx, E):
    x = _c(x)
    out = torch.empty(E, x.shape[1], dtype=torch.float32, device=x.device)
    call("cmp_gather_rows", ptr(x), ptr(idx), E, x.shape[1], ptr(out))
    return out


class _LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        x = _c(x)
        M, H = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        call("cmp_layernorm_fwd", ptr(x), ptr(_c(w)), ptr(_c(b)), M, H, float(eps), ptr(y), ptr(mean), ptr(rstd))
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dy = _c(dy)
        M, H = x.shape
        dx = torch.empty_like(x)
        dyxhat = torch.empty_like(x)
        call("cmp_layernorm_bwd", ptr(dy), ptr(x), ptr(_c(w)), ptr(mean), ptr(rstd), M, H, ptr(dx), ptr(dyxhat))
        return dx, ops.colsum(dyxhat), ops.colsum(dy), None


class _SegSumFn(Function):
    """x_agg[i] = sum over the edges that end at i (scatter-sum of tgv.py:671)."""

    @staticmethod
    def forward(ctx, x, graph):
        ctx.graph = graph
        return _segsum(x, graph.rowptr, None, graph.N)

    @staticmethod
    def backward(ctx, g):
        graph = ctx.graph
        return _gather(g, graph.erow(), graph.E), None


class _EdgeMessageFn(Function):
    """out[i] = sum_e x[j] * filt[e] * scale[e]   (NeighborEmbedding)."""

    @staticmethod
    def forward(ctx, x, filt, scale, graph):
        x, filt = _c(x), _c(filt)
        N, F = x.shape
        out = torch.empty(N, F, dtype=torch.float32, device=x.device)
        call("cmp_edge_message_fwd", ptr(x), ptr(filt), ptr(scale), ptr(graph.rowptr), ptr(graph.col), N, F, ptr(out))
        ctx.graph = graph
        ctx.save_for_backward(x, filt, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        x, filt, scale = ctx.saved_tensors
        graph = ctx.graph
        g = _c(g)
        N, F = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dfilt = torch.empty_like(filt) if ctx.needs_input_grad[1] else None
        call("cmp_edge_message_bwd", ptr(g), ptr(x), ptr(filt), ptr(scale), ptr(graph.rowptr), ptr(graph.col),
             ptr(graph.rowptr_t), ptr(graph.col_t), ptr(graph.eid_t), N, F, ptr(dfilt), ptr(dx))
        return dx, dfilt, None, None


class _EdgeEmbedFn(Function):
    @staticmethod
    def forward(ctx, x, ep, graph):
        x, ep = _c(x), _c(ep)
        E, H = ep.shape
        f = torch.empty_like(ep)
        call("cmp_vis_edge_embed_fwd", ptr(x), ptr(ep), ptr(graph.col), ptr(graph.erow()), E, H, ptr(f))
        ctx.graph = graph
        ctx.save_for_backward(x, ep)
        return f

    @staticmethod
    def backward(ctx, g):
        x, ep = ctx.saved_tensors
        graph = ctx.graph
        g = _c(g)
        E, H = ep.shape
        gep = torch.empty_like(ep)
        t = torch.empty_like(ep)
        call("cmp_vis_edge_embed_bwd", ptr(g), ptr(x), ptr(ep), ptr(graph.col), ptr(graph.erow()), E, H, ptr(gep), ptr(t))
        dx = _segsum(t, graph.rowptr, None, graph.N) + _segsum(t, graph.rowptr_t, graph.eid_t, graph.N)
        return dx, gep, None


class _MessageFn(Function):
    @staticmethod
    def forward(ctx, q, k, v, dk, dv, C, graph, heads, pre_act=False):
        """``pre_act``: ``dk`` / ``dv`` are the outputs of ``dk_proj`` / ``dv_proj``; their SiLU (tgv.py:622-629) is applied
        inside the kernels and the returned gradients are those of the pre-activations."""
        q, k, v, dk, dv = (_c(t) for t in (q, k, v, dk, dv))
        E, H = dk.shape
        m = torch.empty_like(dk)
        pre = torch.empty(E, heads, dtype=torch.float32, device=dk.device)
        call("cmp_vis_message_fwd", ptr(q), ptr(k), ptr(v), ptr(dk), ptr(dv), ptr(C), ptr(graph.col), ptr(graph.erow()),
             E, H, heads, int(pre_act), ptr(m), ptr(pre))
        ctx.graph, ctx.heads, ctx.pre_act = graph, heads, int(pre_act)
        ctx.save_for_backward(q, k, v, dk, dv, C, pre)
        return m

    @staticmethod
    def backward(ctx, gm):
        q, k, v, dk, dv, C, pre = ctx.saved_tensors
        graph = ctx.graph
        gm = _c(gm)
        E, H = dk.shape
        g_dk, g_dv, geq, gek, gev = (torch.empty_like(dk) for _ in range(5))
        call("cmp_vis_message_bwd", ptr(gm), ptr(q), ptr(k), ptr(v), ptr(dk), ptr(dv), ptr(C), ptr(pre), ptr(graph.col),
             ptr(graph.erow()), E, H, ctx.heads, ctx.pre_act, ptr(g_dk), ptr(g_dv), ptr(geq), ptr(gek), ptr(gev))
        dq = _segsum(geq, graph.rowptr, None, graph.N)                     # q is gathered at the target
        dkn = _segsum(gek, graph.rowptr_t, graph.eid_t, graph.N)          # k, v at the source
        dvn = _segsum(gev, graph.rowptr_t, graph.eid_t, graph.N)
        return dq, dkn, dvn, g_dk, g_dv, None, None, None, None


class _VecAggFn(Function):
    @staticmethod
    def forward(ctx, vec, s12, dhat, graph, pre_act=False):
        """``pre_act``: ``s12`` is the output of ``s_proj`` (tgv.py:649), SiLU applied inside the kernels."""
        vec, s12 = _c(vec), _c(s12)
        N, _, H = vec.shape
        out = torch.empty_like(vec)
        call("cmp_vis_vecagg_fwd", ptr(vec), ptr(s12), ptr(dhat), ptr(graph.rowptr), ptr(graph.col), N, H, int(pre_act),
             ptr(out))
        ctx.graph, ctx.pre_act = graph, int(pre_act)
        ctx.save_for_backward(vec, s12, dhat)
        return out

    @staticmethod
    def backward(ctx, g):
        vec, s12, dhat = ctx.saved_tensors
        graph = ctx.graph
        g = _c(g)
        N, _, H = vec.shape
        g_s12 = torch.empty_like(s12)
        g_vec = torch.empty_like(vec)
        call("cmp_vis_vecagg_bwd", ptr(g), ptr(vec), ptr(s12), ptr(dhat), ptr(graph.rowptr), ptr(graph.col),
             ptr(graph.rowptr_t), ptr(graph.col_t), ptr(graph.eid_t), N, H, ctx.pre_act, ptr(g_s12), ptr(g_vec))
        return g_vec, g_s12, None, None, None


class _EdgeUpdateFn(Function):
    @staticmethod
    def forward(ctx, wt, ws, dhat, fpa, graph, pre_act=False):
        """``pre_act``: ``fpa`` is the output of ``f_proj`` (tgv.py:659), SiLU applied inside the kernels."""
        wt, ws, fpa = _c(wt), _c(ws), _c(fpa)
        E, H = fpa.shape
        df = torch.empty_like(fpa)
        wdot = torch.empty_like(fpa)
        call("cmp_vis_edge_update_fwd", ptr(wt), ptr(ws), ptr(dhat), ptr(fpa), ptr(graph.col), ptr(graph.erow()), E, H,
             int(pre_act), ptr(df), ptr(wdot))
        ctx.graph, ctx.pre_act = graph, int(pre_act)
        ctx.save_for_backward(wt, ws, dhat, fpa, wdot)
        return df

    @staticmethod
    def backward(ctx, g):
        wt, ws, dhat, fpa, wdot = ctx.saved_tensors
        graph = ctx.graph
        g = _c(g)
        N, _, H = wt.shape
        g_fpa, gw = torch.empty_like(fpa), torch.empty_like(fpa)
        call("cmp_vis_edge_update_bwd_prep", ptr(g), ptr(fpa), ptr(wdot), fpa.numel(), ctx.pre_act, ptr(g_fpa), ptr(gw))
        g_wt = torch.empty_like(wt)
        g_ws = torch.empty_like(ws)
        call("cmp_vis_edge_update_bwd", ptr(gw), ptr(wt), ptr(ws), ptr(dhat), ptr(graph.rowptr), ptr(graph.col),
             ptr(graph.rowptr_t), ptr(graph.col_t), ptr(graph.eid_t), N, H, ptr(g_wt), ptr(g_ws))
        return g_wt, g_ws, None, g_fpa, None, None


class _NodeUpdateFn(Function):
    """``dx = (vec1 * vec2).sum(1) * o2 + o3``,  ``dvec = vec3 * o1[:, None] + vec_agg`` (tgv.py:616-627) with
    ``vec1 | vec2 | vec3 = vec_proj(vec)`` and ``o1 | o2 | o3 = o_proj(x_agg)`` in one kernel per direction."""

    @staticmethod
    def forward(ctx, vp, o, vec_agg):
        vp, o, vec_agg = _c(vp), _c(o), _c(vec_agg)
        N, _, H = vec_agg.shape
        dx = torch.empty(N, H, dtype=torch.float32, device=o.device)
        dvec = torch.empty_like(vec_agg)
        call("cmp_vis_node_update_fwd", ptr(vp), ptr(o), ptr(vec_agg), N, H, ptr(dx), ptr(dvec))
        ctx.save_for_backward(vp, o)
        return dx, dvec

    @staticmethod
    def backward(ctx, g_dx, g_dvec):
        vp, o = ctx.saved_tensors
        N, H3 = o.shape
        H = H3 // 3
        g_dx, g_dvec = _c(g_dx), _c(g_dvec)
        g_vp, g_o = torch.empty_like(vp), torch.empty_like(o)
        call("cmp_vis_node_update_bwd", ptr(g_dx), ptr(g_dvec), ptr(vp), ptr(o), N, H, ptr(g_vp), ptr(g_o))
        return g_vp, g_o, g_dvec


def layer_norm(x, module: nn.LayerNorm):
    return _LayerNormFn.apply(x, module.weight, module.bias, module.eps)


def edge_geometry(graph: NeighborList, cutoff: float, means, betas, alpha: float):
    """rbf[E,R], unit vectors dhat[E,3] and masked cosine cutoff C[E] of a loop=True neighbour list."""
    E = graph.E
    dev = graph.rowptr.device
    R = means.numel()
    rbf = torch.empty(E, R, dtype=torch.float32, device=dev)
    dhat = torch.empty(E, 3, dtype=torch.float32, device=dev)
    C = torch.empty(E, dtype=torch.float32, device=dev)
    if graph.evec is None:
        raise _lib.ConanMPError("edge_geometry: the neighbour list was built without edge vectors")
    call("cmp_vis_edge_geometry", ptr(graph.evec), ptr(graph.dist), ptr(graph.col), ptr(graph.erow()), E, float(cutoff),
         float(alpha), ptr(_c(means)), ptr(_c(betas)), R, ptr(rbf), ptr(dhat), ptr(C))
    return rbf, dhat, C


# ------------------------------------------------------------------------------------------------
# modules (tgv.py names)
# ------------------------------------------------------------------------------------------------

class CosineCutoff(nn.Module):
    def __init__(self, cutoff: float) -> None:
        super().__init__()
        self.cutoff = cutoff

    def forward(self, distances):
        c = 0.5 * ((distances * math.pi / self.cutoff).cos() + 1.0)
        return c * (distances < self.cutoff).float()


class ExpNormalSmearing(nn.Module):
    def __init__(self, cutoff: float = 5.0, num_rbf: int = 128, trainable: bool = True) -> None:
        super().__init__()
        self.cutoff, self.num_rbf, self.trainable = cutoff, num_rbf, trainable
        self.cutoff_fn = CosineCutoff(cutoff)
        self.alpha = 5.0 / cutoff
        means, betas = self._initial_params()
        if trainable:
            self.register_parameter("means", nn.Parameter(means))
            self.register_parameter("betas", nn.Parameter(betas))
        else:
            self.register_buffer("means", means)
            self.register_buffer("betas", betas)

    def _initial_params(self):
        start = torch.exp(torch.tensor(-self.cutoff))
        means = torch.linspace(start, 1, self.num_rbf)
        betas = torch.tensor([(2 / self.num_rbf * (1 - start)) ** -2] * self.num_rbf)
        return means, betas

    def reset_parameters(self):
        means, betas = self._initial_params()
        self.means.data.copy_(means)
        self.betas.data.copy_(betas)

    def forward(self, dist):
        graph = getattr(dist, "_cmp_graph", None)
        if graph is not None and graph.evec is not None and not self.trainable:
            key = ("vis_geom", float(self.cutoff), self.num_rbf)
            cache = graph.__dict__.setdefault("_vis_cache", {})
            if key not in cache:
                cache[key] = edge_geometry(graph, self.cutoff, self.means, self.betas, self.alpha)
            return cache[key][0]
        d = dist.unsqueeze(-1)
        return self.cutoff_fn(d) * (-self.betas * ((self.alpha * (-d)).exp() - self.means) ** 2).exp()


class Sphere(nn.Module):
    def __init__(self, lmax: int = 2) -> None:
        super().__init__()
        self.lmax = lmax

    def forward(self, edge_vec):
        if self.lmax != 1:
            raise NotImplementedError("only lmax = 1 (what ConAN instantiates, visnet.py:84-86) is provided")
        return edge_vec    # l = 1 real spherical harmonics are (x, y, z)


class VecLayerNorm(nn.Module):
    def __init__(self, hidden_channels: int, trainable: bool, norm_type: Optional[str] = "max_min") -> None:
        super().__init__()
        if norm_type is not None:
            raise NotImplementedError("vecnorm_type='max_min' is never enabled by ConAN and is not provided")
        self.hidden_channels, self.norm_type, self.eps = hidden_channels, norm_type, 1e-12
        weight = torch.ones(hidden_channels)
        if trainable:
            self.register_parameter("weight", nn.Parameter(weight))
        else:
            self.register_buffer("weight", weight)

    def reset_parameters(self):
        torch.nn.init.ones_(self.weight)

    def forward(self, vec):
        if vec.size(1) != 3:
            raise ValueError(f"'{self.__class__.__name__}' only support 3 channels here (got {vec.size(1)})")
        return vec * self.weight.unsqueeze(0).unsqueeze(0)


class Distance(nn.Module):
    def __init__(self, cutoff: float, max_num_neighbors: int = 32, add_self_loops: bool = True) -> None:
        super().__init__()
        self.cutoff, self.max_num_neighbors, self.add_self_loops = cutoff, max_num_neighbors, add_self_loops

    def neighbor_list(self, pos, batch, num_graphs=None, num_edges=None, status=None) -> NeighborList:
        return build_neighbor_list(pos, batch, self.cutoff, self.max_num_neighbors, loop=self.add_self_loops,
                                   num_graphs=num_graphs, want_evec=True, num_edges=num_edges, status=status)

    def forward(self, pos, batch):
        nl = self.neighbor_list(pos, batch)
        E = nl.E
        edge_vec = nl.evec[:E]
        edge_vec._cmp_graph = nl
        return nl.edge_index(), nl.edge_weight(), edge_vec


class NeighborEmbedding(nn.Module):
    def __init__(self, hidden_channels: int, num_rbf: int, cutoff: float, max_z: int = 100) -> None:
        super().__init__()
        self.embedding = Embedding(max_z, hidden_channels)
        self.distance_proj = Linear(num_rbf, hidden_channels)
        self.combine = Linear(hidden_channels * 2, hidden_channels)
        self.cutoff = CosineCutoff(cutoff)
        self.reset_parameters()

    def reset_parameters(self):
        self.embedding.reset_parameters()
        torch.nn.init.xavier_uniform_(self.distance_proj.weight)
        torch.nn.init.xavier_uniform_(self.combine.weight)
        self.distance_proj.bias.data.zero_()
        self.combine.bias.data.zero_()

    def forward(self, z, x, edge_index, edge_weight, edge_attr):
        graph = _graph_of(edge_index)
        E = graph.E
        # self loops are excluded (tgv.py:408-412): their per-edge scale is zero
        C = self.cutoff(edge_weight) * (graph.col[:E] != graph.erow()[:E]).float()
        W = self.distance_proj(edge_attr)
        x_nb = _EdgeMessageFn.apply(self.embedding(z), W, C.contiguous(), graph)
        return self.combine(torch.cat([x, x_nb], dim=1))


class EdgeEmbedding(nn.Module):
    def __init__(self, num_rbf: int, hidden_channels: int) -> None:
        super().__init__()
        self.edge_proj = Linear(num_rbf, hidden_channels)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.edge_proj.weight)
        self.edge_proj.bias.data.zero_()

    def forward(self, edge_index, edge_attr, x):
        return _EdgeEmbedFn.apply(x, self.edge_proj(edge_attr), _graph_of(edge_index))


class ViS_MP(nn.Module):
    def __init__(self, num_heads: int, hidden_channels: int, cutoff: float, vecnorm_type: Optional[str],
                 trainable_vecnorm: bool, last_layer: bool = False) -> None:
        super().__init__()
        if hidden_channels % num_heads != 0:
            raise ValueError(f"The number of hidden channels (got {hidden_channels}) must be evenly divisible by the "
                             f"number of attention heads (got {num_heads})")
        H = hidden_channels
        self.num_heads, self.hidden_channels, self.head_dim, self.last_layer = num_heads, H, H // num_heads, last_layer
        self.layernorm = nn.LayerNorm(H)
        self.vec_layernorm = VecLayerNorm(H, trainable=trainable_vecnorm, norm_type=vecnorm_type)
        self.act = nn.SiLU()
        self.attn_activation = nn.SiLU()
        self.cutoff = CosineCutoff(cutoff)
        self.vec_proj = Linear(H, H * 3, False)
        self.q_proj = Linear(H, H)
        self.k_proj = Linear(H, H)
        self.v_proj = Linear(H, H)
        self.dk_proj = Linear(H, H)
        self.dv_proj = Linear(H, H)
        self.s_proj = Linear(H, H * 2)
        if not self.last_layer:
            self.f_proj = Linear(H, H)
            self.w_src_proj = Linear(H, H, False)
            self.w_trg_proj = Linear(H, H, False)
        self.o_proj = Linear(H, H * 3)
        self.reset_parameters()

    def reset_parameters(self):
        self.layernorm.reset_parameters()
        self.vec_layernorm.reset_parameters()
        for name in ("q_proj", "k_proj", "v_proj", "o_proj", "s_proj", "dk_proj", "dv_proj"):
            lin = getattr(self, name)
            torch.nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.zero_()
        if not self.last_layer:
            torch.nn.init.xavier_uniform_(self.f_proj.weight)
            self.f_proj.bias.data.zero_()
            torch.nn.init.xavier_uniform_(self.w_src_proj.weight)
            torch.nn.init.xavier_uniform_(self.w_trg_proj.weight)
        torch.nn.init.xavier_uniform_(self.vec_proj.weight)

    def forward(self, x, vec, edge_index, r_ij, f_ij, d_ij, cutoff_values=None):
        """``cutoff_values``: ``CosineCutoff(r_ij)`` when the caller already has it (every layer of a block uses the same
        cutoff on the same distances - ViSNetBlock evaluates it once instead of once per layer, tgv.py:645)."""
        graph = _graph_of(edge_index)
        H = self.hidden_channels
        C = cutoff_values if cutoff_values is not None else self.cutoff(r_ij).contiguous()
        d_ij = _c(d_ij)
        x = layer_norm(x, self.layernorm)
        vec = self.vec_layernorm(vec)
        q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        # the SiLUs on the four edge-sized projections (dk, dv, s12, f_proj: tgv.py:622-629,649,659) are applied INSIDE
        # the kernels that consume them: the activated [E, H] / [E, 2H] tensors never exist and 8 activation launches per
        # layer (forward + backward) disappear
        dk = self.dk_proj(f_ij)
        dv = self.dv_proj(f_ij)
        vp = self.vec_proj(vec)                                                  # [N, 3, 3H] = vec1 | vec2 | vec3

        m = _MessageFn.apply(q, k, v, dk, dv, C, graph, self.num_heads, True)    # [E, H]  v_j * silu(dv) * attn
        s12 = self.s_proj(m)                                                     # [E, 2H] = pre-activations of [s1 | s2]
        x_agg = _SegSumFn.apply(m, graph)
        vec_agg = _VecAggFn.apply(vec, s12, d_ij, graph, True)

        # (vec1 * vec2).sum(1) * o2 + o3  and  vec3 * o1 + vec_agg  (tgv.py:616-627) in one kernel per direction
        dx, dvec = _NodeUpdateFn.apply(vp, self.o_proj(x_agg), vec_agg)
        if self.last_layer:
            return dx, dvec, None
        # w_trg / w_src are bias-free linears: apply them once per atom instead of once per edge (tgv.py:657-658)
        wt, ws = self.w_trg_proj(vec), self.w_src_proj(vec)
        df = _EdgeUpdateFn.apply(wt, ws, d_ij, self.f_proj(f_ij), graph, True)
        return dx, dvec, df


class ViSNetBlock(nn.Module):
    def __init__(self, lmax: int = 1, vecnorm_type: Optional[str] = None, trainable_vecnorm: bool = False,
                 num_heads: int = 8, num_layers: int = 6, hidden_channels: int = 128, num_rbf: int = 32,
                 trainable_rbf: bool = False, max_z: int = 100, cutoff: float = 5.0, max_num_neighbors: int = 32,
                 vertex: bool = False) -> None:
        super().__init__()
        if vertex:
            raise NotImplementedError("ViS_MP_Vertex is never enabled by ConAN and is not provided")
        if lmax != 1:
            raise NotImplementedError("only lmax = 1 is provided")
        self.lmax, self.vecnorm_type, self.trainable_vecnorm = lmax, vecnorm_type, trainable_vecnorm
        self.num_heads, self.num_layers, self.hidden_channels = num_heads, num_layers, hidden_channels
        self.num_rbf, self.trainable_rbf, self.max_z, self.cutoff = num_rbf, trainable_rbf, max_z, cutoff
        self.max_num_neighbors = max_num_neighbors
        self.embedding = Embedding(max_z, hidden_channels)
        self.distance = Distance(cutoff, max_num_neighbors=max_num_neighbors)
        self.sphere = Sphere(lmax=lmax)
        self.distance_expansion = ExpNormalSmearing(cutoff, num_rbf, trainable_rbf)
        self.neighbor_embedding = NeighborEmbedding(hidden_channels, num_rbf, cutoff, max_z)
        self.edge_embedding = EdgeEmbedding(num_rbf, hidden_channels)
        self.vis_mp_layers = nn.ModuleList()
        kw = dict(num_heads=num_heads, hidden_channels=hidden_channels, cutoff=cutoff, vecnorm_type=vecnorm_type,
                  trainable_vecnorm=trainable_vecnorm)
        for _ in range(num_layers - 1):
            self.vis_mp_layers.append(ViS_MP(last_layer=False, **kw))
        self.vis_mp_layers.append(ViS_MP(last_layer=True, **kw))
        self.out_norm = nn.LayerNorm(hidden_channels)
        self.vec_out_norm = VecLayerNorm(hidden_channels, trainable=trainable_vecnorm, norm_type=vecnorm_type)
        # persistent device-side error word of every neighbour list this block builds (not part of the state_dict)
        self.register_buffer("status", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.reset_parameters()

    def reset_parameters(self):
        self.embedding.reset_parameters()
        self.distance_expansion.reset_parameters()
        self.neighbor_embedding.reset_parameters()
        self.edge_embedding.reset_parameters()
        for layer in self.vis_mp_layers:
            layer.reset_parameters()
        self.out_norm.reset_parameters()
        self.vec_out_norm.reset_parameters()

    def forward(self, z, pos, batch, num_graphs=None, num_edges=None):
        nl = self.distance.neighbor_list(pos, batch, num_graphs, num_edges, status=self.status)
        x = self.embedding(z, nl.status)
        edge_index, edge_weight = nl.edge_index(), nl.edge_weight()
        if self.trainable_rbf:
            raise NotImplementedError("trainable RBFs are not instantiated by ConAN and are not provided")
        rbf, dhat, _ = edge_geometry(nl, self.cutoff, self.distance_expansion.means, self.distance_expansion.betas,
                                     self.distance_expansion.alpha)
        edge_vec = self.sphere(dhat)
        x = self.neighbor_embedding(z, x, edge_index, edge_weight, rbf)
        vec = torch.zeros(x.size(0), ((self.lmax + 1) ** 2) - 1, x.size(1), dtype=x.dtype, device=x.device)
        edge_attr = self.edge_embedding(edge_index, rbf, x)
        # every layer applies the same CosineCutoff to the same distances (tgv.py:645): evaluated once per forward
        cut = self.vis_mp_layers[0].cutoff(edge_weight).contiguous()
        for attn in self.vis_mp_layers[:-1]:
            dx, dvec, dedge = attn(x, vec, edge_index, edge_weight, edge_attr, edge_vec, cutoff_values=cut)
            x = x + dx
            vec = vec + dvec
            edge_attr = edge_attr + dedge
        dx, dvec, _ = self.vis_mp_layers[-1](x, vec, edge_index, edge_weight, edge_attr, edge_vec, cutoff_values=cut)
        x = x + dx
        vec = vec + dvec
        x = layer_norm(x, self.out_norm)
        vec = self.vec_out_norm(vec)
        self._last_graph = nl
        return x, vec


class GatedEquivariantBlock(nn.Module):
    def __init__(self, hidden_channels: int, out_channels: int, intermediate_channels: Optional[int] = None,
                 scalar_activation: bool = False) -> None:
        super().__init__()
        self.out_channels = out_channels
        if intermediate_channels is None:
            intermediate_channels = hidden_channels
        self.vec1_proj = Linear(hidden_channels, hidden_channels, bias=False)
        self.vec2_proj = Linear(hidden_channels, out_channels, bias=False)
        self.update_net = nn.Sequential(Linear(hidden_channels * 2, intermediate_channels), nn.SiLU(),
                                        Linear(intermediate_channels, out_channels * 2))
        self.act = nn.SiLU() if scalar_activation else None
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.vec1_proj.weight)
        torch.nn.init.xavier_uniform_(self.vec2_proj.weight)
        torch.nn.init.xavier_uniform_(self.update_net[0].weight)
        self.update_net[0].bias.data.zero_()
        torch.nn.init.xavier_uniform_(self.update_net[2].weight)
        self.update_net[2].bias.data.zero_()

    def forward(self, x, v):
        vec1 = torch.norm(self.vec1_proj(v), dim=-2)
        vec2 = self.vec2_proj(v)
        h = self.update_net[2](ops.silu(self.update_net[0](torch.cat([x, vec1], dim=-1))))
        x, gate = torch.split(h, self.out_channels, dim=-1)
        v = gate.unsqueeze(1) * vec2
        if self.act is not None:
            x = ops.silu(x.contiguous())
        return x, v


class EquivariantScalar(nn.Module):
    def __init__(self, hidden_channels: int, output_channels: int) -> None:
        super().__init__()
        self.output_network = nn.ModuleList([
            GatedEquivariantBlock(hidden_channels, hidden_channels // 2, scalar_activation=True),
            GatedEquivariantBlock(hidden_channels // 2, output_channels, scalar_activation=False),
        ])
        self.reset_parameters()

    def reset_parameters(self):
        for layer in self.output_network:
            layer.reset_parameters()

    def pre_reduce(self, x, v):
        for layer in self.output_network:
            x, v = layer(x, v)
        return x + v.sum() * 0


class Atomref(nn.Module):
    def __init__(self, atomref=None, max_z: int = 100) -> None:
        super().__init__()
        if atomref is None:
            atomref = torch.zeros(max_z, 1)
        else:
            atomref = torch.as_tensor(atomref)
        if atomref.ndim == 1:
            atomref = atomref.view(-1, 1)
        self.register_buffer("initial_atomref", atomref)
        self.atomref = nn.Embedding(len(atomref), 1)
        self.reset_parameters()

    def reset_parameters(self):
        self.atomref.weight.data.copy_(self.initial_atomref)

    def forward(self, x, z):
        # parameters stay a torch.nn.Embedding (state_dict key ``atomref.weight``); lookup and the deterministic weight
        # gradient run on the library's embedding kernels (torch's embedding backward alone took 139 us per step)
        return x + ops.embedding(z, self.atomref.weight)


class TorchGeometricViSNet(nn.Module):
    """The vendored ``ViSNet`` of tgv.py:1061-1229 (constructor order unchanged)."""

    def __init__(self, lmax: int = 1, vecnorm_type: Optional[str] = None, trainable_vecnorm: bool = False,
                 num_heads: int = 8, num_layers: int = 6, hidden_channels: int = 128, num_rbf: int = 32,
                 trainable_rbf: bool = False, max_z: int = 100, cutoff: float = 5.0, max_num_neighbors: int = 32,
                 vertex: bool = False, atomref=None, reduce_op: str = "sum", mean: float = 0.0, std: float = 1.0,
                 derivative: bool = False) -> None:
        super().__init__()
        if derivative:
            raise NotImplementedError("derivative=True (forces) is never enabled by ConAN and is not provided")
        self.representation_model = ViSNetBlock(lmax=lmax, vecnorm_type=vecnorm_type, trainable_vecnorm=trainable_vecnorm,
                                                num_heads=num_heads, num_layers=num_layers,
                                                hidden_channels=hidden_channels, num_rbf=num_rbf,
                                                trainable_rbf=trainable_rbf, max_z=max_z, cutoff=cutoff,
                                                max_num_neighbors=max_num_neighbors, vertex=vertex)
        self.output_model = EquivariantScalar(hidden_channels=hidden_channels, output_channels=hidden_channels // 2)
        self.prior_model = Atomref(atomref=atomref, max_z=max_z)
        self.output_model_bary = EquivariantScalar(hidden_channels=hidden_channels, output_channels=hidden_channels // 2)
        self.prior_model_bary = Atomref(atomref=atomref, max_z=max_z)
        self.reduce_op = reduce_op
        self.derivative = derivative
        self.register_buffer("mean", torch.tensor(mean))
        self.register_buffer("std", torch.tensor(std))
        self.reset_parameters()

    def reset_parameters(self):
        self.representation_model.reset_parameters()
        self.output_model.reset_parameters()
        if self.prior_model is not None:
            self.prior_model.reset_parameters()

    def set_precision(self, precision: str):
        """"fp32": every Linear on the exact SIMT GEMM (1e-5 parity).  "bf16": Linears on the tcgen05 kernels with
        split-bf16 operands (hi + lo, three passes: ~2e-5 per GEMM); everything else stays exact fp32."""
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        for m in self.modules():
            if isinstance(m, Linear):
                m.tc = precision == "bf16"
        return self

    def check_status(self):
        """Raise for device-detected input errors of the forward passes since the last call (unsorted batch, atomic
        number out of range, capacity overflow, an edge count different from a promised ``num_edges``)."""
        from .graph import raise_for_status

        raise_for_status(self.representation_model.status, reset=True)

    def _per_atom(self, z, pos, batch, bary=False, num_graphs=None, num_edges=None):
        x, v = self.representation_model(z, pos, batch, num_graphs, num_edges)
        out = self.output_model.pre_reduce(x, v) * self.std
        if self.prior_model is not None:
            out = self.prior_model(out, z)
        if not bary:
            return out
        out_b = self.output_model_bary.pre_reduce(x, v) * self.std
        if self.prior_model_bary is not None:
            out_b = self.prior_model_bary(out_b, z)
        return out, out_b

    def forward(self, z, pos, batch, num_graphs=None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        x = self._per_atom(z, pos, batch, num_graphs=num_graphs)
        if self.reduce_op != "sum":
            raise NotImplementedError("only reduce_op='sum' is provided")
        y = SumAggregation()(x, batch, seg_ptr=self.representation_model._last_graph.seg_ptr)
        return y + self.mean, None


class ViSNet(TorchGeometricViSNet):
    """ConAN's wrapper (``conan_fgw/src/model/graph_embeddings/visnet.py:82-158``)."""

    def __init__(self, device=None, hidden_channels: int = 128, cutoff: float = 5.0):
        super().__init__(hidden_channels=hidden_channels)
        self.device = device
        self.hidden_channels = hidden_channels
        self.cutoff = cutoff
        self.readout = SumAggregation()
        self.barycenter_fn = None

    def forward(self, z, pos, batch, num_graphs=None, num_edges=None):
        x = self._per_atom(z, pos, batch, num_graphs=num_graphs, num_edges=num_edges)
        return self.readout(x, batch, dim=0, seg_ptr=self.representation_model._last_graph.seg_ptr)

    def forward_3d_bary(self, z, pos, batch, num_graphs=None):
        return self._per_atom(z, pos, batch, bary=True, num_graphs=num_graphs)

    def forward_w_barycenter(self, z, pos, num_conformers: int, batch=None, data_batch=None, max_iter: int = 100,
                             epsilon: float = 0.1, num_graphs=None):
        batch = torch.zeros_like(z) if batch is None else batch
        h_3d, h_bary = self.forward_3d_bary(z, pos, batch, num_graphs=num_graphs)
        if self.barycenter_fn is None:
            raise RuntimeError("forward_w_barycenter: the FGW barycenter stays on the reference path; set "
                               "`model.barycenter_fn` (see INTEGRATION.md)")
        # the reference builds a second radius graph here with self.cutoff and no self loops (visnet.py:276)
        nl = build_neighbor_list(pos, batch, self.cutoff, 32, loop=False, num_graphs=num_graphs)
        batch_size = int(nl.G / num_conformers)
        _, h_bary = self.barycenter_fn(node_feature=h_bary, edge_index=nl.edge_index(), batch=batch,
                                       batch_size=batch_size, num_conformers=num_conformers)
        h_3d = self.readout(h_3d, batch, dim=0, seg_ptr=nl.seg_ptr)
        return h_3d, h_bary

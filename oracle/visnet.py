"""Oracle (test infrastructure): ViSNet as vendored by the reference.

Restates ``conan_fgw/src/model/graph_embeddings/torch_geometric_visnet.py``
(vendored PyG ViSNet) for the configuration ConAN instantiates
(``visnet.py:83-86``: ``lmax=1``, ``vecnorm_type=None``, ``vertex=False``,
8 heads, 6 layers, 32 RBFs, cutoff 5.0) and the ConAN wrapper
``visnet.py:93-158``.  ``MessagePassing.propagate`` / ``edge_updater`` are
replaced by explicit gathers and ``index_add_``:
``*_j = t[edge_index[0]]``, ``*_i = t[edge_index[1]]``, sum at ``edge_index[1]``.

Formula sources (file = torch_geometric_visnet.py):
  cosine cutoff :44-46 - exp-normal RBF :82-111 - Distance :331-347 -
  NeighborEmbedding :408-423 - EdgeEmbedding :463-465 - ViS_MP :605-661 -
  ViSNetBlock.forward :861-886 - GatedEquivariantBlock :950-960 -
  EquivariantScalar.pre_reduce :1011-1014 - Atomref :1058.

The parameter tree carries exactly the reference's ``state_dict`` names, so the
reference's weights load with ``strict=True``.  **Pinned**: checked against the
reference file itself (run through ``oracle/pyg_shim.py`` in the build
container) by ``tests/golden/make_golden.py`` -> ``tests/golden/visnet_*.pt``.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .radius import radius_graph_ref


def cosine_cutoff(d, cutoff):
    c = 0.5 * (torch.cos(d * math.pi / cutoff) + 1.0)
    return c * (d < cutoff).to(d.dtype)


class _ExpNormalRBF(nn.Module):
    def __init__(self, cutoff, num_rbf):
        super().__init__()
        self.cutoff, self.num_rbf, self.alpha = cutoff, num_rbf, 5.0 / cutoff
        lo = torch.exp(torch.tensor(-float(cutoff)))
        self.register_buffer("means", torch.linspace(lo, 1, num_rbf))
        self.register_buffer("betas", torch.tensor([(2 / num_rbf * (1 - lo)) ** -2] * num_rbf))

    def forward(self, d):
        d = d.unsqueeze(-1)
        return cosine_cutoff(d, self.cutoff) * torch.exp(
            -self.betas * (torch.exp(self.alpha * (-d)) - self.means) ** 2)


class _NeighborEmbedding(nn.Module):
    def __init__(self, H, num_rbf, cutoff, max_z):
        super().__init__()
        self.embedding = nn.Embedding(max_z, H)
        self.distance_proj = nn.Linear(num_rbf, H)
        self.combine = nn.Linear(2 * H, H)
        self.cutoff = cutoff
        nn.init.xavier_uniform_(self.distance_proj.weight)
        nn.init.xavier_uniform_(self.combine.weight)
        self.distance_proj.bias.data.zero_()
        self.combine.bias.data.zero_()


class _EdgeEmbedding(nn.Module):
    def __init__(self, num_rbf, H):
        super().__init__()
        self.edge_proj = nn.Linear(num_rbf, H)
        nn.init.xavier_uniform_(self.edge_proj.weight)
        self.edge_proj.bias.data.zero_()


class _VecNorm(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.register_buffer("weight", torch.ones(H))      # norm_type None, not trainable


class _VisMP(nn.Module):
    def __init__(self, heads, H, cutoff, last_layer):
        super().__init__()
        self.heads, self.H, self.cutoff, self.last_layer = heads, H, cutoff, last_layer
        self.layernorm = nn.LayerNorm(H)
        self.vec_layernorm = _VecNorm(H)
        self.vec_proj = nn.Linear(H, 3 * H, bias=False)
        self.q_proj = nn.Linear(H, H)
        self.k_proj = nn.Linear(H, H)
        self.v_proj = nn.Linear(H, H)
        self.dk_proj = nn.Linear(H, H)
        self.dv_proj = nn.Linear(H, H)
        self.s_proj = nn.Linear(H, 2 * H)
        if not last_layer:
            self.f_proj = nn.Linear(H, H)
            self.w_src_proj = nn.Linear(H, H, bias=False)
            self.w_trg_proj = nn.Linear(H, H, bias=False)
        self.o_proj = nn.Linear(H, 3 * H)
        for name, m in self.named_children():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    m.bias.data.zero_()

    @staticmethod
    def _reject(v, d):
        # component of v[E,3,H] orthogonal to the unit vector d[E,3]
        return v - (v * d.unsqueeze(2)).sum(dim=1, keepdim=True) * d.unsqueeze(2)

    def forward(self, x, vec, src, dst, r, f, d):
        H, nh = self.H, self.heads
        hd = H // nh
        xn = self.layernorm(x)
        vn = vec * self.vec_layernorm.weight.view(1, 1, -1)
        q = self.q_proj(xn).view(-1, nh, hd)
        k = self.k_proj(xn).view(-1, nh, hd)
        v = self.v_proj(xn).view(-1, nh, hd)
        dk = F.silu(self.dk_proj(f)).view(-1, nh, hd)
        dv = F.silu(self.dv_proj(f)).view(-1, nh, hd)
        v1, v2, v3 = torch.split(self.vec_proj(vn), H, dim=-1)
        vec_dot = (v1 * v2).sum(dim=1)

        attn = (q[dst] * k[src] * dk).sum(dim=-1)
        attn = F.silu(attn) * cosine_cutoff(r, self.cutoff).unsqueeze(1)
        m = (v[src] * dv * attn.unsqueeze(2)).reshape(-1, H)
        s1, s2 = torch.split(F.silu(self.s_proj(m)), H, dim=1)
        mvec = vn[src] * s1.unsqueeze(1) + s2.unsqueeze(1) * d.unsqueeze(2)
        xa = torch.zeros_like(x).index_add_(0, dst, m)
        va = torch.zeros_like(vec).index_add_(0, dst, mvec)

        o1, o2, o3 = torch.split(self.o_proj(xa), H, dim=1)
        dx = vec_dot * o2 + o3
        dvec = v3 * o1.unsqueeze(1) + va
        df = None
        if not self.last_layer:
            w1 = self._reject(self.w_trg_proj(vn[dst]), d)
            w2 = self._reject(self.w_src_proj(vn[src]), -d)
            df = F.silu(self.f_proj(f)) * (w1 * w2).sum(dim=1)
        return dx, dvec, df


class _Block(nn.Module):
    def __init__(self, heads, layers, H, num_rbf, max_z, cutoff, max_num_neighbors):
        super().__init__()
        self.H, self.cutoff, self.max_num_neighbors = H, cutoff, max_num_neighbors
        self.embedding = nn.Embedding(max_z, H)
        self.distance_expansion = _ExpNormalRBF(cutoff, num_rbf)
        self.neighbor_embedding = _NeighborEmbedding(H, num_rbf, cutoff, max_z)
        self.edge_embedding = _EdgeEmbedding(num_rbf, H)
        self.vis_mp_layers = nn.ModuleList(
            _VisMP(heads, H, cutoff, last_layer=(l == layers - 1)) for l in range(layers))
        self.out_norm = nn.LayerNorm(H)
        self.vec_out_norm = _VecNorm(H)

    def forward(self, z, pos, batch):
        x = self.embedding(z)
        ei = radius_graph_ref(pos, self.cutoff, batch, loop=True,
                              max_num_neighbors=self.max_num_neighbors).to(pos.device)
        src, dst = ei[0], ei[1]
        evec = pos[src] - pos[dst]
        nonloop = src != dst
        r = torch.zeros(evec.size(0), dtype=pos.dtype, device=pos.device)
        r[nonloop] = evec[nonloop].norm(dim=-1)
        rbf = self.distance_expansion(r)
        d = evec.clone()
        d[nonloop] = evec[nonloop] / evec[nonloop].norm(dim=1, keepdim=True)

        # neighbour embedding (self loops excluded)
        ne = self.neighbor_embedding
        W = ne.distance_proj(rbf[nonloop]) * cosine_cutoff(r[nonloop], ne.cutoff).view(-1, 1)
        nb = torch.zeros_like(x).index_add_(0, dst[nonloop], ne.embedding(z)[src[nonloop]] * W)
        x = ne.combine(torch.cat([x, nb], dim=1))

        vec = x.new_zeros(x.size(0), 3, x.size(1))
        f = (x[dst] + x[src]) * self.edge_embedding.edge_proj(rbf)
        for layer in self.vis_mp_layers:
            dx, dvec, df = layer(x, vec, src, dst, r, f, d)
            x = x + dx
            vec = vec + dvec
            if df is not None:
                f = f + df
        x = self.out_norm(x)
        vec = vec * self.vec_out_norm.weight.view(1, 1, -1)
        return x, vec


class _GatedBlock(nn.Module):
    def __init__(self, H, out, act):
        super().__init__()
        self.out, self.use_act = out, act
        self.vec1_proj = nn.Linear(H, H, bias=False)
        self.vec2_proj = nn.Linear(H, out, bias=False)
        self.update_net = nn.Sequential(nn.Linear(2 * H, H), nn.SiLU(), nn.Linear(H, 2 * out))
        nn.init.xavier_uniform_(self.vec1_proj.weight)
        nn.init.xavier_uniform_(self.vec2_proj.weight)
        nn.init.xavier_uniform_(self.update_net[0].weight)
        self.update_net[0].bias.data.zero_()
        nn.init.xavier_uniform_(self.update_net[2].weight)
        self.update_net[2].bias.data.zero_()

    def forward(self, x, v):
        n1 = torch.norm(self.vec1_proj(v), dim=-2)
        v2 = self.vec2_proj(v)
        xs, gate = torch.split(self.update_net(torch.cat([x, n1], dim=-1)), self.out, dim=-1)
        v = gate.unsqueeze(1) * v2
        if self.use_act:
            xs = F.silu(xs)
        return xs, v


class _EquivariantScalar(nn.Module):
    def __init__(self, H, out):
        super().__init__()
        self.output_network = nn.ModuleList([_GatedBlock(H, H // 2, True), _GatedBlock(H // 2, out, False)])

    def pre_reduce(self, x, v):
        for blk in self.output_network:
            x, v = blk(x, v)
        return x + v.sum() * 0


class _Atomref(nn.Module):
    def __init__(self, max_z):
        super().__init__()
        self.register_buffer("initial_atomref", torch.zeros(max_z, 1))
        self.atomref = nn.Embedding(max_z, 1)
        self.atomref.weight.data.copy_(self.initial_atomref)

    def forward(self, x, z):
        return x + self.atomref(z)


class ViSNet(nn.Module):
    """ConAN ``ViSNet(device, hidden_channels, cutoff)`` (``visnet.py:82-158``) on the vendored trunk."""

    def __init__(self, device=None, hidden_channels=128, cutoff=5.0, num_heads=8, num_layers=6,
                 num_rbf=32, max_z=100, max_num_neighbors=32):
        super().__init__()
        H = hidden_channels
        self.device, self.hidden_channels = device, H
        # NB visnet.py:84-86 forwards only hidden_channels: the trunk keeps the vendored default cutoff 5.0
        self.representation_model = _Block(num_heads, num_layers, H, num_rbf, max_z, 5.0, max_num_neighbors)
        self.output_model = _EquivariantScalar(H, H // 2)
        self.prior_model = _Atomref(max_z)
        self.output_model_bary = _EquivariantScalar(H, H // 2)
        self.prior_model_bary = _Atomref(max_z)
        self.register_buffer("mean", torch.tensor(0.0))
        self.register_buffer("std", torch.tensor(1.0))
        self.cutoff = cutoff

    def per_atom(self, z, pos, batch, bary=False):
        x, v = self.representation_model(z, pos, batch)
        out = self.prior_model(self.output_model.pre_reduce(x, v) * self.std, z)
        if not bary:
            return out
        out_b = self.prior_model_bary(self.output_model_bary.pre_reduce(x, v) * self.std, z)
        return out, out_b

    def forward(self, z, pos, batch):
        x = self.per_atom(z, pos, batch)
        G = int(batch.max()) + 1
        return x.new_zeros(G, x.size(1)).index_add_(0, batch, x)

    def forward_3d_bary(self, z, pos, batch):
        return self.per_atom(z, pos, batch, bary=True)

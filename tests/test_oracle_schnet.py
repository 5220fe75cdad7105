"""Known-answer and invariance checks that pin the SchNet restatement (CPU).

The reference holds no tests for this path and PyG is absent ("parity unpinned"), so the oracle
is anchored on closed forms of the published algorithm."""
import math

import pytest
import torch

import conan_fgw_b200 as cmp
from oracle import schnet as osn
from conftest import load_golden, rel_err

syn = cmp.synthetic


def small(seed=0, **kw):
    torch.manual_seed(seed)
    cfg = dict(hidden_channels=16, num_filters=16, num_interactions=2, num_gaussians=10, cutoff=5.0)
    cfg.update(kw)
    m = osn.SchNetNoSum(None, **cfg)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return m


def test_gaussian_smearing_known_answers():
    g = osn.GaussianSmearing(0.0, 10.0, 50)
    assert g.offset.shape == (50,) and abs(g.offset[1].item() - 10.0 / 49) < 1e-6
    assert abs(g.coeff - (-0.5 / (10.0 / 49) ** 2)) < 1e-3   # -12.005
    out = g(g.offset.clone())
    assert torch.allclose(out.diagonal(), torch.ones(50))     # rbf at d = offset_k is exactly 1
    d = torch.tensor([1.234])
    want = torch.exp(g.coeff * (d - g.offset) ** 2)
    assert torch.equal(g(d)[0], want)


def test_shifted_softplus_known_answers():
    a = osn.ShiftedSoftplus()
    assert a(torch.zeros(1)).abs().item() < 1e-7                         # ssp(0) = 0
    assert abs(a(torch.tensor([30.0])).item() - (30.0 - math.log(2))) < 1e-5  # identity branch above 20
    x = torch.linspace(-5, 5, 11)
    assert torch.allclose(a(x), torch.log1p(torch.exp(x)) - math.log(2.0), atol=1e-6)


def test_cfconv_isolated_pair_closed_form():
    torch.manual_seed(1)
    F, Ng, cutoff = 8, 6, 4.0
    mlp = torch.nn.Sequential(torch.nn.Linear(Ng, F), osn.ShiftedSoftplus(), torch.nn.Linear(F, F))
    conv = osn.CFConv(5, 7, F, mlp, cutoff)
    x = torch.randn(2, 5)
    d = torch.tensor([1.7, 1.7])
    ei = torch.tensor([[0, 1], [1, 0]])
    rbf = osn.GaussianSmearing(0.0, cutoff, Ng)(d)
    out = conv(x, ei, d, rbf)
    C = 0.5 * (math.cos(1.7 * math.pi / cutoff) + 1.0)
    W = mlp(rbf[0]) * C
    want1 = conv.lin2(conv.lin1(x[0]) * W)      # atom 1 receives from atom 0
    want0 = conv.lin2(conv.lin1(x[1]) * W)
    assert torch.allclose(out[1], want1, atol=1e-6) and torch.allclose(out[0], want0, atol=1e-6)
    # cosine cutoff endpoints: C(0) = 1, C(cutoff) = 0 (no d < cutoff mask in SchNet)
    assert abs(0.5 * (math.cos(0.0) + 1) - 1) < 1e-12 and abs(0.5 * (math.cos(math.pi) + 1)) < 1e-12


def test_state_dict_contract():
    m = osn.SchNetNoSum(None)
    keys = set(m.state_dict().keys())
    blk = {k for k in keys if k.startswith("interactions.0.")}
    assert len(blk) == 13                       # SURVEY.md A.2 [V]: mlp.* and conv.nn.* both present
    for k in ("interactions.0.conv.nn.0.weight", "interactions.0.mlp.0.weight", "interactions.0.conv.lin1.weight",
              "interactions.0.conv.lin2.bias", "interactions.0.lin.bias", "embedding.weight",
              "distance_expansion.offset", "lin1.weight", "lin2.bias", "lin1_bary.weight", "lin2_bary.bias"):
        assert k in keys
    assert "interactions.0.conv.lin1.bias" not in keys
    assert m.interactions[0].conv.nn is m.interactions[0].mlp
    per_block = sum(p.numel() for p in m.interactions[0].parameters())
    assert per_block == 72448
    assert m.lin2.weight.shape == (64, 64)      # ConAN override (sns.py:130)
    assert m.embedding.weight[0].abs().sum() == 0  # padding_idx = 0
    # the CUDA modules expose the identical contract
    c = cmp.SchNetNoSum(None)
    assert set(c.state_dict().keys()) == keys
    assert {k: tuple(v.shape) for k, v in c.state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}
    c.load_state_dict(m.state_dict(), strict=True)


def test_e3_invariance_and_permutation():
    m = small(2)
    b = syn.make_batch(2, 2, 9, seed=5)
    out = m(b.z, b.pos, b.batch)
    # random rotation + translation
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))
    pos2 = (b.pos.double() @ q).float() + torch.tensor([1.5, -2.0, 0.3])
    assert rel_err(m(b.z, pos2, b.batch), out) < 5e-6
    # permute atoms inside every conformer (same permutation keeps batch sorted)
    n = b.atoms_per_conformer
    perm = torch.randperm(n)
    idx = (torch.arange(b.num_graphs)[:, None] * n + perm[None, :]).reshape(-1)
    assert rel_err(m(b.z[idx], b.pos[idx], b.batch), out) < 5e-6


def test_fp32_vs_fp64_twin_and_heads():
    m = small(3)
    b = syn.make_batch(2, 2, 10, seed=6)
    out = m(b.z, b.pos, b.batch)
    out64 = osn.to_double(m)(b.z, b.pos.double(), b.batch)
    assert rel_err(out, out64) < 1e-5
    h, hb = m.forward_3d_bary(b.z, b.pos, b.batch)
    assert h.shape == (b.z.numel(), 8) and hb.shape == h.shape
    assert rel_err(osn.segment_sum(h, b.batch), out) < 1e-6    # forward == readout(first head)
    m.use_readout = False
    assert torch.equal(m(b.z, b.pos, b.batch), h)


def test_conan_head_order_differs_from_pyg():
    m = small(4)
    b = syn.make_batch(1, 1, 6, seed=7)
    h = m.trunk(b.z, b.pos, b.batch)
    conan = m.act(m.lin2(m.lin1(h)))
    pyg_like = m.lin2(m.act(m.lin1(h)))
    assert not torch.allclose(conan, pyg_like)
    assert torch.allclose(m.forward_3d_bary(b.z, b.pos, b.batch)[0], conan)


def test_self_golden_regression():
    g = load_golden("schnet_oracle.pt")
    m = osn.SchNetNoSum(None, **g["config"])
    m.load_state_dict(g["state_dict"], strict=True)
    out = m(g["z"], g["pos"], g["batch"])
    assert rel_err(out, g["out"]) < 1e-6
    out.pow(2).mean().backward()
    for k, p in m.named_parameters():
        if k in g["grads"]:
            assert rel_err(p.grad, g["grads"][k]) < 1e-5, k


def _bond_graph(batch, seed=0):
    """A synthetic covalent graph: a chain inside every conformer, both directions, 3 bond attributes per edge
    (the layout of ``data_batch.edge_index / edge_attr`` consumed at schnet_no_sum.py:166-175)."""
    import types

    g = torch.Generator().manual_seed(seed)
    src, dst = [], []
    b = batch.tolist()
    for a in range(len(b) - 1):
        if b[a] == b[a + 1]:
            src += [a, a + 1]
            dst += [a + 1, a]
    ei = torch.tensor([src, dst], dtype=torch.int64)
    ea = torch.rand(ei.shape[1], 3, generator=g)
    return types.SimpleNamespace(edge_index=ei, edge_attr=ea)


def test_covalent_trunk_and_multiple_returns_contract():
    """use_covalent=True doubles the head input (schnet_no_sum.py:131-142) and adds one InteractionBlock stack on the
    3 bond attributes; SchNetWithMultipleReturns returns (ssp(lin1(h)), edge_index, rbf) (schnet_no_sum.py:405-450)."""
    from conan_fgw_b200 import synthetic as syn

    torch.manual_seed(0)
    b = syn.make_batch(2, 2, 9, seed=3)
    db = _bond_graph(b.batch)
    m = osn.SchNetNoSum(None, hidden_channels=32, num_filters=32, num_interactions=2, num_gaussians=10, cutoff=5.0,
                        use_covalent=True)
    sd = m.state_dict()
    assert sd["lin1.weight"].shape == (16, 64) and sd["lin1_bary.weight"].shape == (16, 64)
    assert sd["interactions_cov.1.mlp.0.weight"].shape == (32, 3)
    out = m(b.z, b.pos, b.batch, data_batch=db)
    assert out.shape == (4, 16) and torch.isfinite(out).all()
    h, hb = m.forward_3d_bary(b.z, b.pos, b.batch, data_batch=db)
    assert h.shape == hb.shape == (b.z.numel(), 16)
    # the covalent stack matters: zeroing its filter MLP output changes the result
    out.sum().backward()
    assert m.interactions_cov[0].mlp[0].weight.grad.abs().sum() > 0

    mr = osn.SchNetWithMultipleReturns(hidden_channels=32, num_filters=32, num_interactions=2, num_gaussians=10, cutoff=5.0)
    h, ei, ea = mr(b.z, b.pos, b.batch)
    assert h.shape == (b.z.numel(), 16) and ei.shape[0] == 2 and ea.shape == (ei.shape[1], 10)
    from oracle.radius import radius_graph_ref

    assert torch.equal(ei, radius_graph_ref(b.pos, 5.0, b.batch))

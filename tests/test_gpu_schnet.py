"""SchNet / SchNetNoSum on the CUDA kernels against the CPU oracle (GPU, through the C ABI).

Tolerance: 1e-5 relative (max|a-b| / max|b|), the fp32 bar of BASELINE.json's north_star."""
import pytest
import torch

import conan_fgw_b200 as cmp
from oracle import schnet as osn
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
syn = cmp.synthetic
DEV = "cuda"
TOL = 1e-5


def pair(seed=0, **cfg):
    torch.manual_seed(seed)
    o = osn.SchNetNoSum(None, **cfg)
    with torch.no_grad():
        for p in o.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    c = cmp.SchNetNoSum(None, **cfg).to(DEV)
    c.load_state_dict(o.state_dict(), strict=True)
    return o, c


def compare(o, c, b, check_grads=True, tol=TOL):
    out_o = o(b.z, b.pos, b.batch)
    out_c = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    assert out_c.shape == out_o.shape
    assert rel_err(out_c, out_o) < tol
    if not check_grads:
        return
    o.zero_grad(), c.zero_grad()
    out_o.pow(2).mean().backward()
    out_c.pow(2).mean().backward()
    po, pc = dict(o.named_parameters()), dict(c.named_parameters())
    for k in po:
        if po[k].grad is None:
            assert pc[k].grad is None, k
            continue
        assert pc[k].grad is not None, k
        assert rel_err(pc[k].grad, po[k].grad) < tol, k


def test_small_model_forward_and_gradients():
    o, c = pair(1, hidden_channels=32, num_filters=48, num_interactions=2, num_gaussians=20, cutoff=6.0)
    compare(o, c, syn.make_batch(3, 2, 11, seed=2))


def test_baseline_shape_cfg1_forward_and_gradients():
    # BASELINE.json configs[0]: 32 molecules x 5 conformers x 26 atoms, H=F=128, T=6, 50 Gaussians, cutoff 10
    o, c = pair(2)
    compare(o, c, syn.make_config_batch("cfg1_esol_fwd", scale=0.25))


def test_baseline_shape_cfg1_full_size_forward():
    # the whole cfg 1 batch (160 conformers, 4 160 atoms, 104 K edges) forward against the CPU oracle
    o, c = pair(2)
    b = syn.make_config_batch("cfg1_esol_fwd")
    with torch.no_grad():
        want = o(b.z, b.pos, b.batch)
        got = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    assert got.shape == want.shape and rel_err(got, want) < 1e-5


def test_conan_regression_shape_and_truncated_graph():
    # ConAN's own instantiation (common.py:524-529): T=3; 65-atom conformers truncate at 32/33 neighbours
    o, c = pair(3, num_interactions=3)
    compare(o, c, syn.make_batch(1, 2, 65, seed=4))


def test_classification_shape():
    # common.py:513-522: H=512, F=256, Ng=10, T=3
    o, c = pair(4, hidden_channels=512, num_filters=256, num_gaussians=10, num_interactions=3)
    compare(o, c, syn.make_batch(2, 2, 20, seed=5))


def test_golden_fixture():
    g = load_golden("schnet_oracle.pt")
    c = cmp.SchNetNoSum(None, **g["config"]).to(DEV)
    c.load_state_dict(g["state_dict"], strict=True)
    z, pos, batch = g["z"].to(DEV), g["pos"].to(DEV), g["batch"].to(DEV)
    out = c(z, pos, batch)
    assert rel_err(out, g["out"]) < TOL
    h, hb = c.forward_3d_bary(z, pos, batch)
    assert rel_err(h, g["h"]) < TOL and rel_err(hb, g["h_bary"]) < TOL
    out.pow(2).mean().backward()
    for k, p in c.named_parameters():
        if k in g["grads"]:
            assert rel_err(p.grad, g["grads"][k]) < TOL, k


def test_deterministic_bitwise():
    _, c = pair(5, num_interactions=2)
    b = syn.make_batch(4, 3, 26, seed=6).to(DEV)
    outs, grads = [], []
    for _ in range(2):
        c.zero_grad()
        out = c(b.z, b.pos, b.batch)
        out.pow(2).mean().backward()
        outs.append(out.detach().clone())
        grads.append(torch.cat([p.grad.reshape(-1) for p in c.parameters() if p.grad is not None]))
    assert torch.equal(outs[0], outs[1]) and torch.equal(grads[0], grads[1])


def test_module_level_api_as_the_reference_calls_it():
    """schnet_no_sum.py:159-164 drives the PyG modules one by one; the same call sequence must work."""
    o, c = pair(6, hidden_channels=32, num_filters=32, num_interactions=2, num_gaussians=16, cutoff=5.0)
    b = syn.make_batch(2, 2, 10, seed=7)
    z, pos, batch = b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV)
    h = c.embedding(z)
    edge_index, edge_weight = c.interaction_graph(pos, batch)
    edge_attr = c.distance_expansion(edge_weight)
    for interaction in c.interactions:
        h = h + interaction(h, edge_index, edge_weight, edge_attr)
    h = c.act(c.lin2(c.lin1(h)))
    out = c.readout(h, batch, dim=0)
    assert rel_err(out, o(b.z, b.pos, b.batch)) < TOL
    # GaussianSmearing / ShiftedSoftplus stand-alone
    assert rel_err(edge_attr, o.distance_expansion(edge_weight.cpu())) < 1e-6
    x = torch.linspace(-25, 25, 101)
    assert rel_err(cmp.ShiftedSoftplus()(x.to(DEV)), osn.ShiftedSoftplus()(x)) < 1e-6


def test_generic_edge_index_any_order_and_custom_edge_attr():
    """InteractionBlock must accept arbitrary edge_attr / unsorted edge_index (covalent branch, sns.py:166-174)."""
    torch.manual_seed(8)
    ob = osn.InteractionBlock(24, 3, 40, 10.0)
    cb = cmp.InteractionBlock(24, 3, 40, 10.0).to(DEV)
    cb.load_state_dict(ob.state_dict(), strict=True)
    N, E = 30, 200
    ei = torch.randint(0, N, (2, E))
    ew = torch.ones(E)
    ea = torch.randn(E, 3)
    x = torch.randn(N, 24, requires_grad=True)
    xc = x.detach().to(DEV).requires_grad_(True)
    yo = ob(x, ei, ew, ea)
    yc = cb(xc, ei.to(DEV), ew.to(DEV), ea.to(DEV))
    assert rel_err(yc, yo) < TOL
    yo.pow(2).sum().backward()
    yc.pow(2).sum().backward()
    assert rel_err(xc.grad, x.grad) < TOL
    for (k, po), (_, pc) in zip(ob.named_parameters(), cb.named_parameters()):
        assert rel_err(pc.grad, po.grad) < TOL, k


def test_padding_atom_and_bad_atomic_number():
    o, c = pair(9, hidden_channels=16, num_filters=16, num_interactions=1, num_gaussians=8, cutoff=5.0)
    b = syn.make_batch(1, 2, 6, seed=10)
    b.z[::3] = 0
    compare(o, c, b)
    assert c.embedding.weight.grad[0].abs().sum() == 0
    z = b.z.clone()
    z[1] = 100
    with pytest.raises(ValueError):
        c(z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))


def test_conformer_mean_and_regression_step_glue():
    """Row (f)-1 of SURVEY.md 8: the K-conformer mean (common.py:414-423, schnet_based_models.py:242)."""
    from conan_fgw_b200 import ops
    x = torch.randn(12, 7, device=DEV, requires_grad=True)
    y = ops.conformers_mean(x, 3)
    want = x.view(4, 3, 7).mean(dim=1)
    assert rel_err(y, want) < 1e-6
    (g,) = torch.autograd.grad(y.sum(), x)
    assert torch.allclose(g, torch.full_like(g, 1.0 / 3))
    with pytest.raises(ValueError):
        ops.conformers_mean(x, 5)


def _bond_graph(batch, seed=0):
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


def test_covalent_trunk_matches_oracle():
    """use_covalent=True (schnet_no_sum.py:131-142,166-175): forward, forward_3d_bary and every gradient, 1e-5."""
    import types

    torch.manual_seed(5)
    cfg = dict(hidden_channels=64, num_filters=64, num_interactions=2, num_gaussians=20, cutoff=6.0, use_covalent=True)
    o = osn.SchNetNoSum(None, **cfg)
    c = cmp.SchNetNoSum(None, **cfg).to(DEV)
    c.load_state_dict(o.state_dict(), strict=True)
    b = syn.make_batch(3, 2, 14, seed=8)
    db = _bond_graph(b.batch)
    dbc = types.SimpleNamespace(edge_index=db.edge_index.to(DEV), edge_attr=db.edge_attr.to(DEV))
    want = o(b.z, b.pos, b.batch, data_batch=db)
    got = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV), data_batch=dbc)
    assert rel_err(got, want) < 1e-5
    want.pow(2).mean().backward()
    got.pow(2).mean().backward()
    for (k, po), (_, pc) in zip(o.named_parameters(), c.named_parameters()):
        if po.grad is not None:
            assert rel_err(pc.grad, po.grad) < 1e-5, k
    h, hb = c.forward_3d_bary(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV), data_batch=dbc)
    ho, hbo = o.forward_3d_bary(b.z, b.pos, b.batch, data_batch=db)
    assert rel_err(h, ho) < 1e-5 and rel_err(hb, hbo) < 1e-5


def test_schnet_with_multiple_returns_matches_oracle():
    """schnet_no_sum.py:405-450: (ssp(lin1(h)), edge_index, rbf); edge_index bit-exact, the rest to 1e-5."""
    torch.manual_seed(6)
    cfg = dict(hidden_channels=64, num_filters=64, num_interactions=2, num_gaussians=20, cutoff=6.0)
    o = osn.SchNetWithMultipleReturns(**cfg)
    c = cmp.SchNetWithMultipleReturns(**cfg).to(DEV)
    c.load_state_dict(o.state_dict(), strict=True)
    b = syn.make_batch(3, 2, 14, seed=9)
    ho, eio, eao = o(b.z, b.pos, b.batch)
    hc, eic, eac = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    assert torch.equal(eic.cpu(), eio)
    assert rel_err(eac, eao) < 1e-5 and rel_err(hc, ho) < 1e-5

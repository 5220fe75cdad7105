"""The ViSNet oracle against outputs of the reference's own vendored file (golden) (CPU)."""
import os

import pytest
import torch

from oracle import visnet as ov
from conftest import load_golden, rel_err


def _load():
    g = load_golden("visnet_ref.pt")
    c = g["config"]
    m = ov.ViSNet(None, hidden_channels=c["hidden_channels"], cutoff=c["cutoff"], num_heads=c["num_heads"],
                  num_layers=c["num_layers"], num_rbf=c["num_rbf"])
    m.load_state_dict(g["state_dict"], strict=True)     # names equal the reference's state_dict
    return g, m


def test_forward_matches_reference_golden():
    g, m = _load()
    x, v = m.representation_model(g["z"], g["pos"], g["batch"])
    assert rel_err(x, g["x_repr"]) < 1e-6 and rel_err(v, g["vec_repr"]) < 1e-6
    a, ab = m.forward_3d_bary(g["z"], g["pos"], g["batch"])
    assert rel_err(a, g["per_atom"]) < 1e-6 and rel_err(ab, g["per_atom_bary"]) < 1e-6
    assert rel_err(m(g["z"], g["pos"], g["batch"]), g["y"]) < 1e-6


def test_gradients_match_reference_golden():
    g, m = _load()
    a, ab = m.forward_3d_bary(g["z"], g["pos"], g["batch"])
    G = int(g["batch"].max()) + 1
    y = a.new_zeros(G, a.size(1)).index_add_(0, g["batch"], a)
    loss = y.pow(2).mean() + 0.5 * ab.pow(2).mean()
    assert rel_err(loss, g["loss"]) < 1e-6
    loss.backward()
    params = dict(m.named_parameters())
    assert set(g["grads"]) <= set(params)
    for k, ref in g["grads"].items():
        assert rel_err(params[k].grad, ref) < 2e-5, k


def test_state_dict_has_the_171_reference_entries():
    m = ov.ViSNet(None, hidden_channels=128)
    sd = m.state_dict()
    assert len(sd) == 171 and sum(p.numel() for p in m.parameters()) == 1798472
    assert "representation_model.vis_mp_layers.5.f_proj.weight" not in sd        # last layer has no edge update
    assert "representation_model.vis_mp_layers.4.w_src_proj.weight" in sd


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="reference tree only exists in the build container")
def test_live_against_reference_file():
    from oracle import pyg_shim
    import conan_fgw_b200 as cmp

    tgv = pyg_shim.load_reference_visnet()
    torch.manual_seed(3)
    ref = tgv.ViSNet(hidden_channels=64, num_layers=2)
    m = ov.ViSNet(None, hidden_channels=64, num_layers=2)
    m.load_state_dict(ref.state_dict(), strict=True)
    b = cmp.synthetic.make_batch(2, 2, 14, seed=8)
    x, v = ref.representation_model(b.z, b.pos.clone(), b.batch)
    want = pyg_shim.scatter(ref.prior_model(ref.output_model.pre_reduce(x, v) * ref.std, b.z), b.batch, dim=0)
    assert rel_err(m(b.z, b.pos, b.batch), want) < 1e-6


def test_cuda_modules_expose_the_reference_state_dict():
    import conan_fgw_b200 as cmp

    o = ov.ViSNet(None, hidden_channels=128)
    c = cmp.ViSNet(None, hidden_channels=128)
    so, sc = o.state_dict(), c.state_dict()
    assert set(so) == set(sc) and len(sc) == 171
    assert {k: tuple(v.shape) for k, v in so.items()} == {k: tuple(v.shape) for k, v in sc.items()}
    c.load_state_dict(so, strict=True)
    assert sum(p.numel() for p in c.parameters()) == 1798472

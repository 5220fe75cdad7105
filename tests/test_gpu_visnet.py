"""ViSNet on the CUDA kernels against (a) outputs of the reference's own vendored file (golden fixture) and
(b) the CPU oracle on BASELINE-shaped batches (GPU, through the C ABI).  Tolerance 1e-5 relative (fp32)."""
import pytest
import torch

import conan_fgw_b200 as cmp
from oracle import visnet as ov
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
syn = cmp.synthetic
DEV = "cuda"
TOL = 1e-5


def test_golden_from_reference_file_forward_and_gradients():
    g = load_golden("visnet_ref.pt")
    c = g["config"]
    m = cmp.TorchGeometricViSNet(hidden_channels=c["hidden_channels"], num_layers=c["num_layers"],
                                 num_heads=c["num_heads"], num_rbf=c["num_rbf"], cutoff=c["cutoff"]).to(DEV)
    m.load_state_dict(g["state_dict"], strict=True)
    z, pos, batch = g["z"].to(DEV), g["pos"].to(DEV), g["batch"].to(DEV)
    x, v = m.representation_model(z, pos, batch)
    assert rel_err(x, g["x_repr"]) < TOL and rel_err(v, g["vec_repr"]) < TOL
    a, ab = m._per_atom(z, pos, batch, bary=True)
    assert rel_err(a, g["per_atom"]) < TOL and rel_err(ab, g["per_atom_bary"]) < TOL
    G = int(g["batch"].max()) + 1
    y = torch.zeros(G, a.size(1), device=DEV).index_add_(0, batch, a)
    assert rel_err(y, g["y"]) < TOL
    loss = y.pow(2).mean() + 0.5 * ab.pow(2).mean()
    assert rel_err(loss, g["loss"]) < TOL
    loss.backward()
    params = dict(m.named_parameters())
    for k, ref in g["grads"].items():
        assert params[k].grad is not None, k
        assert rel_err(params[k].grad, ref) < 5e-5, k


@pytest.mark.parametrize("B,K,n,H,L", [(4, 2, 18, 128, 6), (2, 2, 40, 64, 3), (3, 1, 5, 32, 2),
                                        (2, 2, 14, 512, 2)])       # the classification width (common.py:513-522)
def test_conan_wrapper_vs_oracle(B, K, n, H, L):
    torch.manual_seed(B + n)
    o = ov.ViSNet(None, hidden_channels=H, num_layers=L)
    with torch.no_grad():
        for name, p in o.named_parameters():
            if p.dim() <= 1 or "atomref" in name:
                p.add_(0.1 * torch.randn_like(p))
    if L == 6 and H == 128:
        c = cmp.ViSNet(None, hidden_channels=H).to(DEV)          # exactly ConAN's construction (visnet.py:84-86)
    else:
        c = cmp.ViSNet(None, hidden_channels=H)
        c.representation_model = cmp.ViSNetBlock(hidden_channels=H, num_layers=L)
        c = c.to(DEV)
    c.load_state_dict(o.state_dict(), strict=True)
    b = syn.make_batch(B, K, n, seed=n)
    out_o = o(b.z, b.pos, b.batch)
    out_c = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    assert out_c.shape == out_o.shape and rel_err(out_c, out_o) < TOL
    ho, hbo = o.forward_3d_bary(b.z, b.pos, b.batch)
    hc, hbc = c.forward_3d_bary(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    assert rel_err(hc, ho) < TOL and rel_err(hbc, hbo) < TOL
    (out_o.pow(2).mean() + hbo.pow(2).mean()).backward()
    (out_c.pow(2).mean() + hbc.pow(2).mean()).backward()
    po, pc = dict(o.named_parameters()), dict(c.named_parameters())
    for k in po:
        if po[k].grad is None:
            continue
        assert pc[k].grad is not None, k
        assert rel_err(pc[k].grad, po[k].grad) < 5e-5, k


def test_deterministic_and_module_signatures():
    c = cmp.ViSNet(None, hidden_channels=64).to(DEV)
    b = syn.make_config_batch("cfg3_freesolv_visnet", scale=0.25).to(DEV)
    outs = []
    for _ in range(2):
        c.zero_grad()
        out = c(b.z, b.pos, b.batch)
        out.pow(2).mean().backward()
        outs.append((out.detach().clone(), torch.cat([p.grad.reshape(-1) for p in c.parameters() if p.grad is not None])))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # tgv-style piecewise use (ViSNetBlock.forward, tgv.py:861-886)
    rm = c.representation_model
    ei, ew, ev = rm.distance(b.pos, b.batch)
    assert ei.shape[0] == 2 and ew.shape[0] == ei.shape[1] and ev.shape == (ei.shape[1], 3)
    assert int((ei[0] == ei[1]).sum()) == b.z.numel()            # loop=True: one self loop per atom
    rbf = rm.distance_expansion(ew)
    assert rbf.shape == (ei.shape[1], 32)


def test_tensor_core_linears_mode():
    """set_precision("bf16"): every Linear (incl. the 2H / 3H wide ones, blocked) on the split-bf16 tcgen05 kernels.
    Stated tolerance 2e-4 on embeddings / 2e-3 on gradients (each GEMM is good to ~2e-5)."""
    if not cmp._lib.lib().cmp_device_is_sm100():
        pytest.skip("tcgen05 kernels need an sm_100 device")
    torch.manual_seed(5)
    o = ov.ViSNet(None, hidden_channels=128)
    with torch.no_grad():
        for name, p in o.named_parameters():
            if p.dim() <= 1 or "atomref" in name:
                p.add_(0.1 * torch.randn_like(p))
    c = cmp.ViSNet(None, hidden_channels=128).to(DEV)
    c.load_state_dict(o.state_dict(), strict=True)
    c.set_precision("bf16")
    b = syn.make_batch(6, 2, 18, seed=3)
    out_o = o(b.z, b.pos, b.batch)
    out_c = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    assert rel_err(out_c, out_o) < 2e-4
    out_o.pow(2).mean().backward()
    out_c.pow(2).mean().backward()
    po, pc = dict(o.named_parameters()), dict(c.named_parameters())
    for k in po:
        if po[k].grad is not None:
            assert rel_err(pc[k].grad, po[k].grad) < 2e-3, k
    # the same step inside prepacked_weights (what dp.RegressionStep does): the block images of the wide Linears come from
    # the grouped pack, dX is chained over the row blocks with the cached transposed images - identical numbers
    from conan_fgw_b200 import ops
    ref_grads = {k: p.grad.clone() for k, p in pc.items() if p.grad is not None}
    c.zero_grad()
    with ops.prepacked_weights([c]):
        out_p = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
        out_p.pow(2).mean().backward()
    assert torch.equal(out_p, out_c)
    for k, g_ref in ref_grads.items():
        assert rel_err(pc[k].grad, g_ref) < 1e-6, k
    # wide / blocked linear against fp64 directly
    x = torch.randn(500, 256, device=DEV, requires_grad=True)
    w = (torch.randn(384, 256, device=DEV) / 16).requires_grad_(True)
    bias = torch.randn(384, device=DEV, requires_grad=True)
    y = ops.linear(x, w, bias, tc=True)
    yr = x.double() @ w.double().t() + bias.double()
    assert rel_err(y, yr) < 3e-5
    go = torch.randn_like(y)
    for a, r in zip(torch.autograd.grad(y, [x, w, bias], go), torch.autograd.grad(yr, [x, w, bias], go.double())):
        assert rel_err(a, r) < 6e-5


def test_deferred_grouped_weight_gradients_cover_the_blocked_linears():
    """ViSNet's wide projections (H -> 2H / 3H) queue one grouped-dW problem per <=128 x <=128 block, written straight
    into the full gradient (lddw): deferred + prepacked training-step gradients equal the immediate per-block launches."""
    import conan_fgw_b200 as cmp
    from conan_fgw_b200 import ops

    if not cmp._lib.lib().cmp_device_is_sm100():
        pytest.skip("tcgen05 kernels need an sm_100 device")
    torch.manual_seed(11)
    m = cmp.ViSNet(None, hidden_channels=128).to("cuda").set_precision("bf16")
    b = cmp.synthetic.make_batch(4, 2, 11, seed=3).to("cuda")
    params = list(m.parameters())

    def grads(deferred):
        for p in params:
            p.grad = None
        out = m(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean()
        if deferred:
            with ops.prepacked_weights([m]), ops.deferred_weight_grads():
                out.backward()
        else:
            out.backward()
        return [None if p.grad is None else p.grad.clone() for p in params]

    ref, got = grads(False), grads(True)
    assert sum(g is not None for g in ref) > 50
    for p, r, g in zip(params, ref, got):
        assert (r is None) == (g is None)
        if r is not None:
            scale = r.abs().max().item()
            assert (g - r).abs().max().item() <= 2e-6 * max(scale, 1e-12) + 1e-12, tuple(p.shape)

"""Dense-block fused CFConv (cfconv_dense.cu) against the exact-fp32 message kernels on the same neighbour list (GPU).

Covers complete graphs (<= 33 atoms), truncated / asymmetric graphs (max_num_neighbors), sparse graphs (small cutoff),
every block-boundary size, the 128-atom limit, the fallback above it, and the transposed (d x') pass.
Stated tolerance of the f16 filter MLP: 5e-3 relative (max|a-b| / max|b|) and 5e-3 per row."""
import pytest
import torch

import conan_fgw_b200 as cmp
from conan_fgw_b200 import ops
from conftest import rel_err, row_rel_err

pytestmark = pytest.mark.gpu
syn = cmp.synthetic
DEV = "cuda"
TOL = 5e-3
F, NG = 128, 50


def _need_sm100():
    if not cmp._lib.lib().cmp_device_is_sm100():
        pytest.skip("tcgen05 kernels need an sm_100 device")


@pytest.fixture(params=[-1, 0, 1], ids=["auto", "warp-specialised", "per-pipeline"])
def variant(request):
    """Both kernels behind cmp_cfconv_dense_fwd serve every conformer size (the default picks one by max_atoms_hint):
    the warp-specialised tile pipeline (registers-only path up to 32 atoms, general path above) and the per-pipeline one."""
    cmp._lib.lib().cmp_debug_set_dense_variant(request.param)
    yield request.param
    cmp._lib.lib().cmp_debug_set_dense_variant(-1)


def _block(cutoff, seed=0):
    torch.manual_seed(seed)
    blk = cmp.InteractionBlock(128, NG, F, cutoff).to(DEV)
    with torch.no_grad():
        blk.mlp[0].bias.add_(0.1 * torch.randn(F, device=DEV))
        blk.mlp[2].bias.add_(0.1 * torch.randn(F, device=DEV))
    gs = cmp.GaussianSmearing(0.0, cutoff, NG).to(DEV)
    return blk, gs


def _exact(blk, gs, nl, xp, cutoff, g=None):
    rbf = gs(nl.edge_weight())
    filt = blk.conv.filter(rbf).detach()
    xq = xp.detach().clone().requires_grad_(True)
    agg = ops.cfconv_message(xq, filt, nl, cutoff)
    if g is None:
        return agg.detach(), None
    (dx,) = torch.autograd.grad(agg, xq, g)
    return agg.detach(), dx


def _dense(blk, gs, nl, x, cutoff, transposed):
    W = (blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias)
    return ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, cutoff, transposed)


@pytest.mark.parametrize("n,B,K,cutoff,max_nb", [
    (27, 6, 3, 10.0, 32),     # cfg 2 shape: complete graphs, two row blocks
    (2, 5, 2, 10.0, 32), (3, 4, 1, 10.0, 32), (16, 3, 2, 10.0, 32), (17, 3, 2, 10.0, 32), (24, 2, 2, 10.0, 32),
    (25, 2, 2, 10.0, 32), (32, 2, 2, 10.0, 32), (33, 2, 2, 10.0, 32),
    (45, 2, 2, 10.0, 32),     # cfg 5 shape: truncated
    (65, 2, 2, 10.0, 32),     # cfg 4 shape: truncated, asymmetric
    (45, 2, 2, 5.0, 32),      # sparse (cutoff 5)
    (100, 1, 2, 10.0, 32), (128, 1, 1, 10.0, 32),
    (40, 2, 2, 10.0, 8),      # hard truncation: mostly one-directional pairs
])
def test_dense_kernel_matches_exact_message_path(n, B, K, cutoff, max_nb, variant):
    _need_sm100()
    b = syn.make_batch(B, K, n, seed=n).to(DEV)
    nl = cmp.build_neighbor_list(b.pos, b.batch, cutoff, max_nb, max_atoms=n)
    blk, gs = _block(cutoff, seed=n)
    torch.manual_seed(n + 1)
    xp = torch.randn(b.z.numel(), F, device=DEV)
    g = torch.randn(b.z.numel(), F, device=DEV)
    want, want_dx = _exact(blk, gs, nl, xp, cutoff, g)
    got = _dense(blk, gs, nl, xp, cutoff, False)
    got_dx = _dense(blk, gs, nl, g, cutoff, True)
    nl.check()
    assert rel_err(got, want) < TOL
    assert rel_err(got_dx, want_dx) < TOL
    assert row_rel_err(got, want) < 2 * TOL
    # deterministic (work is handed out by an atomic counter: the assignment of conformers to pipelines varies)
    for _ in range(3):
        assert torch.equal(got, _dense(blk, gs, nl, xp, cutoff, False))


def test_ragged_batch_and_isolated_atoms(variant):
    """Conformers of different sizes in one batch, single-atom conformers and atoms without any neighbour."""
    _need_sm100()
    sizes = [1, 27, 5, 64, 1, 33, 18, 2, 17, 32, 1]
    torch.manual_seed(0)
    pos, batch = [], []
    for g, n in enumerate(sizes):
        p = syn.make_batch(1, 1, n, seed=10 + g).pos
        if n == 18:
            p[3] += 100.0          # an atom out of everyone's range
        pos.append(p)
        batch.append(torch.full((n,), g, dtype=torch.int64))
    pos, batch = torch.cat(pos).to(DEV), torch.cat(batch).to(DEV)
    nl = cmp.build_neighbor_list(pos, batch, 10.0, max_atoms=max(sizes))
    blk, gs = _block(10.0, seed=3)
    xp = torch.randn(pos.size(0), F, device=DEV)
    want, _ = _exact(blk, gs, nl, xp, 10.0)
    got = _dense(blk, gs, nl, xp, 10.0, False)
    assert rel_err(got, want) < TOL
    iso = (nl.rowptr[1:] == nl.rowptr[:-1]).nonzero().flatten()
    assert iso.numel() >= 4 and bool((got[iso] == 0).all())
    # both kernels produce the same bits
    cmp._lib.lib().cmp_debug_set_dense_variant(1 if variant != 1 else 0)
    assert torch.equal(got, _dense(blk, gs, nl, xp, 10.0, False))


def test_conformers_above_the_dense_limit_use_the_per_edge_kernel(variant):
    _need_sm100()
    sizes = [140, 20, 131]
    pos = torch.cat([syn.make_batch(1, 1, n, seed=n).pos for n in sizes]).to(DEV)
    batch = torch.cat([torch.full((n,), g, dtype=torch.int64) for g, n in enumerate(sizes)]).to(DEV)
    nl = cmp.build_neighbor_list(pos, batch, 10.0)            # no max_atoms promise: fallback launch is issued
    blk, gs = _block(10.0, seed=5)
    xp = torch.randn(pos.size(0), F, device=DEV)
    g = torch.randn(pos.size(0), F, device=DEV)
    want, want_dx = _exact(blk, gs, nl, xp, 10.0, g)
    assert rel_err(_dense(blk, gs, nl, xp, 10.0, False), want) < TOL
    assert rel_err(_dense(blk, gs, nl, g, 10.0, True), want_dx) < TOL
    # a broken promise is reported through the status word, not by wrong numbers going unnoticed
    nl2 = cmp.build_neighbor_list(pos, batch, 10.0, max_atoms=64)
    _dense(blk, gs, nl2, xp, 10.0, False)
    with pytest.raises(RuntimeError):
        nl2.check()


def test_adjacency_bits_equal_the_edge_list():
    b = syn.make_batch(3, 2, 50, seed=5).to(DEV)
    nl = cmp.build_neighbor_list(b.pos, b.batch, 10.0)
    adj = nl.adjacency().cpu()
    ei = nl.edge_index().cpu()
    seg = nl.seg_ptr.cpu()
    conf = torch.bucketize(torch.arange(nl.N), seg[1:], right=True)
    want = torch.zeros(nl.N, 4, dtype=torch.int64)
    for j, i in zip(ei[0].tolist(), ei[1].tolist()):
        jl = j - int(seg[conf[i]])
        want[i, jl >> 5] |= 1 << (jl & 31)
    assert torch.equal(adj.to(torch.int64) & 0xFFFFFFFF, want)


@pytest.fixture(params=[0, 1], ids=["two-group", "warp-specialised"])
def bwd_variant(request):
    """Both kernels behind cmp_cfconv_dense_bwd_weights."""
    cmp._lib.lib().cmp_debug_set_dense_bwd_variant(request.param)
    yield request.param
    cmp._lib.lib().cmp_debug_set_dense_bwd_variant(0)


def _exact_weight_grads(blk, gs, nl, xp, g, cutoff):
    """Filter-MLP gradients through the exact-fp32 message kernels + autograd over materialised [E, F] rows."""
    ps = [t.detach().clone().requires_grad_(True) for t in (blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight,
                                                              blk.mlp[2].bias)]
    rbf = gs(nl.edge_weight())
    filt = ops.linear(ops.linear(rbf, ps[0], ps[1], cmp._lib.ACT_SSP), ps[2], ps[3])
    agg = ops.cfconv_message(xp.detach(), filt, nl, cutoff)
    return torch.autograd.grad(agg, ps, g)


@pytest.mark.parametrize("n,B,K,cutoff,max_nb", [
    (27, 6, 3, 10.0, 32), (2, 5, 2, 10.0, 32), (12, 3, 2, 10.0, 32), (16, 3, 2, 10.0, 32), (17, 3, 2, 10.0, 32),
    (33, 2, 2, 10.0, 32), (45, 2, 2, 10.0, 32), (65, 2, 2, 10.0, 32), (45, 2, 2, 5.0, 32), (128, 1, 1, 10.0, 32),
    (40, 2, 2, 10.0, 8),
])
def test_dense_weight_gradients_match_the_exact_path(n, B, K, cutoff, max_nb, bwd_variant):
    """cmp_cfconv_dense_bwd_weights (fp32 rows of g / x' in registers, columns over the dense blocks) against the exact
    gradients and against the pair-list kernel.  Stated tolerance of the bf16 operand images: 1e-2 relative."""
    _need_sm100()
    b = syn.make_batch(B, K, n, seed=n).to(DEV)
    nl = cmp.build_neighbor_list(b.pos, b.batch, cutoff, max_nb, max_atoms=n)
    blk, gs = _block(cutoff, seed=n)
    torch.manual_seed(n + 2)
    xp = torch.randn(b.z.numel(), F, device=DEV)
    g = torch.randn(b.z.numel(), F, device=DEV)
    want = _exact_weight_grads(blk, gs, nl, xp, g, cutoff)
    W = (blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight)
    res = {}
    for dense in (True, False):
        ops.FUSED_DENSE_GRADS = dense
        try:
            before = cmp._lib.launches()
            res[dense] = ops._fused_weight_grads(g, xp, *W, nl, gs.offset, gs.coeff, cutoff)
            res[dense] = tuple(t.clone() for t in res[dense]) + (cmp._lib.launches() - before,)
        finally:
            ops.FUSED_DENSE_GRADS = True
    nl.check()
    for name, got, pair, ref in zip(("dW1", "db1", "dW2", "db2"), res[True], res[False], want):
        assert rel_err(got, ref) < 1e-2, name
        assert rel_err(got, pair) < 1e-2, name
    # deterministic, and no conversion launches
    again = ops._fused_weight_grads(g, xp, *W, nl, gs.offset, gs.coeff, cutoff)
    for a, c in zip(again, res[True]):
        assert torch.equal(a, c)
    assert res[True][4] < res[False][4]


def test_dense_weight_gradients_on_a_ragged_batch(bwd_variant):
    _need_sm100()
    sizes = [1, 27, 5, 64, 1, 33, 18, 2, 17, 32, 1, 128, 3]
    pos = torch.cat([syn.make_batch(1, 1, n, seed=20 + i).pos for i, n in enumerate(sizes)]).to(DEV)
    batch = torch.cat([torch.full((n,), i, dtype=torch.int64) for i, n in enumerate(sizes)]).to(DEV)
    nl = cmp.build_neighbor_list(pos, batch, 10.0, max_atoms=max(sizes))
    blk, gs = _block(10.0, seed=7)
    xp = torch.randn(pos.size(0), F, device=DEV)
    g = torch.randn(pos.size(0), F, device=DEV)
    want = _exact_weight_grads(blk, gs, nl, xp, g, 10.0)
    got = ops._fused_weight_grads(g, xp, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, nl, gs.offset, gs.coeff,
                                  10.0)
    nl.check()
    for name, a, ref in zip(("dW1", "db1", "dW2", "db2"), got, want):
        assert rel_err(a, ref) < 1e-2, name

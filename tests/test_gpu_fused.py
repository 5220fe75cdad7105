"""Fused tcgen05 CFConv (bf16 filter MLP) against the exact-fp32 kernels and the CPU oracle (GPU).

Stated tolerance of the bf16 mode: 5e-3 relative (max|a-b|/max|b|) on embeddings and gradients
(bf16 operands in the filter MLP alone give 7e-4 .. 2e-3)."""
import pytest
import torch

import conan_fgw_b200 as cmp
from conan_fgw_b200 import ops
from oracle import schnet as osn
from conftest import rel_err

pytestmark = pytest.mark.gpu
syn = cmp.synthetic
DEV = "cuda"
TOL_BF16 = 5e-3


def _need_sm100():
    if not cmp._lib.lib().cmp_device_is_sm100():
        pytest.skip("tcgen05 kernels need an sm_100 device")


def make(seed=0, **cfg):
    torch.manual_seed(seed)
    o = osn.SchNetNoSum(None, **cfg)
    with torch.no_grad():
        for p in o.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    c = cmp.SchNetNoSum(None, **cfg).to(DEV)
    c.load_state_dict(o.state_dict(), strict=True)
    return o, c


def test_tiles_cover_every_edge_once():
    _need_sm100()
    for n, B in ((27, 6), (65, 2), (5, 7), (1, 3)):
        b = syn.make_batch(B, 2, n, seed=n).to(DEV)
        nl = cmp.build_neighbor_list(b.pos, b.batch, 10.0)
        tiles, num = nl.tiles()
        T = int(num.item())
        tl = tiles[:T].cpu()
        rp = nl.rowptr.cpu()
        covered = 0
        for fr, er, cs, cn, e0, ne_d, _, _ in tl.tolist():
            ne = int(rp[er] - rp[fr])
            assert e0 == int(rp[fr]) and ne_d == ne
            assert 0 < ne <= 128 and cs <= fr < er <= cs + cn
            covered += ne
        assert covered == nl.E
        if T > 1:
            assert bool((tl[1:, 0] >= tl[:-1, 1]).all())        # ordered, non-overlapping rows


def test_fused_kernel_matches_exact_message_path():
    _need_sm100()
    torch.manual_seed(1)
    b = syn.make_batch(6, 3, 27, seed=2).to(DEV)
    nl = cmp.build_neighbor_list(b.pos, b.batch, 10.0)
    F, Ng = 128, 50
    blk = cmp.InteractionBlock(128, Ng, F, 10.0).to(DEV)
    with torch.no_grad():
        blk.mlp[0].bias.add_(0.1 * torch.randn(F, device=DEV))
        blk.mlp[2].bias.add_(0.1 * torch.randn(F, device=DEV))
    gs = cmp.GaussianSmearing(0.0, 10.0, Ng).to(DEV)
    xp = torch.randn(b.z.numel(), F, device=DEV)
    rbf = gs(nl.edge_weight())
    filt = blk.conv.filter(rbf)
    want = ops.cfconv_message(xp, filt, nl, 10.0)
    got = ops.cfconv_fused(xp, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias, nl, gs.offset,
                           gs.coeff, 10.0)
    assert rel_err(got, want) < TOL_BF16
    # deterministic
    got2 = ops.cfconv_fused(xp, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias, nl, gs.offset,
                            gs.coeff, 10.0)
    assert torch.equal(got, got2)


@pytest.mark.parametrize("n,B,K,T", [(27, 8, 5, 6), (65, 2, 2, 3), (7, 5, 2, 2), (100, 1, 2, 2)])
def test_model_bf16_mode_vs_oracle(n, B, K, T):
    _need_sm100()
    o, c = make(3 + n, num_interactions=T)
    c.set_precision("bf16")
    b = syn.make_batch(B, K, n, seed=n)
    out_o = o(b.z, b.pos, b.batch)
    out_c = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    assert rel_err(out_c, out_o) < TOL_BF16
    out_o.pow(2).mean().backward()
    out_c.pow(2).mean().backward()
    for (k, po), (_, pc) in zip(o.named_parameters(), c.named_parameters()):
        if po.grad is None:
            continue
        assert rel_err(pc.grad, po.grad) < 2e-2, k      # gradients: d x' also runs through the bf16 filter
    # the exact mode of the same module still meets the fp32 bar
    c.set_precision("fp32")
    assert rel_err(c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV)), out_o) < 1e-5


def test_module_api_picks_fused_kernel():
    """Reference-style call sequence (sns.py:159-164) reaches the fused kernel through the tensor tags."""
    _need_sm100()
    o, c = make(9, num_interactions=2)
    c.set_precision("bf16")
    b = syn.make_batch(3, 2, 20, seed=4)
    z, pos, batch = b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV)
    h = c.embedding(z)
    edge_index, edge_weight = c.interaction_graph(pos, batch)
    edge_attr = c.distance_expansion(edge_weight)
    before = cmp._lib.launches()
    for interaction in c.interactions:
        h = h + interaction(h, edge_index, edge_weight, edge_attr)
    out = c.readout(c.act(c.lin2(c.lin1(h))), batch, dim=0)
    assert rel_err(out, o(b.z, b.pos, b.batch)) < TOL_BF16
    assert cmp._lib.launches() - before < 40


def test_pair_list_has_one_representative_per_undirected_pair():
    """cmp_build_pair_list against a set-based restatement: every directed edge is covered exactly once, either as a
    primary edge or as the reverse of one; unpaired edges (neighbour cap) are their own representative."""
    b = syn.make_batch(3, 2, 50, seed=5).to(DEV)           # 50 atoms, cap 32: truncated, asymmetric graph
    b2 = syn.make_batch(4, 3, 20, seed=6).to(DEV)          # complete graphs
    for bb in (b, b2):
        nl = cmp.build_neighbor_list(bb.pos, bb.batch, 10.0)
        E = nl.E
        src, dst, dist, rev, tiles, num, conf_ptr = nl.pair_tiles()
        nl.check()
        P = int(conf_ptr[-1])
        ei = nl.edge_index().cpu()
        edges = {(int(j), int(i)): k for k, (j, i) in enumerate(zip(ei[0], ei[1]))}
        ps, pd, pr, pdist = src[:P].cpu(), dst[:P].cpu(), rev[:P].cpu(), dist[:P].cpu()
        covered = set()
        for j, i, r, d in zip(ps.tolist(), pd.tolist(), pr.tolist(), pdist.tolist()):
            assert (j, i) in edges and (j, i) not in covered
            assert d == float(nl.dist[edges[(j, i)]])
            covered.add((j, i))
            assert bool(r) == ((i, j) in edges)
            if r:
                assert j > i and (i, j) not in covered
                covered.add((i, j))
        assert covered == set(edges) and len(covered) == E
        # sorted by (dst, src) inside a conformer, tiles cover [0, P) in 64-column chunks of one conformer
        key = pd * (bb.z.numel() + 1) + ps
        assert bool((key[1:] > key[:-1]).all())
        t = tiles[: int(num)].cpu()
        assert int(t[:, 5].sum()) == P and int(t[:, 5].max()) <= 64


@pytest.mark.parametrize("case", ["complete27", "tiny", "capped", "mixed", "single_pair"])
def test_pair_kernel_matches_per_edge_kernel_and_exact_path(case):
    """cmp_cfconv_pair_fwd (one filter evaluation per undirected pair, small conformers) against the per-edge fused
    kernel and the exact-fp32 message path, forward and d x' (transposed pass); unpaired edges (neighbour cap) and
    batches that mix small and large conformers included."""
    _need_sm100()
    torch.manual_seed(3)
    mnn = 32
    if case == "complete27":
        b = syn.make_batch(5, 3, 27, seed=11).to(DEV)
        z, pos, batch, G = b.z, b.pos, b.batch, b.num_graphs
    elif case == "tiny":
        b = syn.make_batch(40, 2, 3, seed=12).to(DEV)
        z, pos, batch, G = b.z, b.pos, b.batch, b.num_graphs
    elif case == "capped":            # 24 atoms but only 6 neighbours kept: many edges lose their reverse
        b = syn.make_batch(4, 3, 24, seed=13).to(DEV)
        z, pos, batch, G = b.z, b.pos, b.batch, b.num_graphs
        mnn = 6
    elif case == "mixed":             # 20- and 32-atom conformers (pair kernel) next to 33- and 50-atom ones (per-edge)
        parts = [syn.make_batch(3, 2, n, seed=20 + n).to(DEV) for n in (20, 50, 32, 33)]
        z, pos, batch, G = parts[0].z, parts[0].pos, parts[0].batch, parts[0].num_graphs
        for q in parts[1:]:
            z, pos, batch, G = (torch.cat([z, q.z]), torch.cat([pos, q.pos]), torch.cat([batch, q.batch + G]),
                                G + q.num_graphs)
    else:                             # two atoms: one pair, one tile of a single column
        pos = torch.tensor([[0.0, 0.0, 0.0], [1.2, 0.1, 0.0]], device=DEV)
        batch = torch.zeros(2, dtype=torch.long, device=DEV)
        G = 1
    nl = cmp.build_neighbor_list(pos, batch, 10.0, max_num_neighbors=mnn, num_graphs=G)
    N = pos.size(0)
    F, Ng = 128, 50
    blk = cmp.InteractionBlock(128, Ng, F, 10.0).to(DEV)
    with torch.no_grad():
        blk.mlp[0].bias.add_(0.1 * torch.randn(F, device=DEV))
        blk.mlp[2].bias.add_(0.1 * torch.randn(F, device=DEV))
    gs = cmp.GaussianSmearing(0.0, 10.0, Ng).to(DEV)
    params = [blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias]
    xp = torch.randn(N, F, device=DEV, requires_grad=True)
    go = torch.randn(N, F, device=DEV)
    res = {}
    for pairs in (False, True):
        ops.FUSED_PAIR_FORWARD = pairs
        try:
            out = ops.cfconv_fused(xp, *params, nl, gs.offset, gs.coeff, 10.0)
            (dx,) = torch.autograd.grad(out, [xp], go)
            out2 = ops.cfconv_fused(xp, *params, nl, gs.offset, gs.coeff, 10.0)
            (dx2,) = torch.autograd.grad(out2, [xp], go)
        finally:
            ops.FUSED_PAIR_FORWARD = True
        assert torch.equal(out, out2) and torch.equal(dx, dx2)          # bit-reproducible
        res[pairs] = (out.detach(), dx)
    nl.check()
    # same bf16 filter, different summation order only
    assert rel_err(res[True][0], res[False][0]) < 2e-5
    assert rel_err(res[True][1], res[False][1]) < 2e-5
    rbf = gs(nl.edge_weight())
    ref = ops.cfconv_message(xp, blk.conv.filter(rbf), nl, 10.0)
    (dref,) = torch.autograd.grad(ref, [xp], go)
    assert rel_err(res[True][0], ref) < TOL_BF16
    assert rel_err(res[True][1], dref) < TOL_BF16


@pytest.mark.parametrize("pairs", [True, False])
@pytest.mark.parametrize("n,B,cutoff", [(27, 8, 10.0), (65, 2, 10.0), (9, 4, 10.0), (45, 3, 5.0), (50, 2, 10.0)])
def test_fused_weight_gradients_match_exact_recompute(n, B, cutoff, pairs):
    """cmp_cfconv_fused_bwd_weights[_pairs] (TMEM-accumulated dW) vs the exact-fp32 recompute path."""
    _need_sm100()
    ops.FUSED_PAIR_GRADS = pairs
    try:
        _check_fused_weight_gradients(n, B, cutoff)
    finally:
        ops.FUSED_PAIR_GRADS = True


def _check_fused_weight_gradients(n, B, cutoff):
    torch.manual_seed(n)
    b = syn.make_batch(B, 3, n, seed=n + 1).to(DEV)
    nl = cmp.build_neighbor_list(b.pos, b.batch, cutoff)
    F, Ng = 128, 50
    blk = cmp.InteractionBlock(128, Ng, F, cutoff).to(DEV)
    with torch.no_grad():
        blk.mlp[0].bias.add_(0.1 * torch.randn(F, device=DEV))
        blk.mlp[2].bias.add_(0.1 * torch.randn(F, device=DEV))
    gs = cmp.GaussianSmearing(0.0, cutoff, Ng).to(DEV)
    xp = torch.randn(b.z.numel(), F, device=DEV, requires_grad=True)
    go = torch.randn(b.z.numel(), F, device=DEV)
    params = [blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias]
    res = {}
    for fused in (False, True):
        ops.FUSED_WEIGHT_GRADS = fused
        try:
            out = ops.cfconv_fused(xp, *params, nl, gs.offset, gs.coeff, cutoff)
            res[fused] = torch.autograd.grad(out, [xp] + params, go)
        finally:
            ops.FUSED_WEIGHT_GRADS = True
    names = ["dx'", "dW1", "db1", "dW2", "db2"]
    for name, a, r in zip(names, res[True], res[False]):
        assert rel_err(a, r) < 1e-2, name
    # and the exact (non-fused forward) autograd as ground truth
    rbf = gs(nl.edge_weight())
    out_ref = ops.cfconv_message(xp, blk.conv.filter(rbf), nl, cutoff)
    ref = torch.autograd.grad(out_ref, [xp] + params, go)
    for name, a, r in zip(names, res[True], ref):
        assert rel_err(a, r) < 2e-2, name
    # deterministic
    out = ops.cfconv_fused(xp, *params, nl, gs.offset, gs.coeff, cutoff)
    again = torch.autograd.grad(out, [xp] + params, go)
    for a, r in zip(again, res[True]):
        assert torch.equal(a, r)


@pytest.mark.parametrize("M,K,Nout", [(1000, 128, 128), (17280, 128, 128), (300, 128, 64), (77, 64, 64), (128, 64, 128)])
def test_node_gemm_tc_forward_and_gradients(M, K, Nout):
    """Split-bf16 tcgen05 node GEMMs (fwd, dX with fused ssp', dW/db with TMEM accumulation) vs fp64."""
    _need_sm100()
    import math
    torch.manual_seed(M + K)
    x = torch.randn(M, K, device=DEV, requires_grad=True)
    w = (torch.randn(Nout, K, device=DEV) / math.sqrt(K)).requires_grad_(True)
    bias = torch.randn(Nout, device=DEV, requires_grad=True)
    res = torch.randn(M, Nout, device=DEV, requires_grad=True)
    for act, use_res in ((cmp._lib.ACT_NONE, True), (cmp._lib.ACT_SSP, False)):
        y = ops.linear(x, w, bias, act, res if use_res else None, tc=True)
        yr = x.double() @ w.double().t() + bias.double()
        if act == cmp._lib.ACT_SSP:
            yr = torch.nn.functional.softplus(yr) - math.log(2.0)
        if use_res:
            yr = yr + res.double()
        assert rel_err(y, yr) < 2e-5
        go = torch.randn_like(y)
        ins = [x, w, bias] + ([res] if use_res else [])
        got = torch.autograd.grad(y, ins, go)
        want = torch.autograd.grad(yr, ins, go.double())
        for a, b in zip(got, want):
            assert rel_err(a, b) < 5e-5
        again = torch.autograd.grad(ops.linear(x, w, bias, act, res if use_res else None, tc=True), [w], go)[0]
        assert torch.equal(again, got[1])


def test_deferred_grouped_weight_gradients_match_the_immediate_launches():
    """ops.deferred_weight_grads: the node-linear dW / db of a whole backward pass in one grouped launch
    (cmp_node_gemm_dw_grouped) against the per-layer launches - same split-bf16 arithmetic, different tile-to-CTA
    partition, so equal to fp32 rounding; bit-reproducible; and every gradient tensor really is filled in place."""
    _need_sm100()
    cfg = dict(hidden_channels=128, num_filters=128, num_interactions=3, num_gaussians=50, cutoff=10.0)
    _, c = make(seed=4, **cfg)
    c.set_precision("bf16")
    b = syn.make_batch(6, 3, 23, seed=8).to(DEV)
    params = [p for p in c.parameters()]

    def grads(deferred):
        for p in params:
            p.grad = None
        out = c(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean()
        if deferred:
            with ops.deferred_weight_grads():
                out.backward()
        else:
            out.backward()
        return [None if p.grad is None else p.grad.clone() for p in params]

    ref = grads(False)
    got = grads(True)
    again = grads(True)
    assert sum(g is not None for g in ref) > 20
    for p, r, g, a in zip(params, ref, got, again):
        assert (r is None) == (g is None)
        if r is not None:
            assert rel_err(g, r) < 2e-6
            assert torch.equal(g, a)
    # an empty queue goes through the same exit path
    with ops.deferred_weight_grads():
        pass
    # gradients that already exist would be ACCUMULATED by autograd before the grouped launch writes them: refused loudly
    out = c(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean()
    with pytest.raises(cmp._lib.ConanMPError):
        with ops.deferred_weight_grads():
            out.backward()
    assert ops._dw_queue() is None
    # a context only takes the parameters it owns: somebody else's Linear launches immediately
    for p in params:
        p.grad = None
    out = c(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean()
    with ops.deferred_weight_grads([torch.nn.Parameter(torch.zeros(1, device=DEV))]) as ctx:
        out.backward()
        assert ctx.queue == []
    for g, r in zip([p.grad for p in params], ref):
        if r is not None:
            assert rel_err(g, r) < 2e-6


def test_training_step_with_grouped_packs_and_deferred_gradients_matches_plain_autograd():
    """dp.RegressionStep (weight images packed up front in grouped launches, node-linear dW deferred into one grouped
    launch, CUDA-graph capture) against the same loss differentiated layer by layer: identical loss, gradients equal to
    fp32 rounding, and the captured graph replays to the same numbers."""
    _need_sm100()
    from conan_fgw_b200.dp import RegressionStep

    cfg = dict(hidden_channels=128, num_filters=128, num_interactions=2, num_gaussians=50, cutoff=10.0)
    b = syn.make_batch(8, 5, 21, seed=9).to(DEV)
    targets = torch.randn(8, 1, device=DEV)

    def fresh():
        torch.manual_seed(5)
        m = cmp.SchNetNoSum(None, **cfg).to(DEV).set_precision("bf16")
        return RegressionStep(m, 64, 5, lr=1e-3)

    plain = fresh()
    plain.flat.zero_grad()
    loss_ref = plain.loss(b.z, b.pos, b.batch, targets, b.num_graphs)
    loss_ref.backward()
    plain.flat.collect_grads()
    g_ref = plain.flat.grad.clone()

    step = fresh()
    loss = step._fwd_bwd(b.z, b.pos, b.batch, targets, b.num_graphs)
    assert torch.equal(loss, loss_ref.detach())
    assert rel_err(step.flat.grad, g_ref) < 2e-6
    assert not ops._stack("packs") and ops._dw_queue() is None

    # two optimizer steps eagerly vs from the captured graph: same parameters afterwards
    eager, graphed = fresh(), fresh()
    graphed.capture(b.z, b.pos, b.batch, targets, b.num_graphs)
    graphed.flat.flat.copy_(eager.flat.flat)                 # capture warm-up ran no optimizer step, but be explicit
    for _ in range(2):
        l_e = eager.step(b.z, b.pos, b.batch, targets, b.num_graphs)
        l_g = graphed.step(b.z, b.pos, b.batch, targets, b.num_graphs)
        assert torch.equal(l_e, l_g)
    assert torch.equal(eager.flat.flat, graphed.flat.flat)


@pytest.mark.parametrize("case", ["single_atoms", "no_edges", "empty_batch", "ragged"])
def test_bf16_mode_edge_cases_against_oracle(case):
    """Degenerate inputs through the whole bf16 model (pair list, pair kernel, per-edge kernel, grouped launches):
    conformers of one atom, a cutoff below every distance, an empty batch, ragged conformer sizes around the 32-atom
    switch between the two forward kernels."""
    _need_sm100()
    cfg = dict(hidden_channels=128, num_filters=128, num_interactions=2, num_gaussians=50, cutoff=10.0)
    torch.manual_seed(6)
    if case == "single_atoms":
        sizes = [1, 5, 1, 1, 12]
    elif case == "ragged":
        sizes = [31, 32, 33, 34, 2, 40]
    elif case == "no_edges":
        sizes = [6, 9]
        cfg["cutoff"] = 0.05
    else:
        sizes = []
    o, c = make(seed=7, **cfg)
    c.set_precision("bf16")
    batch = torch.cat([torch.full((n,), g, dtype=torch.long) for g, n in enumerate(sizes)]) if sizes \
        else torch.zeros(0, dtype=torch.long)
    N = batch.numel()
    pos = torch.rand(N, 3) * 4.0
    z = torch.randint(1, 10, (N,))
    ref = o(z, pos, batch) if N else torch.zeros(0, 64)
    out = c(z.to(DEV), pos.to(DEV), batch.to(DEV), num_graphs=len(sizes))
    assert out.shape == (len(sizes), 64)
    if N:
        assert rel_err(out, ref) < TOL_BF16
        ref.pow(2).sum().backward()
        out.pow(2).sum().backward()
        for (n_, po), (_, pc) in zip(o.named_parameters(), c.named_parameters()):
            if po.grad is not None and po.grad.abs().max() > 0:
                assert rel_err(pc.grad, po.grad) < 2e-2, n_


def test_node_chain_kernel_matches_fp64_reference():
    """cmp_node_chain_fwd: 1 - 3 chained split-bf16 linears with every epilogue option (bias, ssp, residual, ssp' scaling
    by a saved activation), ragged tile (M not a multiple of 64), 64- and 128-wide stages; fp32-grade: 5e-5 vs fp64."""
    _need_sm100()
    torch.manual_seed(11)
    M = 64 * 5 + 23

    def ssp(t):
        return torch.nn.functional.softplus(t) - 0.6931471805599453

    def run(dims, opts):
        X = torch.randn(M, dims[0], device=DEV)
        stages, u = [], X.double()
        keep = []
        for s, (K, Nout) in enumerate(zip(dims[:-1], dims[1:])):
            W = torch.randn(Nout, K, device=DEV) / K ** 0.5
            o = opts[s]
            bias = torch.randn(Nout, device=DEV) if o.get("bias") else None
            res = torch.randn(M, Nout, device=DEV) if o.get("residual") else None
            ysv = torch.randn(M, Nout, device=DEV) if o.get("scale_y") else None
            out = torch.empty(M, Nout, device=DEV) if (o.get("out", True) or s == len(dims) - 2) else None
            img = ops._pack_node_weight(W, False)
            stages.append(dict(img=img, K=K, Nout=Nout, bias=bias, act=ops.ACT_SSP if o.get("ssp") else ops.ACT_NONE,
                               residual=res, scale_y=ysv, out=out))
            v = u @ W.double().t()
            if bias is not None:
                v = v + bias.double()
            if o.get("ssp"):
                v = ssp(v)
            if ysv is not None:
                v = v * (1.0 - 0.5 * torch.exp(-ysv.double()))
            if res is not None:
                v = v + res.double()
            keep.append((out, v))
            u = v
        ops._chain(X, stages)
        for out, v in keep:
            if out is not None:
                assert rel_err(out, v) < 5e-5

    run([128, 128], [dict(bias=True, ssp=True)])
    run([128, 128, 128], [dict(bias=True, ssp=True), dict(bias=True, residual=True)])
    run([128, 128, 128, 128], [dict(bias=True, ssp=True), dict(bias=True, residual=True), dict()])
    run([128, 128, 128, 128], [dict(residual=True), dict(scale_y=True), dict()])            # the backward chain
    run([128, 128, 128, 128], [dict(bias=True, ssp=True, out=False), dict(bias=True, residual=True, out=False), dict()])
    run([128, 64, 64], [dict(bias=True), dict(bias=True, ssp=True)])                        # head-shaped stages
    run([64, 128, 32], [dict(bias=True, ssp=True), dict(residual=True)])


def test_chained_trunk_equals_per_layer_launches():
    """nn.CHAIN_NODE_LINEARS: the same split-bf16 arithmetic in one kernel per block tail - embeddings and every gradient
    agree with the one-launch-per-Linear path to fp32 rounding."""
    _need_sm100()
    from conan_fgw_b200 import nn as cnn

    _, c = make(seed=6, num_interactions=3)
    c.set_precision("bf16")
    b = syn.make_batch(5, 3, 23, seed=9).to(DEV)
    res = {}
    for chained in (False, True):
        cnn.CHAIN_NODE_LINEARS = chained
        try:
            c.zero_grad()
            before = cmp._lib.launches()
            out = c(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
            out.pow(2).mean().backward()
            res[chained] = (out.detach().clone(), {k: p.grad.clone() for k, p in c.named_parameters() if p.grad is not None},
                            cmp._lib.launches() - before)
        finally:
            cnn.CHAIN_NODE_LINEARS = True
    assert rel_err(res[True][0], res[False][0]) < 2e-5
    for k, g in res[False][1].items():
        # the filter-MLP weight gradients are accumulated from bf16-staged copies of g and x' (cfconv_tc_bwd.cu): an fp32-
        # rounding difference of those inputs flips bf16 roundings, so these parameters agree to 2e-4, the rest to 5e-5
        tol = 2e-4 if ".mlp." in k else 5e-5
        assert rel_err(res[True][1][k], g) < tol, k
    assert res[True][2] < res[False][2]


# ---- ConAN's classification shape (common.py:513-522: H = 512, F = 256, Ng = 10, T = 3) through the fused kernels -----------
# num_filters = 256 is composed from 128-channel blocks of the same kernels (ops.cfconv_fused); Ng = 10 is one K step.
@pytest.mark.parametrize("F,Ng", [(256, 10), (256, 50), (384, 10), (128, 10)])
def test_wide_filter_layer_matches_exact_message_path(F, Ng):
    _need_sm100()
    torch.manual_seed(F + Ng)
    b = syn.make_batch(3, 2, 40, seed=F).to(DEV)
    nl = cmp.build_neighbor_list(b.pos, b.batch, 10.0, max_atoms=40)
    blk = cmp.InteractionBlock(64, Ng, F, 10.0).to(DEV)
    with torch.no_grad():
        blk.mlp[0].bias.add_(0.1 * torch.randn(F, device=DEV))
        blk.mlp[2].bias.add_(0.1 * torch.randn(F, device=DEV))
    gs = cmp.GaussianSmearing(0.0, 10.0, Ng).to(DEV)
    params = [blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias]
    xp = torch.randn(b.z.numel(), F, device=DEV)
    g = torch.randn(b.z.numel(), F, device=DEV)
    assert ops.fused_supported(F, Ng)

    xq = xp.clone().requires_grad_(True)
    want = ops.cfconv_message(xq, blk.conv.filter(gs(nl.edge_weight())), nl, 10.0)
    want_g = torch.autograd.grad(want, [xq] + params, g)
    for x3, tol, tol_g in ((False, TOL_BF16, 2e-2), (True, 1e-5, 1e-5)):
        xr = xp.clone().requires_grad_(True)
        got = ops.cfconv_fused(xr, *params, nl, gs.offset, gs.coeff, 10.0, x3=x3)
        got_g = torch.autograd.grad(got, [xr] + params, g)
        nl.check()
        assert got.shape == want.shape
        assert rel_err(got, want) < tol, x3
        for name, a, r in zip(("dx", "dW1", "db1", "dW2", "db2"), got_g, want_g):
            assert a.shape == r.shape
            assert rel_err(a, r) < tol_g, (x3, name)


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_classification_shape_runs_the_fused_kernels(prec):
    """tests/test_gpu_schnet.py::test_classification_shape in the fused modes: no [E, *] tensor, no exact message kernel."""
    _need_sm100()
    o, c = make(4, hidden_channels=512, num_filters=256, num_gaussians=10, num_interactions=3)
    c.set_precision(prec)
    c.max_atoms_hint = 20
    b = syn.make_batch(2, 2, 20, seed=5)
    lib = cmp._lib
    lib.timer = lib.KernelTimer(["cmp_cfconv_dense_fwd", "cmp_cfconv_dense_x3_fwd", "cmp_cfconv_message_fwd"])
    try:
        out_o = o(b.z, b.pos, b.batch)
        out_c = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
        out_o.pow(2).mean().backward()
        out_c.pow(2).mean().backward()
        c.check_status()
        torch.cuda.synchronize()
        seen = {k: v[0] for k, v in lib.timer.summary().items()}
    finally:
        lib.timer = None
    fused = "cmp_cfconv_dense_fwd" if prec == "bf16" else "cmp_cfconv_dense_x3_fwd"
    assert seen.get(fused, 0) == 3 * 4 * 2          # 3 blocks x (256 / 128)^2 launches x (forward + d x')
    assert seen.get("cmp_cfconv_message_fwd", 0) == 0
    tol, tol_g = (TOL_BF16, 2e-2) if prec == "bf16" else (1e-5, 2e-5)
    assert rel_err(out_c, out_o) < tol
    for (k, po), (_, pc) in zip(o.named_parameters(), c.named_parameters()):
        if po.grad is not None:
            assert rel_err(pc.grad, po.grad) < tol_g, k

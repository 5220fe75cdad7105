"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there): E(3) invariance,
independence of a conformer's embedding from the rest of the batch, equivariance under re-ordering of molecules.
GPU, through the C ABI, both numerics modes."""
import pytest
import torch

import conan_fgw_b200 as cmp
from conftest import rel_err

pytestmark = pytest.mark.gpu
syn = cmp.synthetic
DEV = "cuda"


def _model(precision, **cfg):
    torch.manual_seed(0)
    m = cmp.SchNetNoSum(None, **cfg).to(DEV)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    if precision == "bf16" and not cmp._lib.lib().cmp_device_is_sm100():
        pytest.skip("tcgen05 kernels need an sm_100 device")
    return m.set_precision(precision)


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-5), ("bf16", 5e-3)])
@pytest.mark.parametrize("cfg,scale", [("cfg2_lipo_train", 1.0), ("cfg4_bace_cls", 0.25)])
def test_e3_invariance_full_size(cfg, scale, precision, tol):
    m = _model(precision)
    b = syn.make_config_batch(cfg, scale=scale)
    z, pos, batch = b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV)
    with torch.no_grad():
        out = m(z, pos, batch)
        q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))
        pos2 = ((b.pos.double() - b.pos.double().mean(0)) @ q).float().to(DEV)   # rotate about the centroid
        out2 = m(z, pos2, batch)
    assert out.shape == (b.num_graphs, 64)
    # the rotated coordinates are re-rounded to fp32, so pairs within 1e-6 of the cutoff may flip: none at these sizes
    assert rel_err(out2, out) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 0.0), ("bf16", 0.0)])
def test_conformers_are_independent_of_the_rest_of_the_batch(precision, tol):
    """Edges never cross conformers (datasets.py:180-199): a sub-batch reproduces its rows of the full batch
    bit for bit (every kernel sums in a fixed, batch-independent order)."""
    m = _model(precision)
    b = syn.make_config_batch("cfg2_lipo_train")
    n = b.atoms_per_conformer
    with torch.no_grad():
        full = m(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
        k = 35 * n       # first 35 conformers = 7 molecules
        part = m(b.z[:k].to(DEV), b.pos[:k].to(DEV), b.batch[:k].to(DEV))
    assert torch.equal(part, full[:35])


def test_molecule_reordering_permutes_the_embeddings():
    m = _model("fp32", num_interactions=3)
    b = syn.make_config_batch("cfg1_esol_fwd")
    n, K, B = b.atoms_per_conformer, b.num_conformers, b.num_molecules
    perm = torch.randperm(B)
    atom_idx = (perm[:, None] * (K * n) + torch.arange(K * n)[None, :]).reshape(-1)
    conf_idx = (perm[:, None] * K + torch.arange(K)[None, :]).reshape(-1)
    with torch.no_grad():
        out = m(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
        out_p = m(b.z[atom_idx].to(DEV), b.pos[atom_idx].to(DEV), b.batch.to(DEV))
    assert torch.equal(out_p, out[conf_idx.to(DEV)])


def test_gradients_are_finite_and_deterministic_at_full_size():
    if not cmp._lib.lib().cmp_device_is_sm100():
        pytest.skip("tcgen05 kernels need an sm_100 device")
    m = _model("bf16")
    b = syn.make_config_batch("cfg2_lipo_train").to(DEV)
    flat = []
    for _ in range(2):
        m.zero_grad()
        m(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean().backward()
        flat.append(torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]))
    assert torch.isfinite(flat[0]).all() and torch.equal(flat[0], flat[1])


@pytest.mark.parametrize("cfg,molecules,cutoff", [("cfg4_bace_cls", 256, 10.0),        # BASELINE configs[3], full batch
                                                  ("cfg5_cov2_stress", 128, 10.0),      # configs[4]: one GPU's shard of 8
                                                  ("cfg5_cov2_stress", 128, 5.0)])      # ... and its short-cutoff arm
def test_fused_mode_agrees_with_exact_mode_at_large_sizes(cfg, molecules, cutoff):
    """The two numerics modes are independent kernel sets (SIMT fp32 with the E x F filter in HBM, tcgen05 with the
    filter in TMEM): at sizes the oracle cannot reach they check each other - forward embedding within the bf16
    tolerance, same neighbour list - and the fused training step stays finite and bit-reproducible."""
    c = syn.CONFIGS[cfg]
    b = syn.make_batch(molecules, c["num_conformers"], c["atoms"], seed=4321).to(DEV)
    exact = _model("fp32", cutoff=cutoff)
    fused = _model("bf16", cutoff=cutoff)
    fused.load_state_dict(exact.state_dict())
    with torch.no_grad():
        ref = exact(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
        out = fused(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
    assert ref.shape == (b.num_graphs, 64) and torch.isfinite(ref).all()
    assert rel_err(out, ref) < 5e-3
    del exact, ref
    torch.cuda.empty_cache()
    grads = []
    for _ in range(2):
        fused.zero_grad()
        fused(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean().backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in fused.parameters() if p.grad is not None]))
    assert torch.isfinite(grads[0]).all() and torch.equal(grads[0], grads[1])

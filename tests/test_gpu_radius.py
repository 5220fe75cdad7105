"""Neighbour lists from the CUDA kernels, bit-exact against the oracle (GPU, through the C ABI)."""
import pytest
import torch

import conan_fgw_b200 as cmp
from oracle.radius import radius_graph_ref
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
syn = cmp.synthetic
DEV = "cuda"


def check(pos, batch, r, loop, mnn=32):
    want = radius_graph_ref(pos, r, batch, loop, mnn)
    nl = cmp.build_neighbor_list(pos.to(DEV), None if batch is None else batch.to(DEV), r, mnn, loop, want_evec=True)
    got = nl.edge_index().cpu()
    assert got.dtype == torch.int64 and got.shape == want.shape, (got.shape, want.shape)
    assert torch.equal(got, want)
    E = want.size(1)
    # distances / edge vectors
    if E:
        vec = pos[want[0]] - pos[want[1]]
        assert torch.allclose(nl.dist[:E].cpu(), vec.norm(dim=-1), rtol=1e-6, atol=1e-7)
        assert torch.allclose(nl.evec[:E].cpu(), vec, rtol=0, atol=0)
    # transpose: same edge multiset, source-major / target ascending, eid_t points at the CSR slot
    N = pos.size(0)
    rt, ct, et = nl.rowptr_t.cpu().long(), nl.col_t[:E].cpu().long(), nl.eid_t[:E].cpu().long()
    assert rt[0] == 0 and rt[-1] == E
    src_t = torch.repeat_interleave(torch.arange(N), rt[1:] - rt[:-1])
    assert torch.equal(want[0][et], src_t) and torch.equal(want[1][et], ct)
    key = src_t * (N + 1) + ct
    assert bool((key[1:] > key[:-1]).all()) if E > 1 else True
    return nl


def test_golden_cases_bit_exact():
    for name, c in load_golden("radius_oracle.pt").items():
        nl = cmp.build_neighbor_list(c["pos"].to(DEV), c["batch"].to(DEV), c["r"], 32, c["loop"])
        assert torch.equal(nl.edge_index().cpu().to(torch.int32), c["edge_index"]), name


@pytest.mark.parametrize("cfg,scale", [("cfg1_esol_fwd", 1.0), ("cfg2_lipo_train", 0.25), ("cfg4_bace_cls", 0.05),
                                       ("cfg5_cov2_stress", 0.01)])
@pytest.mark.parametrize("loop", [False, True])
def test_config_shapes(cfg, scale, loop):
    b = syn.make_config_batch(cfg, scale=scale)
    for r in (5.0, 10.0):
        check(b.pos, b.batch, r, loop)


def test_visnet_shape_loop_true():
    b = syn.make_config_batch("cfg3_freesolv_visnet")
    nl = check(b.pos, b.batch, 5.0, True)
    assert nl.E <= 51840


def test_edge_cases():
    check(torch.zeros(1, 3), None, 5.0, False)                      # single atom, no edges
    check(torch.zeros(1, 3), None, 5.0, True)                       # single self loop
    pos = torch.tensor([[0.0, 0, 0], [3.0, 0, 0], [0.0, 0, 0], [0.0, 4.0, 0]])
    check(pos, None, 3.0, False)                                    # atom exactly at the cutoff, duplicates (d = 0)
    check(pos, None, 5.0, False)                                    # |1-3| = 5 exactly
    g = torch.Generator().manual_seed(0)
    check(torch.rand(50, 3, generator=g), None, 10.0, False)        # truncation regime, 33-neighbour atoms
    check(torch.rand(50, 3, generator=g), None, 10.0, True)
    check(torch.rand(300, 3, generator=g) * 12, None, 4.0, False, mnn=8)   # one big conformer, small cap
    # ragged conformers incl. an empty id in the middle
    batch = torch.tensor([0] * 5 + [1] * 1 + [3] * 7)
    check(torch.rand(13, 3, generator=g) * 3, batch, 2.0, False)
    # empty input
    nl = cmp.build_neighbor_list(torch.zeros(0, 3, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV), 5.0)
    assert nl.E == 0 and nl.edge_index().shape == (2, 0)


def test_unsorted_batch_raises():
    pos = torch.rand(6, 3, device=DEV)
    with pytest.raises(ValueError):
        cmp.radius_graph(pos, 5.0, torch.tensor([0, 1, 0, 1, 2, 2], device=DEV))


def test_radius_graph_signature_and_interaction_graph():
    b = syn.make_batch(2, 2, 12, seed=3)
    ei = cmp.radius_graph(b.pos.to(DEV), r=4.0, batch=b.batch.to(DEV), loop=False, max_num_neighbors=32,
                          flow="source_to_target", num_workers=1)
    assert torch.equal(ei.cpu(), radius_graph_ref(b.pos, 4.0, b.batch))
    t = cmp.radius_graph(b.pos.to(DEV), 4.0, b.batch.to(DEV), flow="target_to_source")
    assert torch.equal(t, ei.flip(0))
    ig = cmp.RadiusInteractionGraph(cutoff=4.0, max_num_neighbors=32)
    ei2, ew = ig(b.pos.to(DEV), b.batch.to(DEV))
    assert torch.equal(ei2, ei)
    want = (b.pos[ei.cpu()[0]] - b.pos[ei.cpu()[1]]).norm(dim=-1)
    assert rel_err(ew, want) < 1e-6


def test_large_sortedness_property():
    # full-size config 4: check structural properties instead of the (slow) oracle
    b = syn.make_config_batch("cfg4_bace_cls", scale=0.5)
    nl = cmp.build_neighbor_list(b.pos.to(DEV), b.batch.to(DEV), 10.0)
    ei = nl.edge_index()
    assert torch.equal(b.batch.to(DEV)[ei[0]], b.batch.to(DEV)[ei[1]])
    key = ei[1] * ei.max().add(1) + ei[0]
    assert bool((key[1:] > key[:-1]).all())
    deg = torch.bincount(ei[1], minlength=b.z.numel())
    assert int(deg.max()) <= 33 and int(deg.min()) >= 1
    d = (b.pos.to(DEV)[ei[0]] - b.pos.to(DEV)[ei[1]]).norm(dim=-1)
    assert float(d.max()) < 10.0

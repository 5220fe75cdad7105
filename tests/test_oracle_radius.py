"""Oracle self-checks for the neighbour search (CPU)."""
import numpy as np
import pytest
import torch

import conan_fgw_b200 as cmp
from oracle.radius import radius_graph_loops_py, radius_graph_ref, radius_interaction_graph_ref
from conftest import load_golden

syn = cmp.synthetic


@pytest.mark.parametrize("n,r,loop", [(26, 10.0, False), (40, 10.0, False), (40, 10.0, True), (18, 5.0, True),
                                      (12, 2.0, False)])
def test_vectorised_equals_literal_loop(n, r, loop):
    b = syn.make_batch(2, 2, n, seed=n)
    assert torch.equal(radius_graph_ref(b.pos, r, b.batch, loop, 32), radius_graph_loops_py(b.pos, r, b.batch, loop, 32))


def test_truncation_rule_first_cap_ascending():
    # 50 atoms all within range: atom i keeps candidates {0..32} (33 incl. self) -> 32 neighbours if i <= 32 else 33
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(50, 3, generator=g)
    ei = radius_graph_ref(pos, 10.0, None, False, 32)
    deg = torch.bincount(ei[1], minlength=50)
    assert deg[:33].eq(32).all() and deg[33:].eq(33).all()
    for i in (0, 10, 40):
        src = ei[0][ei[1] == i]
        want = [j for j in range(33) if j != i]
        assert src.tolist() == want
    # loop=True: cap 32 including self
    ei = radius_graph_ref(pos, 10.0, None, True, 32)
    deg = torch.bincount(ei[1], minlength=50)
    assert deg.eq(32).all()
    assert ei[0][ei[1] == 5].tolist() == list(range(32))
    assert ei[0][ei[1] == 45].tolist() == list(range(32))  # self (45) is beyond the cap and dropped


def test_edges_never_cross_conformers_and_order():
    b = syn.make_batch(3, 2, 20, seed=4)
    ei = radius_graph_ref(b.pos, 10.0, b.batch, False, 32)
    assert torch.equal(b.batch[ei[0]], b.batch[ei[1]])
    key = ei[1] * 10_000 + ei[0]
    assert bool((key[1:] > key[:-1]).all())  # target-major, source ascending, no duplicates
    assert bool((ei[0] != ei[1]).all())


def test_strict_cutoff_and_duplicates():
    pos = torch.tensor([[0.0, 0, 0], [3.0, 0, 0], [0.0, 0, 0], [0.0, 4.0, 0]])
    # |0-1| = 3 exactly: strict '<' excludes it at r=3, includes at r=3.0001
    ei = radius_graph_ref(pos, 3.0, None, False, 32)
    pairs = set(map(tuple, ei.t().tolist()))
    assert (1, 0) not in pairs and (0, 2) in pairs and (2, 0) in pairs  # duplicates at d=0 are neighbours
    ei = radius_graph_ref(pos, 3.0001, None, False, 32)
    assert (1, 0) in set(map(tuple, ei.t().tolist()))
    _, ew = radius_interaction_graph_ref(pos, None, 5.5, 32)
    assert ew.min().item() == 0.0


def test_single_atom_empty_and_unsorted():
    assert radius_graph_ref(torch.zeros(1, 3), 5.0, None, False, 32).shape == (2, 0)
    assert radius_graph_ref(torch.zeros(1, 3), 5.0, None, True, 32).tolist() == [[0], [0]]
    assert radius_graph_ref(torch.zeros(0, 3), 5.0, torch.zeros(0, dtype=torch.long), False, 32).shape == (2, 0)
    with pytest.raises(ValueError):
        radius_graph_ref(torch.zeros(3, 3), 5.0, torch.tensor([1, 0, 1]), False, 32)


def test_flow_and_golden():
    b = syn.make_batch(1, 2, 10, seed=9)
    a = radius_graph_ref(b.pos, 4.0, b.batch)
    t = radius_graph_ref(b.pos, 4.0, b.batch, flow="target_to_source")
    assert torch.equal(a.flip(0), t)
    for name, case in load_golden("radius_oracle.pt").items():
        ei = radius_graph_ref(case["pos"], case["r"], case["batch"], case["loop"], 32)
        assert torch.equal(ei.to(torch.int32), case["edge_index"]), name


def test_matches_kdtree_set_without_truncation():
    # without truncation the rule is just "all pairs with d < r": cross-check with scipy's KD-tree (the CPU
    # torch-cluster path is a KD-tree too; SURVEY.md A.1: both paths return the same SET when nothing truncates)
    from scipy.spatial import cKDTree

    b = syn.make_batch(1, 1, 30, seed=13)
    r = 3.0
    ei = radius_graph_ref(b.pos, r, b.batch, False, 32)
    assert torch.bincount(ei[1]).max() < 32
    tree = cKDTree(b.pos.double().numpy())
    pairs = tree.query_pairs(r - 1e-9)
    want = {(i, j) for i, j in pairs} | {(j, i) for i, j in pairs}
    got = set(map(tuple, ei.t().tolist()))
    # points closer than 1e-6 to the boundary could differ between fp32/fp64; none here
    d = (b.pos[:, None] - b.pos[None]).norm(dim=-1)
    assert ((d - r).abs() > 1e-5).all()
    assert got == want

"""Oracle of the conformer-level glue (CPU): known answers and the reference's call-site contracts."""
import torch

from oracle import aggregation as oag


def test_aggregation_index_is_floor_division_by_k():
    # common.py:414-423 appends whole groups of K: 10 graphs, K = 5 -> [0]*5 + [1]*5
    assert oag.create_aggregation_index(10, 5).tolist() == [0] * 5 + [1] * 5
    assert oag.create_aggregation_index(0, 3).numel() == 0
    # a ragged tail still gets a full group (the reference's loop does not stop inside a molecule)
    assert oag.create_aggregation_index(7, 3).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2]


def test_mean_aggregation_known_answers():
    x = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0], [7.0, 8.0]])
    out = oag.mean_aggregation(x, torch.tensor([0, 0, 1, 1]))
    assert torch.equal(out, torch.tensor([[2.0, 3.0], [6.0, 7.0]]))
    # y_true = mean of K identical labels = the label (common.py:410)
    y = torch.tensor([[0.3]] * 5 + [[-1.2]] * 5)
    assert torch.allclose(oag.mean_aggregation(y, oag.create_aggregation_index(10, 5)), torch.tensor([[0.3], [-1.2]]))
    # empty segment -> zeros (PyG scatter-mean semantics)
    assert torch.equal(oag.mean_aggregation(x[:2], torch.tensor([0, 2]), dim_size=3)[1], torch.zeros(2))


def test_head_is_linear_in_its_branches_and_state_dict_names_follow_the_reference():
    torch.manual_seed(0)
    h = oag.ConformerAggregationHead(8, use_covalent=True, use_barycenter=True, agg_weight=0.2)
    names = set(h.state_dict())
    assert {"transformation_matrix_3d.weight", "transformation_matrix_cov.bias", "transformation_matrix_bary.weight",
            "molecular_regression_lin.weight", "molecular_regression_lin.bias"} <= names
    idx = oag.create_aggregation_index(6, 3)
    a, b, c = torch.randn(6, 8), torch.randn(6, 8), torch.randn(6, 8)
    full = h(a, idx, b, c)
    # by hand: mean over conformers of (W3d a + Wcov b + 0.2 Wbary c), then the regression linear
    x = h.transformation_matrix_3d(a) + h.transformation_matrix_cov(b) + 0.2 * h.transformation_matrix_bary(c)
    ref = h.molecular_regression_lin(x.view(2, 3, 8).mean(1))
    assert torch.allclose(full, ref, atol=1e-6)
    assert full.shape == (2, 1)
    # classification head with the deeper MLP (build_mlp_class, is_complex)
    hc = oag.ConformerAggregationHead(8, use_covalent=False, classification=True, is_complex=True)
    assert hc(a, idx).shape == (2, 1)
    assert [type(m).__name__ for m in hc.molecular_regression_lin] == ["Linear", "ReLU", "Linear", "ReLU", "Linear"]


def test_dense_views_known_answers():
    from oracle import dense as od

    x = torch.arange(10.0).view(5, 2)
    batch = torch.tensor([0, 0, 1, 1, 1])
    out, mask = od.to_dense_batch(x, batch)
    assert out.shape == (2, 3, 2) and mask.tolist() == [[True, True, False], [True, True, True]]
    assert torch.equal(out[0, 2], torch.zeros(2)) and torch.equal(out[1, 2], x[4])
    ei = torch.tensor([[0, 1, 2, 4, 4], [1, 0, 3, 2, 2]])
    adj = od.to_dense_adj(ei, batch)
    assert adj.shape == (2, 3, 3)
    assert adj[0, 0, 1] == 1 and adj[0, 1, 0] == 1 and adj[1, 0, 1] == 1 and adj[1, 2, 0] == 2   # duplicates add
    assert adj.sum() == 5

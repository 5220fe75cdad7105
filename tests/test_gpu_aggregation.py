"""Conformer-level glue on the GPU against its oracle, through the C ABI."""
import pytest
import torch

import conan_fgw_b200 as cmp
from oracle import aggregation as oag
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_aggregation_index_matches_the_reference_loop():
    for G, K in [(10, 5), (640, 5), (7, 3), (0, 4), (1, 1)]:
        got = cmp.create_aggregation_index(G, K, DEV)
        assert got.dtype == torch.int64 and got.is_cuda
        assert torch.equal(got.cpu(), oag.create_aggregation_index(G, K))


def test_mean_aggregation_matches_oracle_and_backpropagates():
    torch.manual_seed(1)
    x = torch.randn(35, 64)
    index = torch.tensor([0] * 5 + [1] * 10 + [2] * 1 + [4] * 19)       # ragged, one empty segment
    ref_in = x.clone().requires_grad_(True)
    ref = oag.mean_aggregation(ref_in, index, dim_size=5)
    ref.pow(2).sum().backward()
    xin = x.to(DEV).requires_grad_(True)
    out = cmp.MeanAggregation()(xin, index.to(DEV), dim_size=5)
    out.pow(2).sum().backward()
    assert rel_err(out, ref) < 1e-6 and rel_err(xin.grad, ref_in.grad) < 1e-6
    with pytest.raises(ValueError):
        cmp.MeanAggregation()(xin, index[:-1].to(DEV))


@pytest.mark.parametrize("cov,bary,cls,cplx", [(True, True, False, False), (True, False, False, True),
                                                (False, False, False, False), (True, True, True, True)])
def test_head_matches_oracle(cov, bary, cls, cplx):
    torch.manual_seed(2)
    C, B, K = 64, 12, 5
    o = oag.ConformerAggregationHead(C, cov, bary, 0.2, cls, cplx).eval()
    c = cmp.ConformerAggregationHead(C, cov, bary, 0.2, cls, cplx).to(DEV).eval()
    c.load_state_dict(o.state_dict(), strict=True)
    x3, xc, xb = torch.randn(B * K, C), torch.randn(B * K, C), torch.randn(B * K, C)
    idx = oag.create_aggregation_index(B * K, K)
    ref = o(x3, idx, xc if cov else None, xb if bary else None)
    ref.pow(2).mean().backward()
    out = c(x3.to(DEV), cmp.create_aggregation_index(B * K, K, DEV), xc.to(DEV) if cov else None,
            xb.to(DEV) if bary else None, num_molecules=B)
    out.pow(2).mean().backward()
    assert out.shape == (B, 1) and rel_err(out, ref) < 1e-5
    for (n, po), (_, pc) in zip(o.named_parameters(), c.named_parameters()):
        assert rel_err(pc.grad, po.grad) < 1e-5, n


def test_dense_batch_and_adjacency_match_oracle():
    """to_dense_batch / to_dense_adj (row f-2) on ragged conformers, from the library's own radius graph."""
    from oracle import dense as od
    from conan_fgw_b200 import utils as cu

    torch.manual_seed(3)
    sizes = [5, 9, 1, 7]
    batch = torch.cat([torch.full((n,), g, dtype=torch.long) for g, n in enumerate(sizes)])
    pos = torch.randn(batch.numel(), 3) * 1.5
    x = torch.randn(batch.numel(), 6)
    ei = cmp.radius_graph(pos.to(DEV), 2.5, batch.to(DEV), max_num_neighbors=4)
    out, mask = cu.to_dense_batch(x.to(DEV), batch.to(DEV))
    ro, rm = od.to_dense_batch(x, batch)
    assert torch.equal(out.cpu(), ro) and torch.equal(mask.cpu(), rm)
    adj = cu.to_dense_adj(ei, batch.to(DEV))
    assert torch.equal(adj.cpu(), od.to_dense_adj(ei.cpu(), batch))
    assert adj.sum().item() == ei.size(1)
    # explicit sizes (no host sync), fill value, duplicates, and the single-graph form
    out2, mask2 = cu.to_dense_batch(x.to(DEV), batch.to(DEV), fill_value=-1.0, max_num_nodes=12, batch_size=4)
    ro2, rm2 = od.to_dense_batch(x, batch, -1.0, 12, 4)
    assert torch.equal(out2.cpu(), ro2) and torch.equal(mask2.cpu(), rm2)
    dup = torch.tensor([[0, 0, 1], [1, 1, 0]])
    assert torch.equal(cu.to_dense_adj(dup.to(DEV)).cpu(), od.to_dense_adj(dup))
    with pytest.raises(cmp._lib.ConanMPError):
        cu.to_dense_batch(x, batch)

"""Fused regression head (conformer mean -> Linear(C, 1) -> MSE, ``cmp_regression_head_fwd/bwd``) against the same three
ops in plain PyTorch fp32 on the GPU (schnet_based_models.py:17-29,242; model/common.py:288)."""
import pytest
import torch

import conan_fgw_b200 as cmp
from conan_fgw_b200 import ops
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("B,K,C,bias", [(128, 5, 64, True), (7, 3, 64, True), (1, 1, 16, False), (300, 10, 256, True)])
def test_fused_head_matches_the_three_separate_ops(B, K, C, bias):
    torch.manual_seed(B + K + C)
    emb = torch.randn(B * K, C, device=DEV)
    w = (0.1 * torch.randn(1, C, device=DEV)).requires_grad_(True)
    b = torch.randn(1, device=DEV).requires_grad_(True) if bias else None
    t = torch.randn(B, 1, device=DEV)
    e1 = emb.clone().requires_grad_(True)
    want = torch.nn.functional.mse_loss(torch.nn.functional.linear(e1.view(B, K, C).mean(dim=1), w, b), t)
    gw = torch.autograd.grad(3.0 * want, [e1, w] + ([b] if bias else []))
    e2 = emb.clone().requires_grad_(True)
    got = ops.regression_head_loss(e2, w, b, t, K)
    gg = torch.autograd.grad(3.0 * got, [e2, w] + ([b] if bias else []))      # upstream gradient != 1
    assert got.shape == want.shape
    assert abs(float(got.detach()) - float(want.detach())) <= 2e-6 * abs(float(want.detach()))
    for a, r in zip(gg, gw):
        assert a.shape == r.shape
        assert rel_err(a, r) < 2e-6
    # deterministic
    got2 = ops.regression_head_loss(e2, w, b, t, K)
    assert torch.equal(got, got2)


def test_training_step_with_and_without_the_fused_head(monkeypatch):
    from conan_fgw_b200 import dp
    syn = cmp.synthetic
    b = syn.make_batch(6, 3, 20, seed=3).to(DEV)
    targets = torch.randn(6, 1, device=DEV)
    losses = {}
    for fused in (True, False):
        monkeypatch.setattr(dp, "FUSED_HEAD", fused)
        torch.manual_seed(0)
        m = cmp.SchNetNoSum(None, num_interactions=2).to(DEV)
        tr = dp.RegressionStep(m, 64, 3, lr=1e-3)
        losses[fused] = [float(tr.step(b.z, b.pos, b.batch, targets, b.num_graphs)) for _ in range(3)]
    for a, r in zip(losses[True], losses[False]):
        assert abs(a - r) <= 1e-5 * abs(r)

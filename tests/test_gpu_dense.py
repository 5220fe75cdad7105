"""Dense building blocks against fp64 PyTorch (GPU, through the C ABI)."""
import math

import pytest
import torch

import conan_fgw_b200 as cmp
from conan_fgw_b200 import ops
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ssp64(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (77, 50, 33), (300, 128, 128), (4160, 128, 50), (129, 65, 17),
                                   (64, 64, 5000), (128, 50, 20000)])
def test_gemm_variants(M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = torch.randn(N, K, generator=g).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = a.double() @ w.double().t()
    assert rel_err(ops.gemm(a, w, False, True), ref) < 2e-6
    assert rel_err(ops.gemm(a, w, False, True, bias=bias, act=cmp._lib.ACT_SSP), ssp64(ref + bias.double())) < 2e-6
    assert rel_err(ops.gemm(a, w, False, True, bias=bias, residual=res), ref + bias.double() + res.double()) < 2e-6
    # NN: a[M,K] @ b[K,N]
    b = w.t().contiguous()
    assert rel_err(ops.gemm(a, b, False, False), ref) < 2e-6
    # TN: at stored [K,M]
    at = a.t().contiguous()
    assert rel_err(ops.gemm(at, b, True, False), ref) < 2e-6


def test_gemm_is_deterministic_with_split_k():
    a = torch.randn(128, 30000, device=DEV)
    b = torch.randn(30000, 128, device=DEV)
    assert torch.equal(ops.gemm(a, b, False, False), ops.gemm(a, b, False, False))


def test_linear_autograd_matches_torch():
    torch.manual_seed(0)
    x = torch.randn(333, 40, device=DEV, requires_grad=True)
    w = torch.randn(24, 40, device=DEV, requires_grad=True)
    bias = torch.randn(24, device=DEV, requires_grad=True)
    res = torch.randn(333, 24, device=DEV, requires_grad=True)
    for act, use_res in ((cmp._lib.ACT_NONE, True), (cmp._lib.ACT_SSP, False), (cmp._lib.ACT_NONE, False)):
        y = ops.linear(x, w, bias, act, res if use_res else None)
        yr = x.double() @ w.double().t() + bias.double()
        if act == cmp._lib.ACT_SSP:
            yr = ssp64(yr)
        if use_res:
            yr = yr + res.double()
        assert rel_err(y, yr) < 2e-6
        go = torch.randn_like(y)
        grads = torch.autograd.grad(y, [x, w, bias] + ([res] if use_res else []), go)
        grads_r = torch.autograd.grad(yr, [x, w, bias] + ([res] if use_res else []), go.double())
        for a, b in zip(grads, grads_r):
            assert rel_err(a, b) < 5e-6


def test_colsum_act_adam():
    x = torch.randn(5000, 70, device=DEV)
    assert rel_err(ops.colsum(x), x.double().sum(0)) < 1e-6
    assert torch.equal(ops.colsum(x), ops.colsum(x))
    t = torch.linspace(-30, 30, 1001, device=DEV, requires_grad=True)
    y = ops.shifted_softplus(t)
    assert rel_err(y, ssp64(t.double())) < 1e-6
    (gy,) = torch.autograd.grad(y.sum(), t)
    assert (gy - torch.sigmoid(t.double())).abs().max() < 2e-7
    s = ops.silu(t)
    assert rel_err(s, torch.nn.functional.silu(t.double())) < 1e-6
    (gs,) = torch.autograd.grad(s.sum(), t)
    (gs_r,) = torch.autograd.grad(torch.nn.functional.silu(t.double()).sum(), t)
    assert rel_err(gs, gs_r) < 2e-6
    # Adam against torch.optim.Adam
    p = torch.randn(1000, device=DEV)
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-2)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(1000, device=DEV)
        p_ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, step, lr=1e-2)
    assert rel_err(p, p_ref) < 1e-6

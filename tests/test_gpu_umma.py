"""Pins the tcgen05 conventions the fused kernels rely on: SWIZZLE_NONE canonical layouts (K-major and
MN-major), LBO/SBO meaning, instruction descriptor, TMEM lane/column mapping (GPU, C ABI probe)."""
import numpy as np
import pytest
import torch

import conan_fgw_b200 as cmp
from conan_fgw_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def image_kmajor(mat, lbo, sbo):
    """[R, K] bf16 -> byte image, byte(r,k) = (r%8)*16 + (k%8)*2 + (r//8)*sbo + (k//8)*lbo."""
    R, K = mat.shape
    r = torch.arange(R)[:, None]
    k = torch.arange(K)[None, :]
    off = (r % 8) * 16 + (k % 8) * 2 + (r // 8) * sbo + (k // 8) * lbo
    size = int(off.max()) + 2
    size = (size + 15) // 16 * 16
    img = torch.zeros(size // 2, dtype=torch.int16)
    img[(off // 2).reshape(-1)] = mat.to(torch.bfloat16).view(torch.int16).reshape(-1)
    return img.view(torch.uint8)


def image_mnmajor(mat, lbo, sbo):
    """[K, R] bf16 -> byte image, byte(k,r) = (r%8)*2 + (k%8)*16 + (r//8)*sbo + (k//8)*lbo."""
    K, R = mat.shape
    k = torch.arange(K)[:, None]
    r = torch.arange(R)[None, :]
    off = (r % 8) * 2 + (k % 8) * 16 + (r // 8) * sbo + (k // 8) * lbo
    size = (int(off.max()) + 2 + 15) // 16 * 16
    img = torch.zeros(size // 2, dtype=torch.int16)
    img[(off // 2).reshape(-1)] = mat.to(torch.bfloat16).view(torch.int16).reshape(-1)
    return img.view(torch.uint8)


def probe(a_img, b_img, N, K, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, kstep=256):
    a_img, b_img = a_img.to(DEV), b_img.to(DEV)
    D = torch.full((128, N), float("nan"), device=DEV)
    _lib.call("cmp_debug_umma_gemm", _lib.ptr(a_img), a_img.numel(), _lib.ptr(b_img), b_img.numel(), _lib.ptr(D), N, K,
              1, a_mn, b_mn, a_lbo, a_sbo, kstep, b_lbo, b_sbo, kstep)
    torch.cuda.synchronize()
    return D.cpu()


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("N,K", [(128, 64), (112, 64), (16, 16), (256, 128)])
def test_kmajor_a_kmajor_b(N, K):
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).bfloat16().float()
    B = torch.randn(N, K, generator=g).bfloat16().float()
    want = A @ B.t()
    sbo = (K // 8) * 128
    got = probe(image_kmajor(A, 128, sbo), image_kmajor(B, 128, sbo), N, K, 0, 0, 128, sbo, 128, sbo)
    msg = ""
    if not rel(got, want) < 1e-5:
        alt = probe(image_kmajor(A, 128, sbo), image_kmajor(B, 128, sbo), N, K, 0, 0, sbo, 128, sbo, 128)
        msg = f"documented LBO/SBO convention fails (rel {rel(got, want):.3e}); swapped gives {rel(alt, want):.3e}"
    assert rel(got, want) < 1e-5, msg


@pytest.mark.parametrize("N,K", [(128, 144), (96, 144), (128, 128), (32, 16)])
def test_kmajor_a_mnmajor_b(N, K):
    g = torch.Generator().manual_seed(N * 3 + K)
    A = torch.randn(128, K, generator=g).bfloat16().float()
    Bt = torch.randn(K, N, generator=g).bfloat16().float()        # stored [K, N]
    want = A @ Bt
    sbo = (K // 8) * 128
    got = probe(image_kmajor(A, 128, sbo), image_mnmajor(Bt, 128, sbo), N, K, 0, 1, 128, sbo, 128, sbo)
    msg = ""
    if not rel(got, want) < 1e-5:
        alt = probe(image_kmajor(A, 128, sbo), image_mnmajor(Bt, 128, sbo), N, K, 0, 1, 128, sbo, sbo, 128)
        msg = f"MN-major: documented convention fails (rel {rel(got, want):.3e}); swapped LBO/SBO gives {rel(alt, want):.3e}"
    assert rel(got, want) < 1e-5, msg


def test_same_bytes_serve_as_transposed_operand():
    """A K-major [rows=a, K=b] image equals an MN-major [K=a, rows=b] image with LBO and SBO exchanged -
    the property the backward kernels use to transpose for free."""
    g = torch.Generator().manual_seed(5)
    X = torch.randn(128, 128, generator=g).bfloat16().float()      # [f, e]
    Y = torch.randn(128, 128, generator=g).bfloat16().float()      # [k, e]
    # dW[f, k] = sum_e X[f, e] Y[k, e]: A = X K-major (K = e), B = Y K-major (K = e)
    sbo = (128 // 8) * 128
    x_img = image_kmajor(X, 128, sbo)
    y_img = image_kmajor(Y, 128, sbo)
    got = probe(x_img, y_img, 128, 128, 0, 0, 128, sbo, 128, sbo)
    assert rel(got, X @ Y.t()) < 1e-5
    # the same y_img read as an MN-major operand [K = k, N = e] (LBO <-> SBO): D[m, e] = sum_k A[m, k] Y[k, e]
    A = torch.randn(128, 128, generator=g).bfloat16().float()
    # stepping K by 16 rows of k in that image moves by 2 * (8-row group stride) = 2 * sbo bytes: use kstep accordingly
    a_img, b_img = image_kmajor(A, 128, sbo).to(DEV), y_img.to(DEV)
    D = torch.zeros(128, 128, device=DEV)
    _lib.call("cmp_debug_umma_gemm", _lib.ptr(a_img), a_img.numel(), _lib.ptr(b_img), b_img.numel(), _lib.ptr(D), 128,
              128, 1, 0, 1, 128, sbo, 256, sbo, 128, 2 * sbo)
    torch.cuda.synchronize()
    assert rel(D.cpu(), A @ Y) < 1e-5


def test_mnmajor_a_and_b():
    """Both operands as MN-major views of K-major [rows = atoms, K = channel] images (weight-gradient GEMMs)."""
    g = torch.Generator().manual_seed(9)
    Y = torch.randn(128, 64, generator=g).bfloat16().float()     # [atoms, Nout]
    X = torch.randn(128, 144, generator=g).bfloat16().float()    # [atoms, K + 16]
    sy, sx = (64 // 8) * 128, (144 // 8) * 128
    y_img = torch.cat([image_kmajor(Y, 128, sy), torch.zeros(2048, dtype=torch.uint8)])   # slack: M = 128 view
    x_img = image_kmajor(X, 128, sx)
    a_img, b_img = y_img.to(DEV), x_img.to(DEV)
    D = torch.zeros(128, 144, device=DEV)
    # D[m = out channel, n = in channel] = sum_atoms Y[a, m] X[a, n]; K step (16 atoms) = 2 atom groups
    _lib.call("cmp_debug_umma_gemm", _lib.ptr(a_img), a_img.numel(), _lib.ptr(b_img), b_img.numel(), _lib.ptr(D), 144,
              128, 1, 1, 1, sy, 128, 2 * sy, sx, 128, 2 * sx)
    torch.cuda.synchronize()
    want = Y.t() @ X
    assert rel(D.cpu()[:64], want) < 1e-5

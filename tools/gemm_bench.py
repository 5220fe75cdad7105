"""Exact-fp32 SIMT GEMM (cmp_gemm_f32) at the node-linear shapes of cfg 2: time per call and difference from an fp64
product.  CMP_GEMM_SCALAR=1 forces the scalar-load kernel (the results must be bit-identical).
Run on the GPU box: python tools/gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conan_fgw_b200 import ops

DEV = "cuda"
torch.manual_seed(0)
M, H = 17280, 128
x = torch.randn(M, H, device=DEV)
dy = torch.randn(M, H, device=DEV)
w = torch.randn(H, H, device=DEV) * 0.1
bias = torch.randn(H, device=DEV)
cases = [
    ("fwd  Y = X W^T + b   [17280,128]x[128,128]^T", lambda: ops.gemm(x, w, False, True, bias=bias), lambda: x.double() @ w.double().t() + bias.double()),
    ("dX   = dY W          [17280,128]x[128,128]  ", lambda: ops.gemm(dy, w, False, False), lambda: dy.double() @ w.double()),
    ("dW   = dY^T X        [128,17280]x[17280,128]", lambda: ops.gemm(dy, x, True, False), lambda: dy.double().t() @ x.double()),
    ("head Y = X W^T       [17280,128]x[64,128]^T ", lambda: ops.gemm(x, w[:64], False, True), lambda: x.double() @ w[:64].double().t()),
]
# the regression head of dp.RegressionStep: [128 molecules, 64] -> [128, 1] and its two gradient GEMMs
xm = torch.randn(128, 64, device=DEV)
wh = torch.randn(1, 64, device=DEV) * 0.1
dyh = torch.randn(128, 1, device=DEV)
cases += [
    ("head fwd  [128,64]x[1,64]^T                  ", lambda: ops.gemm(xm, wh, False, True), lambda: xm.double() @ wh.double().t()),
    ("head dX   [128,1]x[1,64]                     ", lambda: ops.gemm(dyh, wh, False, False), lambda: dyh.double() @ wh.double()),
    ("head dW   [128,1]^T x [128,64]               ", lambda: ops.gemm(dyh, xm, True, False), lambda: dyh.double().t() @ xm.double()),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, fn, ref in cases:
    out = fn()
    want = ref()
    err = float((out.double() - want).abs().max() / want.abs().max())
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    flops = 2.0 * out.numel() * (x.shape[0] if name.startswith("dW") else H)
    print(f"{name}: median {ts[5]:7.1f} us  ({flops / ts[5] * 1e-6:5.1f} TFLOP/s)  rel err vs fp64 {err:.1e}  checksum {float(out.double().sum()):.10e}")

"""Autograd wrappers around the C ABI (one ``torch.autograd.Function`` per kernel family).

PyTorch owns every tensor; the wrappers only pass device pointers and sizes on
the current stream.  Nothing here falls back to a PyTorch implementation.
"""

from __future__ import annotations

import ctypes
import threading

import torch
from torch.autograd import Function

from . import _lib
from ._lib import ACT_NONE, ACT_SILU, ACT_SSP, call, ptr


def _f32c(t):
    if t.dtype != torch.float32:
        raise _lib.ConanMPError(f"conanmp ops compute in float32, got {t.dtype}")
    return t.contiguous()


# ------------------------------------------------------------------------------------------
# dense
# ------------------------------------------------------------------------------------------

def gemm(a, b, trans_a=False, trans_b=True, bias=None, act=ACT_NONE, residual=None, out=None):
    """``out[M,N] = act(op(a) @ op(b) + bias) + residual`` through ``cmp_gemm_f32``."""
    a, b = _f32c(a), _f32c(b)
    if trans_a:
        K, M = a.shape
    else:
        M, K = a.shape
    if trans_b:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    if K != Kb:
        raise ValueError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    if bias is not None:
        bias = _f32c(bias)
    if residual is not None:
        residual = _f32c(residual)
    ws_bytes = _lib.size_query("cmp_gemm_workspace", M, N, K)
    ws = _lib.workspace(ws_bytes, a.device) if ws_bytes else None
    call("cmp_gemm_f32", int(trans_a), int(trans_b), M, N, K, ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out),
         out.stride(0), ptr(bias), int(act), ptr(residual), residual.stride(0) if residual is not None else 0,
         ptr(ws), ws.numel() if ws is not None else 0, work=2.0 * M * N * K)
    return out


def colsum(x):
    x = _f32c(x)
    M, N = x.shape
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    ws_bytes = _lib.size_query("cmp_colsum_workspace", M, N)
    ws = _lib.workspace(ws_bytes, x.device)
    call("cmp_colsum_f32", ptr(x), x.stride(0), M, N, ptr(out), ptr(ws), ws.numel())
    return out


def act_bwd(dy, saved, act):
    dy = _f32c(dy)
    dx = torch.empty_like(dy)
    call("cmp_act_bwd", ptr(dy), ptr(_f32c(saved)) if saved is not None else None, ptr(dx), dy.numel(), int(act))
    return dx


class _LinearFn(Function):
    """y = act(x W^T + b) [+ residual]; replaces torch.nn.Linear (+ ShiftedSoftplus) on the path."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, residual):
        if act != ACT_NONE and residual is not None:
            raise ValueError("linear: an activation and a residual cannot be fused in the same call")
        if act == ACT_SILU:
            raise ValueError("linear: SiLU needs the pre-activation saved; use linear(...) then silu(...)")
        lead = x.shape[:-1]
        x2 = _f32c(x.reshape(-1, x.shape[-1]))
        res2 = residual.reshape(-1, weight.shape[0]) if residual is not None else None
        y = gemm(x2, weight, False, True, bias, act, res2)
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.lead = lead
        ctx.save_for_backward(x2, weight, y if act != ACT_NONE else None)
        return y.reshape(*lead, weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight, y = ctx.saved_tensors
        dy2 = _f32c(dy.reshape(-1, weight.shape[0]))
        g = act_bwd(dy2, y, ctx.act) if ctx.act != ACT_NONE else dy2
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = gemm(g, weight, False, False).reshape(*ctx.lead, weight.shape[1])   # g[M,N] @ W[N,K]
        if ctx.needs_input_grad[1]:
            dw = gemm(g, x2, True, False)                                            # g^T[N,M] @ x[M,K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(g)
        if ctx.has_res and ctx.needs_input_grad[4]:
            dres = dy
        return dx, dw, db, None, dres


def linear(x, weight, bias=None, act=ACT_NONE, residual=None, tc=False):
    """``tc=True``: split-bf16 tcgen05 kernels (bf16 mode); else the exact-fp32 SIMT GEMM."""
    if tc and node_tc_supported(weight.shape[1], weight.shape[0]):
        return _LinearTCFn.apply(x, weight, bias, act, residual)
    if (tc and weight.shape[0] % 16 == 0 and weight.shape[1] % 16 == 0 and bool(_lib.lib().cmp_device_is_sm100())):
        if act == ACT_NONE:
            return _LinearTCBlockedFn.apply(x, weight, bias, residual)
        # wide layer with an activation (CFConv.lin2 of the H = 512 models): blocked GEMM, then the activation kernel
        y = _ActFn.apply(_LinearTCBlockedFn.apply(x, weight, bias, None), act)
        return y if residual is None else y + residual
    return _LinearFn.apply(x, weight, bias, act, residual)


def node_tc_supported(K, Nout) -> bool:
    return bool(_lib.lib().cmp_node_gemm_tc_supported(int(K), int(Nout))) and bool(_lib.lib().cmp_device_is_sm100())


# Open contexts (weight-image caches, deferred weight-gradient queues).  Process-wide on purpose: autograd runs
# ``Function.backward`` on its own device thread, so a thread-local registry would be invisible exactly where the queue is
# filled.  Isolation between models comes from ownership instead: weight images are keyed by the weight's storage
# address, and a deferred-gradient context only accepts the parameters it was opened for.
_registry = {"packs": [], "dw": []}
_registry_lock = threading.Lock()


def _stack(name):
    return _registry[name]


class prepacked_weights:
    """Pack every weight image the enclosed forward + backward will need in a few grouped launches (node linears,
    filter MLPs forward / dense / backward) instead of one or two small launches per layer.  Weights only change in
    the optimizer step, so a training step opens this context right after it (``dp.RegressionStep``); the images are
    dropped when the context closes, so a later weight update can never meet a stale image.  Contexts nest (inner ones
    are searched first); images are keyed by the weight's storage address, so two models never collide."""

    def __init__(self, modules, dense_only=False):
        self.modules = list(modules)
        self.dense_only = bool(dense_only)
        self.cache = None

    def __enter__(self):
        self.cache = _prepack(self.modules, self.dense_only, x3_hint=self.dense_only)
        with _registry_lock:
            _stack("packs").append(self)
        return self

    def __exit__(self, exc_type, exc, tb):
        with _registry_lock:
            _stack("packs").remove(self)
        self.cache = None
        return False


def _prepack(modules, dense_only=False, x3_hint=True):
    """``dense_only``: the caller vouches that no conformer exceeds the dense kernel's atom limit, so the weight image of
    the per-edge forward kernel is not needed."""
    from .nn import InteractionBlock, Linear    # late import: nn imports ops

    cache = {}
    node, filt, filt_x3 = [], [], []
    gmax = 32
    skip = set()     # filter-MLP Linears of fused blocks never run as node GEMMs
    for root in modules:
        for m in root.modules():
            if isinstance(m, InteractionBlock) and m.conv.precision == "fp32" and m.mlp[0].weight.is_cuda and x3_hint:
                # fp32-grade fused kernels: hi + lo images of the dense forward and of the weight-gradient kernel
                W1, b1, W2, b2 = m.mlp[0].weight, m.mlp[0].bias, m.mlp[2].weight, m.mlp[2].bias
                if fused_native(W1.shape[0], W1.shape[1]) and m.conv._standard_mlp():
                    skip.update((W1.data_ptr(), W2.data_ptr()))
                    if W1.data_ptr() not in cache:
                        dev = W1.device
                        px = torch.empty(_lib.size_query("cmp_cfconv_dense_x3_weights_bytes"), dtype=torch.uint8,
                                         device=dev)
                        pbx = torch.empty(_lib.size_query("cmp_cfconv_dense_bwd_x3_weights_bytes"), dtype=torch.uint8,
                                          device=dev)
                        cache[W1.data_ptr()] = (None, None, None, px, pbx)
                        filt_x3.append(tuple(_f32c(t.detach()) for t in (W1, b1, W2, b2)) + (px, pbx))
            if isinstance(m, InteractionBlock) and m.conv.precision == "bf16" and m.mlp[0].weight.is_cuda:
                W1, b1, W2, b2 = m.mlp[0].weight, m.mlp[0].bias, m.mlp[2].weight, m.mlp[2].bias
                if fused_native(W1.shape[0], W1.shape[1]):
                    skip.update((W1.data_ptr(), W2.data_ptr()))
                    if W1.data_ptr() not in cache:
                        dev = W1.device
                        pf = torch.empty(_lib.size_query("cmp_cfconv_tc_weights_bytes"), dtype=torch.uint8, device=dev)
                        pb = torch.empty(_lib.size_query("cmp_cfconv_tc_bwd_weights_bytes"), dtype=torch.uint8,
                                         device=dev)
                        pd = torch.empty(_lib.size_query("cmp_cfconv_dense_weights_bytes"), dtype=torch.uint8,
                                         device=dev)
                        # the per-edge forward image is only packed when conformers above the dense limit may occur
                        cache[W1.data_ptr()] = (None if (FUSED_DENSE and dense_only) else pf, pb, pd)
                        filt.append(tuple(_f32c(t.detach()) for t in (W1, b1, W2, b2)) + (pf, pb, pd))
    for root in modules:
        for m in root.modules():
            if isinstance(m, Linear) and m.tc and m.weight.is_cuda and m.weight.data_ptr() not in skip and \
                    m.weight.data_ptr() not in cache and \
                    bool(_lib.lib().cmp_node_gemm_tc_supported(m.weight.shape[1], m.weight.shape[0])):
                rows, cols = m.weight.shape
                n_norm = _lib.size_query("cmp_node_gemm_weight_bytes", cols)
                n_tr = _lib.size_query("cmp_node_gemm_weight_bytes", rows)
                packed = torch.empty(n_norm + n_tr, dtype=torch.uint8, device=m.weight.device)
                cache[m.weight.data_ptr()] = (packed[:n_norm], packed[n_norm:])
                node.append((_f32c(m.weight.detach()), rows, cols, packed))
            elif isinstance(m, Linear) and m.tc and m.weight.is_cuda and m.weight.data_ptr() not in skip and \
                    m.weight.shape[0] > 128 and m.weight.shape[1] <= 128 and m.weight.is_contiguous() and \
                    m.weight.dtype == torch.float32:
                # a layer wider than the kernel (ViSNet: H -> 2H / 3H): its 128-row blocks are contiguous [nb, K]
                # matrices, so every block gets the image of W_b (forward) and of W_b^T (dX) from the same grouped launch
                rows, cols = m.weight.shape
                w = m.weight.detach()
                for n0 in range(0, rows, 128):
                    nb = min(128, rows - n0)
                    blk = w[n0:n0 + nb]
                    if blk.data_ptr() in cache or not bool(_lib.lib().cmp_node_gemm_tc_supported(cols, nb)):
                        continue
                    n_norm = _lib.size_query("cmp_node_gemm_weight_bytes", cols)
                    n_tr = _lib.size_query("cmp_node_gemm_weight_bytes", nb)
                    packed = torch.empty(n_norm + n_tr, dtype=torch.uint8, device=w.device)
                    cache[blk.data_ptr()] = (packed[:n_norm], packed[n_norm:])
                    node.append((blk, nb, cols, packed))
    for lo in range(0, len(node), gmax):
        chunk = node[lo:lo + gmax]
        arr = (_lib.PackNodeJob * len(chunk))()
        for slot, (w, rows, cols, packed) in zip(arr, chunk):
            slot.W, slot.rows, slot.cols, slot.packed = w.data_ptr(), rows, cols, packed.data_ptr()
        call("cmp_node_gemm_pack_weights_grouped", ctypes.addressof(arr), len(chunk))
    for lo in range(0, len(filt), gmax):
        chunk = filt[lo:lo + gmax]
        F, Ng = chunk[0][0].shape
        if any(c[0].shape != (F, Ng) for c in chunk):
            raise _lib.ConanMPError("prepacked_weights: interaction blocks of different shapes in one model")
        arr = (_lib.PackFilterJob * len(chunk))()
        darr = (_lib.DensePackJob * len(chunk))()
        for slot, dslot, (W1, b1, W2, b2, pf, pb, pd) in zip(arr, darr, chunk):
            slot.W1, slot.b1, slot.W2, slot.b2 = W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr()
            slot.packed_fwd, slot.packed_bwd = pf.data_ptr(), pb.data_ptr()
            dslot.W1, dslot.b1, dslot.W2, dslot.b2 = slot.W1, slot.b1, slot.W2, slot.b2
            dslot.packed = pd.data_ptr()
        if not (FUSED_DENSE and dense_only):
            call("cmp_cfconv_tc_pack_weights_grouped", ctypes.addressof(arr), len(chunk), F, Ng)
        call("cmp_cfconv_tc_pack_bwd_weights_grouped", ctypes.addressof(arr), len(chunk), F, Ng)
        if FUSED_DENSE:
            call("cmp_cfconv_dense_pack_weights_grouped", ctypes.addressof(darr), len(chunk), F, Ng)
    for lo in range(0, len(filt_x3), gmax):
        chunk = filt_x3[lo:lo + gmax]
        F, Ng = chunk[0][0].shape
        if any(c[0].shape != (F, Ng) for c in chunk):
            raise _lib.ConanMPError("prepacked_weights: interaction blocks of different shapes in one model")
        darr = (_lib.DensePackJob * len(chunk))()
        barr = (_lib.BwdX3PackJob * len(chunk))()
        for dslot, bslot, (W1, b1, W2, b2, px, pbx) in zip(darr, barr, chunk):
            dslot.W1, dslot.b1, dslot.W2, dslot.b2 = W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr()
            dslot.packed = px.data_ptr()
            bslot.W1, bslot.b1, bslot.W2, bslot.packed = dslot.W1, dslot.b1, dslot.W2, pbx.data_ptr()
        call("cmp_cfconv_dense_x3_pack_weights_grouped", ctypes.addressof(darr), len(chunk), F, Ng)
        call("cmp_cfconv_dense_bwd_x3_pack_weights_grouped", ctypes.addressof(barr), len(chunk), F, Ng)
    cache["_keepalive"] = (node, filt, filt_x3)
    return cache


def _cached_images(weight):
    key = weight.data_ptr()
    with _registry_lock:
        open_ctx = list(_stack("packs"))
    for ctx in reversed(open_ctx):
        hit = ctx.cache.get(key) if ctx.cache is not None else None
        if hit is not None:
            return hit
    return None


def _pack_node_weight(weight, transpose):
    rows, cols = weight.shape
    image_k = rows if transpose else cols
    packed = torch.empty(_lib.size_query("cmp_node_gemm_weight_bytes", image_k), dtype=torch.uint8,
                         device=weight.device)
    call("cmp_node_gemm_pack_weight", ptr(_f32c(weight)), rows, cols, int(transpose), ptr(packed))
    return packed


def _pack_node_weight_both(weight):
    """(image of W, image of W^T) from ONE launch: the forward uses the first, the dX backward the second."""
    rows, cols = weight.shape
    n_norm = _lib.size_query("cmp_node_gemm_weight_bytes", cols)
    n_tr = _lib.size_query("cmp_node_gemm_weight_bytes", rows)
    packed = torch.empty(n_norm + n_tr, dtype=torch.uint8, device=weight.device)
    call("cmp_node_gemm_pack_weight", ptr(_f32c(weight)), rows, cols, 2, ptr(packed))
    return packed[:n_norm], packed[n_norm:]


def _node_gemm(x2, w_img, K, Nout, bias=None, act=ACT_NONE, residual=None, saved_y=None):
    M = x2.shape[0]
    y = torch.empty(M, Nout, dtype=torch.float32, device=x2.device)
    call("cmp_node_gemm_fwd", ptr(x2), x2.stride(0), ptr(saved_y), saved_y.stride(0) if saved_y is not None else 0,
         ptr(w_img), ptr(bias), int(act), ptr(residual), residual.stride(0) if residual is not None else 0, ptr(y),
         y.stride(0), M, K, Nout, work=2.0 * M * K * Nout)
    return y


def _tc_matmul_blocked(x2, W, bias=None, residual=None, saved_y=None):
    """y[M, Nout] = x2'[M, K] @ W[Nout, K]^T + bias + residual for any K, Nout that are multiples of 16, as a sequence
    of <=128 x <=128 blocks of the tcgen05 node kernel: column blocks of the output are independent launches, K blocks
    are chained through the kernel's residual input (no activation: callers apply it separately).  Row blocks of a
    contiguous W with K <= 128 take their image from the prepack cache when a training step opened one."""
    M, K = x2.shape
    Nout = W.shape[0]
    y = torch.empty(M, Nout, dtype=torch.float32, device=x2.device)
    f4 = 4
    for n0 in range(0, Nout, 128):
        nb = min(128, Nout - n0)
        for bi, k0 in enumerate(range(0, K, 128)):
            kb = min(128, K - k0)
            blk = W[n0:n0 + nb, k0:k0 + kb]
            hit = _cached_images(blk) if (K <= 128 and blk.is_contiguous()) else None
            img = hit[0] if hit is not None else _pack_node_weight(blk.contiguous(), False)
            first = bi == 0
            if first:
                res_ptr = residual.data_ptr() + n0 * f4 if residual is not None else None
                ldr = residual.stride(0) if residual is not None else 0
            else:
                res_ptr, ldr = y.data_ptr() + n0 * f4, Nout
            b_ptr = bias.data_ptr() + n0 * f4 if (bias is not None and first) else None
            sy_ptr = saved_y.data_ptr() + k0 * f4 if saved_y is not None else None
            call("cmp_node_gemm_fwd", x2.data_ptr() + k0 * f4, K, sy_ptr, K if saved_y is not None else 0, ptr(img),
                 b_ptr, ACT_NONE, res_ptr, ldr, y.data_ptr() + n0 * f4, Nout, M, kb, nb, work=2.0 * M * kb * nb)
    return y


def _tc_dx_blocked(dy2, W):
    """dX[M, K] = dY[M, Nout] @ W[Nout, K] for a wide layer: the row blocks W_b of W are the K blocks of this product,
    chained through the kernel's residual input; their W_b^T images come from the prepack cache when there is one
    (else the transposed weight is materialised and packed block by block as before)."""
    M, Nout = dy2.shape
    K = W.shape[1]
    if not (K <= 128 and W.is_contiguous() and _cached_images(W[0:min(128, Nout)]) is not None):
        return _tc_matmul_blocked(dy2, W.t().contiguous())
    dx = torch.empty(M, K, dtype=torch.float32, device=dy2.device)
    f4 = 4
    for bi, n0 in enumerate(range(0, Nout, 128)):
        nb = min(128, Nout - n0)
        hit = _cached_images(W[n0:n0 + nb])
        img = hit[1] if hit is not None else _pack_node_weight(W[n0:n0 + nb].contiguous(), True)
        res_ptr, ldr = (None, 0) if bi == 0 else (dx.data_ptr(), K)
        call("cmp_node_gemm_fwd", dy2.data_ptr() + n0 * f4, Nout, None, 0, ptr(img), None, ACT_NONE, res_ptr, ldr,
             dx.data_ptr(), K, M, nb, K, work=2.0 * M * nb * K)
    return dx


def _tc_dw_blocked(dy2, saved_y, x2, want_db, param=None):
    """dW[Nout, K] = dY'^T X (+ db = column sums of dY') in <=128 x <=128 blocks of cmp_node_gemm_dw."""
    M, Nout = dy2.shape
    K = x2.shape[1]
    dev = dy2.device
    dw = torch.empty(Nout, K, dtype=torch.float32, device=dev)
    db = torch.empty(Nout, dtype=torch.float32, device=dev) if want_db else None
    f4 = 4
    dwq = _dw_queue(param)
    if dwq is not None:
        # queued for the grouped launch: every <=128 x <=128 block writes straight into dw (leading dimension K)
        for n0 in range(0, Nout, 128):
            nb = min(128, Nout - n0)
            for k0 in range(0, K, 128):
                kb = min(128, K - k0)
                dwq.append(dict(
                    dY=dy2.data_ptr() + n0 * f4, lddy=Nout,
                    saved_y=saved_y.data_ptr() + n0 * f4 if saved_y is not None else None,
                    ldys=Nout if saved_y is not None else 0, X=x2.data_ptr() + k0 * f4, ldx=K, M=M, K=kb, Nout=nb,
                    dW=dw.data_ptr() + (n0 * K + k0) * f4, lddw=K,
                    db=db.data_ptr() + n0 * f4 if (want_db and k0 == 0) else None, keep=(dy2, saved_y, x2),
                    param=param, dW_base=dw.data_ptr(), first=(n0 == 0 and k0 == 0)))
        return dw, db
    for n0 in range(0, Nout, 128):
        nb = min(128, Nout - n0)
        for k0 in range(0, K, 128):
            kb = min(128, K - k0)
            blk = torch.empty(nb, kb, dtype=torch.float32, device=dev)
            ws = _lib.workspace(_lib.size_query("cmp_node_gemm_dw_workspace", kb), dev)
            db_ptr = db.data_ptr() + n0 * f4 if (want_db and k0 == 0) else None
            sy_ptr = saved_y.data_ptr() + n0 * f4 if saved_y is not None else None
            call("cmp_node_gemm_dw", dy2.data_ptr() + n0 * f4, Nout, sy_ptr, Nout if saved_y is not None else 0,
                 x2.data_ptr() + k0 * f4, K, M, kb, nb, ptr(blk), db_ptr, ptr(ws), ws.numel(), work=2.0 * M * kb * nb)
            if nb == Nout and kb == K:
                dw = blk
            else:
                dw[n0:n0 + nb, k0:k0 + kb].copy_(blk)
    return dw, db


class _LinearTCBlockedFn(Function):
    """linear(tc) for layers wider than 128 (ViSNet: H -> 2H / 3H, 2H -> H): blocked over the 128-wide kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        Nout, K = weight.shape
        lead = x.shape[:-1]
        x2 = _f32c(x.reshape(-1, K))
        res2 = _f32c(residual.reshape(-1, Nout)) if residual is not None else None
        y = _tc_matmul_blocked(x2, weight, _f32c(bias) if bias is not None else None, res2)
        ctx.lead, ctx.has_bias, ctx.has_res = lead, bias is not None, residual is not None
        ctx.save_for_backward(x2, weight)
        return y.reshape(*lead, Nout)

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        Nout, K = weight.shape
        dy2 = _f32c(dy.reshape(-1, Nout))
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = _tc_dx_blocked(dy2, weight).reshape(*ctx.lead, K)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = _tc_dw_blocked(dy2, None, x2, ctx.has_bias, param=weight if ctx.needs_input_grad[1] else None)
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = dy
        return dx, dw, db, dres


class deferred_weight_grads:
    """Inside this context the tensor-core Linears do not launch their weight-gradient GEMMs one by one during
    ``backward()``: they return (still unwritten) gradient tensors and queue the problem; leaving the context issues
    ALL of them as one grouped launch (``cmp_node_gemm_dw_grouped``) that fills the SMs instead of 20 small launches
    of ~17 us each.  Only valid when nothing reads or accumulates into the parameter gradients before the context
    exits, i.e. every parameter is used once, has no hooks and ``p.grad`` was ``None`` (``dp.RegressionStep``
    guarantees all three; the flush verifies that autograd adopted every weight AND bias gradient tensor).
    ``params``: the parameters this context owns (None = every parameter).  A Linear whose weight belongs to no open
    context launches its gradient GEMM immediately, so two models stepping at the same time never share a queue."""

    def __init__(self, params=None):
        self.queue = None
        self.owned = None if params is None else {p.data_ptr() for p in params}

    def __enter__(self):
        self.queue = []
        with _registry_lock:
            _stack("dw").append(self)
        return self

    def __exit__(self, exc_type, exc, tb):
        with _registry_lock:
            _stack("dw").remove(self)
        queue, self.queue = self.queue, None
        if exc_type is None:
            _flush_deferred_dw(queue)
        return False


def _dw_queue(weight=None):
    """The queue of the innermost open context that owns ``weight`` (called from autograd's backward thread)."""
    key = None if weight is None else weight.data_ptr()
    with _registry_lock:
        for ctx in reversed(_stack("dw")):
            if ctx.queue is not None and (ctx.owned is None or key is None or key in ctx.owned):
                return ctx.queue
    return None


def _flush_deferred_dw(queue):
    if not queue:
        return
    dev = queue[0]["keep"][0].device
    gmax = _lib.size_query("cmp_node_gemm_dw_group_max")
    ws = _lib.workspace(_lib.size_query("cmp_node_gemm_dw_grouped_workspace"), dev)
    for lo in range(0, len(queue), gmax):
        chunk = queue[lo:lo + gmax]
        arr = (_lib.DwProblem * len(chunk))()
        work = 0.0
        for slot, q in zip(arr, chunk):
            slot.dY, slot.lddy, slot.saved_y, slot.ldys = q["dY"], q["lddy"], q["saved_y"], q["ldys"]
            slot.X, slot.ldx, slot.M, slot.K, slot.Nout = q["X"], q["ldx"], q["M"], q["K"], q["Nout"]
            slot.dW, slot.lddw, slot.db = q["dW"], q["lddw"], q["db"]
            work += 2.0 * q["M"] * q["K"] * q["Nout"]
        call("cmp_node_gemm_dw_grouped", ctypes.addressof(arr), len(chunk), ptr(ws), ws.numel(), work=work)
    # autograd must have ADOPTED the tensors that were returned unwritten (it does when .grad was None and nothing else
    # references them); had it copied or accumulated them instead, the values written above would be lost: fail loudly
    for q in queue:
        for key, base in (("param", "dW_base"), ("bias_param", "db_base")):
            w = q.get(key)
            if w is not None and w.is_leaf and q.get("first", True):
                if w.grad is None or w.grad.data_ptr() != q[base]:
                    raise _lib.ConanMPError(
                        "deferred_weight_grads: a parameter gradient was copied or accumulated before the grouped launch "
                        "wrote it (use it only with p.grad = None, no hooks, and parameters that are used once)")


class _LinearTCFn(Function):
    """Same contract as _LinearFn on the tcgen05 split-bf16 node GEMM kernels (node_gemm_tc.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, residual):
        if act not in (ACT_NONE, ACT_SSP):
            raise ValueError("linear(tc): only NONE / SSP epilogues are provided")
        if act != ACT_NONE and residual is not None:
            raise ValueError("linear: an activation and a residual cannot be fused in the same call")
        Nout, K = weight.shape
        lead = x.shape[:-1]
        x2 = _f32c(x.reshape(-1, K))
        res2 = _f32c(residual.reshape(-1, Nout)) if residual is not None else None
        cached = _cached_images(weight)
        if cached is not None:
            w_img, ctx.w_img_t = cached
        elif ctx.needs_input_grad[0]:
            w_img, ctx.w_img_t = _pack_node_weight_both(weight)
        else:
            w_img, ctx.w_img_t = _pack_node_weight(weight, False), None
        y = _node_gemm(x2, w_img, K, Nout, _f32c(bias) if bias is not None else None, act, res2)
        ctx.act, ctx.lead = act, lead
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.bias_ref = bias          # the Parameter itself: the deferred-gradient flush checks that autograd adopted db
        ctx.save_for_backward(x2, weight, y if act != ACT_NONE else None)
        return y.reshape(*lead, Nout)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, y = ctx.saved_tensors
        Nout, K = weight.shape
        dy2 = _f32c(dy.reshape(-1, Nout))
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            # dX = (dY * ssp') W : the forward kernel with the transposed weight image
            w_t = ctx.w_img_t if ctx.w_img_t is not None else _pack_node_weight(weight, True)
            dx = _node_gemm(dy2, w_t, Nout, K, saved_y=y).reshape(*ctx.lead, K)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = _weight_grads(dy2, y, x2, weight, ctx.bias_ref if ctx.has_bias else None, ctx.needs_input_grad[1],
                                   ctx.needs_input_grad[2])
        if ctx.has_res and ctx.needs_input_grad[4]:
            dres = dy
        return dx, dw, db, None, dres



def _weight_grads(dy2, saved_y, x2, weight, bias, want_w=True, want_b=True):
    """dW = dY'^T X and db = column sums of dY' of one tensor-core Linear (dY' = dY * ssp'(saved_y) when given):
    queued for the grouped launch inside ``deferred_weight_grads`` (see there), launched immediately otherwise."""
    Nout, K = weight.shape
    M = x2.shape[0]
    dev = dy2.device
    dw = torch.empty(Nout, K, dtype=torch.float32, device=dev)
    db = torch.empty(Nout, dtype=torch.float32, device=dev) if bias is not None else None
    dwq = _dw_queue(weight)
    if dwq is not None:
        # queued: raw pointers only for the outputs (autograd must stay the sole owner of dw / db, or it would clone
        # them - unwritten - instead of adopting them as .grad); inputs are kept alive here
        dwq.append(dict(
            dY=dy2.data_ptr(), lddy=dy2.stride(0), saved_y=saved_y.data_ptr() if saved_y is not None else None,
            ldys=saved_y.stride(0) if saved_y is not None else 0, X=x2.data_ptr(), ldx=x2.stride(0), M=M, K=K, Nout=Nout,
            dW=dw.data_ptr(), lddw=K, db=db.data_ptr() if db is not None else None, keep=(dy2, saved_y, x2),
            param=weight if want_w else None, dW_base=dw.data_ptr(),
            bias_param=bias if (db is not None and want_b) else None, db_base=db.data_ptr() if db is not None else None))
    else:
        ws = _lib.workspace(_lib.size_query("cmp_node_gemm_dw_workspace", K), dev)
        call("cmp_node_gemm_dw", ptr(dy2), dy2.stride(0), ptr(saved_y), saved_y.stride(0) if saved_y is not None else 0,
             ptr(x2), x2.stride(0), M, K, Nout, ptr(dw), ptr(db), ptr(ws), ws.numel(), work=2.0 * M * K * Nout)
    return dw, db


def _node_images(weight):
    """(image of W, image of W^T) of a node Linear: from the prepack cache, else packed now."""
    cached = _cached_images(weight)
    return cached if cached is not None else _pack_node_weight_both(weight)


def _chain(x2, stages):
    """Run ``cmp_node_chain_fwd``.  stages: list of dicts {img, K, Nout, bias, act, residual, scale_y, out}."""
    M = x2.shape[0]
    arr = (_lib.ChainStage * len(stages))()
    work = 0.0
    for slot, st in zip(arr, stages):
        slot.w_img = st["img"].data_ptr()
        slot.bias = st["bias"].data_ptr() if st.get("bias") is not None else None
        r, y, o = st.get("residual"), st.get("scale_y"), st.get("out")
        slot.residual, slot.ldr = (r.data_ptr(), r.stride(0)) if r is not None else (None, 0)
        slot.scale_y, slot.lds = (y.data_ptr(), y.stride(0)) if y is not None else (None, 0)
        slot.out, slot.ldo = (o.data_ptr(), o.stride(0)) if o is not None else (None, 0)
        slot.K, slot.Nout, slot.act = st["K"], st["Nout"], int(st.get("act", ACT_NONE))
        work += 2.0 * M * st["K"] * st["Nout"]
    call("cmp_node_chain_fwd", ptr(x2), x2.stride(0), M, ctypes.addressof(arr), len(stages), work=work)


class _BlockTailFn(Function):
    """The node-level tail of an interaction block and the head of the next one in ONE kernel per direction
    (``cmp_node_chain_fwd``; PyG ``CFConv.lin2 -> ShiftedSoftplus -> InteractionBlock.lin``, the residual of
    ``sns.py:164`` and the next block's ``CFConv.lin1``):

        y = ssp(agg W2^T + b2);   h' = h + y Wl^T + bl;   x'' = h' W1n^T        (x'' only when W1n is given)

    backward: ``dx'' -> [W1n, + dh'] -> dh -> [Wl, * ssp'(y)] -> dpre -> [W2] -> dagg`` on the same kernel with the
    transposed weight images; the five weight / bias gradients go to the grouped weight-gradient launch."""

    @staticmethod
    def forward(ctx, agg, h, W2, b2, Wl, bl, W1n):
        agg, h = _f32c(agg), _f32c(h)
        N = agg.shape[0]
        H, F = W2.shape
        dev = agg.device
        img2, img2_t = _node_images(W2)
        imgl, imgl_t = _node_images(Wl)
        y = torch.empty(N, H, dtype=torch.float32, device=dev)
        hn = torch.empty(N, H, dtype=torch.float32, device=dev)
        stages = [dict(img=img2, K=F, Nout=H, bias=b2, act=ACT_SSP, out=y),
                  dict(img=imgl, K=H, Nout=H, bias=bl, residual=h, out=hn)]
        xn = None
        ctx.imgs_t = [img2_t, imgl_t, None]
        if W1n is not None:
            img1, img1_t = _node_images(W1n)
            xn = torch.empty(N, W1n.shape[0], dtype=torch.float32, device=dev)
            stages.append(dict(img=img1, K=H, Nout=W1n.shape[0], out=xn))
            ctx.imgs_t[2] = img1_t
        _chain(agg, stages)
        ctx.has_next = W1n is not None
        ctx.refs = (b2, bl)         # the Parameters themselves (bias-gradient adoption check of the deferred launch)
        ctx.save_for_backward(agg, y, hn, W2, Wl, W1n)
        if xn is None:
            return hn
        return hn, xn

    @staticmethod
    def backward(ctx, dhn, dxn=None):
        agg, y, hn, W2, Wl, W1n = ctx.saved_tensors
        b2, bl = ctx.refs
        N = agg.shape[0]
        H, F = W2.shape
        dev = agg.device
        img2_t, imgl_t, img1_t = ctx.imgs_t
        dpre = torch.empty(N, H, dtype=torch.float32, device=dev)
        dagg = torch.empty(N, F, dtype=torch.float32, device=dev)
        tail = [dict(img=imgl_t, K=H, Nout=H, scale_y=y, out=dpre), dict(img=img2_t, K=H, Nout=F, out=dagg)]
        if ctx.has_next and dxn is not None:
            dxn = _f32c(dxn)
            dh = torch.empty(N, H, dtype=torch.float32, device=dev)
            res = _f32c(dhn) if dhn is not None else None
            _chain(dxn, [dict(img=img1_t, K=W1n.shape[0], Nout=H, residual=res, out=dh)] + tail)
        else:
            dh = _f32c(dhn)
            _chain(dh, tail)
        dW2, db2 = _weight_grads(dpre, None, agg, W2, b2, ctx.needs_input_grad[2], ctx.needs_input_grad[3])
        dWl, dbl = _weight_grads(dh, None, y, Wl, bl, ctx.needs_input_grad[4], ctx.needs_input_grad[5])
        dW1n = None
        if ctx.has_next and dxn is not None:
            dW1n, _ = _weight_grads(dxn, None, hn, W1n, None, ctx.needs_input_grad[6], False)
        return dagg, dh, dW2, db2, dWl, dbl, dW1n


def block_tail(agg, h, lin2, lin, lin1_next=None):
    """``(h', x'')`` (or ``h'`` alone) from the aggregate of an interaction block: see ``_BlockTailFn``."""
    W1n = lin1_next.weight if lin1_next is not None else None
    return _BlockTailFn.apply(agg, h, lin2.weight, lin2.bias, lin.weight, lin.bias, W1n)


def block_tail_supported(lin2, lin, lin1_next=None) -> bool:
    mods = [lin2, lin] + ([lin1_next] if lin1_next is not None else [])
    return (all(m.tc and node_tc_supported(m.weight.shape[1], m.weight.shape[0]) for m in mods)
            and lin2.bias is not None and lin.bias is not None and (lin1_next is None or lin1_next.bias is None))


class _ActFn(Function):
    @staticmethod
    def forward(ctx, x, act):
        x = _f32c(x)
        y = torch.empty_like(x)
        call("cmp_act_fwd", ptr(x), ptr(y), x.numel(), int(act))
        ctx.act = act
        ctx.save_for_backward(y if act == ACT_SSP else x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (saved,) = ctx.saved_tensors
        return act_bwd(dy, saved, ctx.act), None


def shifted_softplus(x):
    return _ActFn.apply(x, ACT_SSP)


def silu(x):
    return _ActFn.apply(x, ACT_SILU)


# ------------------------------------------------------------------------------------------
# SchNet pieces
# ------------------------------------------------------------------------------------------

def gaussian_rbf(dist, offset, coeff):
    """GaussianSmearing.forward: ``exp(coeff * (d - offset_k)^2)`` -> ``[E, Ng]`` (no gradient: the
    reference never differentiates through positions, ``derivative=False``)."""
    d = _f32c(dist.detach().reshape(-1))
    E, Ng = d.numel(), offset.numel()
    out = torch.empty(E, Ng, dtype=torch.float32, device=d.device)
    call("cmp_rbf_gaussian_fwd", ptr(d), E, ptr(_f32c(offset)), Ng, float(coeff), ptr(out), Ng)
    return out


class _EmbeddingFn(Function):
    @staticmethod
    def forward(ctx, z, weight, padding_idx, status):
        z = z.to(torch.int64).contiguous()
        weight = _f32c(weight)
        N, (V, H) = z.numel(), weight.shape
        out = torch.empty(N, H, dtype=torch.float32, device=weight.device)
        call("cmp_embedding_fwd", ptr(z), N, ptr(weight), V, H, ptr(out), ptr(status))
        ctx.save_for_backward(z)
        ctx.shape = (V, H)
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        return out

    @staticmethod
    def backward(ctx, dout):
        (z,) = ctx.saved_tensors
        V, H = ctx.shape
        dout = _f32c(dout)
        dw = torch.empty(V, H, dtype=torch.float32, device=dout.device)
        ws_bytes = _lib.size_query("cmp_embedding_bwd_workspace", z.numel(), V, H)
        ws = _lib.workspace(ws_bytes, dout.device)
        call("cmp_embedding_bwd", ptr(z), z.numel(), ptr(dout), V, H, ctx.padding_idx, ptr(dw), ptr(ws), ws.numel())
        return None, dw, None, None


def embedding(z, weight, padding_idx=None, status=None):
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=weight.device)
    return _EmbeddingFn.apply(z, weight, padding_idx, status)


class _CFConvMessageFn(Function):
    """agg_i = sum_{j->i} x'_j * filt_ij * C(d_ij)  (PyG CFConv.propagate/message)."""

    @staticmethod
    def forward(ctx, xprime, filt, graph, cutoff):
        xprime, filt = _f32c(xprime), _f32c(filt)
        N, F = xprime.shape
        if filt.shape[0] != graph.E or filt.shape[1] != F:
            raise ValueError("cfconv: filter must be [E, num_filters] in CSR edge order")
        agg = torch.empty(N, F, dtype=torch.float32, device=xprime.device)
        call("cmp_cfconv_message_fwd", ptr(xprime), ptr(filt), ptr(graph.dist), ptr(graph.rowptr), ptr(graph.col),
             N, F, float(cutoff), ptr(agg))
        ctx.graph, ctx.cutoff = graph, float(cutoff)
        ctx.save_for_backward(xprime, filt)
        return agg

    @staticmethod
    def backward(ctx, g):
        xprime, filt = ctx.saved_tensors
        graph = ctx.graph
        g = _f32c(g)
        N, F = xprime.shape
        dx = torch.empty_like(xprime) if ctx.needs_input_grad[0] else None
        dfilt = torch.empty_like(filt) if ctx.needs_input_grad[1] else None
        if dx is not None and graph.rowptr_t is None:
            raise _lib.ConanMPError("cfconv backward needs the transposed neighbour list")
        call("cmp_cfconv_message_bwd", ptr(g), ptr(xprime), ptr(filt), ptr(graph.dist), ptr(graph.rowptr),
             ptr(graph.col), ptr(graph.rowptr_t), ptr(graph.col_t), ptr(graph.eid_t), N, F, ctx.cutoff,
             ptr(dfilt), ptr(dx))
        return dx, dfilt, None, None


def cfconv_message(xprime, filt, graph, cutoff):
    return _CFConvMessageFn.apply(xprime, filt, graph, cutoff)


def _fused_fwd_launch(xin, dist, rowptr, col, tiles, num_tiles, packed, offset, coeff, cutoff, n_edges_hint):
    N, F = xin.shape
    Ng = offset.numel()
    out = torch.empty(N, F, dtype=torch.float32, device=xin.device)
    # algorithmic FLOPs of this launch: 2*(Ng*F + F*F) per edge
    work = 2.0 * (Ng * F + F * F) * float(n_edges_hint)
    call("cmp_cfconv_fused_fwd", ptr(xin), ptr(dist), ptr(rowptr), ptr(col), ptr(tiles), ptr(num_tiles), ptr(packed),
         ptr(offset), Ng, float(coeff), float(cutoff), N, F, ptr(out), work=work)
    return out


# Forward / d x' aggregation kernel of the fused path:
#   FUSED_DENSE (default): dense-block kernel (cfconv_dense.cu) for conformers of <= 128 atoms, per-edge kernel for larger
#   FUSED_PAIR_FORWARD:    round-1 pair-list kernel for conformers of <= 30 atoms + per-edge kernel (cross-checking)
#   neither:               per-edge kernel for everything (cross-checking)
FUSED_DENSE = True
FUSED_PAIR_FORWARD = True


def _offset_host(offset):
    """Gaussian centres as a host float array (they become kernel parameters of the dense kernel).  Cached on the
    buffer tensor: the one device -> host copy happens on first use, never inside a captured region afterwards."""
    key = (offset.data_ptr(), offset._version, offset.numel())
    cached = getattr(offset, "_cmp_host", None)
    if cached is None or cached[0] != key:
        vals = offset.detach().to(torch.float32).cpu().tolist()
        cached = (key, (ctypes.c_float * len(vals))(*vals))
        offset._cmp_host = cached
    return cached[1]


def _per_edge_aggregate(xin, graph, packed, offset, coeff, cutoff, transposed, min_atoms, e_hint):
    if transposed:
        tiles, num, dist = graph.tiles_t(min_atoms)
        rowptr, col = graph.rowptr_t, graph.col_t
    else:
        tiles, num = graph.tiles(min_atoms)
        dist, rowptr, col = graph.dist, graph.rowptr, graph.col
    return _fused_fwd_launch(xin, dist, rowptr, col, tiles, num, packed, offset, coeff, cutoff, e_hint)


def x3_graph_ok(graph) -> bool:
    """The fp32-grade fused kernels serve radius-built graphs whose conformers all fit the dense kernels."""
    return (graph is not None and graph.G > 0 and graph.pos is not None and not graph.loop
            and graph.max_atoms is not None and graph.max_atoms <= _lib.size_query("cmp_cfconv_dense_max_atoms"))


def _fused_aggregate(xin, graph, W, offset, coeff, cutoff, transposed, x3=False):
    """agg = sum_j x_j * W(d_ij) C(d_ij) over the neighbour list (or its transpose) on the fused tcgen05 kernels.
    ``W`` = (W1, b1, W2, b2) of the filter MLP (weight images come from the prepack cache or are packed here).
    ``x3``: the fp32-grade kernel (f16 hi + lo operand images, three MMA passes, fp32 epilogues)."""
    e_hint = graph._E if graph._E is not None else graph.cap_E
    N, F = xin.shape
    Ng = offset.numel()
    flops = 2.0 * (Ng * F + F * F) * float(e_hint)
    if x3:
        if not x3_graph_ok(graph):
            raise _lib.ConanMPError("fp32-grade fused CFConv: needs a radius-built graph with max_atoms <= 128")
        out = torch.empty(N, F, dtype=torch.float32, device=xin.device)
        call("cmp_cfconv_dense_x3_fwd", ptr(xin), ptr(graph.pos), ptr(graph.seg_ptr), ptr(graph.adjacency()), graph.G,
             ptr(pack_dense_x3_weights(*W)), ctypes.addressof(_offset_host(offset)), Ng, float(coeff), float(cutoff), F,
             int(bool(transposed)), 0, ptr(out), ptr(graph._counter), ptr(graph.status), work=flops)
        return out
    if FUSED_DENSE and graph.G > 0 and graph.pos is not None and not graph.loop:
        cap = _lib.size_query("cmp_cfconv_dense_max_atoms")
        dense_only = graph.max_atoms is not None and graph.max_atoms <= cap
        if dense_only:
            out = torch.empty(N, F, dtype=torch.float32, device=xin.device)
        else:
            # zero-fills the output and serves the conformers above the dense kernel's atom limit (their share of the
            # work is not known on the host without a sync: the algorithmic FLOPs are booked on the dense launch)
            out = _per_edge_aggregate(xin, graph, pack_filter_weights(*W), offset, coeff, cutoff, transposed, cap + 1, 0)
        adj = graph.adjacency()
        call("cmp_cfconv_dense_fwd", ptr(xin), ptr(graph.pos), ptr(graph.seg_ptr), ptr(adj), graph.G,
             ptr(pack_dense_weights(*W)), ctypes.addressof(_offset_host(offset)), Ng, float(coeff), float(cutoff), F,
             int(bool(transposed)), int(not dense_only), int(graph.max_atoms or 0), ptr(out), ptr(graph._counter),
             ptr(graph.status), work=flops)
        return out
    packed = pack_filter_weights(*W)
    pairs = FUSED_PAIR_FORWARD and graph.G > 0
    min_atoms = _lib.size_query("cmp_cfconv_pair_max_atoms") + 1 if pairs else 0
    out = _per_edge_aggregate(xin, graph, packed, offset, coeff, cutoff, transposed, min_atoms, 0 if pairs else e_hint)
    if pairs:
        psrc, pdst, pdist, prev, _, _, conf_ptr = graph.pair_tiles()
        call("cmp_cfconv_pair_fwd", ptr(xin), ptr(graph.seg_ptr), ptr(conf_ptr), ptr(psrc), ptr(pdst), ptr(pdist),
             ptr(prev), graph.G, ptr(packed), ptr(offset), Ng, float(coeff), float(cutoff), F, int(bool(transposed)),
             ptr(out), work=flops)
    return out


def pack_filter_weights(W1, b1, W2, b2):
    cached = _cached_images(W1)
    if cached is not None and cached[0] is not None:
        return cached[0]
    F, Ng = W1.shape
    nbytes = _lib.size_query("cmp_cfconv_tc_weights_bytes")
    packed = torch.empty(nbytes, dtype=torch.uint8, device=W1.device)
    call("cmp_cfconv_tc_pack_weights", ptr(_f32c(W1)), ptr(_f32c(b1)), ptr(_f32c(W2)), ptr(_f32c(b2)), F, Ng,
         ptr(packed))
    return packed


def pack_dense_weights(W1, b1, W2, b2):
    cached = _cached_images(W1)
    if cached is not None and len(cached) > 2 and cached[2] is not None:
        return cached[2]
    F, Ng = W1.shape
    packed = torch.empty(_lib.size_query("cmp_cfconv_dense_weights_bytes"), dtype=torch.uint8, device=W1.device)
    call("cmp_cfconv_dense_pack_weights", ptr(_f32c(W1.detach())), ptr(_f32c(b1.detach())), ptr(_f32c(W2.detach())),
         ptr(_f32c(b2.detach())), F, Ng, ptr(packed))
    return packed


def pack_dense_x3_weights(W1, b1, W2, b2):
    cached = _cached_images(W1)
    if cached is not None and len(cached) > 3 and cached[3] is not None:
        return cached[3]
    F, Ng = W1.shape
    packed = torch.empty(_lib.size_query("cmp_cfconv_dense_x3_weights_bytes"), dtype=torch.uint8, device=W1.device)
    call("cmp_cfconv_dense_x3_pack_weights", ptr(_f32c(W1.detach())), ptr(_f32c(b1.detach())), ptr(_f32c(W2.detach())),
         ptr(_f32c(b2.detach())), F, Ng, ptr(packed))
    return packed


def pack_bwd_x3_weights(W1, b1, W2):
    cached = _cached_images(W1)
    if cached is not None and len(cached) > 4 and cached[4] is not None:
        return cached[4]
    F, Ng = W1.shape
    packed = torch.empty(_lib.size_query("cmp_cfconv_dense_bwd_x3_weights_bytes"), dtype=torch.uint8, device=W1.device)
    job = (_lib.BwdX3PackJob * 1)()
    keep = [_f32c(t.detach()) for t in (W1, b1, W2)]
    job[0].W1, job[0].b1, job[0].W2, job[0].packed = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), \
        packed.data_ptr()
    call("cmp_cfconv_dense_bwd_x3_pack_weights_grouped", ctypes.addressof(job), 1, F, Ng)
    return packed


class _CFConvFusedFn(Function):
    """Geometric CFConv in one tcgen05 kernel (bf16 filter MLP):  agg = sum_j x'_j * W(d_ij) * C(d_ij).

    backward: d x' runs the SAME fused kernel over the transposed neighbour list with the upstream
    gradient as its input (W depends on the edge only through d_ij); the filter-MLP weight gradients
    are recomputed on the exact-fp32 kernels (rbf -> Linear+ssp -> Linear, then GEMMs with K = E)."""

    @staticmethod
    def forward(ctx, xprime, W1, b1, W2, b2, graph, offset, coeff, cutoff, x3=False):
        xprime = _f32c(xprime)
        agg = _fused_aggregate(xprime, graph, (W1, b1, W2, b2), offset, coeff, cutoff, transposed=False, x3=x3)
        ctx.graph, ctx.coeff, ctx.cutoff, ctx.x3 = graph, float(coeff), float(cutoff), bool(x3)
        ctx.save_for_backward(xprime, W1, b1, W2, b2, offset)
        return agg

    @staticmethod
    def backward(ctx, g):
        xprime, W1, b1, W2, b2, offset = ctx.saved_tensors
        graph = ctx.graph
        g = _f32c(g)
        N, F = xprime.shape
        dx = None
        if ctx.needs_input_grad[0]:
            if graph.rowptr_t is None:
                raise _lib.ConanMPError("cfconv backward needs the transposed neighbour list")
            dx = _fused_aggregate(g, graph, (W1, b1, W2, b2), offset, ctx.coeff, ctx.cutoff, transposed=True,
                                  x3=ctx.x3)
        grads = [None, None, None, None]
        if any(ctx.needs_input_grad[1:5]):
            if FUSED_WEIGHT_GRADS and (not ctx.x3 or X3_WEIGHT_GRADS):
                grads = list(_fused_weight_grads(g, xprime, W1, b1, W2, graph, offset, ctx.coeff, ctx.cutoff,
                                                 x3=ctx.x3))
            else:   # exact-fp32 recompute of the filter MLP (kept for cross-checking the fused kernel)
                E = graph.E
                dfilt = torch.empty(E, F, dtype=torch.float32, device=g.device)
                call("cmp_cfconv_message_bwd", ptr(g), ptr(xprime), None, ptr(graph.dist), ptr(graph.rowptr),
                     ptr(graph.col), None, None, None, N, F, ctx.cutoff, ptr(dfilt), None)
                with torch.enable_grad():
                    ps = [t.detach().requires_grad_(True) for t in (W1, b1, W2, b2)]
                    rbf = gaussian_rbf(graph.dist[:E], offset, ctx.coeff)
                    filt = linear(linear(rbf, ps[0], ps[1], ACT_SSP), ps[2], ps[3])
                    grads = list(torch.autograd.grad(filt, ps, dfilt))
        return (dx, *grads, None, None, None, None, None)


FUSED_WEIGHT_GRADS = True
# fp32-grade mode: weight gradients on the three-pass tcgen05 kernel; False = exact-fp32 recompute on [E, *] tensors
X3_WEIGHT_GRADS = True
# one column per undirected pair in the weight-gradient pass (both directions share the filter); False = one column
# per directed edge with fp32 g (kept for cross-checking)
FUSED_PAIR_GRADS = True


# weight gradients over the dense blocks of the forward kernel (conformers of <= 128 atoms, promised by max_atoms): fp32
# rows of g and x' in registers instead of a pair list and bf16 copies; False = always the pair-list kernel
FUSED_DENSE_GRADS = True


def _fused_weight_grads(g, xprime, W1, b1, W2, graph, offset, coeff, cutoff, x3=False):
    """dW1, db1, dW2, db2 of the filter MLP through the tcgen05 weight-gradient kernels (K = edges / pairs)."""
    N, F = xprime.shape
    Ng = offset.numel()
    dev = g.device
    cached = _cached_images(W1)
    if x3:
        packed = None
    elif cached is not None and cached[1] is not None:
        packed = cached[1]
    else:
        packed = torch.empty(_lib.size_query("cmp_cfconv_tc_bwd_weights_bytes"), dtype=torch.uint8, device=dev)
        call("cmp_cfconv_tc_pack_bwd_weights", ptr(_f32c(W1)), ptr(_f32c(b1)), ptr(_f32c(W2)), F, Ng, ptr(packed))
    dW1 = torch.empty(F, Ng, dtype=torch.float32, device=dev)
    db1 = torch.empty(F, dtype=torch.float32, device=dev)
    dW2 = torch.empty(F, F, dtype=torch.float32, device=dev)
    db2 = torch.empty(F, dtype=torch.float32, device=dev)
    e_hint = graph._E if graph._E is not None else graph.cap_E
    if x3:
        if not x3_graph_ok(graph):
            raise _lib.ConanMPError("fp32-grade fused CFConv: needs a radius-built graph with max_atoms <= 128")
        ws = _lib.workspace(_lib.size_query("cmp_cfconv_dense_bwd_workspace"), dev)
        call("cmp_cfconv_dense_bwd_x3_weights", ptr(_f32c(g)), ptr(_f32c(xprime)), ptr(graph.pos), ptr(graph.seg_ptr),
             ptr(graph.adjacency()), ptr(graph.dense_bwd_tiles()), graph.G, ptr(pack_bwd_x3_weights(W1, b1, W2)),
             ptr(offset), Ng, float(coeff), float(cutoff), F, ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), ptr(ws),
             ws.numel(), work=2.0 * (Ng * F + F * F) * float(e_hint))
        return dW1, db1, dW2, db2
    if (FUSED_DENSE_GRADS and FUSED_DENSE and graph.G > 0 and graph.pos is not None and not graph.loop
            and graph.max_atoms is not None and graph.max_atoms <= _lib.size_query("cmp_cfconv_dense_max_atoms")):
        ws = _lib.workspace(_lib.size_query("cmp_cfconv_dense_bwd_workspace"), dev)
        call("cmp_cfconv_dense_bwd_weights", ptr(_f32c(g)), ptr(_f32c(xprime)), ptr(graph.pos), ptr(graph.seg_ptr),
             ptr(graph.adjacency()), ptr(graph.dense_bwd_tiles()), graph.G, ptr(packed), ptr(offset), Ng, float(coeff),
             float(cutoff), F, ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), ptr(ws), ws.numel(),
             work=2.0 * (Ng * F + F * F) * float(e_hint))
        return dW1, db1, dW2, db2
    xb = torch.empty(N, F, dtype=torch.bfloat16, device=dev)
    call("cmp_f32_to_bf16", ptr(xprime), N * F, ptr(xb))
    ws = _lib.workspace(_lib.size_query("cmp_cfconv_fused_bwd_workspace"), dev)
    # algorithmic FLOPs: the weight-gradient half of "bwd = 2 x fwd" = 2*(Ng*F + F*F) per edge
    if FUSED_PAIR_GRADS:
        gb = torch.empty(N, F, dtype=torch.bfloat16, device=dev)
        call("cmp_f32_to_bf16", ptr(_f32c(g)), N * F, ptr(gb))
        psrc, pdst, pdist, prev, tiles, num, _ = graph.pair_tiles()
        call("cmp_cfconv_fused_bwd_weights_pairs", ptr(gb), ptr(xb), ptr(pdist), ptr(psrc), ptr(pdst), ptr(prev),
             ptr(tiles), ptr(num), ptr(packed), ptr(offset), Ng, float(coeff), float(cutoff), F, ptr(dW1), ptr(db1),
             ptr(dW2), ptr(db2), ptr(ws), ws.numel(), work=2.0 * (Ng * F + F * F) * float(e_hint))
        return dW1, db1, dW2, db2
    erow, tiles, num = graph.flat_tiles()
    call("cmp_cfconv_fused_bwd_weights", ptr(g), ptr(xb), ptr(graph.dist), ptr(graph.col), ptr(erow), ptr(tiles),
         ptr(num), ptr(packed), ptr(offset), Ng, float(coeff), float(cutoff), F, ptr(dW1), ptr(db1), ptr(dW2),
         ptr(db2), ptr(ws), ws.numel(), work=2.0 * (Ng * F + F * F) * float(e_hint))
    return dW1, db1, dW2, db2


FB = 128     # filter channels of the fused kernels (one TMEM lane / one epilogue thread per channel)


def cfconv_fused(xprime, W1, b1, W2, b2, graph, offset, coeff, cutoff, x3=False):
    """Fused geometric CFConv.  ``num_filters`` = 128 runs one kernel per direction; 256 / 384 / 512 (ConAN's classification
    models use F = 256, ``conan_fgw/src/model/common.py:513-522``) are composed from the same kernels, because the
    aggregate is linear in the filter and the filter's second Linear splits over blocks of 128 hidden channels:

        agg[:, o] = sum_k  CFConv(x'[:, o];  W1[k], b1[k],  W2[o, k],  b2[o] if k == 0 else 0)

    (o, k = 128-channel blocks of the output / hidden filter channels).  (F / 128)^2 launches per direction instead of
    one - the same tensor-pipe work as a native F-wide kernel, the Gaussians and the first Linear evaluated F / 128 times
    too often - but still no ``[E, *]`` tensor and no host sync; gradients follow through autograd over the slices."""
    F = W1.shape[0]
    if F == FB:
        return _CFConvFusedFn.apply(xprime, W1, b1, W2, b2, graph, offset, coeff, cutoff, x3)
    if F % FB != 0:
        raise _lib.ConanMPError(f"fused CFConv: num_filters must be a multiple of {FB} (got {F})")
    nb = F // FB
    zero_b2 = torch.zeros(FB, dtype=torch.float32, device=xprime.device)
    w1k = [(W1[k * FB:(k + 1) * FB], b1[k * FB:(k + 1) * FB]) for k in range(nb)]      # row slices: contiguous views
    outs = []
    for o in range(nb):
        xo = xprime[:, o * FB:(o + 1) * FB].contiguous()
        acc = None
        for k in range(nb):
            W2ok = W2[o * FB:(o + 1) * FB, k * FB:(k + 1) * FB].contiguous()
            b2o = b2[o * FB:(o + 1) * FB] if k == 0 else zero_b2
            term = _CFConvFusedFn.apply(xo, w1k[k][0], w1k[k][1], W2ok, b2o, graph, offset, coeff, cutoff, x3)
            acc = term if acc is None else acc + term
        outs.append(acc)
    return torch.cat(outs, dim=1)


def fused_native(num_filters, num_gaussians) -> bool:
    """Shapes the fused kernels serve in ONE launch (and the grouped weight packs prepare images for)."""
    return bool(_lib.lib().cmp_cfconv_tc_supported(int(num_filters), int(num_gaussians))) and \
        bool(_lib.lib().cmp_device_is_sm100())


def fused_supported(num_filters, num_gaussians) -> bool:
    """Shapes ``cfconv_fused`` serves: the native one, or a multiple of it composed from 128-channel blocks."""
    F = int(num_filters)
    return F % FB == 0 and FB <= F <= 4 * FB and fused_native(FB, num_gaussians)


class _SegmentSumFn(Function):
    @staticmethod
    def forward(ctx, x, seg_ptr, G):
        x = _f32c(x)
        C = x.shape[1]
        out = torch.empty(G, C, dtype=torch.float32, device=x.device)
        call("cmp_segment_sum_fwd", ptr(x), ptr(seg_ptr), G, C, ptr(out))
        ctx.save_for_backward(seg_ptr)
        ctx.N, ctx.G, ctx.C = x.shape[0], G, C
        return out

    @staticmethod
    def backward(ctx, dout):
        (seg_ptr,) = ctx.saved_tensors
        dout = _f32c(dout)
        dx = torch.zeros(ctx.N, ctx.C, dtype=torch.float32, device=dout.device)
        call("cmp_segment_sum_bwd", ptr(dout), ptr(seg_ptr), ctx.G, ctx.C, ptr(dx))
        return dx, None, None


def segment_sum(x, seg_ptr, G):
    """Sum readout over sorted, contiguous segments (``seg_ptr int32[G+1]``)."""
    return _SegmentSumFn.apply(x, seg_ptr, int(G))


_uniform_seg_cache = {}


def conformers_mean(x, num_conformers: int):
    """Mean over the K consecutive conformers of each molecule: [B*K, C] -> [B, C].

    Replaces ``create_aggregation_index`` (a Python loop + host->device copy every step,
    ``conan_fgw/src/model/common.py:414-423``) and ``conformers_mean_aggr``
    (``schnet_based_models.py:242``): the index is implicit because conformers of a molecule are consecutive."""
    G = x.shape[0]
    K = int(num_conformers)
    if G % K != 0:
        raise ValueError("conformers_mean: the number of conformers must be a multiple of K")
    key = (G, K, x.device)
    seg = _uniform_seg_cache.get(key)
    if seg is None:
        seg = torch.arange(0, G + 1, K, dtype=torch.int32, device=x.device)
        _uniform_seg_cache[key] = seg
    return segment_sum(x, seg, G // K) * (1.0 / K)


class _RegressionHeadLossFn(Function):
    """conformer mean -> Linear(C, 1) -> MSE as ONE forward and ONE backward launch (``cmp_regression_head_fwd/bwd``)."""

    @staticmethod
    def forward(ctx, emb, weight, bias, targets, K):
        emb = _f32c(emb)
        G, C = emb.shape
        if G % K != 0:
            raise ValueError("regression_head_loss: the number of conformers must be a multiple of K")
        B = G // K
        w, t = _f32c(weight.reshape(-1)), _f32c(targets.reshape(-1))
        if w.numel() != C or t.numel() != B:
            raise ValueError("regression_head_loss: weight must be [1, C] and targets [B] or [B, 1]")
        err = torch.empty(B, dtype=torch.float32, device=emb.device)
        loss = torch.empty((), dtype=torch.float32, device=emb.device)
        call("cmp_regression_head_fwd", ptr(emb), emb.stride(0), B, int(K), C, ptr(w), ptr(bias), ptr(t), ptr(err), ptr(loss))
        ctx.K, ctx.has_bias, ctx.wshape = int(K), bias is not None, weight.shape
        ctx.save_for_backward(emb, w, err)
        return loss

    @staticmethod
    def backward(ctx, g):
        emb, w, err = ctx.saved_tensors
        G, C = emb.shape
        B = G // ctx.K
        g = _f32c(g.reshape(1))
        d_emb = torch.empty_like(emb)
        dw = torch.empty(C, dtype=torch.float32, device=emb.device)
        db = torch.empty(1, dtype=torch.float32, device=emb.device) if ctx.has_bias else None
        call("cmp_regression_head_bwd", ptr(emb), emb.stride(0), B, ctx.K, C, ptr(w), ptr(err), ptr(g), ptr(d_emb),
             d_emb.stride(0), ptr(dw), ptr(db))
        return d_emb, dw.reshape(ctx.wshape), db, None, None


def regression_head_loss(emb, weight, bias, targets, num_conformers: int):
    """``mse_loss(linear(conformers_mean(emb, K), weight, bias), targets)`` for a one-output head (ConAN's regression
    head: ``schnet_based_models.py:17-29,242``, ``model/common.py:288``) in two kernel launches."""
    return _RegressionHeadLossFn.apply(emb, weight, bias, targets, int(num_conformers))


def segments_from_batch(batch, num_graphs=None, status=None):
    """Segment pointers of a SORTED index vector.  An unsorted index cannot be served by the segment kernels (PyG's
    scatter would accept it): it is reported - through ``status`` when the caller supplies its own word (sync-free
    paths read it later), otherwise by one device -> host read here."""
    from .graph import num_graphs_of, raise_for_status

    batch = batch.to(torch.int64).contiguous()
    G = num_graphs_of(batch, num_graphs)
    seg = torch.empty(G + 1, dtype=torch.int32, device=batch.device)
    own = status is None
    if own:
        status = torch.zeros(1, dtype=torch.int32, device=batch.device)
    call("cmp_batch_to_segments", ptr(batch), batch.numel(), G, ptr(seg), ptr(status))
    if own and not torch.cuda.is_current_stream_capturing():
        raise_for_status(status)
    return seg, G


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              grad_scale=1.0):
    """Fused Adam on flat fp32 buffers (in place)."""
    call("cmp_adam_step", ptr(param, torch.float32), ptr(grad, torch.float32), ptr(exp_avg, torch.float32),
         ptr(exp_avg_sq, torch.float32), param.numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
         float(weight_decay), int(step), float(grad_scale))

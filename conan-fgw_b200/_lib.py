"""ctypes binding of ``libconanmp.so`` (the C ABI declared in ``include/conanmp.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``build.build_library()``.
There is no fallback of any kind: if the shared object is missing, or a kernel
is asked to run without a CUDA device, the call raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libconanmp.so")

P, I, L, F, D, S = c_void_p, c_int, c_int64, c_float, c_double, c_size_t

# name -> (restype, argtypes); mirrors include/conanmp.h one to one
SIGNATURES = {
    "cmp_last_error_string": (c_char_p, []),
    "cmp_version": (I, []),
    "cmp_launch_count": (ctypes.c_longlong, []),
    "cmp_launch_count_reset": (None, []),
    "cmp_device_is_sm100": (I, []),
    "cmp_batch_to_segments": (I, [P, L, L, P, P, P]),
    "cmp_radius_csr_workspace": (S, [L, L]),
    "cmp_radius_csr": (I, [P, P, L, L, D, I, I, L, P, P, P, P, P, P, P, P, P, S, P, P]),
    "cmp_csr_to_edge_index": (I, [P, P, L, L, P, P]),
    "cmp_check_edge_count": (I, [P, L, L, P, P]),
    "cmp_gemm_workspace": (S, [L, L, L]),
    "cmp_gemm_f32": (I, [I, I, L, L, L, P, L, P, L, P, L, P, I, P, L, P, S, P]),
    "cmp_colsum_workspace": (S, [L, L]),
    "cmp_colsum_f32": (I, [P, L, L, L, P, P, S, P]),
    "cmp_act_fwd": (I, [P, P, L, I, P]),
    "cmp_act_bwd": (I, [P, P, P, L, I, P]),
    "cmp_rbf_gaussian_fwd": (I, [P, L, P, I, F, P, L, P]),
    "cmp_embedding_fwd": (I, [P, L, P, I, I, P, P, P]),
    "cmp_embedding_bwd_workspace": (S, [L, I, I]),
    "cmp_embedding_bwd": (I, [P, L, P, I, I, I, P, P, S, P]),
    "cmp_cfconv_message_fwd": (I, [P, P, P, P, P, L, I, F, P, P]),
    "cmp_cfconv_message_bwd": (I, [P, P, P, P, P, P, P, P, P, L, I, F, P, P, P]),
    "cmp_regression_head_max_channels": (I, []),
    "cmp_regression_head_fwd": (I, [P, L, L, I, I, P, P, P, P, P, P]),
    "cmp_regression_head_bwd": (I, [P, L, L, I, I, P, P, P, P, L, P, P, P]),
    "cmp_segment_sum_fwd": (I, [P, P, L, I, P, P]),
    "cmp_segment_sum_bwd": (I, [P, P, L, I, P, P]),
    "cmp_adam_step": (I, [P, P, P, P, L, F, F, F, F, F, I, F, P]),
    "cmp_build_tiles_workspace": (S, [L]),
    "cmp_build_tiles": (I, [P, P, L, I, P, L, P, P, S, P, P]),
    "cmp_build_tiles_min_atoms": (I, [P, P, L, I, I, P, L, P, P, S, P, P]),
    "cmp_gather_f32": (I, [P, P, P, L, P, P]),
    "cmp_cfconv_tc_supported": (I, [I, I]),
    "cmp_cfconv_tc_weights_bytes": (S, []),
    "cmp_cfconv_tc_tile_edges": (I, []),
    "cmp_cfconv_tc_pack_weights": (I, [P, P, P, P, I, I, P, P]),
    "cmp_cfconv_fused_fwd": (I, [P, P, P, P, P, P, P, P, I, F, F, L, I, P, P]),
    "cmp_debug_umma_gemm": (I, [P, L, P, L, P, I, I, I, I, I, I, I, I, I, I, I, P]),
    "cmp_node_gemm_tc_supported": (I, [I, I]),
    "cmp_node_gemm_weight_bytes": (S, [I]),
    "cmp_node_gemm_pack_weight": (I, [P, I, I, I, P, P]),
    "cmp_node_gemm_fwd": (I, [P, L, P, L, P, P, I, P, L, P, L, L, I, I, P]),
    "cmp_node_chain_max_stages": (I, []),
    "cmp_node_chain_fwd": (I, [P, L, L, P, I, P]),
    "cmp_node_gemm_dw_workspace": (S, [I]),
    "cmp_node_gemm_dw": (I, [P, L, P, L, P, L, L, I, I, P, P, P, S, P]),
    "cmp_edge_message_fwd": (I, [P, P, P, P, P, L, I, P, P]),
    "cmp_edge_message_bwd": (I, [P, P, P, P, P, P, P, P, P, L, I, P, P, P]),
    "cmp_vis_edge_geometry": (I, [P, P, P, P, L, F, F, P, P, I, P, P, P, P]),
    "cmp_layernorm_fwd": (I, [P, P, P, L, I, F, P, P, P, P]),
    "cmp_layernorm_bwd": (I, [P, P, P, P, P, L, I, P, P, P]),
    "cmp_csr_segment_sum": (I, [P, P, P, L, I, P, P]),
    "cmp_gather_rows": (I, [P, P, L, I, P, P]),
    "cmp_vis_edge_embed_fwd": (I, [P, P, P, P, L, I, P, P]),
    "cmp_vis_edge_embed_bwd": (I, [P, P, P, P, P, L, I, P, P, P]),
    "cmp_vis_message_fwd": (I, [P, P, P, P, P, P, P, P, L, I, I, I, P, P, P]),
    "cmp_vis_message_bwd": (I, [P, P, P, P, P, P, P, P, P, P, L, I, I, I, P, P, P, P, P, P]),
    "cmp_vis_vecagg_fwd": (I, [P, P, P, P, P, L, I, I, P, P]),
    "cmp_vis_vecagg_bwd": (I, [P, P, P, P, P, P, P, P, P, L, I, I, P, P, P]),
    "cmp_vis_edge_update_fwd": (I, [P, P, P, P, P, P, L, I, I, P, P, P]),
    "cmp_vis_edge_update_bwd_prep": (I, [P, P, P, L, I, P, P, P]),
    "cmp_vis_node_update_fwd": (I, [P, P, P, L, I, P, P, P]),
    "cmp_vis_node_update_bwd": (I, [P, P, P, P, L, I, P, P, P]),
    "cmp_vis_edge_update_bwd": (I, [P, P, P, P, P, P, P, P, P, L, I, P, P, P]),
    "cmp_debug_set_fwd_timestamps": (None, [P]),
    "cmp_debug_set_bwd_timestamps": (None, [P]),
    "cmp_debug_set_pair_timestamps": (None, [P]),
    "cmp_debug_set_dense_timestamps": (None, [P]),
    "cmp_debug_set_dense_stagger": (None, [I]),
    "cmp_debug_set_dense_pipes": (None, [I]),
    "cmp_debug_set_dense_mode": (None, [I]),
    "cmp_debug_set_dense_variant": (None, [I]),
    "cmp_debug_set_dense_bwd_variant": (None, [I]),
    "cmp_csr_expand_rows": (I, [P, L, P, P]),
    "cmp_cfconv_tc_bwd_tile_edges": (I, []),
    "cmp_build_flat_tiles_workspace": (S, [L]),
    "cmp_build_flat_tiles": (I, [P, P, P, L, I, P, L, P, P, S, P, P]),
    "cmp_f32_to_bf16": (I, [P, L, P, P]),
    "cmp_cfconv_tc_bwd_weights_bytes": (S, []),
    "cmp_cfconv_fused_bwd_workspace": (S, []),
    "cmp_cfconv_tc_pack_bwd_weights": (I, [P, P, P, I, I, P, P]),
    "cmp_cfconv_fused_bwd_weights": (I, [P, P, P, P, P, P, P, P, P, I, F, F, I, P, P, P, P, P, S, P]),
    "cmp_node_gemm_pack_weights_grouped": (I, [P, I, P]),
    "cmp_cfconv_tc_pack_weights_grouped": (I, [P, I, I, I, P]),
    "cmp_cfconv_tc_pack_bwd_weights_grouped": (I, [P, I, I, I, P]),
    "cmp_dense_batch": (I, [P, P, L, L, I, F, P, P, P]),
    "cmp_dense_adj": (I, [P, L, P, P, L, L, P, P]),
    "cmp_node_gemm_dw_group_max": (I, []),
    "cmp_node_gemm_dw_grouped_workspace": (S, []),
    "cmp_node_gemm_dw_grouped": (I, [P, I, P, S, P]),
    "cmp_cfconv_pair_max_atoms": (I, []),
    "cmp_cfconv_pair_fwd": (I, [P, P, P, P, P, P, P, L, P, P, I, F, F, I, I, P, P]),
    "cmp_build_pair_list_workspace": (S, [L, L]),
    "cmp_build_pair_list": (I, [P, P, P, P, L, L, I, L, P, P, P, P, P, P, S, P, P]),
    "cmp_cfconv_fused_bwd_weights_pairs": (I, [P, P, P, P, P, P, P, P, P, P, I, F, F, I, P, P, P, P, P, S, P]),
    "cmp_cfconv_dense_max_atoms": (I, []),
    "cmp_cfconv_dense_supported": (I, [I, I]),
    "cmp_cfconv_dense_weights_bytes": (S, []),
    "cmp_build_adjacency": (I, [P, P, P, L, L, P, P]),
    "cmp_cfconv_dense_pack_weights": (I, [P, P, P, P, I, I, P, P]),
    "cmp_cfconv_dense_pack_weights_grouped": (I, [P, I, I, I, P]),
    "cmp_cfconv_dense_fwd": (I, [P, P, P, P, L, P, P, I, F, F, I, I, I, I, P, P, P, P]),
    "cmp_cfconv_dense_x3_weights_bytes": (S, []),
    "cmp_cfconv_dense_x3_pack_weights": (I, [P, P, P, P, I, I, P, P]),
    "cmp_cfconv_dense_x3_pack_weights_grouped": (I, [P, I, I, I, P]),
    "cmp_cfconv_dense_x3_fwd": (I, [P, P, P, P, L, P, P, I, F, F, I, I, I, P, P, P, P]),
    "cmp_cfconv_dense_bwd_x3_weights_bytes": (S, []),
    "cmp_cfconv_dense_bwd_x3_pack_weights_grouped": (I, [P, I, I, I, P]),
    "cmp_cfconv_dense_bwd_x3_weights": (I, [P, P, P, P, P, P, L, P, P, I, F, F, I, P, P, P, P, P, S, P]),
    "cmp_cfconv_dense_bwd_workspace": (S, []),
    "cmp_build_dense_bwd_tiles": (I, [P, L, P, P, P]),
    "cmp_cfconv_dense_bwd_weights": (I, [P, P, P, P, P, P, L, P, P, I, F, F, I, P, P, P, P, P, S, P]),
}

ACT_NONE, ACT_SSP, ACT_SILU = 0, 1, 2
PREC_FP32, PREC_BF16 = 0, 1
STATUS_UNSORTED_BATCH, STATUS_BAD_ATOMIC_NUMBER, STATUS_EDGE_OVERFLOW = 1, 2, 4

_lib = None
_launches = 0  # number of C-ABI compute calls issued (bench.py reports it)


class ConanMPError(RuntimeError):
    pass


class PackNodeJob(ctypes.Structure):
    """``cmp_pack_node_job_t``."""
    _fields_ = [("W", ctypes.c_void_p), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32), ("packed", ctypes.c_void_p)]


class PackFilterJob(ctypes.Structure):
    """``cmp_pack_filter_job_t``."""
    _fields_ = [("W1", ctypes.c_void_p), ("b1", ctypes.c_void_p), ("W2", ctypes.c_void_p), ("b2", ctypes.c_void_p),
                ("packed_fwd", ctypes.c_void_p), ("packed_bwd", ctypes.c_void_p)]


class DensePackJob(ctypes.Structure):
    """``cmp_dense_pack_job_t``."""
    _fields_ = [("W1", ctypes.c_void_p), ("b1", ctypes.c_void_p), ("W2", ctypes.c_void_p), ("b2", ctypes.c_void_p),
                ("packed", ctypes.c_void_p)]


class BwdX3PackJob(ctypes.Structure):
    """``cmp_bwd_x3_pack_job_t``."""
    _fields_ = [("W1", ctypes.c_void_p), ("b1", ctypes.c_void_p), ("W2", ctypes.c_void_p), ("packed", ctypes.c_void_p)]


class ChainStage(ctypes.Structure):
    """``cmp_chain_stage_t``."""
    _fields_ = [("w_img", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p),
                ("scale_y", ctypes.c_void_p), ("out", ctypes.c_void_p), ("ldr", ctypes.c_int64),
                ("lds", ctypes.c_int64), ("ldo", ctypes.c_int64), ("K", ctypes.c_int32), ("Nout", ctypes.c_int32),
                ("act", ctypes.c_int32)]


class DwProblem(ctypes.Structure):
    """``cmp_dw_problem_t`` of include/conanmp.h (one weight-gradient problem of a grouped launch)."""
    _fields_ = [("dY", ctypes.c_void_p), ("lddy", ctypes.c_int64), ("saved_y", ctypes.c_void_p),
                ("ldys", ctypes.c_int64), ("X", ctypes.c_void_p), ("ldx", ctypes.c_int64), ("M", ctypes.c_int64),
                ("K", ctypes.c_int32), ("Nout", ctypes.c_int32), ("dW", ctypes.c_void_p), ("lddw", ctypes.c_int64),
                ("db", ctypes.c_void_p)]


def register(extra: dict):
    """Let sibling modules (fused kernels added later) declare more entry points."""
    SIGNATURES.update(extra)
    if _lib is not None:
        _bind(_lib, extra)


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args


def lib():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ConanMPError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for these kernels.")
        handle = ctypes.CDLL(LIB_PATH)
        _bind(handle, SIGNATURES)
        _lib = handle
    return _lib


def launches() -> int:
    """Kernels launched by libconanmp in this process (counted inside the library)."""
    return int(lib().cmp_launch_count())


def reset_launches():
    lib().cmp_launch_count_reset()


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ConanMPError("conanmp kernels need CUDA tensors (there is no CPU path); got a tensor on " + str(t.device))
    if dtype is not None and t.dtype != dtype:
        raise ConanMPError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ConanMPError("expected a contiguous tensor")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


class KernelTimer:
    """Optional CUDA-event timing of selected entry points on the launching stream (bench.py uses it
    to measure the dominant kernel live, inside the timed region)."""

    def __init__(self, names, external=False):
        self.names = set(names)
        self.records = []   # (name, start_event, end_event, work)
        # external=True: the events become event-record NODES when the calls are captured into a CUDA graph
        # (cudaEventRecordExternal), so the same records time every replay: call accumulate() after each replay + sync
        self.external = bool(external)
        self.acc = {}       # name -> [launches, total_ms, total_work, [durations]]

    def summary(self):
        """{name: (launches, total_ms, total_work)} - call after a device synchronise."""
        out = {}
        for name, e0, e1, work in self.records:
            n, ms, w = out.get(name, (0, 0.0, 0.0))
            out[name] = (n + 1, ms + e0.elapsed_time(e1), w + float(work))
        return out

    def accumulate(self):
        """Add the durations of the last graph replay (external mode) - call after a device synchronise."""
        for name, e0, e1, work in self.records:
            slot = self.acc.setdefault(name, [0, 0.0, 0.0, []])
            ms = e0.elapsed_time(e1)
            slot[0] += 1
            slot[1] += ms
            slot[2] += float(work)
            slot[3].append(ms)

    def summary_accumulated(self):
        return {k: (v[0], v[1], v[2]) for k, v in self.acc.items()}

    def medians_accumulated(self):
        return {k: sorted(v[3])[len(v[3]) // 2] for k, v in self.acc.items()}

    def medians(self):
        """{name: median launch duration in ms} - call after a device synchronise."""
        per = {}
        for name, e0, e1, _ in self.records:
            per.setdefault(name, []).append(e0.elapsed_time(e1))
        return {k: sorted(v)[len(v) // 2] for k, v in per.items()}


timer: "KernelTimer | None" = None

# NVTX ranges around the phases of a step (radius graph + forward, backward, gradient all-reduce, Adam) and around every
# interaction block: off by default (a range costs a few hundred ns on the launching thread); CMP_NVTX=1 turns them on
# for nsys / ncu --nvtx timelines.
NVTX = os.environ.get("CMP_NVTX", "0") not in ("", "0")


class nvtx_range:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def call(name, *args, work=0.0):
    """Invoke an int-returning entry point on the current stream; raise on a non-zero code.
    ``work`` = algorithmic FLOPs (or bytes) of this launch, recorded when a KernelTimer is active."""
    global _launches
    fn = getattr(lib(), name)
    t = timer
    if t is not None and name in t.names:
        e0 = torch.cuda.Event(enable_timing=True, external=t.external)
        e1 = torch.cuda.Event(enable_timing=True, external=t.external)
        e0.record()
        rc = fn(*args, stream())
        e1.record()
        t.records.append((name, e0, e1, work))
    else:
        rc = fn(*args, stream())
    _launches += 1
    if rc != 0:
        msg = lib().cmp_last_error_string().decode("utf-8", "replace")
        exc = ValueError if rc in (-1, -2) else ConanMPError
        raise exc(f"{name} failed ({rc}): {msg}")


def size_query(name, *args) -> int:
    return int(getattr(lib(), name)(*args))


def workspace(nbytes: int, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)

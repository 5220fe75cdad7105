"""Ad-hoc: CUDA-event timing of the forward aggregation kernels (dense-block vs round-1 pair / per-edge) on a workload.

usage: python tools/dense_timing.py [workload ...]        (default cfg2_lipo_train)
Prints us per launch and the fraction of the measured burst bf16 peak (algorithmic FLOPs per DIRECTED edge)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp  # noqa: E402
from conan_fgw_b200 import ops  # noqa: E402

dev = "cuda"
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except OSError:
    pass
PEAK = float(peaks.get("bf16_tflops", 1590.0))


def time_it(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for wl in (sys.argv[1:] or ["cfg2_lipo_train"]):
    b = cmp.synthetic.make_config_batch(wl).to(dev)
    cutoff = 10.0
    n_max = int(torch.bincount(b.batch).max())
    nl = cmp.build_neighbor_list(b.pos, b.batch, cutoff, max_atoms=n_max, num_graphs=b.num_graphs)
    E = nl.E
    torch.manual_seed(0)
    blk = cmp.InteractionBlock(128, 50, 128, cutoff).to(dev)
    gs = cmp.GaussianSmearing(0.0, cutoff, 50).to(dev)
    W = (blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias)
    x = torch.randn(b.z.numel(), 128, device=dev)
    flops = 45568.0 * E
    res = {}
    sweep = [int(v) for v in os.environ.get("DENSE_STAGGER", "600").split(",")]
    cases = [(f"dense st={ns}", True, True, ns) for ns in sweep] + [("pair/per-edge", False, True, None)]
    for name, dense, pair, ns in cases:
        ops.FUSED_DENSE, ops.FUSED_PAIR_FORWARD = dense, pair
        if ns is not None:
            cmp._lib.lib().cmp_debug_set_dense_stagger(ns)
        with torch.no_grad():
            med, best = time_it(lambda: ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, cutoff, False))
        res[name] = med
        print(f"{wl}: n_max={n_max} E={E} {name:14s} median {med:8.1f} us  best {best:8.1f} us  "
              f"{flops / med * 1e-6:7.1f} TFLOP/s = {flops / med * 1e-6 / PEAK:.3f} of burst peak {PEAK}")
    ops.FUSED_DENSE, ops.FUSED_PAIR_FORWARD = True, True

"""Ad-hoc: CUDA-event timing of cmp_cfconv_dense_fwd, warp-specialised (variant 0) vs per-pipeline (variant 1) kernel,
plus their maximum difference.   usage: python tools/dense_variants.py [workload ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp  # noqa: E402
from conan_fgw_b200 import ops  # noqa: E402

dev = "cuda"
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"])
except OSError:
    PEAK = 1590.0


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
    outs = {}
    for variant in (1, 0):
        cmp._lib.lib().cmp_debug_set_dense_variant(variant)
        for tr in (False, True):
            with torch.no_grad():
                outs[(variant, tr)] = ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, cutoff, tr).clone()
                med, best = time_it(lambda: ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, cutoff, tr))
            print(f"{wl}: n_max={n_max} E={E} variant {variant} transposed={int(tr)} median {med:8.1f} us  best {best:8.1f} us  "
                  f"{flops / med * 1e-6:7.1f} TFLOP/s = {flops / med * 1e-6 / PEAK:.3f} of burst peak {PEAK}", flush=True)
    cmp._lib.lib().cmp_debug_set_dense_variant(0)
    for tr in (False, True):
        d = (outs[(0, tr)] - outs[(1, tr)]).abs().max().item()
        print(f"{wl}: transposed={int(tr)} max |variant 0 - variant 1| = {d:.3e} (max |out| {outs[(1, tr)].abs().max().item():.3e})")

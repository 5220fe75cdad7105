"""fp32-grade fused mode: measured errors against the fp64-grade CPU oracle and the exact GPU mode, and kernel timings.
Run on the GPU box: python tools/x3_errors.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import conan_fgw_b200 as cmp
from conan_fgw_b200 import nn as cnn, ops
from oracle import schnet as osn

syn = cmp.synthetic
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def model_errors(cfg, batch, hint, node_tc, x3_wgrads=True):
    ops.X3_WEIGHT_GRADS = x3_wgrads
    torch.manual_seed(0)
    o = osn.SchNetNoSum(None, **cfg)
    with torch.no_grad():
        for p in o.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    o64 = osn.SchNetNoSum(None, **cfg).double()
    o64.load_state_dict({k: v.double() for k, v in o.state_dict().items()})
    c = cmp.SchNetNoSum(None, **cfg).to(DEV)
    c.load_state_dict(o.state_dict(), strict=True)
    c.max_atoms_hint = hint
    cnn.FP32_NODE_TC = node_tc
    b = batch
    out_o = o64(b.z, b.pos.double(), b.batch)
    out_c = c(b.z.to(DEV), b.pos.to(DEV), b.batch.to(DEV))
    o64.zero_grad(), c.zero_grad()
    out_o.pow(2).mean().backward()
    out_c.pow(2).mean().backward()
    c.check_status()
    po, pc = dict(o64.named_parameters()), dict(c.named_parameters())
    errs = sorted(((rel(pc[k].grad, po[k].grad), k) for k in po if po[k].grad is not None), reverse=True)
    return rel(out_c, out_o), errs


if __name__ == "__main__":
    full = dict(hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=10.0)
    for name, cfg, batch, n in [
        ("cfg2-like 8x5x27, T=6", full, syn.make_batch(8, 5, 27, seed=1), 27),
        ("T=3, 2 x 65 atoms", dict(full, num_interactions=3), syn.make_batch(1, 2, 65, seed=4), 65),
        ("T=6, 4x3x45", full, syn.make_batch(4, 3, 45, seed=5), 45),
    ]:
        for hint, tc, xw in ((None, False, True), (n, False, False), (n, False, True), (n, True, True)):
            e, errs = model_errors(cfg, batch, hint, tc, xw)
            mode = "exact" if hint is None else ("x3 fwd" + (" + x3 wgrads" if xw else " + exact wgrads")
                                                 + (" + tcgen05 node linears" if tc else " + exact node linears"))
            top = ", ".join(f"{k.replace('interactions.', 'i')} {g:.1e}" for g, k in errs[:4])
            print(f"{name:22s} {mode:48s} out {e:.2e}  grads: {top}", flush=True)
    # timings at cfg 2
    b = syn.make_config_batch("cfg2_lipo_train").to(DEV)
    nl = cmp.build_neighbor_list(b.pos, b.batch, 10.0, 32, max_atoms=27)
    torch.manual_seed(0)
    blk = cmp.InteractionBlock(128, 50, 128, 10.0).to(DEV)
    gs = cmp.GaussianSmearing(0.0, 10.0, 50).to(DEV)
    W = [blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias]
    xp = torch.randn(b.z.numel(), 128, device=DEV)
    g = torch.randn(b.z.numel(), 128, device=DEV)
    for x3 in (False, True):
        for what in ("fwd", "wgrad"):
            def run():
                if what == "fwd":
                    return ops._fused_aggregate(xp, nl, W, gs.offset, gs.coeff, 10.0, False, x3=x3)
                return ops._fused_weight_grads(g, xp, W[0], W[1], W[2], nl, gs.offset, gs.coeff, 10.0, x3=x3)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            print(f"cfg2 {what:6s} x3={x3}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (incl. weight packing)")

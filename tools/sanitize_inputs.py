"""Small forward + backward of both backbones in every numerics mode (exact kernels, fp32-grade fused x3 kernels, f16 fused
kernels) - the workload of tools/sanitize.sh."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
syn = cmp.synthetic
for n, B, K in ((9, 2, 2), (27, 2, 1), (40, 1, 2)):
    b = syn.make_batch(B, K, n, seed=n).to(dev)
    m = cmp.SchNetNoSum(None, num_interactions=2).to(dev)
    for prec in ("exact", "fp32", "bf16"):
        if prec != "exact" and not cmp._lib.lib().cmp_device_is_sm100():
            continue
        m.set_precision(prec)
        # both kernels behind cmp_cfconv_dense_fwd: 0 = warp-specialised tile pipeline, 1 = per-pipeline; with the atom
        # bound promised (variant 0 pass) the weight gradients run on the dense-block kernel, without it on the pair list
        for variant in ((0, 1) if prec == "bf16" else (-1,)):
            cmp._lib.lib().cmp_debug_set_dense_variant(variant)
            m.max_atoms_hint = n if variant == 0 or prec == "fp32" else None     # fp32 + bound: the x3 kernels
            m.zero_grad()
            out = m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
            out.pow(2).mean().backward()
            m.check_status()
        cmp._lib.lib().cmp_debug_set_dense_variant(-1)
    print(f"schnet n={n}: ok", flush=True)
v = cmp.ViSNet(None, hidden_channels=32).to(dev)
b = syn.make_batch(2, 2, 8, seed=3).to(dev)
v(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean().backward()
torch.cuda.synchronize()
print("visnet: ok", flush=True)

"""Ad-hoc: which phase is the warp-specialised dense kernel's critical path?  Times the DBG build with phases switched off
(results are garbage, only the timing matters).  bit 0: no epilogue 1, 1: no Gaussians, 2: no epilogue 2, 3: no staging loads,
4: no proxy fence after epilogue 1."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp
from conan_fgw_b200 import _lib, ops
dev = "cuda"
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_lipo_train"
b = cmp.synthetic.make_config_batch(wl).to(dev)
n_max = int(torch.bincount(b.batch).max())
nl = cmp.build_neighbor_list(b.pos, b.batch, 10.0, max_atoms=n_max, num_graphs=b.num_graphs)
torch.manual_seed(0)
blk = cmp.InteractionBlock(128, 50, 128, 10.0).to(dev)
gs = cmp.GaussianSmearing(0.0, 10.0, 50).to(dev)
W = (blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias)
x = torch.randn(b.z.numel(), 128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in [0, 15, 15 + 64, 15 + 128, 15 + 64 + 128, 64, 128]:
    _lib.lib().cmp_debug_set_dense_mode(mode)
    ts = []
    with torch.no_grad():
        for i in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, 10.0, False)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{wl} mode {mode:2d}: median {ts[len(ts) // 2]:7.1f} us  best {ts[0]:7.1f} us", flush=True)
_lib.lib().cmp_debug_set_dense_mode(0)

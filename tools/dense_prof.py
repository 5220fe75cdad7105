"""Ad-hoc: run the dense forward kernel a few times on a workload (for ncu)."""
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
_lib.lib().cmp_debug_set_dense_pipes(int(os.environ.get("DENSE_PIPES", "4")))
_lib.lib().cmp_debug_set_dense_stagger(int(os.environ.get("DENSE_STAGGER", "600")))
with torch.no_grad():
    for _ in range(5):
        ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, 10.0, False)
torch.cuda.synchronize()

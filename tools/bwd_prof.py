"""Ad-hoc: run the filter-MLP weight-gradient kernels a few times on a workload (for ncu / timing).
usage: python tools/bwd_prof.py [workload] ;  BWD_DENSE=0 selects the pair-list kernel"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp
from conan_fgw_b200 import ops
dev = "cuda"
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_lipo_train"
b = cmp.synthetic.make_config_batch(wl).to(dev)
n_max = int(torch.bincount(b.batch).max())
nl = cmp.build_neighbor_list(b.pos, b.batch, 10.0, max_atoms=n_max, num_graphs=b.num_graphs)
E = nl.E
torch.manual_seed(0)
blk = cmp.InteractionBlock(128, 50, 128, 10.0).to(dev)
gs = cmp.GaussianSmearing(0.0, 10.0, 50).to(dev)
W = (blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight)
x = torch.randn(b.z.numel(), 128, device=dev)
g = torch.randn(b.z.numel(), 128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for dense in ([True, False] if "BWD_DENSE" not in os.environ else [os.environ["BWD_DENSE"] != "0"]):
    ops.FUSED_DENSE_GRADS = dense
    ts = []
    with torch.no_grad(), ops.prepacked_weights([blk]):
        for i in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops._fused_weight_grads(g, x, *W, nl, gs.offset, gs.coeff, 10.0)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    fl = 45568.0 * E
    print(f"{wl}: weight gradients, {'dense-block' if dense else 'pair-list (incl. 2 bf16 conversions)'} kernel: median {ts[len(ts) // 2]:.1f} us "
          f"(incl. the partial reduction) = {fl / ts[len(ts) // 2] * 1e-6:.0f} TFLOP/s algorithmic", flush=True)

"""Ad-hoc: clock64 timeline of the dense forward kernel (CTA 0, pipeline 0) on a workload."""
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
_lib.lib().cmp_debug_set_dense_mode(int(os.environ.get("DENSE_MODE", "0")))
_lib.lib().cmp_debug_set_dense_stagger(int(os.environ.get("DENSE_STAGGER", "600")))
with torch.no_grad():
    for _ in range(2):
        ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, 10.0, False)
    buf = torch.zeros(512, dtype=torch.int64, device=dev)
    _lib.lib().cmp_debug_set_dense_timestamps(buf.data_ptr())
    ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, 10.0, False)
    torch.cuda.synchronize()
    _lib.lib().cmp_debug_set_dense_timestamps(None)
fine = buf.cpu()[256:272].tolist()
t = buf.cpu()[:256].view(32, 8)
names = ["masks+bar", "rbf", "bar+MMA1 wait", "ep1", "bar+MMA2 wait", "ep2"]
base = int(t[0][0])
for i in range(32):
    r = t[i]
    if r[0] == 0:
        break
    d = [int(r[k + 1] - r[k]) for k in range(6)]
    gap = int(r[0] - t[i - 1][6]) if i else 0
    print(i, "t0=%d" % (int(r[0]) - base), "npad=%d" % int(r[7]), "gap=%d" % gap, " ".join(f"{n}={v}" for n, v in zip(names, d)),
          "total=%d" % int(r[6] - r[0]))

print("ep1 of tile 1: cycles per 32-column iteration:", [fine[i + 1] - fine[i] for i in range(0, 4) if fine[i + 1]], "tail", fine[8] - fine[3] if fine[8] else None,
      "start->first iter", fine[0] - int(t[1][3]))

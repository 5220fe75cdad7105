"""Ad-hoc: clock64 timeline of the warp-specialised dense kernel (CTA 0: XU sets, MMA thread, EP2) on a workload."""
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
_lib.lib().cmp_debug_set_dense_mode(int(os.environ.get("DENSE_MODE", "0")))
with torch.no_grad():
    for _ in range(2):
        ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, 10.0, False)
    buf = torch.zeros(3 * 64 * 8 + 8, dtype=torch.int64, device=dev)
    _lib.lib().cmp_debug_set_dense_timestamps(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops._fused_aggregate(x, nl, W, gs.offset, gs.coeff, 10.0, False)
    e1.record()
    torch.cuda.synchronize()
    _lib.lib().cmp_debug_set_dense_timestamps(None)
g = buf.cpu()[1536:].tolist()
print(f"event-timed launch {e0.elapsed_time(e1) * 1e3:.1f} us | CTA 0: set-up {(g[1] - g[0]) / 1e3:.2f} us, body {(g[2] - g[1]) / 1e3:.2f} us | "
      f"last CTA: entry +{(g[4] - g[0]) / 1e3:.2f} us, set-up {(g[5] - g[4]) / 1e3:.2f} us, body {(g[6] - g[5]) / 1e3:.2f} us, exit +{(g[6] - g[0]) / 1e3:.2f} us")
t = buf.cpu()[:1536].view(3, 64, 8)
base = int(t[1][0][0])
r = lambda v: int(v) - base
print("cycles relative to the MMA thread's start; tile k is served by XU set k & 1")
for k in range(40):
    if t[0][k][5] == 0:
        break
    s, m, e = t[0][k], t[1][k], t[2][k]
    print(f"k={k:2d} npad={int(s[7]):3d} | Gaussians {r(s[6])} (stage) {r(s[5])} -> {r(s[4])} ({int(s[4] - s[5])}) | MMA1 issued {r(m[3])} | "
          f"epilogue 1: waits from {r(s[0])}, D1 at {r(s[1])}, A2 free {r(s[2])}, done {r(s[3])} ({int(s[3] - s[2])}) | "
          f"MMA2 issued {r(t[1][k + 2][1]) if k + 2 < 64 and t[1][k + 2][1] else '-'} | "
          f"EP2: waits from {r(e[0])}, D2 at {r(e[1])}, done {r(e[2])} ({int(e[2] - e[1])})")

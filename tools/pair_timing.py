"""Ad-hoc: clock64 timeline of the pair forward kernel (CTA 0, pipeline 0)."""
import sys, torch
sys.path.insert(0, ".")
import conan_fgw_b200 as cmp
from conan_fgw_b200 import _lib
dev = "cuda"
b = cmp.synthetic.make_config_batch("cfg2_lipo_train").to(dev)
torch.manual_seed(0)
m = cmp.SchNetNoSum(None).to(dev).set_precision("bf16")
with torch.no_grad():
    for _ in range(2): m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
    buf = torch.zeros(240, dtype=torch.int64, device=dev)
    _lib.lib().cmp_debug_set_pair_timestamps(buf.data_ptr())
    m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
    torch.cuda.synchronize()
    _lib.lib().cmp_debug_set_pair_timestamps(None)
t = buf.cpu().view(24, 10)
names = ["conf start->tile top", "group bar", "meta+rbf", "MMA1 wait", "ep1", "MMA2 wait", "xwait+ep2"]
base = int(t[0][0])
for i in range(12):
    r = t[i]
    if r[1] == 0: break
    d = [int(r[k+1]-r[k]) for k in range(7)]
    fin = int(r[8]-r[7]) if r[8] else 0
    print(i, "t0=%d" % (int(r[1])-base), " ".join(f"{n}={v}" for n, v in zip(names, d)), "->finalize done", fin, "| ep2 cycles in TMEM ld+wait:", int(r[9]))

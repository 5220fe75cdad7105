"""Ad-hoc: per-phase clock64 timestamps of the fused forward kernel (CTA 0, pipeline 0)."""
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
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    _lib.lib().cmp_debug_set_fwd_timestamps(buf.data_ptr())
    m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
    torch.cuda.synchronize()
    _lib.lib().cmp_debug_set_fwd_timestamps(None)
t = buf.cpu().view(32, 8)
names = ["barA+rowptr+barB", "meta+rbf", "fence+MMA1 wait", "ep1", "fence+MMA2 wait", "x wait+ep2", "loop"]
for i in range(12):
    r = t[i]
    if r[0] == 0: break
    d = [int(r[k+1]-r[k]) for k in range(6)]
    nxt = int(t[i+1][0]-r[6]) if t[i+1][0] else 0
    print(i, " ".join(f"{n}={v}" for n, v in zip(names, d)), "to-next-top=", nxt, "total", int(r[6]-r[0]))

# ---- backward weight-gradient kernel ----
m.zero_grad()
out = m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
buf2 = torch.zeros(240, dtype=torch.int64, device=dev)
_lib.lib().cmp_debug_set_bwd_timestamps(buf2.data_ptr())
out.pow(2).mean().backward()
torch.cuda.synchronize()
_lib.lib().cmp_debug_set_bwd_timestamps(None)
t = buf2.cpu().view(20, 12)
names = ["w_done wait", "meta+rbf", "bar+xwait", "dF build", "MMA1 wait", "ep1", "MMA da wait", "ep3"]
print("bwd kernel (last launch = block 0):")
for i in range(10):
    r = t[i]
    if r[0] == 0: break
    d = [int(r[k+1]-r[k]) for k in range(8)]
    print(i, " ".join(f"{n}={v}" for n, v in zip(names, d)), "total", int(r[8]-r[0]))

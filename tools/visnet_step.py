"""Ad-hoc: ConAN-ViSNet cfg 3 training step (dp.RegressionStep, CUDA graph): ms per step and C-ABI launches per step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp
from conan_fgw_b200 import _lib
from conan_fgw_b200.dp import RegressionStep
dev = "cuda"
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
b = cmp.synthetic.make_config_batch("cfg3_freesolv_visnet")
G = b.num_graphs
K = cmp.synthetic.CONFIGS["cfg3_freesolv_visnet"]["num_conformers"]
torch.manual_seed(0)
model = cmp.ViSNet(None, hidden_channels=H).to(dev).set_precision("bf16")
d = b.to(dev)
E = model.representation_model.distance.neighbor_list(d.pos, d.batch, G).E
tr = RegressionStep(model, H // 2, K, lr=1e-3, backbone_kwargs={"num_edges": E})
targets = torch.randn(G // K, 1, device=dev)
for _ in range(3):
    tr.step(d.z, d.pos, d.batch, targets, G)
torch.cuda.synchronize()
n0 = _lib.launches()
tr.step(d.z, d.pos, d.batch, targets, G)
torch.cuda.synchronize()
launches = _lib.launches() - n0
def timeit(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        tr.step(d.z, d.pos, d.batch, targets, G)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
eager = timeit(5)
tr.capture(d.z, d.pos, d.batch, targets, G)
tr.step(d.z, d.pos, d.batch, targets, G)
graph = timeit(20)
print(f"ViSNet H={H} N={d.z.numel()} E={E}: eager {eager:.2f} ms/step, CUDA graph {graph:.2f} ms/step = {G / graph * 1e3:.0f} conformers/s, "
      f"{launches} C-ABI launches/step")

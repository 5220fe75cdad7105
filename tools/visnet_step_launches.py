"""Ad-hoc: one ConAN-ViSNet cfg 3 TRAINING step (dp.RegressionStep, eager) between cudaProfilerStart / Stop for an ncu
launch list:  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv ..."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp
from conan_fgw_b200.dp import RegressionStep
dev = "cuda"
b = cmp.synthetic.make_config_batch("cfg3_freesolv_visnet")
G = b.num_graphs
K = cmp.synthetic.CONFIGS["cfg3_freesolv_visnet"]["num_conformers"]
torch.manual_seed(0)
model = cmp.ViSNet(None, hidden_channels=128).to(dev).set_precision("bf16")
d = b.to(dev)
E = model.representation_model.distance.neighbor_list(d.pos, d.batch, G).E
tr = RegressionStep(model, 64, K, lr=1e-3, backbone_kwargs={"num_edges": E})
targets = torch.randn(G // K, 1, device=dev)
for _ in range(3):
    tr.step(d.z, d.pos, d.batch, targets, G)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(d.z, d.pos, d.batch, targets, G)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("N", d.z.numel(), "E", E)

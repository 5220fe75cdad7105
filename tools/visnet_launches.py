"""Ad-hoc: one ViSNet training step (cfg 3) for an ncu launch list."""
import sys, torch
sys.path.insert(0, ".")
import conan_fgw_b200 as cmp
dev = "cuda"
b = cmp.synthetic.make_config_batch("cfg3_freesolv_visnet")
G = b.num_graphs
torch.manual_seed(0)
model = cmp.ViSNet(None, hidden_channels=128).to(dev).set_precision("bf16")
d = b.to(dev)
E = model.representation_model.distance.neighbor_list(d.pos, d.batch, G).E
for p_ in model.parameters():
    p_.grad = torch.zeros_like(p_)
def step():
    for p_ in model.parameters():
        p_.grad.zero_()
    model(d.z, d.pos, d.batch, num_graphs=G, num_edges=E).pow(2).mean().backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("N", d.z.numel(), "E", E)

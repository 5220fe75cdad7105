"""Ad-hoc profiling driver (not a test): a few bf16-mode forward(+backward) passes of cfg2 for ncu."""
import sys, torch
sys.path.insert(0, ".")
import conan_fgw_b200 as cmp
dev = "cuda"
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
b = cmp.synthetic.make_config_batch("cfg2_lipo_train").to(dev)
torch.manual_seed(0)
m = cmp.SchNetNoSum(None).to(dev).set_precision("bf16")
for _ in range(2):
    if bwd:
        m.zero_grad(); m(b.z, b.pos, b.batch, num_graphs=b.num_graphs).pow(2).mean().backward()
    else:
        with torch.no_grad():
            m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
torch.cuda.synchronize()
print("done")

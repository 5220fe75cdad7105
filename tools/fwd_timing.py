"""Ad-hoc timing helper (not a test): forward-only and fwd+bwd step times per precision mode."""
import sys, time, torch
sys.path.insert(0, ".")
import conan_fgw_b200 as cmp
dev = "cuda"
for cfgname in ("cfg2_lipo_train", "cfg4_bace_cls"):
    b = cmp.synthetic.make_config_batch(cfgname).to(dev)
    for prec in ("fp32", "bf16"):
        torch.manual_seed(0)
        m = cmp.SchNetNoSum(None).to(dev).set_precision(prec)
        G = b.num_graphs
        def fwd():
            with torch.no_grad():
                return m(b.z, b.pos, b.batch, num_graphs=G)
        def fb():
            m.zero_grad(); out = m(b.z, b.pos, b.batch, num_graphs=G); out.pow(2).mean().backward()
        for name, fn in (("fwd", fwd), ("fwd+bwd", fb)):
            for _ in range(3): fn()
            torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{cfgname} {prec} {name}: {ms:.3f} ms  {G/ms*1e3:.0f} conformers/s", flush=True)

# ViSNet, BASELINE.json configs[2]: FreeSolv-shaped, 32 x 5 conformers x 18 atoms, hidden 128, 6 layers
b = cmp.synthetic.make_config_batch("cfg3_freesolv_visnet").to(dev)
torch.manual_seed(0)
m = cmp.ViSNet(None, hidden_channels=128).to(dev)
G = b.num_graphs
def vfwd():
    with torch.no_grad():
        return m(b.z, b.pos, b.batch, num_graphs=G)
def vfb():
    m.zero_grad(); m(b.z, b.pos, b.batch, num_graphs=G).pow(2).mean().backward()
for name, fn in (("fwd", vfwd), ("fwd+bwd", vfb)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"cfg3_freesolv_visnet fp32 {name}: {ms:.3f} ms  {G/ms*1e3:.0f} conformers/s", flush=True)

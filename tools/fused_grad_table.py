"""Per-parameter gradient error of the fused (tcgen05) mode against the exact-fp32 GPU mode on the FULL cfg 2 batch
(640 conformers, 449 K edges, T = 6) - the table behind the stated tolerance of the fused mode."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp
dev = "cuda"
torch.manual_seed(0)
m = cmp.SchNetNoSum(None, hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=10.0).to(dev)
b = cmp.synthetic.make_config_batch("cfg2_lipo_train").to(dev)
want = m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
want.pow(2).mean().backward()
ref = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
m.set_precision("bf16")
m.max_atoms_hint = 27
m.zero_grad()
got = m(b.z, b.pos, b.batch, num_graphs=b.num_graphs)
got.pow(2).mean().backward()
m.check_status()
rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
rms = lambda a, r: float((a - r).pow(2).mean().sqrt() / r.pow(2).mean().sqrt())
print(f"embeddings [640, 64]: max-rel {rel(got, want):.2e}, rms-rel {rms(got, want):.2e}")
print("| parameter | max-rel | rms-rel |\n|---|---|---|")
rows = [(rel(p.grad, ref[k]), rms(p.grad, ref[k]), k) for k, p in m.named_parameters() if k in ref]
for e, r, k in sorted(rows, reverse=True):
    print(f"| `{k}` | {e:.2e} | {r:.2e} |")
print(f"worst: {max(rows)[0]:.2e} ({max(rows)[2]})")

"""Ad-hoc: per-parameter gradient error of the fused (tcgen05) mode against the CPU oracle on the smoke() inputs, with
and without the max_atoms hint (hint -> warp-specialised dense forward + dense-block weight gradients; no hint ->
per-pipeline dense forward + pair-list weight gradients with bf16 copies of g / x')."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp
from oracle import schnet as osn
dev = "cuda"
torch.manual_seed(0)
cfg = dict(hidden_channels=128, num_filters=128, num_interactions=2, num_gaussians=50, cutoff=10.0)
ref = osn.SchNetNoSum(None, **cfg)
model = cmp.SchNetNoSum(None, **cfg).to(dev)
model.load_state_dict(ref.state_dict(), strict=True)
b = cmp.synthetic.make_batch(4, 2, 26, seed=1)
want = ref(b.z, b.pos, b.batch)
want.pow(2).mean().backward()
model.set_precision("bf16")
for hint in (None, 26):
    model.max_atoms_hint = hint
    model.zero_grad()
    out = model(b.z.to(dev), b.pos.to(dev), b.batch.to(dev), num_graphs=b.num_graphs)
    out.pow(2).mean().backward()
    err = (out.detach().cpu() - want.detach()).abs().max().item() / want.detach().abs().max().item()
    rows = []
    for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        if q.grad is not None:
            rows.append(((p.grad.cpu() - q.grad).abs().max().item() / (q.grad.abs().max().item() + 1e-30), k))
    rows.sort(reverse=True)
    print(f"hint={hint}: fwd {err:.2e}; worst gradients:", ", ".join(f"{k} {e:.2e}" for e, k in rows[:5]))

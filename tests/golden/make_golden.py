"""Regenerate the committed golden fixtures.  Runs ONLY in the build container, where
/root/reference is mounted (the GPU box never sees the reference tree).

  python tests/golden/make_golden.py

Outputs (all small, committed):
  cfm_geometry.pt    real conformer coordinates from the reference's only fixture,
                     notebooks/data/cfm_log.pt (`node_feature`, `batch`: 10 conformers x 22 atoms)
  visnet_ref.pt      inputs, a randomised state_dict and the outputs / gradients obtained by running
                     the reference's OWN vendored ViSNet file
                     (conan_fgw/src/model/graph_embeddings/torch_geometric_visnet.py, unmodified,
                     through oracle/pyg_shim.py) with the forward of visnet.py:115-121,143-156
  schnet_oracle.pt   inputs, state_dict and outputs of oracle.schnet.SchNetNoSum (self-golden:
                     regression guard only - the SchNet arithmetic is un-vendored PyG, "parity
                     unpinned")
  radius_oracle.pt   edge lists of oracle.radius for truncating / non-truncating geometries
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyg_shim  # noqa: E402
from oracle import schnet as oschnet  # noqa: E402
from oracle.radius import radius_graph_ref  # noqa: E402

import conan_fgw_b200  # noqa: E402

syn = conan_fgw_b200.synthetic

REF = "/root/reference"


def randomise(module, seed):
    """Make every parameter non-trivial (biases, LayerNorm affine and atomref start at 0/1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() <= 1 or "atomref" in name:
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def main():
    # ---- real geometry -----------------------------------------------------------------
    cfm = torch.load(os.path.join(REF, "notebooks/data/cfm_log.pt"), weights_only=False)
    geom = {"pos": cfm["node_feature"].clone().float(), "batch": cfm["batch"].clone().long()}
    torch.save(geom, os.path.join(HERE, "cfm_geometry.pt"))

    # ---- ViSNet from the reference's own file ----------------------------------------------
    tgv = pyg_shim.load_reference_visnet(REF)
    torch.manual_seed(11)
    ref = tgv.ViSNet(hidden_channels=32, num_layers=3, num_heads=8, num_rbf=32, cutoff=5.0)
    randomise(ref, 12)
    b = syn.make_batch(3, 2, 9, seed=21)
    z = b.z.clone()
    z[::7] = 0  # exercise z = 0 too
    x, v = ref.representation_model(z, b.pos.clone(), b.batch)
    a = ref.prior_model(ref.output_model.pre_reduce(x, v) * ref.std, z)
    a_b = ref.prior_model_bary(ref.output_model_bary.pre_reduce(x, v) * ref.std, z)
    y = pyg_shim.scatter(a, b.batch, dim=0)
    loss = y.pow(2).mean() + 0.5 * a_b.pow(2).mean()
    loss.backward()
    torch.save({
        "config": dict(hidden_channels=32, num_layers=3, num_heads=8, num_rbf=32, cutoff=5.0),
        "state_dict": {k: t.detach().clone() for k, t in ref.state_dict().items()},
        "z": z, "pos": b.pos, "batch": b.batch,
        "x_repr": x.detach(), "vec_repr": v.detach(), "per_atom": a.detach(), "per_atom_bary": a_b.detach(),
        "y": y.detach(), "loss": loss.detach(),
        "grads": {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None},
    }, os.path.join(HERE, "visnet_ref.pt"))

    # ---- SchNet oracle self-golden -----------------------------------------------------------
    torch.manual_seed(5)
    m = oschnet.SchNetNoSum(None, hidden_channels=32, num_filters=32, num_interactions=2, num_gaussians=20,
                            cutoff=6.0)
    randomise(m, 6)
    b2 = syn.make_batch(2, 3, 12, seed=22)
    out = m(b2.z, b2.pos, b2.batch)
    h, hb = m.forward_3d_bary(b2.z, b2.pos, b2.batch)
    (out.pow(2).mean()).backward()
    torch.save({
        "config": dict(hidden_channels=32, num_filters=32, num_interactions=2, num_gaussians=20, cutoff=6.0),
        "state_dict": {k: t.detach().clone() for k, t in m.state_dict().items()},
        "z": b2.z, "pos": b2.pos, "batch": b2.batch, "out": out.detach(), "h": h.detach(), "h_bary": hb.detach(),
        "grads": {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None},
    }, os.path.join(HERE, "schnet_oracle.pt"))

    # ---- radius graphs ---------------------------------------------------------------------------
    cases = {}
    for name, (B, K, n, r, loop) in {
        "dense26_r10": (2, 2, 26, 10.0, False),
        "trunc65_r10": (1, 2, 65, 10.0, False),
        "loop18_r5": (2, 2, 18, 5.0, True),
        "trunc65_r10_loop": (1, 1, 65, 10.0, True),
    }.items():
        bb = syn.make_batch(B, K, n, seed=31)
        cases[name] = dict(pos=bb.pos, batch=bb.batch, r=r, loop=loop,
                           edge_index=radius_graph_ref(bb.pos, r, bb.batch, loop, 32).to(torch.int32))
    cases["cfm_r10"] = dict(pos=geom["pos"], batch=geom["batch"], r=10.0, loop=False,
                            edge_index=radius_graph_ref(geom["pos"], 10.0, geom["batch"], False, 32).to(torch.int32))
    cases["cfm_r3_loop"] = dict(pos=geom["pos"], batch=geom["batch"], r=3.0, loop=True,
                                edge_index=radius_graph_ref(geom["pos"], 3.0, geom["batch"], True, 32).to(torch.int32))
    torch.save(cases, os.path.join(HERE, "radius_oracle.pt"))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

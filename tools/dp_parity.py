"""Hardware data-parallel correctness check (SURVEY.md 4 "Multi-GPU", trainer.py:308-325, datamodules.py:40-41).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_parity.py

Every rank builds the same model, takes its molecules by the DistributedSampler(shuffle=False) rule (dp.shard_molecules)
and runs dp.RegressionStep over NCCL.  Rank 0 also runs the SAME model on the whole batch alone.  Checked:
  1. the flat gradient after the all-reduce, times 1 / W, equals the 1-GPU gradient on the concatenated batch;
  2. the loss trajectory of 3 Adam steps (mean of the rank losses) equals the 1-GPU trajectory;
both in the exact-fp32 mode (tolerance 1e-5 relative) and in the fused tcgen05 mode (bf16 tolerance 5e-3)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conan_fgw_b200 as cmp  # noqa: E402
from conan_fgw_b200.dp import RegressionStep, shard_molecules  # noqa: E402

CFG = dict(hidden_channels=128, num_filters=128, num_interactions=3, num_gaussians=50, cutoff=10.0)
B, K, NATOMS = 16, 5, 27


def take(batch, mols):
    """The conformers of the molecules `mols` (in that order) as a new sorted batch."""
    n, k = batch.atoms_per_conformer, batch.num_conformers
    idx = torch.cat([torch.arange(m * k * n, (m + 1) * k * n) for m in mols])
    z, pos = batch.z[idx], batch.pos[idx]
    bt = torch.arange(len(mols) * k).repeat_interleave(n)
    return z, pos, bt, len(mols) * k


def run(precision, dev, rank, world, full, targets):
    def trainer():
        torch.manual_seed(0)
        model = cmp.SchNetNoSum(None, **CFG).to(dev).set_precision(precision)
        model.max_atoms_hint = NATOMS
        return RegressionStep(model, CFG["hidden_channels"] // 2, K, lr=1e-3)

    mine = shard_molecules(B, rank, world)
    z, pos, bt, G = take(full, mine)
    z, pos, bt = z.to(dev), pos.to(dev), bt.to(dev)
    tg = targets[mine].to(dev)
    out = {}
    # ---- 1. gradient equivalence ----
    tr = trainer()
    tr._fwd_bwd(z, pos, bt, tg, G)
    w = tr.flat.all_reduce()
    g_dp = (tr.flat.grad / w).clone()
    # ---- 2. three Adam steps ----
    tr = trainer()
    losses_dp = []
    for _ in range(3):
        loss = tr.step(z, pos, bt, tg, G).clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        losses_dp.append(float(loss.item()) / world)
    if rank == 0:
        zf, pf, bf, Gf = take(full, list(range(B)))
        zf, pf, bf, tf = zf.to(dev), pf.to(dev), bf.to(dev), targets.to(dev)
        one = trainer()
        one.flat.all_reduce = lambda group=None: 1          # the 1-GPU run: no collective
        one._fwd_bwd(zf, pf, bf, tf, Gf)
        g_one = one.flat.grad.clone()
        one = trainer()
        one.flat.all_reduce = lambda group=None: 1
        losses_one = [float(one.step(zf, pf, bf, tf, Gf).item()) for _ in range(3)]
        out = {"precision": precision, "world": world, "molecules": B,
               "grad_rel_err": float((g_dp - g_one).abs().max() / g_one.abs().max()),
               "grad_rms_rel_err": float((g_dp - g_one).pow(2).mean().sqrt() / g_one.pow(2).mean().sqrt()),
               "losses_dp": losses_dp, "losses_1gpu": losses_one,
               "loss_rel_err": max(abs(a - b) / abs(b) for a, b in zip(losses_dp, losses_one))}
    dist.barrier()
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    full = cmp.synthetic.make_batch(B, K, NATOMS, seed=77)
    targets = torch.randn(B, 1, generator=torch.Generator().manual_seed(5))
    results = []
    for precision, tol in (("fp32", 1e-5), ("bf16", 5e-3)):
        r = run(precision, dev, rank, world, full, targets)
        if rank == 0:
            r["tolerance"] = tol
            r["ok"] = r["grad_rel_err"] < tol and r["loss_rel_err"] < tol
            results.append(r)
            print(json.dumps(r), flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(results, open("gpurun_out/dp_parity.json", "w"), indent=1)
        assert all(r["ok"] for r in results), "data-parallel gradients / losses differ from the 1-GPU run"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

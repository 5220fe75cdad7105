"""Synthetic conformer batches with the layout ConAN's collate produces.

Layout contract (reference ``conan_fgw/src/data/datasets.py:171-199`` and
``data/conformers/features.py:196-205``): ``z int64[N]`` atomic numbers,
``pos float32[N, 3]`` coordinates in Angstrom, ``batch int64[N]`` = conformer id
``0..G-1`` sorted non-decreasing; the K conformers of one molecule are
consecutive and share ``n`` and ``z``; ``conformers_index int64[G]`` maps a
conformer to its molecule (``model/common.py:414-423``).

Geometry follows SURVEY.md 8(d): conformer 0 = atoms uniform in a cube of side
``(n / 0.10)**(1/3)`` Angstrom with a 0.9 Angstrom minimum separation,
conformers 1..K-1 = conformer 0 + N(0, 0.3 Angstrom) per coordinate.
Everything is generated on the CPU from one ``torch.Generator`` so that the CPU
oracle and the GPU path see identical tensors.
"""

from __future__ import annotations

import dataclasses

import torch

# atomic numbers and their sampling weights: H, C, N, O, F, S, Cl
_Z = torch.tensor([1, 6, 7, 8, 9, 16, 17], dtype=torch.int64)
_W = torch.tensor([0.50, 0.35, 0.05, 0.08, 0.02 / 3, 0.02 / 3, 0.02 / 3], dtype=torch.float64)


@dataclasses.dataclass
class ConformerBatch:
    z: torch.Tensor                  # int64 [N]
    pos: torch.Tensor                # float32 [N, 3]
    batch: torch.Tensor              # int64 [N]   conformer id, sorted
    conformers_index: torch.Tensor   # int64 [G]   molecule id of each conformer
    num_molecules: int
    num_conformers: int              # K
    atoms_per_conformer: int         # n (fixed-size variant)

    @property
    def num_graphs(self) -> int:
        return self.num_molecules * self.num_conformers

    def to(self, device, non_blocking=False):
        return dataclasses.replace(
            self,
            z=self.z.to(device, non_blocking=non_blocking),
            pos=self.pos.to(device, non_blocking=non_blocking),
            batch=self.batch.to(device, non_blocking=non_blocking),
            conformers_index=self.conformers_index.to(device, non_blocking=non_blocking),
        )

    def pin(self):
        return dataclasses.replace(
            self, z=self.z.pin_memory(), pos=self.pos.pin_memory(), batch=self.batch.pin_memory(),
            conformers_index=self.conformers_index.pin_memory())


def make_batch(num_molecules: int, num_conformers: int, atoms: int, seed: int = 1234,
               density: float = 0.10, min_dist: float = 0.9, noise: float = 0.3) -> ConformerBatch:
    """Build one batch of ``num_molecules x num_conformers`` conformers with ``atoms`` atoms each."""
    g = torch.Generator().manual_seed(seed)
    B, K, n = int(num_molecules), int(num_conformers), int(atoms)
    side = (n / density) ** (1.0 / 3.0)

    z_mol = _Z[torch.multinomial(_W, B * n, replacement=True, generator=g)].view(B, n)

    # sequential insertion with rejection, vectorised over molecules
    base = torch.zeros(B, n, 3, dtype=torch.float32)
    for a in range(n):
        cand = torch.rand(B, 3, generator=g) * side
        if a > 0:
            for _ in range(64):
                d = (base[:, :a, :] - cand[:, None, :]).norm(dim=-1).min(dim=1).values
                bad = d < min_dist
                if not bool(bad.any()):
                    break
                fresh = torch.rand(B, 3, generator=g) * side
                cand = torch.where(bad[:, None], fresh, cand)
        base[:, a, :] = cand

    pos = base[:, None, :, :].repeat(1, K, 1, 1)
    if K > 1:
        pos[:, 1:] += noise * torch.randn(B, K - 1, n, 3, generator=g)

    G = B * K
    return ConformerBatch(
        z=z_mol[:, None, :].expand(B, K, n).reshape(-1).contiguous(),
        pos=pos.reshape(-1, 3).contiguous(),
        batch=torch.arange(G, dtype=torch.int64).repeat_interleave(n),
        conformers_index=torch.arange(B, dtype=torch.int64).repeat_interleave(K),
        num_molecules=B, num_conformers=K, atoms_per_conformer=n,
    )


# The BASELINE.json configurations (SURVEY.md 8, table at the top of the section).
CONFIGS = {
    "cfg1_esol_fwd": dict(num_molecules=32, num_conformers=5, atoms=26, cutoff=10.0),
    "cfg2_lipo_train": dict(num_molecules=128, num_conformers=5, atoms=27, cutoff=10.0),
    "cfg3_freesolv_visnet": dict(num_molecules=32, num_conformers=5, atoms=18, cutoff=5.0),
    "cfg4_bace_cls": dict(num_molecules=256, num_conformers=10, atoms=65, cutoff=10.0),
    "cfg5_cov2_stress": dict(num_molecules=1024, num_conformers=20, atoms=45, cutoff=10.0),
}


def make_config_batch(name: str, seed: int = 1234, scale: float = 1.0) -> ConformerBatch:
    c = CONFIGS[name]
    B = max(1, int(round(c["num_molecules"] * scale)))
    return make_batch(B, c["num_conformers"], c["atoms"], seed=seed)

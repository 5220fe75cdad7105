"""Oracle (test infrastructure): neighbour search with torch-cluster semantics.

Restates ``torch_cluster.radius_graph`` 1.6.1 as reached through
``torch_geometric.nn.radius_graph`` - the function the reference calls at
``conan_fgw/src/model/graph_embeddings/schnet_no_sum.py:160`` (through PyG's
``RadiusInteractionGraph``) and ``torch_geometric_visnet.py:331-337``.
torch-cluster is an un-vendored dependency (``environment.yml:162``), so the
rule below is the published upstream algorithm:

* queries and candidates come from the same conformer (``batch`` sorted);
* candidate ``j`` matches query ``i`` iff ``d2(i, j) < float32(r * r)`` with
  ``d2 = ((dx*dx) + (dy*dy)) + (dz*dz)`` evaluated in float32, every product
  and sum individually rounded (no FMA contraction);
* CUDA truncation rule: per query keep the first ``cap`` matches in ascending
  ``j`` (self included), ``cap = max_num_neighbors + (0 if loop else 1)``;
* self pairs are dropped afterwards when ``loop=False``;
* result ``edge_index[0] = j`` (source), ``edge_index[1] = i`` (target),
  grouped by ``i`` ascending, ``j`` ascending inside a group.
"""

from __future__ import annotations

import numpy as np
import torch


def _segments(batch: np.ndarray):
    """Start offsets of each conformer (batch must be sorted non-decreasing)."""
    if batch.size == 0:
        return np.zeros(1, dtype=np.int64)
    if np.any(np.diff(batch) < 0):
        raise ValueError("radius_graph: 'batch' must be sorted non-decreasing")
    num = int(batch.max()) + 1
    # ptr[g] = first atom with batch >= g  (bucketize semantics; empty ids allowed)
    return np.searchsorted(batch, np.arange(num + 1), side="left").astype(np.int64)


def radius_graph_ref(
    pos: torch.Tensor,
    r: float,
    batch: torch.Tensor | None = None,
    loop: bool = False,
    max_num_neighbors: int = 32,
    flow: str = "source_to_target",
) -> torch.Tensor:
    """Return ``edge_index int64[2, E]`` exactly as torch-cluster's CUDA path would."""
    assert flow in ("source_to_target", "target_to_source")
    p = pos.detach().cpu().to(torch.float32).contiguous().numpy()
    n_atoms = p.shape[0]
    if batch is None:
        b = np.zeros(n_atoms, dtype=np.int64)
    else:
        b = batch.detach().cpu().numpy().astype(np.int64)
        if b.shape[0] != n_atoms:
            raise ValueError("radius_graph: batch and pos disagree on the atom count")
    ptr = _segments(b)
    cap = int(max_num_neighbors) if loop else int(max_num_neighbors) + 1
    r2 = np.float32(float(r) * float(r))

    src_parts, dst_parts = [], []
    # group conformers of equal size so the dense d2 blocks can be batched
    sizes = np.diff(ptr)
    for n in np.unique(sizes):
        n = int(n)
        if n == 0:
            continue
        starts = ptr[:-1][sizes == n]
        for c0 in range(0, len(starts), 4096):
            st = starts[c0 : c0 + 4096]
            idx = st[:, None] + np.arange(n)[None, :]              # [g, n] global ids
            q = p[idx]                                             # [g, n, 3]
            dx = q[:, :, None, 0] - q[:, None, :, 0]               # [g, i, j]; sign irrelevant after squaring
            dy = q[:, :, None, 1] - q[:, None, :, 1]
            dz = q[:, :, None, 2] - q[:, None, :, 2]
            d2 = ((dx * dx) + (dy * dy)) + (dz * dz)               # float32, individually rounded
            hit = d2 < r2
            rank = np.cumsum(hit, axis=2)                          # 1-based rank among matches, ascending j
            keep = hit & (rank <= cap)
            if not loop:
                eye = np.eye(n, dtype=bool)[None]
                keep = keep & ~eye
            g_i, i_i, j_i = np.nonzero(keep)                       # row-major: conformer, i, j ascending
            dst_parts.append(idx[g_i, i_i])
            src_parts.append(idx[g_i, j_i])
    if src_parts:
        src = np.concatenate(src_parts)
        dst = np.concatenate(dst_parts)
        order = np.lexsort((src, dst))                             # destination-major, source ascending
        src, dst = src[order], dst[order]
    else:
        src = np.zeros(0, dtype=np.int64)
        dst = np.zeros(0, dtype=np.int64)
    if flow == "source_to_target":
        ei = np.stack([src, dst])
    else:
        ei = np.stack([dst, src])
    return torch.from_numpy(ei.astype(np.int64))


def radius_interaction_graph_ref(pos, batch, cutoff=10.0, max_num_neighbors=32):
    """PyG ``RadiusInteractionGraph.forward``: edges + Euclidean lengths."""
    ei = radius_graph_ref(pos, cutoff, batch, loop=False, max_num_neighbors=max_num_neighbors)
    row, col = ei[0], ei[1]
    p = pos.detach().cpu()
    ew = (p[row] - p[col]).norm(dim=-1)
    return ei, ew


def radius_graph_loops_py(pos, r, batch, loop, max_num_neighbors):
    """Literal per-query loop of the CUDA kernel (pure Python; small inputs only).

    Used by the tests to cross-check the vectorised version above.
    """
    p = pos.detach().cpu().to(torch.float32).numpy()
    b = np.zeros(len(p), dtype=np.int64) if batch is None else batch.cpu().numpy()
    ptr = _segments(b)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    r2 = np.float32(float(r) * float(r))
    rows, cols = [], []
    for i in range(len(p)):
        g = int(b[i])
        count = 0
        for j in range(int(ptr[g]), int(ptr[g + 1])):
            dx = np.float32(p[j, 0] - p[i, 0])
            dy = np.float32(p[j, 1] - p[i, 1])
            dz = np.float32(p[j, 2] - p[i, 2])
            d2 = np.float32(np.float32(np.float32(dx * dx) + np.float32(dy * dy)) + np.float32(dz * dz))
            if d2 < r2:
                rows.append(i)
                cols.append(j)
                count += 1
            if count >= cap:
                break
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    src, dst = cols, rows
    if not loop:
        m = src != dst
        src, dst = src[m], dst[m]
    return torch.from_numpy(np.stack([src, dst]))

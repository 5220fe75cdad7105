"""ORACLE (test infrastructure only) - ``torch_geometric.utils.to_dense_batch`` / ``to_dense_adj`` (PyG 2.3.0) as
ConAN calls them (``schnet_no_sum.py:242-253``).  Parity unpinned against PyG (un-vendored); the semantics restated
here: ``out[b, i - ptr[b]] = x[i]``, ``mask`` marks real nodes; ``adj[batch[row], row - ptr, col - ptr] += 1`` with
``row = edge_index[0]``."""
import torch


def to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    if batch is None:
        batch = torch.zeros(x.size(0), dtype=torch.long)
    B = int(batch.max()) + 1 if batch_size is None else batch_size
    counts = torch.bincount(batch, minlength=B)
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    n_max = int(counts.max()) if max_num_nodes is None else max_num_nodes
    out = x.new_full((B, n_max, x.size(1)), fill_value)
    mask = torch.zeros(B, n_max, dtype=torch.bool)
    for i in range(x.size(0)):
        b = int(batch[i]); a = i - int(ptr[b])
        if a < n_max:
            out[b, a] = x[i]
            mask[b, a] = True
    return out, mask


def to_dense_adj(edge_index, batch=None, max_num_nodes=None, batch_size=None):
    if batch is None:
        n = int(edge_index.max()) + 1 if edge_index.numel() else 0
        batch = torch.zeros(n, dtype=torch.long)
    B = int(batch.max()) + 1 if batch_size is None else batch_size
    counts = torch.bincount(batch, minlength=B)
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    n_max = int(counts.max()) if max_num_nodes is None else max_num_nodes
    adj = torch.zeros(B, n_max, n_max)
    for e in range(edge_index.size(1)):
        r, c = int(edge_index[0, e]), int(edge_index[1, e])
        b = int(batch[r]); a, d = r - int(ptr[b]), c - int(ptr[b])
        if 0 <= a < n_max and 0 <= d < n_max:
            adj[b, a, d] += 1
    return adj

"""Oracle tooling: the three PyG symbols the reference's vendored ViSNet imports.

``torch_geometric_visnet.py:9-10`` needs ``torch_geometric.nn.MessagePassing``,
``torch_geometric.nn.radius_graph`` and ``torch_geometric.utils.scatter``.  PyG is
not installable here, so ``install()`` registers a minimal stand-in under those
module names; the reference file can then be imported *unmodified* from
``/root/reference`` in the build container to pin ``oracle/visnet.py`` and to
produce ``tests/golden/visnet_*.pt``.  Never used on the GPU box (the reference
tree does not travel) and never imported by the product.

Semantics follow SURVEY.md A.4: ``flow='source_to_target'``; ``foo_j`` gathers
``foo`` at ``edge_index[0]``, ``foo_i`` at ``edge_index[1]`` along ``node_dim``;
outputs of ``message`` are handed to ``aggregate(features, index, ptr, dim_size)``
when the subclass overrides it, else summed at ``edge_index[1]``.
"""

from __future__ import annotations

import inspect
import sys
import types

import torch

from .radius import radius_graph_ref


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    out = src.new_zeros(shape).index_add_(dim, index, src)
    if reduce in ("sum", "add"):
        return out
    if reduce == "mean":
        cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
        view = [1] * src.dim()
        view[dim] = -1
        return out / cnt.view(view)
    raise NotImplementedError(reduce)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", node_dim=-2, **kwargs):
        super().__init__()
        assert aggr == "add"
        self.node_dim = node_dim

    def _collect(self, fn, edge_index, kwargs):
        args = {}
        for name in list(inspect.signature(fn).parameters):
            if name.endswith("_j") or name.endswith("_i"):
                t = kwargs[name[:-2]]
                sel = edge_index[0] if name.endswith("_j") else edge_index[1]
                dim = self.node_dim if self.node_dim >= 0 else t.dim() + self.node_dim
                args[name] = t.index_select(dim, sel)
            else:
                args[name] = kwargs[name]
        return args

    def propagate(self, edge_index, size=None, **kwargs):
        out = self.message(**self._collect(self.message, edge_index, kwargs))
        first = next(v for v in kwargs.values() if torch.is_tensor(v))
        ref = kwargs.get("x", first)
        dim = self.node_dim if self.node_dim >= 0 else ref.dim() + self.node_dim
        dim_size = ref.size(dim)
        if type(self).aggregate is not MessagePassing.aggregate:
            return self.aggregate(out, edge_index[1], None, dim_size)
        return scatter(out, edge_index[1], dim=dim, dim_size=dim_size)

    def aggregate(self, inputs, index, ptr=None, dim_size=None):  # default: sum
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size)

    def edge_updater(self, edge_index, **kwargs):
        return self.edge_update(**self._collect(self.edge_update, edge_index, kwargs))


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target",
                 num_workers=1):
    return radius_graph_ref(x, r, batch, loop, max_num_neighbors, flow).to(x.device)


def install():
    """Register the stand-in as ``torch_geometric`` (no-op if a real PyG is importable)."""
    try:
        import torch_geometric  # noqa: F401
        return False
    except ModuleNotFoundError:
        pass
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_nn.MessagePassing = MessagePassing
    tg_nn.radius_graph = radius_graph
    tg_utils.scatter = scatter
    tg.nn, tg.utils = tg_nn, tg_utils
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.nn"] = tg_nn
    sys.modules["torch_geometric.utils"] = tg_utils
    return True


def load_reference_visnet(reference_root="/root/reference"):
    """Import the reference's vendored ViSNet file as a module (build container only)."""
    import importlib.util
    import os

    path = os.path.join(reference_root, "conan_fgw/src/model/graph_embeddings/torch_geometric_visnet.py")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    install()
    spec = importlib.util.spec_from_file_location("_ref_tgv", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

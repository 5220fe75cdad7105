"""Neighbour lists: the drop-in for ``radius_graph`` / ``RadiusInteractionGraph``.

Reference call sites replaced (paths in the reference tree):
``conan_fgw/src/model/graph_embeddings/schnet_no_sum.py:160,208,342`` (through
PyG ``RadiusInteractionGraph``), ``torch_geometric_visnet.py:331-347``
(``Distance``), ``visnet.py:90,276``.

The CUDA kernels emit a destination-sorted CSR (``rowptr``/``col``/``dist``) plus
its source-sorted transpose.  ``radius_graph`` keeps PyG's signature and returns
the PyG-format ``edge_index`` (canonical order: target-major, source ascending =
the order torch-cluster's CUDA kernel produces); the returned tensor carries the
CSR as ``edge_index._cmp_graph`` so that downstream modules of this package do
not rebuild it.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def raise_for_status(status: torch.Tensor, reset: bool = False):
    """Read a device-side status word (one device -> host copy: synchronises) and raise for the errors it reports.
    The lazily built tile / pair lists and the fused kernels report through the same word as the radius search, so it is
    read every time this is called; sync-free callers (CUDA-graph steps) call it where they synchronise anyway."""
    s = int(status.item())
    if reset and s:
        status.zero_()
    if s & _lib.STATUS_UNSORTED_BATCH:
        raise ValueError("radius_graph: 'batch' must be sorted non-decreasing with ids in [0, num_graphs)")
    if s & _lib.STATUS_EDGE_OVERFLOW:
        raise RuntimeError("radius_graph: neighbour / tile capacity overflow, an edge count different from the promised "
                           "num_edges, or a conformer above the promised max_atoms bound")
    if s & _lib.STATUS_BAD_ATOMIC_NUMBER:
        raise ValueError("atomic numbers must lie in [0, 100)")


class NeighborList:
    """Device-resident CSR neighbour list of a batch of conformers."""

    def __init__(self, N, G, cap_E, device, status=None, zero_fill=False):
        i32 = dict(dtype=torch.int32, device=device)
        self._alloc = torch.zeros if zero_fill else torch.empty
        self.N, self.G, self.cap_E = int(N), int(G), int(cap_E)
        self.seg_ptr = torch.empty(G + 1, **i32)
        self.conf_edge_ptr = torch.empty(G + 1, **i32)
        self.rowptr = torch.empty(N + 1, **i32)
        # zero_fill: with a promised edge count (num_edges) kernels run over the promised number of rows; if the real
        # geometry has fewer edges the surplus rows must still hold valid atom indices (the mismatch itself is reported
        # through the status word, never as an out-of-bounds access)
        self.col = self._alloc(max(cap_E, 1), **i32)
        self.dist = self._alloc(max(cap_E, 1), dtype=torch.float32, device=device)
        self.evec = None
        self.rowptr_t = self.col_t = self.eid_t = None
        # device-side error word (CMP_STATUS_* bits); a model passes its own persistent word so that a sync-free training
        # step can read it where it synchronises anyway (nn.SchNet.check_status / dp.RegressionStep.check)
        self.status = status if status is not None else torch.zeros(1, **i32)
        self.cutoff = None
        self.loop = False
        self.pos = None              # float32 [N, 3] the list was built from (radius-built lists only)
        self.max_atoms = None        # caller's bound on atoms per conformer (None: unknown)
        self.sym_atoms = 0           # 0: nothing is known about symmetry (graph built from a raw edge_index)
        self._E = None
        self._edge_index = None
        self._checked = False

    @property
    def E(self) -> int:
        """Edge count (first access synchronises with the device)."""
        if self._E is None:
            self._E = int(self.rowptr[self.N].item()) if self.N > 0 else 0
            self.check()
        return self._E

    def check(self):
        raise_for_status(self.status)

    def edge_index(self) -> torch.Tensor:
        if self._edge_index is None:
            E = self.E
            ei = torch.empty(2, E, dtype=torch.int64, device=self.rowptr.device)
            if E:
                _lib.call("cmp_csr_to_edge_index", _lib.ptr(self.rowptr), _lib.ptr(self.col), self.N, E, _lib.ptr(ei))
            ei._cmp_graph = self
            self._edge_index = ei
        return self._edge_index

    # ---- edge tiles for the fused tensor-core kernels (built on first use, no host sync) ----------
    def _build_tiles(self, rowptr, min_atoms=0):
        dev = rowptr.device
        tile_e = _lib.size_query("cmp_cfconv_tc_tile_edges")
        cap = self.cap_E // 32 + self.G + 1
        tiles = torch.empty(max(cap, 1), 8, dtype=torch.int32, device=dev)
        num = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = _lib.workspace(_lib.size_query("cmp_build_tiles_workspace", self.G), dev)
        _lib.call("cmp_build_tiles_min_atoms", _lib.ptr(rowptr), _lib.ptr(self.seg_ptr), self.G, tile_e, int(min_atoms),
                  _lib.ptr(tiles), cap, _lib.ptr(num), _lib.ptr(ws), ws.numel(), _lib.ptr(self.status))
        return tiles, num

    def tiles(self, min_atoms=0):
        """(tiles int32[cap,8], num_tiles int32[1]) over the target-sorted CSR; conformers with fewer than
        ``min_atoms`` atoms are left out (they go to the pair kernel)."""
        cache = self.__dict__.setdefault("_tiles", {})
        if min_atoms not in cache:
            if self.G == 0 and self.N > 0:
                raise _lib.ConanMPError("edge tiles need conformer segments (graph was built from a raw edge_index)")
            cache[min_atoms] = self._build_tiles(self.rowptr, min_atoms)
        return cache[min_atoms]

    def tiles_t(self, min_atoms=0):
        """Same for the source-sorted transpose, plus ``dist_t`` (distances in transposed edge order)."""
        cache = self.__dict__.setdefault("_tiles_t", {})
        if min_atoms not in cache:
            if self.rowptr_t is None:
                raise _lib.ConanMPError("the transposed neighbour list was not built")
            tiles, num = self._build_tiles(self.rowptr_t, min_atoms)
            if getattr(self, "_dist_t", None) is None:
                self._dist_t = torch.empty_like(self.dist)
                _lib.call("cmp_gather_f32", _lib.ptr(self.dist), _lib.ptr(self.eid_t), _lib.ptr(self.rowptr[self.N:]),
                          self.cap_E, _lib.ptr(self._dist_t))
            cache[min_atoms] = (tiles, num, self._dist_t)
        return cache[min_atoms]

    def adjacency(self) -> torch.Tensor:
        """uint32 [N, 4] bit matrix per conformer (bit j of row i: edge j -> i, conformer-local indices) for the dense
        fused CFConv kernels (``cmp_build_adjacency``; conformers above 128 atoms are left out).  Built once."""
        if getattr(self, "_adj", None) is None:
            if self.G == 0 and self.N > 0:
                raise _lib.ConanMPError("the adjacency matrix needs conformer segments (graph built from a raw edge_index)")
            adj = torch.empty(max(self.N, 1), 4, dtype=torch.int32, device=self.rowptr.device)
            _lib.call("cmp_build_adjacency", _lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.seg_ptr), self.N,
                      self.G, _lib.ptr(adj))
            self._adj = adj
            self._counter = torch.zeros(4, dtype=torch.int32, device=adj.device)
        return self._adj

    def dense_bwd_tiles(self) -> torch.Tensor:
        """int32 [G + 1]: tiles of the dense weight-gradient kernel before every conformer (``cmp_build_dense_bwd_tiles``;
        a conformer above 128 atoms raises the status bit).  Built once per neighbour list, shared by all blocks."""
        if getattr(self, "_dense_bwd_tiles", None) is None:
            tp = torch.empty(self.G + 1, dtype=torch.int32, device=self.rowptr.device)
            _lib.call("cmp_build_dense_bwd_tiles", _lib.ptr(self.seg_ptr), self.G, _lib.ptr(tp), _lib.ptr(self.status))
            self._dense_bwd_tiles = tp
        return self._dense_bwd_tiles

    def erow(self) -> torch.Tensor:
        """Target atom of every edge (``edge_index[1]`` as int32), built once."""
        if getattr(self, "_erow", None) is None:
            erow = self._alloc(self.col.shape, dtype=self.col.dtype, device=self.col.device)
            _lib.call("cmp_csr_expand_rows", _lib.ptr(self.rowptr), self.N, _lib.ptr(erow))
            self._erow = erow
        return self._erow

    def flat_tiles(self):
        """(erow int32[cap_E], tiles int32[cap,8], num_tiles int32[1]): 64-edge chunks per conformer for the
        fused weight-gradient kernel (edges are independent there, so no row alignment)."""
        if getattr(self, "_flat_tiles", None) is None:
            if self.G == 0 and self.N > 0:
                raise _lib.ConanMPError("edge tiles need conformer segments (graph was built from a raw edge_index)")
            dev = self.rowptr.device
            erow = self.erow()
            tile_e = _lib.size_query("cmp_cfconv_tc_bwd_tile_edges")
            cap = self.cap_E // tile_e + self.G + 1
            tiles = torch.empty(max(cap, 1), 8, dtype=torch.int32, device=dev)
            num = torch.zeros(1, dtype=torch.int32, device=dev)
            ws = _lib.workspace(_lib.size_query("cmp_build_flat_tiles_workspace", self.G), dev)
            _lib.call("cmp_build_flat_tiles", _lib.ptr(self.conf_edge_ptr), _lib.ptr(self.seg_ptr), _lib.ptr(erow),
                      self.G, tile_e, _lib.ptr(tiles), cap, _lib.ptr(num), _lib.ptr(ws), ws.numel(),
                      _lib.ptr(self.status))
            self._flat_tiles = (erow, tiles, num)
        return self._flat_tiles

    def pair_tiles(self):
        """(pair_src, pair_dst, pair_dist, pair_rev, tiles, num_tiles): one primary edge per undirected pair
        (``cmp_build_pair_list``) cut into 64-column chunks per conformer - the work list of the pair mode of the
        fused weight-gradient kernel (the filter is shared by both directions of a pair)."""
        if getattr(self, "_pair_tiles", None) is None:
            if self.G == 0 and self.N > 0:
                raise _lib.ConanMPError("pair tiles need conformer segments (graph was built from a raw edge_index)")
            dev = self.rowptr.device
            cap = max(self.cap_E, 1)      # every edge may be unpaired
            src = torch.empty(cap, dtype=torch.int32, device=dev)
            dst = torch.empty(cap, dtype=torch.int32, device=dev)
            dist = torch.empty(cap, dtype=torch.float32, device=dev)
            rev = torch.empty(cap, dtype=torch.int32, device=dev)
            conf_ptr = torch.empty(self.G + 1, dtype=torch.int32, device=dev)
            ws = _lib.workspace(_lib.size_query("cmp_build_pair_list_workspace", self.N, self.G), dev)
            _lib.call("cmp_build_pair_list", _lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.dist),
                      _lib.ptr(self.seg_ptr), self.N, self.G, int(self.sym_atoms), cap, _lib.ptr(src), _lib.ptr(dst),
                      _lib.ptr(dist),
                      _lib.ptr(rev), _lib.ptr(conf_ptr), _lib.ptr(ws), ws.numel(), _lib.ptr(self.status))
            tile_e = _lib.size_query("cmp_cfconv_tc_bwd_tile_edges")
            cap_t = cap // tile_e + self.G + 1
            tiles = torch.empty(max(cap_t, 1), 8, dtype=torch.int32, device=dev)
            num = torch.zeros(1, dtype=torch.int32, device=dev)
            ws = _lib.workspace(_lib.size_query("cmp_build_flat_tiles_workspace", self.G), dev)
            _lib.call("cmp_build_flat_tiles", _lib.ptr(conf_ptr), _lib.ptr(self.seg_ptr), _lib.ptr(dst),
                      self.G, tile_e, _lib.ptr(tiles), cap_t, _lib.ptr(num), _lib.ptr(ws), ws.numel(),
                      _lib.ptr(self.status))
            self._pair_tiles = (src, dst, dist, rev, tiles, num, conf_ptr)
        return self._pair_tiles

    def edge_weight(self) -> torch.Tensor:
        w = self.dist[: self.E]
        w._cmp_graph = self
        return w


def num_graphs_of(batch: torch.Tensor, num_graphs: Optional[int] = None) -> int:
    if num_graphs is not None:
        return int(num_graphs)
    if batch.numel() == 0:
        return 0
    return int(batch[-1].item()) + 1  # batch is sorted: the last id is the largest (host sync)


def build_neighbor_list(pos: torch.Tensor, batch: Optional[torch.Tensor], r: float, max_num_neighbors: int = 32,
                        loop: bool = False, num_graphs: Optional[int] = None, want_evec: bool = False,
                        want_transpose: bool = True, num_edges: Optional[int] = None,
                        max_atoms: Optional[int] = None, status: Optional[torch.Tensor] = None) -> NeighborList:
    """``num_edges``: the caller vouches for the edge count (e.g. a CUDA-graph replay of the same geometry); it
    removes the one host synchronisation that code sizing tensors by E otherwise needs."""
    if not pos.is_cuda:
        raise _lib.ConanMPError("build_neighbor_list: pos must be a CUDA tensor (no CPU path)")
    if pos.dim() != 2 or pos.size(1) != 3:
        raise ValueError("pos must be [N, 3]")
    pos = pos.detach().to(torch.float32).contiguous()
    N = pos.size(0)
    dev = pos.device
    if batch is None:
        batch = torch.zeros(N, dtype=torch.int64, device=dev)
        num_graphs = 1 if N else 0
    if batch.numel() != N:
        raise ValueError("batch and pos disagree on the number of atoms")
    batch = batch.to(torch.int64).contiguous()
    G = num_graphs_of(batch, num_graphs)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    nl = NeighborList(N, G, N * cap, dev, status=status, zero_fill=num_edges is not None)
    nl.cutoff, nl.loop = float(r), bool(loop)
    nl.pos = pos
    nl.max_atoms = None if max_atoms is None else int(max_atoms)
    nl.sym_atoms = int(cap)      # conformers of at most this many atoms cannot have been truncated: symmetric lists
    i32 = dict(dtype=torch.int32, device=dev)
    if want_evec:
        nl.evec = nl._alloc(max(nl.cap_E, 1), 3, dtype=torch.float32, device=dev)
    if want_transpose:
        nl.rowptr_t = torch.empty(N + 1, **i32)
        nl.col_t = nl._alloc(max(nl.cap_E, 1), **i32)
        nl.eid_t = nl._alloc(max(nl.cap_E, 1), **i32)
    _lib.call("cmp_batch_to_segments", _lib.ptr(batch), N, G, _lib.ptr(nl.seg_ptr), _lib.ptr(nl.status))
    ws_bytes = _lib.size_query("cmp_radius_csr_workspace", N, G)
    ws = _lib.workspace(ws_bytes, dev)
    _lib.call("cmp_radius_csr", _lib.ptr(pos), _lib.ptr(nl.seg_ptr), N, G, float(r), int(max_num_neighbors),
              int(bool(loop)), nl.cap_E, _lib.ptr(nl.rowptr), _lib.ptr(nl.col), _lib.ptr(nl.dist),
              _lib.ptr(nl.evec), _lib.ptr(nl.rowptr_t), _lib.ptr(nl.col_t), _lib.ptr(nl.eid_t),
              _lib.ptr(nl.conf_edge_ptr), _lib.ptr(ws), ws.numel(), _lib.ptr(nl.status))
    if num_edges is not None:
        # trusted for sizing (no host sync); a geometry whose edge count differs sets CMP_STATUS_EDGE_OVERFLOW
        nl._E = int(num_edges)
        _lib.call("cmp_check_edge_count", _lib.ptr(nl.rowptr), N, int(num_edges), _lib.ptr(nl.status))
    return nl


def radius_graph(x: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32, flow: str = "source_to_target", num_workers: int = 1) -> torch.Tensor:
    """``torch_geometric.nn.radius_graph`` drop-in (CUDA truncation rule, see ``oracle/radius.py``)."""
    assert flow in ("source_to_target", "target_to_source")
    nl = build_neighbor_list(x, batch, r, max_num_neighbors, loop)
    ei = nl.edge_index()
    if flow == "target_to_source":
        return ei.flip(0)
    return ei


def graph_from_edge_index(edge_index: torch.Tensor, edge_weight: torch.Tensor, num_nodes: int):
    """CSR view of an arbitrary ``edge_index`` (generic ``InteractionBlock`` entry, e.g. the covalent
    graph of ``schnet_no_sum.py:166-174``).  Returns ``(graph, perm)``; ``perm`` reorders per-edge
    tensors into CSR order (None when the input already is target-major)."""
    tagged = getattr(edge_index, "_cmp_graph", None)
    if tagged is not None and tagged.N == num_nodes:
        return tagged, None
    dev = edge_index.device
    src, dst = edge_index[0], edge_index[1]
    E = src.numel()
    key = dst * num_nodes + src
    perm = None
    if E > 1 and bool((key[1:] < key[:-1]).any()):
        perm = torch.argsort(key, stable=True)
        src, dst = src[perm], dst[perm]
    nl = NeighborList(num_nodes, 0, E, dev)
    nl._E = E
    nl._checked = True
    counts = torch.bincount(dst, minlength=num_nodes)
    nl.rowptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)]).to(torch.int32)
    nl.col = src.to(torch.int32).contiguous() if E else nl.col
    w = edge_weight if perm is None else edge_weight[perm]
    nl.dist = w.to(torch.float32).contiguous() if E else nl.dist
    # transpose (source-major, target ascending)
    tkey = src * num_nodes + dst
    tperm = torch.argsort(tkey, stable=True)
    tcounts = torch.bincount(src, minlength=num_nodes)
    nl.rowptr_t = torch.cat([tcounts.new_zeros(1), tcounts.cumsum(0)]).to(torch.int32)
    nl.col_t = dst[tperm].to(torch.int32).contiguous() if E else nl.col
    nl.eid_t = tperm.to(torch.int32).contiguous() if E else nl.col
    return nl, perm

"""Data-parallel training step for the backbone: shard molecules, flat gradient all-reduce, fused Adam.

What it mirrors in the reference: Lightning DDP (``conan_fgw/src/trainer.py:308-325``,
strategy ``ddp_find_unused_parameters_false``) with ``DistributedSampler(shuffle=False)``
(``conan_fgw/src/data/datamodules.py:40-41``) and Adam (``model/common.py:368-370``).
The forward path needs no inter-GPU traffic (conformer graphs are independent); the only
collective is one all-reduce of a single flat fp32 gradient buffer per step.
"""

from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops


def shard_molecules(num_molecules: int, rank: int, world_size: int) -> List[int]:
    """Indices of the molecules rank ``rank`` owns: ``rank, rank+W, ...`` padded by wrap-around to equal
    length on every rank - the rule of ``DistributedSampler(shuffle=False, drop_last=False)``."""
    if num_molecules == 0:
        return []
    per = -(-num_molecules // world_size)
    total = per * world_size
    idx = list(range(num_molecules))
    while len(idx) < total:
        idx += idx[: total - len(idx)]
    return idx[rank:total:world_size]


class FlatParameters:
    """Re-homes a module's parameters (and gradients) as views of one contiguous fp32 buffer so the
    gradient all-reduce is a single collective and Adam a single kernel."""

    def __init__(self, modules):
        params, seen = [], set()
        for m in modules:
            for p in m.parameters():
                if id(p) not in seen and p.requires_grad:
                    seen.add(id(p))
                    params.append(p)
        if not params:
            raise ValueError("no trainable parameters")
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.params = params
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self._grad_views = []
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            self._grad_views.append(self.grad[off:off + k].view_as(p))
            off += k
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0

    def zero_grad(self):
        """Gradients are produced as fresh tensors by the backward kernels (``p.grad = None`` lets autograd adopt them
        without an accumulation kernel per parameter); ``collect_grads`` then moves them into the flat buffer with one
        multi-tensor copy."""
        for p in self.params:
            p.grad = None

    def collect_grads(self):
        self.grad.zero_()          # parameters that received no gradient (e.g. an unused head) contribute zeros
        dst, src = [], []
        for p, view in zip(self.params, self._grad_views):
            if p.grad is not None:
                dst.append(view)
                src.append(p.grad)
        if dst:
            torch._foreach_copy_(dst, src)

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def adam(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        self.step_count += 1
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, lr, betas, eps,
                      weight_decay, grad_scale)


# conformer mean -> linear head -> MSE of RegressionStep.loss as one fused forward / backward launch each
# (ops.regression_head_loss); False = the separate segment-mean / GEMM / torch MSE ops (kept for cross-checking)
FUSED_HEAD = True


class RegressionStep:
    """One ConAN-style regression training step around the backbone:
    backbone -> [G, H/2] -> mean over the K conformers of a molecule (``schnet_based_models.py:242``)
    -> linear head -> MSE (``model/common.py:288``) -> backward -> [all-reduce] -> Adam."""

    def __init__(self, backbone, hidden_half: int, num_conformers: int, lr: float = 1e-3, group=None,
                 backbone_kwargs=None):
        self.backbone = backbone
        self.backbone_kwargs = dict(backbone_kwargs or {})   # e.g. {"num_edges": E} for a sync-free ViSNet forward
        dev = next(backbone.parameters()).device
        from .nn import Linear

        self.head = Linear(hidden_half, 1).to(dev)     # regression head (schnet_based_models.py:17-29) on cmp_gemm_f32
        self.K = int(num_conformers)
        self.flat = FlatParameters([backbone, self.head])
        self.lr = lr
        self.group = group

    def loss(self, z, pos, batch, targets, num_graphs):
        emb = self.backbone(z, pos, batch, num_graphs=num_graphs, **self.backbone_kwargs)          # [G, H/2]
        if (FUSED_HEAD and emb.is_cuda and self.head.out_features == 1 and emb.shape[1] <= 512
                and emb.shape[0] // self.K <= 1024):     # one CTA walks the molecules: larger batches keep the separate ops
            # conformer mean -> linear head -> MSE (and the whole backward of the three) in two launches
            return ops.regression_head_loss(emb, self.head.weight, self.head.bias, targets, self.K)
        mol = ops.conformers_mean(emb, self.K)                               # conformers of a molecule are consecutive
        pred = self.head(mol)
        return torch.nn.functional.mse_loss(pred, targets)

    def _fwd_bwd(self, z, pos, batch, targets, num_graphs):
        self.flat.zero_grad()
        # weights only change in adam(): every tensor-core weight image of this step is packed up front, grouped
        hint = getattr(self.backbone, "max_atoms_hint", None)
        with ops.prepacked_weights([self.backbone, self.head], dense_only=hint is not None and hint <= 128):
            with _lib.nvtx_range("cmp/forward"):
                loss = self.loss(z, pos, batch, targets, num_graphs)
            # parameter gradients are None here and every parameter is used once: the node-linear weight gradients
            # can be queued during backward and issued as ONE grouped launch at its end
            with _lib.nvtx_range("cmp/backward"), ops.deferred_weight_grads(self.flat.params):
                loss.backward()
        self.flat.collect_grads()
        return loss.detach()

    def check(self):
        """Raise for device-side input errors reported since the last call (the captured / sync-free step never reads
        the status word itself).  Call it where the host synchronises anyway, e.g. next to ``loss.item()``."""
        chk = getattr(self.backbone, "check_status", None)
        if chk is not None:
            chk()

    def capture(self, z, pos, batch, targets, num_graphs, timer=None):
        """Capture zero-grad + radius graph + forward + loss + backward into ONE CUDA graph (every entry point of
        the library is stream-ordered and sync-free on the bf16 path).  Later ``step`` calls with tensors of the same
        shapes copy their inputs into the static buffers and replay it; the gradient all-reduce and Adam follow
        eagerly.  Shapes are fixed per capture: a training loop keeps one graph per (N, G) bucket."""
        # the static inputs are views of ONE device buffer (256-byte aligned slots): a loader that collates into the pinned
        # twin of that layout (``host_staging``) uploads a whole step with a single H2D copy (``upload``)
        srcs = [t.contiguous() for t in (z, pos, batch, targets)]
        offs, total = [], 0
        for t in srcs:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        self._static_buf = torch.empty(max(total, 256), dtype=torch.uint8, device=srcs[0].device)
        self._static_layout = [(o, t.numel() * t.element_size(), t.dtype, tuple(t.shape)) for o, t in zip(offs, srcs)]
        self._static = [self._static_buf[o:o + nb].view(dt).view(shape) for o, nb, dt, shape in self._static_layout]
        for dst, src in zip(self._static, srcs):
            dst.copy_(src)
        self._static_G = int(num_graphs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._fwd_bwd(*self._static, self._static_G)
        torch.cuda.current_stream().wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        # timer: a _lib.KernelTimer(external=True) whose event records are captured as graph nodes around the selected
        # launches (bench.py: per-kernel durations of the replayed step itself)
        prev, _lib.timer = _lib.timer, timer
        try:
            with torch.cuda.graph(self._graph):
                self._static_loss = self._fwd_bwd(*self._static, self._static_G)
        finally:
            _lib.timer = prev
        return self

    def host_staging(self):
        """``(buffer, [z, pos, batch, targets])``: a pinned host buffer with the layout of the captured step's input
        buffer and typed views of its four slots.  Fill the views (e.g. in the collate function), then ``upload(buffer)``."""
        buf = torch.empty(self._static_buf.numel(), dtype=torch.uint8).pin_memory()
        views = [buf[o:o + nb].view(dt).view(shape) for o, nb, dt, shape in self._static_layout]
        return buf, views

    def upload(self, host_buffer):
        """One asynchronous H2D copy of a whole step's inputs into the captured graph's input buffers."""
        self._static_buf.copy_(host_buffer, non_blocking=True)
        return self._static

    def step(self, z, pos, batch, targets, num_graphs):
        g = getattr(self, "_graph", None)
        if g is not None and int(num_graphs) == self._static_G and all(
                a.shape == b.shape for a, b in zip((z, pos, batch, targets), self._static)):
            for dst, src in zip(self._static, (z, pos, batch, targets)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            g.replay()
            loss = self._static_loss
        else:
            loss = self._fwd_bwd(z, pos, batch, targets, num_graphs)
        with _lib.nvtx_range("cmp/grad_allreduce"):
            world = self.flat.all_reduce(self.group)
        with _lib.nvtx_range("cmp/adam"):
            self.flat.adam(lr=self.lr, grad_scale=1.0 / world)
        return loss

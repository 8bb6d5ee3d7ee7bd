"""Data-parallel train step for the bag models: one flat fp32 parameter buffer, one flat gradient
buffer that the weight-gradient kernels write into directly, ONE NCCL all-reduce per step, and a
fused Adam over the flat buffer.

Replaces what ``accelerate`` + DistributedDataParallel do around the reference's loop
(2d_FPE/train_fno.py:75-77,116-123,139-145): DDP ships every parameter (13.5 M floats, of which
9.9 M belong to the Encoder2D branch NIOFP2D_FNO never calls, hence find_unused_parameters=True)
in 25 MiB buckets; here only the live gradients (3.56 M floats for the 2-D NIO-FNO) travel, as plain
NCCL all-reduce calls on slices of one flat buffer (one call, or two to three when the heads' region is
reduced early under the encoder backward).  Samples are sharded across ranks (pure data parallel, weak
scaling), each rank draws its own bag subsample from its own NumPy stream (seed + rank,
train_fno.py:78-81).

Replica consistency: the reference seeds every process differently (seed + process_index,
train_fno.py:78-81) and relies on DistributedDataParallel broadcasting rank 0's parameters and buffers
when it wraps the model.  FlatTrainer does the same at construction (``sync_from_rank0``): the flat
parameter buffer, the parameters outside it (the detached ``fc0``, the never-called branch) and every
module buffer (BatchNorm statistics of the NIO / BlinDNO models) are broadcast from rank 0.  After that
BatchNorm running statistics evolve per rank: DDP's default re-broadcasts rank 0's buffers before every
forward, but they do not enter train-mode outputs or gradients, so the trained weights are identical;
``sync_buffers()`` restores DDP's state (e.g. before saving a checkpoint or switching to eval()).

The optimiser maths is torch.optim.Adam's (lr, betas, eps; no weight decay); parameters that never
receive a gradient in the reference (``fc0`` used through ``.data``; the unused branch) are left
untouched exactly as Adam leaves parameters whose ``.grad`` is None.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

import numpy as np

from . import ops
from .surface.fno import _FnoBase


def live_parameters(model: nn.Module):
    """(name, parameter) pairs the train step updates, in registration order."""
    dead: Iterable[str] = ("fc0.",)
    if hasattr(model, "FNO_input"):
        dead = ("fc0.", "branch.")          # NIO-FNO never calls its branch encoder (Q8)
    seen = set()
    out = []
    for name, p in model.named_parameters():   # named_parameters de-duplicates shared tensors (Q10)
        if name.startswith(tuple(dead)) or not p.requires_grad or id(p) in seen:
            continue
        seen.add(id(p))
        out.append((name, p))
    return out


class FlatTrainer:
    """Owns the flat parameter / gradient / Adam-state buffers of ``model`` and runs train steps."""

    def __init__(self, model: nn.Module, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group: Optional[dist.ProcessGroup] = None, loss_fn=None, sync_from_rank0: bool = True):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.loss_fn = loss_fn or nn.functional.mse_loss
        # the default criterion (MSELoss on the concatenated head outputs, train_fno.py:116,146) runs as one kernel that
        # reads the head outputs where they are (ops.heads_mse); a caller-supplied loss_fn sees the usual tensor
        self.fused_mse = loss_fn is None and hasattr(model, "_heads_as_list")
        self.step_count = 0
        self.adam_fn = ops.adam_step_flat
        # CUDA-graph replay of the device work of a step (see step()): one graph per bag size
        self.use_graphs = False
        self._graphs = {}
        self._graph_pool = None
        self.replayed_launches = 0
        # Backward split (NIO-FNO): the heads' gradients (94 % of the buffer) are complete before the
        # per-snapshot net starts back-propagating, so their all-reduce runs on a side stream under that
        # backward; only the small FNO_input region is reduced at the end.  With one process there is nothing to reduce,
        # but the split still pays under graph replay: the heads' Adam update (94 % of the parameters, ~19 us) runs on
        # the side stream under the per-snapshot net's backward instead of after it.
        self.split_backward = hasattr(model, "FNO_input") and hasattr(model, "_expose_lifted")
        self._comm_stream = None
        self._late_span = None
        # The heads' Adam update follows their early all-reduce on the communication stream, so that only the
        # small FNO_input region is reduced and updated after the encoder backward (off: one Adam over everything).
        self.early_adam = True
        self.timing = None                 # set to a dict by enable_timing(): per-step event pairs of the step's tail

        named = live_parameters(model)
        if not named:
            raise RuntimeError("model has no trainable parameters")
        self.device = named[0][1].device
        by_id = {id(p): n for n, p in named}

        # FNO nets get one contiguous region each, in the slot layout their backward kernels expect
        regions = []          # (module or None, [params])
        claimed = set()
        for mod in model.modules():
            if isinstance(mod, _FnoBase):
                ps = mod._params()
                if all(id(p) in by_id for p in ps):
                    regions.append((mod, ps))
                    claimed.update(id(p) for p in ps)
        rest = [p for _, p in named if id(p) not in claimed]
        if rest:
            regions.append((None, rest))

        sizes_per_region = []
        total = 0
        for _, ps in regions:
            sizes = [p.numel() * (2 if p.is_complex() else 1) for p in ps]
            offs, n = ops.slot_layout(sizes)
            sizes_per_region.append((total, offs, sizes, n))
            total += n
        self.numel = total
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_grad = self._alloc_grad_buffer(total)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.n_live = 0

        for (mod, ps), (base, offs, sizes, n) in zip(regions, sizes_per_region):
            for p, off, size in zip(ps, offs, sizes):
                pv = self.flat_param[base + off: base + off + size]
                gv = self.flat_grad[base + off: base + off + size]
                if p.is_complex():
                    pv = torch.view_as_complex(pv.view(*p.shape, 2))
                    gv = torch.view_as_complex(gv.view(*p.shape, 2))
                else:
                    pv, gv = pv.view(p.shape), gv.view(p.shape)
                with torch.no_grad():
                    pv.copy_(p.data)
                p.data = pv
                p.grad = gv
                self.n_live += p.numel() * (2 if p.is_complex() else 1)
            if mod is not None:
                mod._grad_sink = self.flat_grad[base: base + n]
                if mod is getattr(model, "FNO_input", None):
                    self._late_span = (base, base + n)
        if sync_from_rank0 and self.world > 1:
            self.sync_from_rank0()

    def _alloc_grad_buffer(self, total: int) -> torch.Tensor:
        """The flat gradient buffer.  With BDN_NCCL_REGISTER=1 (and world > 1, NCCL) it is allocated from NCCL's own
        allocator (ncclMemAlloc through torch's MemPool) and registered with the communicator, which lets NCCL reduce it
        in place over NVLink SHARP (NVLS) without staging copies; ``self.nccl_registered`` says whether that happened
        (any failure falls back to a plain allocation and says so on stderr -- the arithmetic is the same)."""
        import os
        import sys
        self.nccl_registered = False
        want = os.environ.get("BDN_NCCL_REGISTER", "0") == "1"
        if want and self.world > 1 and self.device.type == "cuda" and dist.get_backend(self.group) == "nccl":
            try:
                pg = self.group if self.group is not None else dist.distributed_c10d._get_default_group()
                backend = pg._get_backend(self.device)
                pool = torch.cuda.MemPool(backend.mem_allocator)
                with torch.cuda.use_mem_pool(pool):
                    buf = torch.zeros(total, dtype=torch.float32, device=self.device)
                backend.register_mem_pool(pool)
                self._nccl_pool = pool
                self.nccl_registered = True
                return buf
            except Exception as e:      # noqa: BLE001 -- optional fast path
                print(f"blindno_b200: NCCL buffer registration unavailable ({type(e).__name__}: {e}); plain allocation", file=sys.stderr)
        return torch.zeros(total, dtype=torch.float32, device=self.device)

    # -- replica consistency ---------------------------------------------------------------
    def _src(self) -> int:
        return dist.get_global_rank(self.group, 0) if self.group is not None else 0

    def sync_from_rank0(self):
        """What DistributedDataParallel does when it wraps a model: every rank continues from rank 0's parameters
        and buffers (the reference builds its replicas from different seeds, 2d_FPE/train_fno.py:78-81)."""
        if self.world == 1:
            return
        src = self._src()
        dist.broadcast(self.flat_param, src=src, group=self.group)
        flat_ids = {id(p) for _, p in live_parameters(self.model)}
        seen = set()
        with torch.no_grad():
            for p in self.model.parameters():
                if id(p) in flat_ids or id(p) in seen:
                    continue
                seen.add(id(p))
                t = torch.view_as_real(p.data) if p.is_complex() else p.data
                dist.broadcast(t, src=src, group=self.group)
        self.sync_buffers()

    def sync_buffers(self):
        """Rank 0's module buffers (BatchNorm running statistics, step counters) to every rank."""
        if self.world == 1:
            return
        src = self._src()
        with torch.no_grad():
            for b in self.model.buffers():
                dist.broadcast(b, src=src, group=self.group)

    # -- pieces of a step ----------------------------------------------------------------
    def zero_grad(self):
        self.flat_grad.zero_()

    def reduce_gradients(self):
        """Sum over ranks in one collective; the 1/world of DDP's mean is folded into the Adam kernel."""
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def _early_spans(self):
        a, b = self._late_span
        return [(lo, hi) for lo, hi in ((0, a), (b, self.numel)) if hi > lo]

    def reduce_early(self):
        """All-reduce everything but the FNO_input region on the communication stream (non-blocking for the
        compute stream)."""
        if self.world == 1 and not self.early_adam:
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        self._comm_stream.wait_stream(main)
        with torch.cuda.stream(self._comm_stream):
            for lo, hi in self._early_spans():
                if self.world > 1:
                    dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                if self.early_adam:
                    # same stream: the update starts when this span's sum has landed, under the encoder backward
                    # (which reads none of these parameters)
                    self._adam(lo, hi, self.step_count + 1)

    def reduce_late(self):
        if self.world > 1:
            a, b = self._late_span
            dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.group)
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)

    def _adam(self, lo: int, hi: int, step: int):
        self.adam_fn(self.flat_param[lo:hi], self.flat_grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], lr=self.lr,
                     betas=self.betas, eps=self.eps, step=step, grad_scale=1.0 / self.world)

    def optimizer_step(self, spans=None):
        """Fused Adam over the flat buffer, or over ``spans`` of it (1/world of DDP's mean folded in).  There is no
        CPU arithmetic here: CPU tensors make the op raise.  (The gloo host-logic test replaces ``adam_fn`` with
        a torch-op restatement of its own.)"""
        self.step_count += 1
        for lo, hi in (spans if spans is not None else [(0, self.numel)]):
            self._adam(lo, hi, self.step_count)

    def _loss(self, run_model, target):
        """Forward + criterion.  Returns (loss, roots, root_grads): backward starts from ``roots`` with ``root_grads``.
        The default criterion on CUDA runs as one kernel that also produces d loss / d (head outputs)
        (ops.heads_mse_grads): the autograd graph then starts at the head outputs, there is no loss node."""
        if self.fused_mse and target.is_cuda:
            self.model._heads_as_list = True
            try:
                outs = run_model()
            finally:
                self.model._heads_as_list = False
            if isinstance(outs, (list, tuple)):
                loss, grads = ops.heads_mse_grads(outs, target)
                return loss, list(outs), grads
            loss = self.loss_fn(outs, target)
        else:
            loss = self.loss_fn(run_model(), target)
        return loss, [loss], [torch.ones_like(loss)]

    # -- the step --------------------------------------------------------------------------
    def step(self, x: torch.Tensor, grid: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """zero_grad -> forward -> loss -> backward -> all-reduce -> Adam.  Returns the (device) loss."""
        if self.use_graphs and (x.is_cuda or x.is_pinned()) and self.flat_param.is_cuda and getattr(self.model, "accepts_idx", False):
            return self._graph_step(x, grid, target)
        self.zero_grad()
        loss, roots, root_grads = self._loss(lambda: self.model(x, grid), target)
        torch.autograd.backward(roots, root_grads)
        self.reduce_gradients()
        self.optimizer_step()
        return loss.detach()

    # -- CUDA graphs -----------------------------------------------------------------------
    # The reference step is a few hundred tiny launches with a fresh bag size L ~ U[50, L0) every step
    # (2d_FPE/NIOModules.py:548-551): eager, the host cannot feed a B200 fast enough.  The bag is still
    # drawn on the host from the NumPy stream in the reference's order; only its indices are copied into
    # a static buffer, and zero_grad + forward + loss + backward are replayed from the graph captured for
    # that bag size (all graphs share one memory pool).  All-reduce and Adam stay outside the graph.
    def enable_graphs(self, enabled: bool = True):
        self.use_graphs = bool(enabled)
        return self

    def _draw(self, n_snapshots: int):
        from .surface.nio import draw_bag
        return draw_bag(n_snapshots, self.model.training)

    def _graph_entry(self, x, grid, target, n_keep):
        key = (n_keep, tuple(x.shape), tuple(target.shape), None if grid is None else tuple(grid.shape),
               bool(self.model.training))
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        dev = self.device
        ent = {
            "x": torch.empty(x.shape, dtype=x.dtype, device=dev),
            "target": torch.empty(target.shape, dtype=target.dtype, device=dev),
            "grid": None if grid is None else grid.detach().to(dev).clone(),      # (BlinDNO models take no grid)
            "idx": torch.zeros(max(n_keep, 1), dtype=torch.int32, device=dev) if n_keep else None,
        }
        ent["x"].copy_(x)
        ent["target"].copy_(target)
        torch.cuda.synchronize(dev)

        split = self.split_backward and self._late_span is not None

        def body_a():
            self.flat_grad.zero_()
            self.model._expose_lifted = split
            loss, roots, root_grads = self._loss(
                lambda: self.model(ent["x"], ent["grid"], idx=ent["idx"]) if ent["idx"] is not None else
                self.model(ent["x"], ent["grid"]), ent["target"])
            if not split:
                torch.autograd.backward(roots, root_grads)
                return loss.detach(), None, None
            lifted = self.model._lifted
            self.model._lifted = None
            # runs the heads' backward (gradients -> flat buffer)
            (g_lifted,) = torch.autograd.grad(roots, [lifted], grad_outputs=root_grads)
            return loss.detach(), lifted, g_lifted

        def body_b(lifted, g_lifted):
            torch.autograd.backward([lifted], [g_lifted])          # the per-snapshot net's backward

        def body():
            loss, lifted, g_lifted = body_a()
            if split:
                body_b(lifted, g_lifted)
            return loss

        # warm-up on a side stream (plans, lazy module loads, autograd buffers), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            if ent["idx"] is None:
                state = np.random.get_state()
            # the warm-up run must leave no trace: BatchNorm running statistics (NIO's conv encoder and trunk) are
            # module buffers that a train-mode forward updates
            buffers = [(b, b.detach().clone()) for b in self.model.buffers()]
            body()
            with torch.no_grad():
                for b, saved in buffers:
                    b.copy_(saved)
            if ent["idx"] is None:
                np.random.set_state(state)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        graph = torch.cuda.CUDAGraph()
        launches0 = ops.kernel_launches()
        ent["graph_b"] = None
        if not split:
            with torch.cuda.graph(graph, pool=self._graph_pool):
                ent["loss"] = body()
        else:
            with torch.cuda.graph(graph, pool=self._graph_pool):
                ent["loss"], lifted, g_lifted = body_a()
            graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_b, pool=self._graph_pool):
                body_b(lifted, g_lifted)
            ent["graph_b"] = graph_b
        ent["launches"] = ops.kernel_launches() - launches0      # kernels of this library inside the graph(s)
        ent["graph"] = graph
        self._graphs[key] = ent
        return ent

    def _graph_step(self, x, grid, target):
        idx = self._draw(x.shape[1])
        n_keep = 0 if idx is None else int(len(idx))
        ent = self._graph_entry(x, grid, target, n_keep)
        ent["x"].copy_(x, non_blocking=True)
        ent["target"].copy_(target, non_blocking=True)
        if n_keep:
            # pageable source: the runtime stages it before returning, so the host array may be reused
            ent["idx"].copy_(torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)))
        ent["graph"].replay()
        late_only = None
        if ent["graph_b"] is not None:
            t = self._tail_events()
            self.reduce_early()            # heads' gradients travel (and are applied) while the per-snapshot net back-propagates
            ent["graph_b"].replay()
            if t:
                t[0].record()
            self.reduce_late()
            if t:
                t[1].record()
            if self.early_adam:
                late_only = [self._late_span]
        else:
            self.reduce_gradients()
        self.replayed_launches += ent["launches"]
        self.optimizer_step(late_only)
        if ent["graph_b"] is not None and t:
            t[2].record()
        return ent["loss"]

    # -- timing of the step's tail (bench.py's scaling_breakdown) ---------------------------------
    def enable_timing(self, enabled: bool = True):
        self.timing = {"events": []} if enabled else None
        return self

    def _tail_events(self):
        if self.timing is None:
            return None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        self.timing["events"].append(ev)
        return ev

    def tail_times_ms(self):
        """(mean ms between the end of the encoder backward and the end of the late all-reduce [includes waiting for the
        slowest rank and for the early all-reduce], mean ms of the Adam that follows) over the recorded steps."""
        if not self.timing or not self.timing["events"]:
            return None
        torch.cuda.synchronize(self.device)
        red = [a.elapsed_time(b) for a, b, _ in self.timing["events"]]
        adam = [b.elapsed_time(c) for _, b, c in self.timing["events"]]
        self.timing["events"].clear()
        return sum(red) / len(red), sum(adam) / len(adam)

    def prepare_graphs(self, x, grid, target, bag_sizes=None):
        """Capture the graphs of every bag size a training run can draw (L in [50, L0)) up front."""
        n0 = x.shape[1]
        sizes = list(bag_sizes) if bag_sizes is not None else (list(range(50, n0)) if self.model.training else [0])
        for n in sizes:
            self._graph_entry(x, grid, target, n)
        return len(self._graphs)


class HostPipeline:
    """Feeds a FlatTrainer from HOST (pinned) batches without stalling the device.

    The reference loop (2d_FPE/train_fno.py:139-145) moves every batch to the device, steps, and calls
    ``loss.item()`` -- a full host<->device round trip per step.  Here the same three things happen for
    every step, pipelined: the host->device copy of batch i+1 runs on a copy stream while step i
    computes (two device staging slots), and the loss of step i is copied to pinned host memory
    asynchronously and read one step later.  ``step`` returns the previous step's loss as a float
    (``None`` on the first call); ``flush`` returns the last one.
    """

    def __init__(self, trainer: FlatTrainer, grid: torch.Tensor):
        self.trainer, self.grid = trainer, grid
        dev = trainer.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self._pending = False
        self.h2d_bytes = 0

    def _issue(self, k: int, hx: torch.Tensor, hy: torch.Tensor):
        dev = self.trainer.device
        if self.slots[k] is None or self.slots[k][0].shape != hx.shape or self.slots[k][1].shape != hy.shape:
            self.slots[k] = (torch.empty(hx.shape, dtype=hx.dtype, device=dev),
                             torch.empty(hy.shape, dtype=hy.dtype, device=dev))
        else:
            self.copy_stream.wait_event(self.consumed[k])      # the step that read this slot has copied it out
        with torch.cuda.stream(self.copy_stream):
            self.slots[k][0].copy_(hx, non_blocking=True)
            self.slots[k][1].copy_(hy, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.h2d_bytes = (hx.numel() * hx.element_size() + hy.numel() * hy.element_size())

    def step(self, hx: torch.Tensor, hy: torch.Tensor, next_batch=None):
        """Step on the host batch (hx, hy); ``next_batch=(hx', hy')`` starts its upload behind this step."""
        k = self.i & 1
        if not self._pending:
            self._issue(k, hx, hy)
        main = torch.cuda.current_stream(self.trainer.device)
        main.wait_event(self.ready[k])
        loss = self.trainer.step(self.slots[k][0], self.grid, self.slots[k][1])
        self.consumed[k].record(main)
        self.loss_host[k].copy_(loss, non_blocking=True)
        self.loss_done[k].record(main)
        self._pending = next_batch is not None
        if self._pending:
            self._issue(k ^ 1, *next_batch)
        prev = None
        if self.i > 0:
            self.loss_done[k ^ 1].synchronize()
            prev = float(self.loss_host[k ^ 1])
        self.i += 1
        return prev

    def flush(self):
        if self.i == 0:
            return None
        k = (self.i - 1) & 1
        self.loss_done[k].synchronize()
        return float(self.loss_host[k])


def shard_batch(n_samples: int, rank: int, world: int):
    """Contiguous, near-even split of a global batch over ranks (samples are independent bags)."""
    base, extra = divmod(n_samples, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)

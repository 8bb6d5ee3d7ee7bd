#!/usr/bin/env python
"""BASELINE.json configs[4] / SURVEY.md C5: the 2D-FPE NIO-FNO train step swept over
bag size L0 x grid resolution n x retained modes, one JSON line per point.

    python tools/sweep.py [--bags 50,100,200,400] [--grids 61,80,128,256] [--modes 12,16,32,64]
                          [--batch 4] [--steps 5] [--warmup 2] [--graphs] [--ref-steps 0]

Model: NIOFP2D_FNO(2,3,100,25, fno_layers=3, width=12, modes=M, out=2) -- the per-snapshot FNO_input stays
(width 4, modes 12) as the reference hard-codes it (2d_FPE/NIOModules.py:528).  A point is valid when the
kept row blocks do not overlap on the padded grid: 2*M <= n + round(n/4).  Steps are eager launches unless
--graphs (one CUDA graph per point with the bag size pinned to its mean, so the number is the kernel-bound
rate without per-bag-size capture cost).  Under torchrun every rank runs the same point on its own bags
(weak scaling, gradient all-reduce included) and rank 0 prints samples/s over all ranks.
--ref-steps K additionally times K steps of the CPU oracle (torch.fft restatement of the reference) for
points whose hidden tensor [B*L, n, n, 128] fits comfortably in host memory, else records "reference: skipped".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def grid2d(n):
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    return torch.tensor(np.stack(np.meshgrid(ax, ax, indexing="ij"), axis=2))


def spectral_train_flops(images, c, hp, wp, m, layers):
    """pruned-DFT GEMM flops of one spectral layer, training = 2 x DFT terms + 3 x mix term."""
    dft = 8 * c * hp * wp * m + 32 * c * m * m * hp
    mix = 16 * m * m * c * c
    return images * layers * (2 * dft + 3 * mix)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bags", default="50,100,200,400")
    ap.add_argument("--grids", default="61,80,128,256")
    ap.add_argument("--modes", default="12,16,32,64")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--graphs", action="store_true")
    ap.add_argument("--ref-steps", type=int, default=0)
    ap.add_argument("--max-act-gb", type=float, default=40.0, help="skip points whose saved activations exceed this")
    args = ap.parse_args()

    import torch.distributed as dist
    from blindno_b200 import ops
    from blindno_b200.parallel import FlatTrainer
    from blindno_b200.surface import nio

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def emit(obj):
        if rank == 0:
            print(json.dumps(obj), flush=True)

    for n in [int(v) for v in args.grids.split(",")]:
        pad = int(round(n * 0.25))
        npad = n + pad
        grid = grid2d(n).to(dev)
        for bag in [int(v) for v in args.bags.split(",")]:
            for modes in [int(v) for v in args.modes.split(",")]:
                point = {"grid": n, "padded": npad, "bag": bag, "modes": modes, "batch_per_gpu": args.batch, "n_gpus": world}
                if 2 * modes > npad:
                    emit(dict(point, status="invalid", why=f"2*modes={2 * modes} > padded grid {npad} (row blocks overlap)"))
                    continue
                mean_keep = (50 + bag - 1) // 2 if bag > 50 else bag
                act_gb = 3 * args.batch * bag * 4 * npad * npad * 4 / 1e9
                if act_gb > args.max_act_gb:
                    emit(dict(point, status="skipped", why=f"saved activations {act_gb:.1f} GB > --max-act-gb"))
                    continue
                try:
                    torch.manual_seed(1)
                    np.random.seed(1 + rank)
                    model = nio.make_models("2d_FPE")["NIOFP2D_FNO"](2, 3, 100, 25, 3, 12, modes, 2).to(dev).train()
                    trainer = FlatTrainer(model, lr=5e-4)
                    g = torch.Generator().manual_seed(100 + rank)
                    pool = [(torch.randn(args.batch, bag, n, n, generator=g).to(dev),
                             torch.randn(args.batch, n, n, 2, generator=g).to(dev)) for _ in range(2)]
                    if args.graphs:
                        # pin the bag size: np.random.randint(50, bag) is replaced by its mean for this measurement
                        trainer.enable_graphs(True)
                        trainer._draw = lambda n_snap, k=mean_keep: np.random.choice(n_snap, k)
                    kept = []
                    l0 = ops.kernel_launches()
                    for i in range(args.warmup):
                        trainer.step(pool[i % 2][0], grid, pool[i % 2][1])
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(args.steps):
                        loss = trainer.step(pool[i % 2][0], grid, pool[i % 2][1])
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.steps
                    if world > 1:
                        t = torch.tensor([ms], device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        ms = t.item()
                    images = args.batch * mean_keep
                    flops = spectral_train_flops(images, 4, npad, npad, 12, 2) + \
                        2 * spectral_train_flops(args.batch, 12, npad, npad, modes, 3)
                    res = dict(point, status="ok", ms_per_step=ms, samples_per_s=world * args.batch / (ms / 1e3),
                               loss=float(loss), mean_kept=mean_keep, mode="graph" if args.graphs else "eager",
                               spectral_gemm_tflops=world * flops / (ms / 1e3) / 1e12,
                               peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 1e9)
                    if args.ref_steps and rank == 0:
                        hidden_gb = args.batch * bag * n * n * 128 * 4 / 1e9
                        if hidden_gb * 6 > 64:
                            res["reference"] = f"skipped: [B*L,n,n,128] hidden tensor is {hidden_gb:.1f} GB per copy on the reference path"
                        else:
                            from oracle import blindno_oracle as O
                            p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()
                                 if not k.startswith("branch.")}
                            opt = torch.optim.Adam(O.trainable(p), lr=5e-4)
                            xc, yc = pool[0][0].cpu(), pool[0][1].cpu()
                            gc = grid.cpu()
                            O.train_step(p, opt, O.niofp2d_fno_forward, xc, gc, yc)
                            t0 = time.perf_counter()
                            for _ in range(args.ref_steps):
                                O.train_step(p, opt, O.niofp2d_fno_forward, xc, gc, yc)
                            sec = (time.perf_counter() - t0) / args.ref_steps
                            res["reference_cpu_samples_per_s"] = args.batch / sec
                            res["reference_cpu_threads"] = torch.get_num_threads()
                    emit(res)
                except Exception as exc:      # a size this build does not serve is a result, not a crash
                    emit(dict(point, status="error", why=f"{type(exc).__name__}: {str(exc)[:300]}"))
                finally:
                    model = trainer = pool = None
                    torch.cuda.empty_cache()
                    torch.cuda.reset_peak_memory_stats(dev)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Run one of the reference's train / eval scripts UNCHANGED against this package.

    python tools/run_reference_script.py oracle/_ref/2d_FPE/train_fno.py --steps 20 --save-ckpt gpurun_out/ckpt.pt
    python tools/run_reference_script.py oracle/_ref/2d_FPE/eval_fno.py --script-args "--ckpt gpurun_out/ckpt.pt --start 0 --end 3"
    python tools/run_reference_script.py /path/to/reference/1d_FPE/train_fno.py --steps 5 --modules reference --device cpu

(oracle/_ref is the staged copy of the reference's files, oracle/stage_reference.py; any path to the reference tree works.)

The scripts are flat files with hard-coded dataset paths, 400 epochs, plots and an `accelerate` launcher.
Without touching them this harness supplies, in-process and only for the run:

  * the model zoo: `blindno_b200.dropin.install(<variant>)` registers NIOModules / FNOModules / ... in
    sys.modules (``--modules reference`` uses the script directory's own files instead: that is how the
    harness itself is validated on a box without a GPU, tests/test_harness_cpu.py);
  * stand-ins for modules this image does not have: `matplotlib` (no-op), `accelerate` (single process per
    GPU: device placement, `prepare`, `backward`, `save`; under torchrun the model is wrapped in DDP exactly
    as accelerate would), `timm`;
  * a synthetic dataset in the schema the script's Dataset class reads, served when `np.load` is asked for
    the script's (absent) hard-coded file;
  * a step limit: after ``--steps`` optimiser steps the run stops and one JSON line reports samples/s;
  * ``--save-ckpt``: when the run stops, the model the script trains is saved the way the scripts save their best
    checkpoint (``accelerator.save(model.state_dict(), ...)``: under DDP the keys carry the ``module.`` prefix that
    eval_fno.py:104-122 strips; ``--ckpt-prefix module.`` reproduces that in a single-process run);
  * eval scripts (argparse, no optimiser): ``--script-args`` is handed to the script as its command line, the run ends
    when the script does, and the rows of the ``metrics.csv`` it wrote are summarised in the JSON line.

Outputs the script writes (checkpoints, curves) go to ``--workdir`` (default gpurun_out/script_run).
"""
from __future__ import annotations

import argparse
import json
import os
import runpy
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _StepLimit(BaseException):
    """Raised from the optimiser step hook; BaseException so that no `except Exception` in a script swallows it."""


class _Anything:
    """An object that accepts every call, attribute, index and iteration (what `plt`, `fig`, `axes[0, 1]` need)."""

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, key):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))


def install_standins(device: torch.device):
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            def _attr(name):
                if name.startswith("__") and name.endswith("__"):
                    raise AttributeError(name)
                return _Anything()
            plt.__getattr__ = mpl.__getattr__ = _attr
            mpl.pyplot = plt
            sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    if "timm" not in sys.modules:
        try:
            import timm  # noqa: F401
        except ImportError:
            timm, models, layers = (types.ModuleType(n) for n in ("timm", "timm.models", "timm.models.layers"))
            layers.trunc_normal_ = torch.nn.init.trunc_normal_
            timm.models, models.layers = models, layers
            sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    try:
        import accelerate  # noqa: F401
        return
    except ImportError:
        pass

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))

    class DistributedDataParallelKwargs:
        def __init__(self, **kw):
            self.kw = kw

    class _DeviceLoader:
        def __init__(self, loader, dev):
            self.loader, self.dev, self.dataset = loader, dev, loader.dataset

        def __len__(self):
            return len(self.loader)

        def __iter__(self):
            for batch in self.loader:
                yield tuple(t.to(self.dev, non_blocking=True) if torch.is_tensor(t) else t for t in batch)

    class Accelerator:
        def __init__(self, kwargs_handlers=None, **kw):
            self.device = device
            self.process_index, self.num_processes = rank, world
            self.is_local_main_process = self.is_main_process = rank == 0
            self._ddp_kw = next((h.kw for h in (kwargs_handlers or []) if isinstance(h, DistributedDataParallelKwargs)), {})
            if world > 1 and not torch.distributed.is_initialized():
                torch.distributed.init_process_group("nccl" if device.type == "cuda" else "gloo")

        def prepare(self, *objs):
            out = []
            for o in objs:
                if isinstance(o, torch.utils.data.DataLoader):
                    o = _DeviceLoader(o, self.device)
                elif isinstance(o, torch.nn.Module):
                    o = o.to(self.device)
                    if world > 1:
                        o = torch.nn.parallel.DistributedDataParallel(
                            o, device_ids=[self.device.index] if self.device.type == "cuda" else None, **self._ddp_kw)
                out.append(o)
            return tuple(out) if len(out) > 1 else out[0]

        def backward(self, loss):
            loss.backward()

        def save(self, obj, path):
            if self.is_main_process:
                torch.save(obj, path)

        def wait_for_everyone(self):
            if world > 1:
                torch.distributed.barrier()

        def unwrap_model(self, m):
            return getattr(m, "module", m)

        def print(self, *a, **k):
            if self.is_main_process:
                print(*a, **k)

    acc = types.ModuleType("accelerate")
    acc.Accelerator, acc.DistributedDataParallelKwargs = Accelerator, DistributedDataParallelKwargs
    sys.modules["accelerate"] = acc


def synthetic_dataset(path: str, n_samples: int, bag: int):
    """The dict-like the script's Dataset class indexes, by the hard-coded file's name."""
    rng = np.random.default_rng(0)
    name = os.path.basename(path)
    f32 = np.float32
    if name.startswith("test_"):                        # eval scripts: a held-out file in the train file's schema
        name = name[len("test_"):]
        rng = np.random.default_rng(1)
    if name == "dataset_2D_drift_diffusion.npz":        # 2d_FPE/train_fno.py:20-23
        return {"trajectories": (rng.standard_normal((n_samples, bag, 61, 61)) * 1e-10).astype(f32),
                "potential": (rng.standard_normal((n_samples, 61, 61)) * 1e-21).astype(f32),
                "drag": (1.0 + 0.1 * rng.standard_normal((n_samples, 61, 61))).astype(f32) * f32(1e-6)}
    if name == "dataset_2D_drift.npz":                  # 2d_Non_conservative_FPE/train_fno.py:20-22, F is [M, 2, Nx, Ny]
        return {"trajectories": (rng.standard_normal((n_samples, bag, 80, 80)) * 1e-10).astype(f32),
                "F": (rng.standard_normal((n_samples, 2, 80, 80)) * 1e-12).astype(f32)}
    if name == "dataset_1D_drift_diffusion.npz":        # 1d_FPE/train_fno.py:17-21
        return {"trajectories": (rng.standard_normal((n_samples, bag, 80)) * 1e-5).astype(f32),
                "potential": (rng.standard_normal((n_samples, 80)) * 1e-20).astype(f32),
                "drag": (1.0 + 0.1 * rng.standard_normal(n_samples)).astype(f32) * f32(1e-5)}     # one scalar per sample (:56)
    if name == "training_data_GPE.npy":                 # 1d_GPE/train_nio_GPE.py:38-42 (np.load(...).item())
        d = {"y": np.abs(rng.standard_normal((n_samples, bag + 1, 128))).astype(f32),
             "V": np.abs(rng.standard_normal((n_samples, 128))).astype(f32),
             "g": np.abs(rng.standard_normal(n_samples)).astype(f32) + 1, "kappa": np.abs(rng.standard_normal(n_samples)).astype(f32) + 1}
        box = np.empty((), dtype=object)
        box[()] = d
        return box
    raise FileNotFoundError(f"{path}: not on disk and no synthetic schema is known for {name!r}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=40, help="synthetic dataset size (the scripts split it 80/20)")
    ap.add_argument("--bag", type=int, default=100)
    ap.add_argument("--modules", default="ours", choices=["ours", "reference"])
    ap.add_argument("--device", default=None)
    ap.add_argument("--workdir", default=os.path.join(ROOT, "gpurun_out", "script_run"))
    ap.add_argument("--save-ckpt", default=None, help="save the trained model's state_dict here when the run stops")
    ap.add_argument("--ckpt-prefix", default="", help="key prefix of the saved state_dict ('module.' = as saved under DDP)")
    ap.add_argument("--script-args", default="", help="command line handed to the script (eval scripts use argparse)")
    ap.add_argument("--seed", type=int, default=None, help="torch / numpy seed set before the script starts")
    args = ap.parse_args()
    for k in ("save_ckpt",):
        if getattr(args, k):
            setattr(args, k, os.path.abspath(getattr(args, k)))

    script = os.path.abspath(args.script)
    variant = os.path.basename(os.path.dirname(script))
    dev = torch.device(args.device or (f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if torch.cuda.is_available() else "cpu"))
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    install_standins(dev)
    if args.modules == "ours":
        from blindno_b200 import dropin
        dropin.install(variant)
    else:
        sys.path.insert(0, os.path.dirname(script))

    real_load = np.load

    def load(path, *a, **k):
        if isinstance(path, (str, os.PathLike)) and not os.path.exists(path):
            return synthetic_dataset(str(path), args.samples, args.bag)
        return real_load(path, *a, **k)

    np.load = load

    stats = {"steps": 0, "samples": 0, "t0": None, "batch": 0}
    real_mse = torch.nn.MSELoss.forward

    def mse(self, inp, target):
        stats["batch"] = int(inp.shape[0])
        return real_mse(self, inp, target)

    torch.nn.MSELoss.forward = mse
    real_step = torch.optim.Adam.step

    def step(self, *a, **k):
        out = real_step(self, *a, **k)
        stats["steps"] += 1
        if stats["steps"] == args.warmup:
            if dev.type == "cuda":
                torch.cuda.synchronize()
            stats["t0"], stats["samples"] = time.perf_counter(), 0
        elif stats["steps"] > args.warmup:
            stats["samples"] += stats["batch"]
        if stats["steps"] >= args.warmup + args.steps:
            raise _StepLimit()
        return out

    torch.optim.Adam.step = step

    # the model the script trains: the first bag model put in train mode (every script calls model.train() per epoch)
    trained = {}
    real_train = torch.nn.Module.train

    def train(self, mode=True):
        if mode and "model" not in trained and type(self).__name__.startswith(("NIOFP", "PermInv")):
            trained["model"] = self
        return real_train(self, mode)

    torch.nn.Module.train = train
    import shlex
    argv = sys.argv
    script_args = shlex.split(args.script_args)
    script_args = [os.path.abspath(a) if (a.endswith(".pt") and not os.path.isabs(a)) else a for a in script_args]
    sys.argv = [script] + script_args
    if args.seed is not None:
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)

    os.makedirs(args.workdir, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(args.workdir)
    launches0 = None
    if args.modules == "ours":
        from blindno_b200 import ops
        launches0 = ops.kernel_launches()
    status = "completed"
    try:
        runpy.run_path(script, run_name="__main__")
    except _StepLimit:
        status = "step limit reached"
    finally:
        os.chdir(cwd)
        np.load, torch.nn.MSELoss.forward, torch.optim.Adam.step = real_load, real_mse, real_step
        torch.nn.Module.train, sys.argv = real_train, argv
    if dev.type == "cuda":
        torch.cuda.synchronize()
    sec = time.perf_counter() - stats["t0"] if stats["t0"] else float("nan")
    world = int(os.environ.get("WORLD_SIZE", 1))
    res = {"script": os.path.join(variant, os.path.basename(script)), "modules": args.modules, "device": str(dev), "status": status,
           "optimizer_steps": stats["steps"], "timed_steps": max(stats["steps"] - args.warmup, 0), "batch_per_process": stats["batch"],
           "world": world, "samples_per_s": world * stats["samples"] / sec if sec == sec and sec > 0 else None,
           "note": "eager loop of the unchanged script (its own DataLoader, torch.optim.Adam, loss.item() per step)"}
    if launches0 is not None:
        from blindno_b200 import ops
        res["gpu_launches"] = ops.kernel_launches() - launches0
    if args.save_ckpt and "model" in trained and int(os.environ.get("RANK", 0)) == 0:
        model = getattr(trained["model"], "module", trained["model"])
        sd = {args.ckpt_prefix + k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        os.makedirs(os.path.dirname(args.save_ckpt) or ".", exist_ok=True)
        torch.save(sd, args.save_ckpt)
        res["checkpoint"] = {"path": os.path.relpath(args.save_ckpt, ROOT), "tensors": len(sd),
                             "model": type(model).__name__, "prefix": args.ckpt_prefix}
    if stats["steps"] == 0:           # an eval script: summarise the metrics.csv it wrote (2d_FPE/eval_fno.py:193-198,275-278)
        import csv
        import glob
        res["note"] = "unchanged eval script: checkpoint load, per-sample forward under no_grad, de-normalised relative L2"
        res.pop("samples_per_s", None)
        outdirs = [a for i, a in enumerate(script_args) if i and script_args[i - 1] == "--outdir"]
        roots = [os.path.join(args.workdir, o) if not os.path.isabs(o) else o for o in outdirs] or [args.workdir]
        for path in [q for r_ in roots for q in glob.glob(os.path.join(r_, "**", "metrics.csv"), recursive=True)][:1]:
            rows = list(csv.reader(open(path)))
            body = [[float(v) for v in r] for r in rows[1:] if r]
            res["metrics"] = {"file": os.path.relpath(path, args.workdir), "columns": rows[0], "rows": len(body),
                              "mean": [sum(c) / len(body) for c in list(zip(*body))[1:]] if body else None,
                              "values": body}
    if int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps(res))


if __name__ == "__main__":
    main()

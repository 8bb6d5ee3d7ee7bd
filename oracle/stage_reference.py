#!/usr/bin/env python
"""Stage the UNMODIFIED reference for use as a checker / CPU baseline where /root/reference does not exist.

TEST INFRASTRUCTURE ONLY (see oracle/blindno_oracle.py's header): nothing in the product imports this.

The reference (yl602019618/Reconstruction-of-PDE-without-Time-Label) is pure Python with no package and no build
system; the GPU box has no /root/reference.  This recipe copies the handful of module files the hot path lives in
(and the train / eval scripts the harness runs unchanged) from the read-only reference tree into the git-ignored
``oracle/_ref/`` -- a build artefact like the compiled .so: it travels with the repo snapshot, it never enters the
history.  ``__graft_entry__.build()`` runs it whenever the reference tree is mounted.

    python oracle/stage_reference.py [--src /root/reference] [--dst oracle/_ref]

``load(variant, module)`` then imports a staged module exactly as the reference's scripts do (from their own
directory, with the 3-line ``timm`` stand-in the 2-D ``NIOModules`` needs at import time).
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
VARIANTS = ("1d_FPE", "1d_GPE", "2d_FPE", "2d_Non_conservative_FPE")
MODULES = ("NIOModules.py", "FNOModules.py", "DeepONetModules.py", "Baselines.py", "debug_tools.py")
SCRIPTS = {          # the scripts BASELINE.json's configs name (run unchanged by tools/run_reference_script.py)
    "1d_FPE": ("train_fno.py", "eval_fno.py"),
    "1d_GPE": ("train_nio_GPE.py",),
    "2d_FPE": ("train_fno.py", "eval_fno.py", "train_nio.py"),
    "2d_Non_conservative_FPE": ("train_fno.py",),
}


def stage(src: str = "/root/reference", dst: str = DST) -> dict:
    """Copy the files; returns {relative path: sha256}.  Idempotent."""
    if not os.path.isdir(os.path.join(src, "2d_FPE")):
        raise FileNotFoundError(f"reference tree not found at {src}")
    manifest = {}
    for var in VARIANTS:
        os.makedirs(os.path.join(dst, var), exist_ok=True)
        names = list(MODULES) + [s for s in SCRIPTS.get(var, ()) if os.path.exists(os.path.join(src, var, s))]
        for name in names:
            a, b = os.path.join(src, var, name), os.path.join(dst, var, name)
            shutil.copyfile(a, b)
            manifest[f"{var}/{name}"] = hashlib.sha256(open(b, "rb").read()).hexdigest()
        model_dir = os.path.join(src, var, "model")          # Transolver package the 2-D NIOModules imports at the top
        if os.path.isdir(model_dir):
            out = os.path.join(dst, var, "model")
            shutil.rmtree(out, ignore_errors=True)
            shutil.copytree(model_dir, out, ignore=shutil.ignore_patterns("__pycache__", "._*", "*.pyc"))
            for root, _, files in os.walk(out):
                for f in files:
                    path = os.path.join(root, f)
                    manifest[os.path.relpath(path, dst)] = hashlib.sha256(open(path, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    return manifest


def available(dst: str = DST) -> bool:
    return os.path.exists(os.path.join(dst, "2d_FPE", "NIOModules.py"))


def root(dst: str = DST) -> str:
    """Directory the staged (or, in the build container, the mounted) reference variants live under."""
    if available(dst):
        return dst
    if os.path.isdir("/root/reference/2d_FPE"):
        return "/root/reference"
    raise FileNotFoundError("no reference: oracle/_ref is not staged and /root/reference is not mounted")


def _timm_stub():
    if "timm" in sys.modules:
        return
    import torch
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


_SIBLINGS = ("NIOModules", "FNOModules", "DeepONetModules", "Baselines", "debug_tools", "model")


def load(variant: str, module: str = "NIOModules"):
    """Import ``module`` from the staged ``<variant>`` directory the way the reference's scripts do (cwd on sys.path).
    The four directories hold same-named modules, so the module cache is cleared of them first."""
    _timm_stub()
    path = os.path.join(root(), variant)
    for name in list(sys.modules):
        if name in _SIBLINGS or name.startswith("model."):
            del sys.modules[name]
    sys.path.insert(0, path)
    try:
        return importlib.import_module(module)
    finally:
        sys.path.remove(path)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=DST)
    a = ap.parse_args()
    m = stage(a.src, a.dst)
    print(f"staged {len(m)} files under {a.dst}")

"""CPU: tools/run_reference_script.py drives the reference's UNCHANGED train scripts -- stand-ins
for matplotlib / accelerate / timm, a synthetic dataset in each script's schema, a step limit.  Here (no GPU) the
harness is validated with the script directory's own modules; on a GPU box the same command with the default
``--modules ours`` routes the script's ``from NIOModules import ...`` to blindno_b200.dropin."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "1d_FPE")), reason="reference tree not mounted")
@pytest.mark.parametrize("script,batch", [("1d_FPE/train_fno.py", 32), ("1d_GPE/train_nio_GPE.py", 8)])
def test_unchanged_script_runs_under_the_harness(script, batch, tmp_path):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), os.path.join(REF, script), "--steps", "2",
           "--warmup", "1", "--samples", "40" if batch == 32 else "10", "--bag", "60", "--modules", "reference", "--device", "cpu",
           "--workdir", str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["status"] == "step limit reached" and out["optimizer_steps"] == 3 and out["batch_per_process"] == batch
    assert out["samples_per_s"] > 0


def test_synthetic_datasets_follow_the_scripts_schemas():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import run_reference_script as H
    finally:
        sys.path.pop(0)
    d = H.synthetic_dataset("/x/dataset_2D_drift_diffusion.npz", 5, 7)
    assert d["trajectories"].shape == (5, 7, 61, 61) and d["potential"].shape == (5, 61, 61) and d["drag"].shape == (5, 61, 61)
    d = H.synthetic_dataset("/x/dataset_2D_drift.npz", 5, 7)
    assert d["F"].shape == (5, 2, 80, 80)
    d = H.synthetic_dataset("/x/dataset_1D_drift_diffusion.npz", 5, 7)
    assert d["trajectories"].shape == (5, 7, 80) and d["drag"].shape == (5,)
    d = H.synthetic_dataset("/x/training_data_GPE.npy", 5, 7).item()
    assert d["y"].shape == (5, 8, 128) and d["V"].shape == (5, 128) and d["g"].shape == (5,)
    with pytest.raises(FileNotFoundError):
        H.synthetic_dataset("/x/unknown.npz", 1, 1)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "2d_FPE")), reason="reference tree not mounted")
def test_eval_script_path_checkpoint_with_ddp_prefix(tmp_path):
    """The eval path of the harness, validated with the reference's own modules: train_fno.py for two steps, the model
    saved as under DDP (``module.`` keys), eval_fno.py unchanged on that file (its own prefix stripping and
    load_state_dict), metrics.csv summarised."""
    h = os.path.join(ROOT, "tools", "run_reference_script.py")
    ckpt = str(tmp_path / "ck.pt")
    res = subprocess.run([sys.executable, h, os.path.join(REF, "2d_FPE", "train_fno.py"), "--steps", "1", "--warmup", "1", "--samples",
                          "10", "--bag", "60", "--modules", "reference", "--device", "cpu", "--workdir", str(tmp_path / "t"),
                          "--save-ckpt", ckpt, "--ckpt-prefix", "module."], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["checkpoint"]["tensors"] > 100 and os.path.exists(ckpt)
    res = subprocess.run([sys.executable, h, os.path.join(REF, "2d_FPE", "eval_fno.py"), "--modules", "reference", "--device", "cpu",
                          "--samples", "4", "--bag", "60", "--workdir", str(tmp_path / "e"), "--script-args",
                          f"--ckpt {ckpt} --outdir out --start 0 --end 1 --device cpu"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["status"] == "completed" and out["metrics"]["rows"] == 2 and out["metrics"]["columns"][1] == "rel_l2_drift"

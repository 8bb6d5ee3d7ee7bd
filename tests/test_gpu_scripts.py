"""GPU (-m gpu): the reference's UNCHANGED scripts against this package (BASELINE.json north_star: "the train_fno.py,
train_nio.py and eval_*.py scripts run unchanged").  The scripts are the staged copies under oracle/_ref
(oracle/stage_reference.py); tools/run_reference_script.py supplies the absent datasets / matplotlib / accelerate.

The eval test is the drop-in proof for the checkpoint contract: train_fno.py trains the drop-in model on the GPU, its
state_dict is saved with DDP's ``module.`` prefix, and eval_fno.py -- unchanged, with its own load_checkpoint_robust /
load_state_dict(strict=False) -- is run twice on that file: with this package's modules on the GPU and with the
reference's own modules on the CPU.  The metrics.csv the script writes must agree."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
HARNESS = os.path.join(ROOT, "tools", "run_reference_script.py")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "2d_FPE", "train_fno.py")),
                               reason="reference not staged (python oracle/stage_reference.py in the build container)")


def _run(args, timeout=900):
    res = subprocess.run([sys.executable, HARNESS] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    return json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])


@needs_ref
@pytest.mark.parametrize("script,batch,extra", [
    ("2d_FPE/train_fno.py", 4, []),
    ("2d_Non_conservative_FPE/train_fno.py", 4, []),
    ("1d_FPE/train_fno.py", 32, ["--samples", "80"]),
    ("1d_GPE/train_nio_GPE.py", 16, ["--samples", "20"]),      # the script's batch of 32, capped by the 80 % train split
    ("2d_FPE/train_nio.py", 4, []),
])
def test_unchanged_train_script_runs_on_the_cuda_path(script, batch, extra, tmp_path):
    out = _run([os.path.join(REF, script), "--steps", "4", "--warmup", "2", "--bag", "100", "--workdir", str(tmp_path)] + extra)
    assert out["modules"] == "ours" and out["device"].startswith("cuda")
    assert out["status"] == "step limit reached" and out["optimizer_steps"] == 6 and out["batch_per_process"] == batch
    assert out["gpu_launches"] > 0 and out["samples_per_s"] > 0


@needs_ref
def test_unchanged_eval_script_agrees_with_the_reference_modules_on_a_trained_checkpoint(tmp_path):
    ckpt = str(tmp_path / "model_checkpoint_best.pt")
    train = _run([os.path.join(REF, "2d_FPE", "train_fno.py"), "--steps", "6", "--warmup", "1", "--samples", "20", "--workdir",
                  str(tmp_path / "train"), "--save-ckpt", ckpt, "--ckpt-prefix", "module.", "--seed", "5"])
    assert train["checkpoint"]["model"] == "NIOFP2D_FNO" and train["checkpoint"]["prefix"] == "module."
    common = ["--samples", "8", "--script-args", f"--ckpt {ckpt} --outdir out --start 0 --end 5"]
    ours = _run([os.path.join(REF, "2d_FPE", "eval_fno.py"), "--workdir", str(tmp_path / "ours")] + common)
    assert ours["modules"] == "ours" and ours["gpu_launches"] > 0 and ours["status"] == "completed"
    common[-1] += " --device cpu"
    ref = _run([os.path.join(REF, "2d_FPE", "eval_fno.py"), "--modules", "reference", "--device", "cpu", "--workdir",
                str(tmp_path / "ref")] + common)
    assert ours["metrics"]["rows"] == ref["metrics"]["rows"] == 6
    for a, b in zip(ours["metrics"]["values"], ref["metrics"]["values"]):
        assert a[0] == b[0]
        for x, y in zip(a[1:], b[1:]):          # the reported relative L2 errors, to the precision the script prints (6 decimals)
            assert abs(x - y) <= 5e-6 * max(abs(y), 1.0), (a, b)

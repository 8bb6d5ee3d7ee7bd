"""CPU: the bench.py contract of the reference arm (the only arm that runs without a GPU) -- one JSON line on stdout
with the keys the driver reads, on the smallest workload."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "1d_FPE",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nio_fno_train_samples_per_sec" and d["unit"] == "samples/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["config"]["workload"] == "1d_FPE" and "model" not in d["config"]
    from oracle import stage_reference as S
    try:
        S.root()
        kind = "reference"          # the unmodified reference modules are staged (or mounted): the arm runs THEM
    except FileNotFoundError:
        kind = "port"
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == kind
    assert d["steps"] == 1 and d["warmup"] == 1, "the arm must time exactly the steps it was asked for"
    # the two arms' config dicts must be equal: both come from make_config (the CUDA arm cannot run here)
    import bench
    args = type("A", (), {"workload": "1d_FPE", "batch_per_gpu": 0})()
    assert d["config"] == bench.make_config(args, bench.WORKLOADS["1d_FPE"], 1)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_default_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)

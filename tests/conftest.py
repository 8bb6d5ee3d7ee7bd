import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# ---------------------------------------------------------------------------------------------------------------
# parity report: every gradient / output comparison the GPU tests make through tests.test_gpu_parity._grad_check is
# recorded (measured error, the reference's own fp32 error, which clause of the rule admitted it) and written to
# gpurun_out/parity_report.json when a -m gpu session ends; the committed copy is profiles/parity_report.json.
# ---------------------------------------------------------------------------------------------------------------
PARITY_RECORDS = []


def pytest_sessionfinish(session, exitstatus):
    if not PARITY_RECORDS:
        return
    import json
    cases = {}
    for r in PARITY_RECORDS:
        cases.setdefault(r["case"], []).append({k: v for k, v in r.items() if k != "case"})
    clauses = {}
    for r in PARITY_RECORDS:
        clauses[r["admitted_by"]] = clauses.get(r["admitted_by"], 0) + 1
    worst = max(PARITY_RECORDS, key=lambda r: r["rel_err"])
    out = {"rule": "|ours - fp64 oracle| <= max(tol * max|fp64|, 3 * |fp32 oracle - fp64 oracle|, floor); "
                   "rel_err = |ours - fp64|_max / max|fp64| of that tensor",
           "comparisons": len(PARITY_RECORDS), "admitted_by": clauses,
           "worst": {k: worst[k] for k in ("case", "tensor", "rel_err", "admitted_by")},
           "worst_admitted_by_1e-5_clause": max((r["rel_err"] for r in PARITY_RECORDS if r["admitted_by"] == "tol*scale"), default=None),
           "exit_status": int(exitstatus), "cases": cases}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as fh:
        json.dump(out, fh, indent=1)

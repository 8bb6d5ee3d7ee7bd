"""CPU: the C-ABI library loads, exports every symbol the header declares, and rejects bad
arguments without touching a GPU."""
import ctypes as C
import os
import re

import pytest

from blindno_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "blindno_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bdn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    handle = C.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.lib()
    names = _declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/blindno_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), "python binding table and header disagree"


def test_abi_version_and_error_channel():
    L = _lib.lib()
    assert L.bdn_abi_version() == _lib.ABI_VERSION
    s = _lib.FnoShape()
    s.ndim = 3
    assert L.bdn_fno_workspace_bytes(C.byref(s)) == 0
    assert b"ndim" in L.bdn_last_error() or b"bad" in L.bdn_last_error()


@pytest.mark.parametrize("n", list(range(0, 400)))
def test_pad_amount_is_bankers_rounding(n):
    assert _lib.lib().bdn_pad_amount(n) == int(round(n * 0.25)) == _lib.pad_amount(n)


def _shape(**kw):
    s = _lib.FnoShape()
    base = dict(ndim=2, images=4, c_in=3, width=4, c_out=1, hidden=128, n_layers=2, h=61, w=61, hp=76, wp=76,
                out_h=61, out_w=61, m1=12, m2=12, prec=0)
    base.update(kw)
    for k, v in base.items():
        setattr(s, k, v)
    return s


def test_shape_validation_without_gpu():
    L = _lib.lib()
    ok = _shape()
    assert L.bdn_fno_workspace_bytes(C.byref(ok)) > 0
    assert L.bdn_fno_act_floats(C.byref(ok)) == 3 * 4 * 4 * 76 * 76
    # kept spectra of every layer + (few-image nets only) the mode-major copy of the spectral weights
    assert L.bdn_fno_spec_floats(C.byref(ok)) == 2 * 4 * 4 * 24 * 12 * 2 + 2 * 12 * 24 * 4 * 4 * 2
    many = _shape(images=400)
    assert L.bdn_fno_spec_floats(C.byref(many)) == 2 * 400 * 4 * 24 * 12 * 2
    # overlapping row blocks (Q14), too many columns, too many layers, 1-D with rows
    for bad in (_shape(m1=39), _shape(m2=40), _shape(n_layers=9), _shape(ndim=1), _shape(width=0)):
        assert L.bdn_fno_workspace_bytes(C.byref(bad)) == 0
        assert len(L.bdn_last_error()) > 0
    one_d = _shape(ndim=1, h=1, hp=1, out_h=1, w=80, wp=100, out_w=80, m1=0, m2=15, width=30, c_in=30)
    assert L.bdn_fno_workspace_bytes(C.byref(one_d)) > 0


def test_null_pointers_are_rejected_not_dereferenced():
    L = _lib.lib()
    s = _shape()
    rc = L.bdn_fno_forward(C.byref(s), None, None, None, None, None, None, 0, None)
    assert rc == -1
    rc = L.bdn_adam_step(None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None)
    assert rc == -1


def test_stage_entry_points_validate_without_gpu():
    """The stage-level entry points (what the fno_lift_pad / fno_layer / fno_project custom ops bind) reject bad
    shapes and null pointers before any CUDA call, and size their workspace on the host."""
    L = _lib.lib()
    s = _shape()
    assert L.bdn_stage_layer_workspace_bytes(C.byref(s)) > 0
    assert L.bdn_stage_layer_workspace_bytes(C.byref(_shape(m1=39))) == 0
    assert L.bdn_stage_lift_forward(C.byref(s), None, None, None, None, None) == -1
    assert L.bdn_stage_layer_forward(C.byref(s), None, 0, None, None, None, None, None, None, None, 0, None) == -1
    assert L.bdn_stage_layer_backward(C.byref(s), None, None, 0, None, None, None, None, None, None, None, None, None, None,
                                      0, None) == -1
    assert L.bdn_stage_project_forward(C.byref(s), None, None, None, None, None, None, None) == -1
    assert L.bdn_stage_project_backward(C.byref(s), None, None, None, None, None, None, 0, 1, None, None, None, None, None,
                                        None) == -1
    assert b"null" in L.bdn_last_error()
    # an empty batch is a no-op, not an error, whatever the pointers are
    empty = _shape(images=0)
    assert L.bdn_stage_project_forward(C.byref(empty), None, None, None, None, None, None, None) == 0
    assert L.bdn_bag_pool_lift_forward(None, None, None, None, None, 0, 5, 10, 2, 4, None) == 0


def test_gelu_formula_accuracy_in_fp32():
    """The exact-GELU evaluation the kernels use (csrc/bdn_internal.cuh: Abramowitz-Stegun 26.2.17, one reciprocal
    and one exponential), restated in NumPy fp32: |Phi error| <= 3.5e-7, |gelu error| <= 5e-7, |phi error| <= 1e-7
    over [-12, 12] -- at the rounding level of fp32 and below the 1e-5 parity bound by a wide margin."""
    import math
    import numpy as np
    f = np.float32
    x = np.linspace(-12, 12, 400001).astype(f)
    t = f(1) / (np.abs(x) * f(0.2316418882663604) + f(1))
    e = np.exp2((x * x) * f(-0.72134752044448170368) + f(-1.3257480647361592)).astype(f)
    p = f(1.3302745)
    for c in (-1.8212559, 1.7814779, -0.35656378, 0.31938154):
        p = (p * t + f(c)).astype(f)
    q = (p * t * e).astype(f)
    cdf = np.where(x >= 0, f(1) - q, q).astype(np.float64)
    xd = x.astype(np.float64)
    ref = np.array([0.5 * (1.0 + math.erf(v / math.sqrt(2.0))) for v in xd[::40]])
    assert np.abs(cdf[::40] - ref).max() <= 3.5e-7
    assert np.abs(xd[::40] * cdf[::40] - xd[::40] * ref).max() <= 5e-7
    # the forward-only form gelu = max(x, 0) - |x| q (no Phi): same bound
    gelu = (np.maximum(x, f(0)) - (np.abs(x) * q).astype(f)).astype(np.float64)
    assert np.abs(gelu[::40] - xd[::40] * ref).max() <= 5e-7
    assert np.abs(e.astype(np.float64) - np.exp(-0.5 * xd * xd) / math.sqrt(2 * math.pi)).max() <= 1e-7


def test_concurrent_loads_see_a_whole_library():
    """One process per GPU imports the package at the same time under torchrun: the check-and-build runs under a file
    lock and the library is put in place by an atomic rename (a 2-GPU run once loaded a half-written .so).  Four
    processes load it at once; every one must get a library that answers."""
    import subprocess
    import sys
    code = ("from blindno_b200 import _lib, build; import os; L = _lib.lib(); "
            "assert os.path.exists(os.path.join(build.LIBDIR, '.build.lock')) or os.path.exists(build.LIBPATH); "
            "print(L.bdn_abi_version())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = [subprocess.Popen([sys.executable, "-c", code], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for _ in range(4)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-400:] for o in outs]
    assert len({o[0].strip() for o in outs}) == 1

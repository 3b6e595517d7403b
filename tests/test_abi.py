"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/coslam_b200.h
declares, keeps the reference's POD layouts, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from coslam_b200 import api
from coslam_b200.ctypes_defs import BaOptions, BaProblem, KltConfig, KltFeature, PoseOpt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "coslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cosl_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    syms = header_symbols()
    assert len(syms) > 40
    lib = C.CDLL(api.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_pod_layouts_match_the_reference_structs():
    # KLT_TrackedFeature {int status; float pos[2]; float gain; int fed;}  (v3d_gpuklt.h:166-176)
    assert C.sizeof(KltFeature) == 20
    assert KltFeature.pos.offset == 4 and KltFeature.gain.offset == 12 and KltFeature.fed.offset == 16
    # KLT_SequenceTrackerConfig: 11 reference fields + compat extension
    assert C.sizeof(KltConfig) == 48
    assert C.sizeof(BaProblem) == 8 * 11 and C.sizeof(BaOptions) == 8 + 4 + 4 + 40 + 4 + 4
    assert C.sizeof(PoseOpt) % 8 == 0


def test_defaults_match_the_reference():
    c = KltConfig()
    api.LIB.cosl_klt_config_default(C.byref(c))
    r = KltConfig.reference_defaults()
    for f, _ in KltConfig._fields_:
        assert getattr(c, f) == getattr(r, f), f
    o = PoseOpt()
    api.LIB.cosl_pose_opt_default(C.byref(o))
    assert (o.maxIterLM, o.maxIterRW, o.lambda0) == (100, 5, 1e-3)
    assert (o.epsErrorChangeLM, o.epsParamChangeLM, o.epsErrorChangeRW) == (1e-7, 1e-6, 1e-6)
    b = BaOptions()
    api.LIB.cosl_ba_options_default(C.byref(b))
    assert (b.max_err, b.outer_iters, b.inner_iters) == (6.0, 5, 10)
    assert list(b.opts) == [1e-3 * 1e-4, 1e-12, 1e-12, 0.0, 1e-16]


def test_version_and_error_string():
    assert b"sm_100a" in api.LIB.cosl_version()
    assert api.kernel_launch_count() >= 0


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu():
    with pytest.raises(api.CoslError):
        api.KltTracker(KltConfig.coslam_live(), 640, 480, 6, 32, 32)
    from coslam_b200 import synth
    prob, _ = synth.make_ba_scene(2, 2, 50, 320, 240, seed=1, m_con=1, n_con=0)
    with pytest.raises(api.CoslError):
        api.ba_solve(prob, BaOptions.defaults())
    K, R0, t0, Ms, ms, _, _ = synth.make_pose_case(n_pts=8)
    with pytest.raises(api.CoslError):
        api.pose_intracam(K, R0, t0, Ms, ms, 10.0)


def test_bad_arguments_return_error_codes():
    h = C.c_void_p()
    cfg = KltConfig.coslam_live()
    rc = api.LIB.cosl_klt_create(C.byref(cfg), 4, 4, 6, 32, 32, 0, 0, 0, C.byref(h))
    assert rc == -1 and b"geometry" in api.LIB.cosl_last_error()
    assert api.LIB.cosl_klt_advance(None) == -1
    assert api.LIB.cosl_ba_solver_run(None, None) == -1


def test_product_tree_does_not_reach_into_the_oracle():
    """oracle/ is test infrastructure: nothing under coslam_b200/ may import, include or load it,
    and the shared library must not be linked against it."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(#\s*include\s*[\"<][^\">]*oracle|from\s+oracle|import\s+oracle|.*liboracle)",
                     re.M)
    for dirpath, _, files in os.walk(os.path.join(root, "coslam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), os.path.join(dirpath, f)
    lib = os.path.join(root, "coslam_b200", "libcoslam_b200.so")
    deps = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    assert "oracle" not in deps

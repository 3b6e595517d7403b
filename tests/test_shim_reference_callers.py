"""Drop-in checks against the reference's OWN sources (syntax-only compiles, CPU, dev container only:
/root/reference does not exist on the GPU boxes, the tests skip there).

* tracking/GPUKLT.cpp -- the only user of V3D_GPU::KLT_SequenceTracker -- must compile UNCHANGED
  against coslam_b200/shim/v3d_gpuklt.h (pre-included: it carries the reference header's include
  guard, so the reference's CGKLT/v3d_gpuklt.h, which sits next to GPUKLT.h, becomes a no-op).
  LibVisualSLAM, which the reference needs but does not ship, is replaced by the minimal
  declarations in tests/stubs/.
* intraCamEstimate / IntraCamPoseOption: the same probe translation unit must compile against the
  reference's slam/SL_IntraCamPose.h and against the shim header.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
STUBS = os.path.join(ROOT, "tests", "stubs")
SHIM = os.path.join(ROOT, "coslam_b200", "shim")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("g++") is None,
                                reason="needs the reference tree and g++")


def _syntax_only(src, *flags):
    cmd = ["g++", "-std=c++11", "-fsyntax-only", *flags, "-I", STUBS, "-I",
           os.path.join(ROOT, "include"), "-I", REF, src]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-4000:]


def test_reference_gpuklt_facade_compiles_unchanged_against_the_shim():
    _syntax_only(os.path.join(REF, "tracking", "GPUKLT.cpp"), "-include",
                 os.path.join(SHIM, "v3d_gpuklt.h"), "-I", os.path.join(REF, "tracking"))


def test_pose_probe_compiles_against_the_reference_header():
    _syntax_only(os.path.join(STUBS, "probe_pose.cpp"))


def test_pose_probe_compiles_against_the_shim_header():
    _syntax_only(os.path.join(STUBS, "probe_pose.cpp"), "-include",
                 os.path.join(SHIM, "SL_IntraCamPose.h"))


def test_posegraph_probe_compiles_against_the_reference_header():
    _syntax_only(os.path.join(STUBS, "probe_posegraph.cpp"))


def test_posegraph_probe_compiles_against_the_shim_header():
    _syntax_only(os.path.join(STUBS, "probe_posegraph.cpp"), "-include",
                 os.path.join(SHIM, "SL_GlobalPoseEstimation.h"))

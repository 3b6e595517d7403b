"""The reference-named C++ shims (coslam_b200/shim/*.h) compile and link against the C-ABI library
from caller code shaped like the reference's own call sites (tests/shim_caller.cpp); on a GPU box
the caller is also executed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "shim_caller")
    lib = os.path.join(ROOT, "coslam_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-Wall", "-Werror",
                           "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "coslam_b200", "shim"),
                           os.path.join(ROOT, "tests", "shim_caller.cpp"), "-o", exe, "-L" + lib,
                           "-lcoslam_b200", "-Wl,-rpath," + lib])
    return exe


def test_shims_compile_and_link(tmp_path):
    exe = _build(str(tmp_path))
    out = subprocess.check_output([exe]).decode()
    assert "coslam_b200" in out


@pytest.mark.gpu
def test_shims_run_on_gpu(tmp_path):
    exe = _build(str(tmp_path))
    out = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert out.returncode == 0, f"rc={out.returncode}\n{out.stdout}{out.stderr}"
    assert "klt:" in out.stdout and "pose: ok 1" in out.stdout and "ba:" in out.stdout
    assert "posegraph: ok" in out.stdout


def test_ba_shim_host_logic_with_a_mock_abi(tmp_path):
    """CPU: bundleAdjustRobust's flattening (CSR by point), option forwarding, write-back of R/t/X
    and outlier flags, and the exception on failure -- against a mock of the C-ABI."""
    exe = os.path.join(str(tmp_path), "mock_ba_shim")
    src = os.path.join(ROOT, "tests", "stubs", "mock_ba_shim.cpp")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "tests", "stubs"),
                           "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "coslam_b200", "shim"), src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "MOCK_BA_SHIM_OK" in out.stdout, out.stdout + out.stderr


def test_klt_shim_host_logic_with_a_mock_abi(tmp_path):
    """CPU: V3D_GPU::KLT_SequenceTracker forwards configuration, geometry, image pitch, feature
    buffers and counts to the C-ABI exactly as the reference's class receives them."""
    exe = os.path.join(str(tmp_path), "mock_klt_shim")
    src = os.path.join(ROOT, "tests", "stubs", "mock_klt_shim.cpp")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "coslam_b200", "shim"), src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "MOCK_KLT_SHIM_OK" in out.stdout, out.stdout + out.stderr


def test_pose_shim_host_logic_with_a_mock_abi(tmp_path):
    """CPU: intraCamEstimate forwards the option fields, copies the diagnostics back and maps the
    two failure modes to `false`."""
    exe = os.path.join(str(tmp_path), "mock_pose_shim")
    src = os.path.join(ROOT, "tests", "stubs", "mock_pose_shim.cpp")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "coslam_b200", "shim"), src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "MOCK_POSE_SHIM_OK" in out.stdout, out.stdout + out.stderr


def test_posegraph_shim_host_logic_with_a_mock_abi(tmp_path):
    """CPU: GlobalPoseGraph flattens nodes / edges for cosl_posegraph_spread_chains, keeps the
    reference's newR / newt semantics across the two method calls with ONE ABI call, refuses
    non-chain graphs and turns ABI failures into exceptions."""
    exe = os.path.join(str(tmp_path), "mock_posegraph_shim")
    src = os.path.join(ROOT, "tests", "stubs", "mock_posegraph_shim.cpp")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "coslam_b200", "shim"), src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "MOCK_POSEGRAPH_SHIM_OK" in out.stdout, out.stdout + out.stderr

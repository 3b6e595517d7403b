"""Two-rank NCCL bundle adjustment against the single-GPU solve (needs >= 2 GPUs on the box; the
host-side sharding logic is covered on CPU by tests/test_dist_cpu.py with gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_ba_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "mgpu_ba_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MGPU_PARITY_OK" in out.stdout + out.stderr


def test_single_process_multi_gpu_drop_in():
    """cosl_ba_solve_multi (the `n_gpus` of SURVEY.md 8b): threads of ONE process drive 2 GPUs; result
    equals the single-GPU drop-in.  Points are in map-creation order, so the two shards see
    different camera pairs (the structure exchange of the solver is exercised)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from coslam_b200 import api, synth
    from coslam_b200.ctypes_defs import BaOptions
    prob, _ = synth.make_ba_scene(4, 30, 5000, 1280, 720, seed=23, m_con=4, n_con=3, sort_by_home=True)
    opt = BaOptions.defaults()
    opt.outer_iters, opt.inner_iters = 2, 6
    p1, p2 = prob.copy(), prob.copy()
    i1 = api.ba_solve(p1, opt)
    i2 = api.ba_solve_multi(p2, opt, 2)
    assert i1[10] == i2[10]
    assert abs(i1[1] - i2[1]) <= 1e-9 * i1[1]
    assert np.abs(p1.X - p2.X).max() < 1e-6 and np.abs(p1.R - p2.R).max() < 1e-8
    assert np.array_equal(p1.outlier, p2.outlier)

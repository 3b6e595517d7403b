"""Repeated cosl_ba_solve calls on the c4 scene with COSL_BA_TIMING=1: shows the host-side setup cost
(index building, pooled allocations, uploads, graph capture) call after call."""
import os, sys, time
os.environ.setdefault("COSL_BA_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions

prob, _ = synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                              seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)
opt = BaOptions.defaults()
opt.outer_iters, opt.inner_iters = 1, 10
for i in range(4):
    p = prob.copy()
    t0 = time.perf_counter()
    info = api.ba_solve(p, opt)
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.1f} ms, trials {info[9]:.0f}", flush=True)

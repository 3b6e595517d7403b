"""Round-2 experiment driver for the reduced-system solve: c4 scene once, then per nested-dissection
depth the per-class timers of a few fixed LM trials, parity of the cost against depth 0, and the task
timeline of the default plan."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions

t0 = time.time()
prob, _ = synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                              seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)
print("scene", prob.m, prob.n, prob.nobs, f"{time.time() - t0:.1f}s", flush=True)
depths = sys.argv[1].split(",") if len(sys.argv) > 1 else ["-1"]
ref = None
for dep in depths:
    os.environ["COSL_BA_ND_DEPTH"] = dep
    for grid in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"]):
        if grid != "0":
            os.environ["COSL_BA_SOLVE_GRID"] = grid
        s = api.BaSolver(prob.copy(), BaOptions.defaults())
        s.run_fixed(3)
        s.reset()
        s.profile_enable(True)
        t1 = time.perf_counter()
        info = s.run_fixed(8)
        wall = time.perf_counter() - t1
        tm = s.timers()
        s.profile_enable(False)
        if ref is None:
            ref = info[1]
        print(f"depth {dep} grid {grid}: plan {s.plan_info()} cost {info[1]:.12g} rel diff {abs(info[1] - ref) / ref:.2e} "
              f"wall/trial {1e3 * wall / 8:.3f} ms", {k: round(1e3 * v[0] / max(1, v[1]), 1) for k, v in tm.items()},
              flush=True)
        s.close()

"""Where a cosl_ba_solve call spends its time (COSL_BA_TIMING=1 makes the library print the host-side
breakdown): c4 global BA and the c3 / c2 local BA, second call of each (warm memory pool)."""
import os, sys, time
os.environ["COSL_BA_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions
cases = {"c4": lambda: synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                                            seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)[0],
         "c3": lambda: synth.make_ba_scene(4, 5, 20000, 1280, 720, seed=synth.BASE_SEED + 3, m_con=8, n_con=2)[0],
         "c2": lambda: synth.make_ba_scene(2, 5, 5000, 640, 480, seed=synth.BASE_SEED + 2, m_con=4, n_con=2)[0]}
for name, mk in cases.items():
    prob = mk()
    o = BaOptions.defaults()
    if name == "c4":
        o.outer_iters, o.inner_iters = 1, 10
    api.ba_solve(prob.copy(), o)
    for rep in range(2):
        p = prob.copy()
        t0 = time.perf_counter()
        info = api.ba_solve(p, o)
        print(f"== {name} call {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms, {int(info[10])} LM trials, "
              f"device loop {1e3 * info[11]:.2f} ms", file=sys.stderr, flush=True)

"""c3 / c2 local BA: per-class device time of an LM trial for the solver variants (env switches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions
for name, (nc, npts, W, H) in {"c3": (4, 20000, 1280, 720), "c2": (2, 5000, 640, 480)}.items():
    prob, _ = synth.make_ba_scene(nc, 5, npts, W, H, seed=synth.BASE_SEED + (3 if nc == 4 else 2), m_con=2 * nc, n_con=2)
    for env in ({}, {"COSL_BA_SCHUR_ROWS": "1"}, {"COSL_BA_SMALL_KERNEL": "1"}):
        for k in ("COSL_BA_SCHUR_ROWS", "COSL_BA_SMALL_KERNEL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        s = api.BaSolver(prob.copy(), BaOptions.defaults())
        s.run_fixed(5); s.reset(); s.profile_enable(True)
        info = s.run_fixed(40)
        tm = s.timers()
        print(name, env, "cost", f"{info[1]:.9g}", "us/trial by class", {k: round(1e3 * v[0] / 40, 1) for k, v in tm.items()},
              "total", round(sum(1e3 * v[0] / 40 for v in tm.values()), 1), flush=True)
        s.close()

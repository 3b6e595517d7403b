"""One c3 (or c2) local-BA solver, a few fixed trials: target of the ncu launch list (tools/r2_local_ncu.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
nc, npts, W, H = {"c3": (4, 20000, 1280, 720), "c2": (2, 5000, 640, 480)}[name]
prob, _ = synth.make_ba_scene(nc, 5, npts, W, H, seed=synth.BASE_SEED + (3 if nc == 4 else 2), m_con=2 * nc, n_con=2)
s = api.BaSolver(prob.copy(), BaOptions.defaults())
s.run_fixed(3); s.reset()
info = s.run_fixed(int(sys.argv[2]) if len(sys.argv) > 2 else 6)
print(name, "cost", info[1], "trials", info[9], s.plan_info(), s.stats())
s.close()

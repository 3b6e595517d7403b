"""Small driver for ncu: BA c4 (or c3 with argv[2] == 'c3') -- a few fixed LM trials."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "c4"
if which == "c4":
    prob, _ = synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                                  seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)
else:
    prob, _ = synth.make_ba_scene(4, 5, 20000, 1280, 720, seed=synth.BASE_SEED + 3, m_con=8, n_con=2)
s = api.BaSolver(prob, BaOptions.defaults())
s.run_fixed(2)
s.reset()
s.profile_enable(True)
info = s.run_fixed(trials)
print({k: (round(v[0], 3), v[1], round(1e3 * v[0] / max(1, v[1]), 2)) for k, v in s.timers().items()})
print(which, "trials", info[9], "cost", info[0], "->", info[1], "launches", api.kernel_launch_count())

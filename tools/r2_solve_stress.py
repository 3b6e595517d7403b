"""Stress the dataflow solve for ordering bugs: the same LM trial from the same start, many times; a task that
read a tile or a solution block too early shows up as a cost that differs by more than atomic-order noise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name, mk in {"c4": lambda: synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                                                     seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)[0],
                 "kf40": lambda: synth.make_ba_scene(4, 40, 8000, 1280, 720, seed=7, m_con=4, n_con=0, sort_by_home=True)[0]}.items():
    prob = mk()
    s = api.BaSolver(prob.copy(), BaOptions.defaults())
    costs = []
    for r in range(n_rep):
        s.reset()
        info = s.run_fixed(2)
        costs.append(info[1])
    c = np.asarray(costs)
    print(name, "runs", n_rep, "cost", c[0], "max rel deviation", float(np.abs(c - c[0]).max() / c[0]), flush=True)
    assert np.abs(c - c[0]).max() <= 1e-9 * c[0]
    s.close()
print("SOLVE_STRESS_OK")

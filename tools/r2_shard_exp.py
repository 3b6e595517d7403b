"""Single-GPU timing of ONE shard of the c4 problem (as a rank of an N-rank run would see it, without
the collectives): separates shard-structure effects from communication effects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions
prob, _ = synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                              seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)
for N in (1, 2, 8):
    shard, (lo, hi) = prob.shard(0, N)
    s = api.BaSolver(shard, BaOptions.defaults())
    s.run_fixed(3); s.reset(); s.profile_enable(True)
    info = s.run_fixed(8)
    print(f"shard 0 of {N}: points {shard.n} obs {shard.nobs}", s.stats(), {k: round(1e3 * v[0] / max(1, v[1]), 1) for k, v in s.timers().items()}, flush=True)
    s.close()

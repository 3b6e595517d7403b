"""Timeline of the persistent reduced-system solve (ba_tile_solve) on the c4 scene: per task type the
execution and wait times, the makespan, and the chain of tasks that determined it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from coslam_b200 import api, synth
from coslam_b200.ctypes_defs import BaOptions

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c4":
    prob, _ = synth.make_ba_scene(bench.BA_CAMS, bench.BA_KF, bench.BA_PTS, bench.KLT_W, bench.KLT_H,
                                  seed=synth.BASE_SEED + 4, m_con=bench.BA_CAMS, n_con=0)
else:
    prob, _ = synth.make_ba_scene(4, 40, 6000, 1280, 720, seed=16, m_con=4, n_con=0)
s = api.BaSolver(prob, BaOptions.defaults())
print("plan", s.plan_info(), s.stats())
s.run_fixed(3)
s.reset()
s.trace_arm()
s.run_fixed(1)
tm, meta = s.trace_get()
t0 = tm[:, 1].min()
tick, ready, done = [(tm[:, k].astype(np.int64) - int(t0)) / 1e3 for k in (1, 2, 3)]
names = ["POTRF", "TRSM", "UPD", "BWD", "SUM"]
print(f"makespan {done.max():.1f} us, {len(tm)} tasks on {len(np.unique(tm[:, 0]))} SMs")
for ty in range(5):
    m = meta[:, 0] == ty
    if m.any():
        ex = done[m] - ready[m]
        wt = ready[m] - tick[m]
        print(f"{names[ty]:6s} n {m.sum():5d}  exec us: mean {ex.mean():6.2f} p50 {np.median(ex):6.2f} max {ex.max():6.2f}"
              f"   wait us: mean {wt.mean():7.2f} max {wt.max():7.2f}")
        ph = tm[m][:, 4:8].astype(np.int64)
        if (ph[:, 0] > 0).any():
            rd = tm[m][:, 2].astype(np.int64)
            dn = tm[m][:, 3].astype(np.int64)
            segs = [ph[:, 0] - rd]
            for k in range(1, 4):
                if (ph[:, k] > 0).any():
                    segs.append(ph[:, k] - ph[:, k - 1])
                    last = ph[:, k]
                else:
                    break
            else:
                last = ph[:, 3]
            last = ph[:, max(k for k in range(4) if (ph[:, k] > 0).any())]
            segs.append(dn - last)
            print("       phases us (mean):", [round(float(np.mean(sg)) / 1e3, 2) for sg in segs])
# chain: walk back from the last task through the latest-finishing task that ended before ready
order = np.argsort(done)
cur = int(np.argmax(done))
chain = []
while True:
    chain.append(cur)
    prev = [t for t in range(len(tm)) if done[t] <= ready[cur] + 0.05 and done[t] > ready[cur] - 1.5 and t != cur]
    if not prev or ready[cur] - tick[cur] < 0.2:
        # did not wait: predecessor is whatever the same SM ran before
        same = [t for t in range(len(tm)) if tm[t, 0] == tm[cur, 0] and done[t] <= tick[cur] + 0.05 and t != cur]
        if not same:
            break
        cur = max(same, key=lambda t: done[t])
    else:
        cur = max(prev, key=lambda t: done[t])
    if len(chain) > 400:
        break
chain.reverse()
cnt = {n: 0 for n in names}
tim = {n: 0.0 for n in names}
for t in chain:
    cnt[names[meta[t, 0]]] += 1
    tim[names[meta[t, 0]]] += done[t] - ready[t]
print("chain length", len(chain), "by type", cnt, "exec us by type", {k: round(v, 1) for k, v in tim.items()})
print("chain head:", [(names[meta[t, 0]], int(meta[t, 1]), int(meta[t, 2]), round(float(done[t]), 1)) for t in chain[:12]])
print("chain tail:", [(names[meta[t, 0]], int(meta[t, 1]), int(meta[t, 2]), round(float(done[t]), 1)) for t in chain[-12:]])

"""Timeline of the persistent reduced-system solve (ba_tile_solve) on the c4 scene: per task type the
execution / wait times and phase stamps, the makespan, and the EXACT critical chain (for every task
the dependency -- counter reaching its required value -- that was satisfied last, or the previous
task of the same CTA if it did not have to wait)."""
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
tasks, lists = s.tasks()
t0 = int(tm[:, 1].min())
tick, ready, done = [(tm[:, k].astype(np.int64) - t0) / 1e3 for k in (1, 2, 3)]
names = ["POTRF", "TRSM", "UPD", "BWD", "SUM"]
print(f"makespan {done.max():.1f} us, {len(tm)} tasks on {len(np.unique(tm[:, 0]))} SMs")
for ty in range(5):
    m = meta[:, 0] == ty
    if not m.any():
        continue
    ex, wt = done[m] - ready[m], ready[m] - tick[m]
    print(f"{names[ty]:6s} n {m.sum():5d}  exec us: mean {ex.mean():6.2f} p50 {np.median(ex):6.2f} max {ex.max():6.2f}"
          f"   wait us: mean {wt.mean():7.2f} max {wt.max():7.2f}")
    ph = (tm[m][:, 4:8].astype(np.int64) - t0) / 1e3
    used = [k for k in range(4) if (tm[m][:, 4 + k] > 0).all()]
    if used:
        pts = [ready[m]] + [ph[:, k] for k in used] + [done[m]]
        print("       phases us (mean):", [round(float(np.mean(pts[i + 1] - pts[i])), 2) for i in range(len(pts) - 1)])
# exact critical chain
n = len(tasks)
inc = {}  # counter -> sorted done times of its incrementers
for t in range(n):
    inc.setdefault(int(tasks[t, 7]), []).append((done[t], t))
for k in inc:
    inc[k].sort()


def binding(t):
    """(time, task) of the last-satisfied dependency of task t."""
    best = (-1.0, -1)
    deps = [(int(tasks[t, 8 + 2 * w]), int(tasks[t, 9 + 2 * w])) for w in range(3) if tasks[t, 8 + 2 * w] >= 0]
    if tasks[t, 0] in (3, 4):
        deps += [(int(lists[e, 0]), int(lists[e, 1])) for e in range(tasks[t, 14], tasks[t, 15])]
    for ci, v in deps:
        if v <= 0:
            continue
        tt, who = inc[ci][v - 1]
        if tt > best[0]:
            best = (tt, who)
    return best


by_sm = {}
for t in np.argsort(tick):
    by_sm.setdefault(int(tm[t, 0]), []).append(int(t))
prev_on_sm = {}
for sm, lst in by_sm.items():
    for a, b in zip(lst[:-1], lst[1:]):
        prev_on_sm[b] = a
cur = int(np.argmax(done))
chain = []
while cur >= 0 and len(chain) < 2000:
    chain.append(cur)
    bt, who = binding(cur)
    if who >= 0 and bt >= tick[cur] - 0.3:   # it waited for (or barely missed waiting for) that dependency
        cur = who
    else:
        cur = prev_on_sm.get(cur, -1)         # ready at once: the CTA itself was the limit
chain.reverse()
cnt = {nm: 0 for nm in names}
tim = {nm: 0.0 for nm in names}
gap = 0.0
for a, b in zip(chain[:-1], chain[1:]):
    gap += max(0.0, ready[b] - done[a])
for t in chain:
    cnt[names[meta[t, 0]]] += 1
    tim[names[meta[t, 0]]] += done[t] - ready[t]
print("critical chain:", len(chain), "tasks", cnt, "exec us", {k: round(v, 1) for k, v in tim.items()},
      "hand-over gaps us", round(gap, 1))
print("chain:", " ".join(f"{names[meta[t, 0]][0]}{meta[t, 1]}" + (f".{meta[t, 2]}" if meta[t, 0] in (1, 2, 4) else "") +
                          f"@{done[t]:.0f}" for t in chain))

"""Generates the committed golden fixtures in tests/golden/ from the CPU oracle.

The reference itself cannot be executed in this environment (its KLT is Cg shaders, its BA solver
lives in LibVisualSLAM/sba-1.6 which is not in the tree), so these vectors freeze the ORACLE's
outputs on small seeded inputs: the CPU suite checks that the oracle still reproduces them
(guards against silent drift of the restatement), the GPU suite checks the CUDA path against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from coslam_b200 import synth  # noqa: E402
from coslam_b200.ctypes_defs import BaOptions, KltConfig  # noqa: E402
from oracle import orc  # noqa: E402

W, H, L, FW, FH = 160, 120, 4, 12, 10


def klt():
    seq = synth.ImageSequence(H, W, 4242, n_frames=3)
    out = {"frames": np.stack(seq.frames)}
    for gain in (1, 0):
        cfg = KltConfig.coslam_live(with_gain=bool(gain))
        cfg.minCornerness = 1200.0
        k = orc.OracleKlt(cfg, W, H, L, FW, FH)
        f0, n0 = k.first(seq.frames[0])
        out[f"g{gain}_f0"], out[f"g{gain}_n0"] = f0, n0
        if gain:
            for l in range(L):
                out[f"pyr{l}"] = k.pyramid(0, l)  # after advance: frame 0's pyramid is pyr0
            out["corn0"] = k.cornerness()
        f1, n1 = k.next(seq.frames[1])
        f2, n2 = k.next(seq.frames[2])
        out[f"g{gain}_f1"], out[f"g{gain}_n1"] = f1, n1
        out[f"g{gain}_f2"], out[f"g{gain}_n2"] = f2, n2
    np.savez_compressed(os.path.join(HERE, "klt_small.npz"), **out)


def pose():
    K, R0, t0, Ms, ms, Rt, tt = synth.make_pose_case(n_pts=48, seed=99)
    ok, R, t, opt = orc.pose_intracam(K, R0, t0, Ms, ms, 10.0)
    np.savez_compressed(os.path.join(HERE, "pose_small.npz"), K=K, R0=R0, t0=t0, Ms=Ms, ms=ms,
                        ok=ok, R=R, t=t, err=opt.err, nIterRW=opt.nIterRW, lam=opt.lambda_)


def ba():
    prob, truth = synth.make_ba_scene(2, 3, 150, 320, 240, seed=31, m_con=2, n_con=2)
    opt = BaOptions.defaults()
    opt.outer_iters, opt.inner_iters = 2, 8
    p = prob.copy()
    info = orc.ba_solve(p, opt)
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), K=prob.K, R=prob.R, t=prob.t, X=prob.X,
                        ptr=prob.ptr, cam=prob.cam, xy=prob.xy, m_con=2, n_con=2, R_out=p.R,
                        t_out=p.t, X_out=p.X, outlier=p.outlier, info=info)


if __name__ == "__main__":
    klt()
    pose()
    ba()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))

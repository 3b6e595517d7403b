"""Second, independent implementation of intraCamEstimate: a line-by-line numpy reading of the
in-tree reference (slam/SL_IntraCamPose.cpp:10-39 exp map, :43-117 forward-difference Jacobians,
:259-303 weighted LM step, :367-380 pose update, :439-456 weighted error, :475-549 LM procedure,
:626-709 Tukey re-weighting loop) against the C++ oracle.  The two share nothing but the reference
text; agreement pins the oracle against transcription errors (the reference has no tests)."""
import numpy as np
import pytest

from coslam_b200 import synth

EPS = 1e-8


def exp_map(w):  # :10-39
    th = np.sqrt(w @ w)
    if th == 0:
        return np.eye(3)
    h = w / th
    st, ct = np.sin(th), 1 - np.cos(th)
    return np.array([
        [-ct * h[1] * h[1] - ct * h[2] * h[2] + 1, ct * h[0] * h[1] - st * h[2], st * h[1] + ct * h[0] * h[2]],
        [st * h[2] + ct * h[0] * h[1], -ct * h[0] * h[0] - ct * h[2] * h[2] + 1, ct * h[1] * h[2] - st * h[0]],
        [ct * h[0] * h[2] - st * h[1], st * h[0] + ct * h[1] * h[2], -ct * h[0] * h[0] - ct * h[1] * h[1] + 1]])


def project(K, R, t, M):
    x = K @ (R @ M + t)
    return x[:2] / x[2]


def lm_step(K, R, t, Ws, Ms, ms, lam):  # :259-303
    sA, sB = np.zeros((6, 6)), np.zeros(6)
    for w, M, m in zip(Ws, Ms, ms):
        rm = project(K, R, t, M)
        J = np.empty((2, 6))
        for a in range(3):  # :43-83
            e = np.zeros(3)
            e[a] = EPS
            J[:, a] = (project(K, R @ exp_map(e), t, M) - rm) / EPS
        for a in range(3):  # :88-117
            t1 = t.copy()
            t1[a] += EPS
            J[:, 3 + a] = (project(K, R, t1, M) - rm) / EPS
        J *= w
        sA += J.T @ J
        sB += J.T @ ((m - rm) * w)
    sA[np.diag_indices(6)] += lam
    return np.linalg.inv(sA) @ sB


def err2w(K, R, t, Ws, Ms, ms):  # :439-456
    return sum(w * ((m - project(K, R, t, M)) ** 2).sum() for w, M, m in zip(Ws, Ms, ms))


class Opt:
    maxIterLM, maxIterRW = 100, 5
    epsErrorChangeLM, epsParamChangeLM, epsErrorChangeRW = 1e-7, 1e-6, 1e-6
    lambda0 = 1e-3


def lm_proc(K, R0, t0, Ws, Ms, ms, o):  # :475-549
    o.lam = o.lambda0
    o.err = err = err2w(K, R0, t0, Ws, Ms, ms)
    R, t = R0.copy(), t0.copy()
    R_opt, t_opt = R, t
    R_tmp, t_tmp = R, t
    ret = 1
    i = 0
    while i < o.maxIterLM:
        p = lm_step(K, R, t, Ws, Ms, ms, o.lam)
        R_opt, t_opt = R @ exp_map(p[:3]), t + p[3:]
        if p @ p < o.epsParamChangeLM:
            ret = 0
            break
        err = err2w(K, R_opt, t_opt, Ws, Ms, ms)
        if abs(err - o.err) < o.epsErrorChangeLM:
            ret = 0
            break
        if err <= o.err:
            R, t = R_opt, t_opt
            R_tmp, t_tmp = R_opt, t_opt
            o.err = err
            o.lam /= 10
        else:
            o.lam *= 10
            if o.lam > 1e18:
                ret = -1
                break
        i += 1
    if ret == -1:
        R_opt, t_opt = R_tmp, t_tmp
    o.err = err
    return ret >= 0, R_opt, t_opt


def tukey(e, tau):
    return 0.0 if e >= tau else (1 - (e / tau) ** 2) ** 2


def intra_cam_estimate(K, R0, t0, Ms, ms, tau, prev_errs=None):  # :626-709
    o = Opt()
    Ws = np.ones(len(Ms)) if prev_errs is None else np.array([tukey(abs(e), tau) for e in prev_errs])
    R, t = R0.copy(), t0.copy()
    errRW = -1.0
    k = 0
    R_opt, t_opt = R, t
    while k < o.maxIterRW:
        ok, R_opt, t_opt = lm_proc(K, R, t, Ws, Ms, ms, o)
        if not ok:
            return False, R_opt, t_opt, k
        o.lambda0 = o.lam
        if errRW < 0:
            errRW = o.err
        elif abs(o.err - errRW) < o.epsErrorChangeRW:
            return True, R_opt, t_opt, k
        else:
            errRW = o.err
        R, t = R_opt, t_opt
        Ws = np.array([tukey(np.sqrt(((project(K, R, t, M) - m) ** 2).sum()), tau) for M, m in zip(Ms, ms)])
        k += 1
    return True, R_opt, t_opt, k


@pytest.mark.parametrize("seed,with_prev", [(5, False), (6, True), (7, False)])
def test_pose_oracle_matches_numpy_reading_of_the_reference(orc, seed, with_prev):
    K, R0, t0, Ms, ms, Rt, tt = synth.make_pose_case(64, 1280, 720, seed=seed)
    prev = None
    if with_prev:
        prev = np.array([np.sqrt(((project(K, R0, t0, M) - m) ** 2).sum()) for M, m in zip(Ms, ms)])
    ok_o, R_o, t_o, opt = orc.pose_intracam(K, R0, t0, Ms, ms, 10.0, prev_errs=prev)
    ok_n, R_n, t_n, k_n = intra_cam_estimate(K, R0, t0, Ms, ms, 10.0, prev)
    assert ok_o and ok_n
    assert opt.nIterRW == k_n
    # forward differences with eps = 1e-8 carry ~1e-8 relative noise into every LM step; both
    # implementations stop on the same tests, so they end within that noise of each other
    assert np.abs(R_o - R_n).max() < 1e-6 and np.abs(t_o - t_n).max() < 1e-6
    assert np.abs(R_n - Rt).max() < 5e-3  # and both are near the truth

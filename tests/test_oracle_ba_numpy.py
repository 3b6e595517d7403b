"""Second implementation of one SBA Levenberg-Marquardt trial: dense normal equations in numpy
(numeric Jacobian of the KRTS projection with the local-quaternion rotation update of
app/SL_CoSLAMBA.cpp:290-378, mu = tau * max diag(J^T J), (J^T J + mu I) delta = J^T eps) against the
oracle's Schur-complement path.  Same step => the Schur contraction, the reduced solve, the back
substitution, the parameterisation and the damping initialisation of the oracle are all consistent
with the textbook dense formulation."""
import numpy as np
import pytest

from coslam_b200 import synth
from coslam_b200.ctypes_defs import BaOptions


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def apply(base, p):
    """Parameters p = [free cameras (v, t) ..., free points X ...] on top of `base`."""
    q = base.copy()
    mf, nf = base.m - base.m_con, base.n - base.n_con
    cams = p[:6 * mf].reshape(mf, 6)
    R0 = base.R.reshape(-1, 3, 3)
    for a in range(mf):
        j = base.m_con + a
        v = cams[a, :3]
        dq = np.concatenate([[np.sqrt(1.0 - v @ v)], v])  # _MK_QUAT_FRM_VEC
        q.R[j] = (quat_to_R(dq) @ R0[j]).ravel()
        q.t[j] = cams[a, 3:]
    q.X[base.n_con:] = p[6 * mf:].reshape(nf, 3)
    return q


def pack(base):
    mf = base.m - base.m_con
    cams = np.zeros((mf, 6))
    cams[:, 3:] = base.t[base.m_con:]
    return np.concatenate([cams.ravel(), base.X[base.n_con:].ravel()])


@pytest.mark.parametrize("max_err", [0.0, 6.0])
def test_one_lm_trial_equals_dense_normal_equations(orc, max_err):
    """max_err = 0: the plain sba_motstr_levmar_x step; max_err = 6: one round of the robust wrapper,
    i.e. Tukey weights w = (1 - (|e| / max_err)^2)^2 (0 beyond max_err) frozen at the start, applied
    as sqrt(w) to residuals and Jacobian."""
    prob, _ = synth.make_ba_scene(3, 3, 60, 640, 480, seed=11, m_con=3, n_con=2)
    opt = BaOptions.defaults()
    opt.max_err = max_err
    p0 = pack(prob)
    r0 = apply(prob, p0).residuals()
    if max_err > 0:
        en = np.sqrt((r0 * r0).sum(1)) / max_err
        sw = np.sqrt(np.where(en < 1.0, (1 - en * en) ** 2, 0.0)).repeat(2)
    else:
        sw = np.ones(r0.size)
    eps = lambda p: apply(prob, p).residuals().ravel() * sw   # sqrt(w) (x - xhat)
    e0 = eps(p0)
    h = 1e-6
    J = np.empty((e0.size, p0.size))                      # d xhat / d p = -d eps / d p
    for k in range(p0.size):
        d = np.zeros_like(p0)
        d[k] = h
        J[:, k] = -(eps(p0 + d) - eps(p0 - d)) / (2 * h)
    JtJ = J.T @ J
    mu = opt.opts[0] * JtJ.diagonal().max()
    delta = np.linalg.solve(JtJ + mu * np.eye(p0.size), J.T @ e0)
    want = apply(prob, p0 + delta)
    cost_want = float((eps(p0 + delta) ** 2).sum())
    assert cost_want < (e0 ** 2).sum()                    # the first trial is a descent step
    got = prob.copy()
    info = orc.ba_run_fixed(got, opt, 1)
    assert int(info[9]) == 1
    assert abs(info[0] - (e0 ** 2).sum()) <= 1e-9 * info[0]
    assert abs(info[1] - cost_want) <= 1e-6 * cost_want
    assert np.abs(got.X - want.X).max() < 1e-6
    assert np.abs(got.t - want.t).max() < 1e-6
    assert np.abs(got.R - want.R).max() < 1e-7

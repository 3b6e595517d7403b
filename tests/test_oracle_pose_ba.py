"""Pins of the pose and BA oracles: analytic-vs-numeric Jacobians, recovery of known ground truth,
and an independent cross-check of the BA optimum against scipy.optimize.least_squares."""
import ctypes as C

import numpy as np
import pytest

from coslam_b200 import synth
from coslam_b200.ctypes_defs import BaOptions, PoseOpt


def test_pose_recovers_ground_truth(orc):
    K, R0, t0, Ms, ms, Rt, tt = synth.make_pose_case(n_pts=192, seed=1, noise_px=0.0,
                                                     outlier_frac=0.0)
    ok, R, t, opt = orc.pose_intracam(K, R0, t0, Ms, ms, 10.0)
    assert ok and np.abs(R - Rt).max() < 1e-6 and np.abs(t - tt).max() < 1e-5
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)


def test_pose_is_robust_to_outliers(orc):
    K, R0, t0, Ms, ms, Rt, tt = synth.make_pose_case(n_pts=192, seed=2, noise_px=0.3,
                                                     outlier_frac=0.15)
    ok, R, t, opt = orc.pose_intracam(K, R0, t0, Ms, ms, 10.0)
    assert ok and np.abs(R - Rt).max() < 2e-3 and np.abs(t - tt).max() < 0.02
    assert opt.nIterRW >= 1 and opt.lambda_ > 0


def test_pose_prev_errs_zero_weight(orc):
    K, R0, t0, Ms, ms, Rt, tt = synth.make_pose_case(n_pts=50, seed=3, noise_px=0.0,
                                                     outlier_frac=0.0)
    ms2 = ms.copy()
    ms2[:10] += 300.0  # grossly wrong, but flagged by prevErrs >= tau -> weight 0 in round 1
    prev = np.zeros(50)
    prev[:10] = 50.0
    ok, R, t, _ = orc.pose_intracam(K, R0, t0, Ms, ms2, 10.0, prev)
    assert ok and np.abs(R - Rt).max() < 1e-5


def test_so3_exp_and_projection(orc):
    L = orc.lib()
    w = np.array([0.3, -0.2, 0.5])
    R = np.empty(9)
    L.orc_so3_exp(w.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
    from scipy.spatial.transform import Rotation
    assert np.allclose(R.reshape(3, 3), Rotation.from_rotvec(w).as_matrix(), atol=1e-14)


def test_ba_jacobians_match_numeric(orc):
    rng = np.random.default_rng(0)
    K = np.array([900.0, 0.3, 640, 0, 910.0, 360, 0, 0, 1])
    for _ in range(5):
        q0 = rng.normal(size=4)
        q0 /= np.linalg.norm(q0)
        v = rng.normal(0, 0.05, 3)
        t = rng.normal(0, 1, 3)
        X = np.array([rng.normal(), rng.normal(), 8.0]) @ orc.quat2mat(q0)  # in front of the camera
        xy, A, B = orc.ba_project(K, q0, v, t, X)
        eps = 1e-6
        An, Bn = np.zeros((2, 6)), np.zeros((2, 3))
        for k in range(3):
            d = np.zeros(3)
            d[k] = eps
            An[:, k] = (orc.ba_project(K, q0, v + d, t, X)[0] - orc.ba_project(K, q0, v - d, t, X)[0]) / (2 * eps)
            An[:, 3 + k] = (orc.ba_project(K, q0, v, t + d, X)[0] - orc.ba_project(K, q0, v, t - d, X)[0]) / (2 * eps)
            Bn[:, k] = (orc.ba_project(K, q0, v, t, X + d)[0] - orc.ba_project(K, q0, v, t, X - d)[0]) / (2 * eps)
        assert np.abs(A - An).max() < 1e-5 * np.abs(A).max()
        assert np.abs(B - Bn).max() < 1e-5 * np.abs(B).max()
    # projection agrees with K (R X + t) for v = 0
    R = orc.quat2mat(q0)
    xy, _, _ = orc.ba_project(K, q0, np.zeros(3), t, X)
    u = K.reshape(3, 3) @ (R @ X + t)
    assert np.allclose(xy, u[:2] / u[2])


def test_ba_recovers_noise_free_scene(orc):
    prob, truth = synth.make_ba_scene(2, 4, 300, 640, 480, seed=2, m_con=2, n_con=0, noise_px=0.0,
                                      outlier_frac=0.0)
    opt = BaOptions.defaults()
    opt.max_err, opt.outer_iters, opt.inner_iters = 0.0, 1, 50
    r0 = prob.rms()
    info = orc.ba_solve(prob, opt)
    assert r0 > 0.5 and prob.rms() < 1e-6
    assert info[6] in (1, 2, 3, 6)


def test_ba_optimum_matches_scipy(orc):
    """Independent check of the unweighted optimum: scipy TRF on the same residual (rotation
    vector parameterisation) reaches the same cost."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    prob, truth = synth.make_ba_scene(2, 3, 60, 640, 480, seed=8, m_con=2, n_con=0, noise_px=0.5,
                                      outlier_frac=0.0)
    p = prob.copy()
    opt = BaOptions.defaults()
    opt.max_err, opt.outer_iters, opt.inner_iters = 0.0, 1, 100
    info = orc.ba_solve(p, opt)
    m, n, mc = prob.m, prob.n, prob.m_con
    pt = np.repeat(np.arange(n), np.diff(prob.ptr))
    Kf = prob.K.reshape(-1, 3, 3)

    def res(z):
        rv = z[:3 * (m - mc)].reshape(-1, 3)
        tv = z[3 * (m - mc):6 * (m - mc)].reshape(-1, 3)
        X = z[6 * (m - mc):].reshape(-1, 3)
        R = np.concatenate([prob.R[:mc].reshape(-1, 3, 3), Rotation.from_rotvec(rv).as_matrix()])
        t = np.concatenate([prob.t[:mc], tv])
        Pc = np.einsum("nij,nj->ni", R[prob.cam], X[pt]) + t[prob.cam]
        u = np.einsum("nij,nj->ni", Kf[prob.cam], Pc)
        return (prob.xy - u[:, :2] / u[:, 2:3]).ravel()

    z0 = np.concatenate([Rotation.from_matrix(p.R[mc:].reshape(-1, 3, 3)).as_rotvec().ravel(),
                         p.t[mc:].ravel(), p.X.ravel()])
    cost_oracle = (res(z0) ** 2).sum()
    assert abs(cost_oracle - info[1]) < 1e-6 * info[1]
    sol = least_squares(res, z0, method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-12)
    assert (sol.fun ** 2).sum() >= cost_oracle * (1 - 1e-6)  # scipy cannot do better from there
    assert abs((sol.fun ** 2).sum() - cost_oracle) < 1e-5 * cost_oracle


def test_ba_robust_rounds_flag_outliers(orc):
    prob, truth = synth.make_ba_scene(2, 5, 2000, 640, 480, seed=11, m_con=4, n_con=2)
    opt = BaOptions.defaults()
    opt.outer_iters, opt.inner_iters = 2, 10
    info = orc.ba_solve(prob, opt)
    flagged = prob.outlier.astype(bool)
    true = truth["is_outlier"]
    assert (flagged & true).sum() > 0.9 * true.sum()
    assert (flagged & ~true).sum() < 0.01 * len(true)
    assert prob.rms(~true) < 0.8
    assert info[13] == flagged.sum()


def test_ba_constraints_and_fixed_mode(orc):
    prob, _ = synth.make_ba_scene(2, 4, 500, 640, 480, seed=12, m_con=3, n_con=5)
    p = prob.copy()
    opt = BaOptions.defaults()
    orc.ba_solve(p, opt)
    assert np.array_equal(p.R[:3], prob.R[:3]) and np.array_equal(p.t[:3], prob.t[:3])
    assert np.array_equal(p.X[:5], prob.X[:5])
    assert not np.array_equal(p.X[5:], prob.X[5:])
    q = prob.copy()
    info = orc.ba_run_fixed(q, opt, 4)
    assert info[9] == 4

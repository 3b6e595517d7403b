"""ctypes loader for oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by anything under coslam_b200/."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

from coslam_b200.ctypes_defs import (COSL_BA_INFOSZ, FEAT_DTYPE, BaOptions, BaProblem, KltConfig,
                                     KltFeature, PoseOpt)

# libgomp spin-waiting fights with OpenBLAS' own worker threads (10x slowdowns measured)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")


def _quota_threads():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, -(-int(q) // int(per))))
    except Exception:
        pass
    return max(1, n)


# OpenBLAS (dlopen'ed below for dpotrf) starts one worker per LOGICAL cpu unless told otherwise; on a
# box with 128 cpus and a 16-cpu quota its idle workers alone eat the quota.
os.environ.setdefault("OPENBLAS_NUM_THREADS", str(_quota_threads()))

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_klt_create.restype = C.c_void_p
    L.orc_klt_create.argtypes = [C.POINTER(KltConfig)] + [C.c_int] * 7
    L.orc_klt_destroy.argtypes = [C.c_void_p]
    L.orc_klt_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                 C.c_void_p, C.POINTER(C.c_int)]
    for f in (L.orc_klt_redetect, L.orc_klt_track):
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int)]
    L.orc_klt_feed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.orc_klt_advance.argtypes = [C.c_void_p]
    for f in (L.orc_klt_set_margin, L.orc_klt_set_conv, L.orc_klt_set_ssd):
        f.argtypes = [C.c_void_p, C.c_float]
        f.restype = None
    L.orc_klt_debug_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_klt_debug_cornerness.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_klt_last_num_candidates.argtypes = [C.c_void_p]
    L.orc_pose_intracam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                    C.POINTER(PoseOpt)]
    L.orc_ba_solve.argtypes = [C.POINTER(BaProblem), C.POINTER(BaOptions), C.c_void_p]
    L.orc_ba_run_fixed.argtypes = [C.POINTER(BaProblem), C.POINTER(BaOptions), C.c_int, C.c_void_p]
    L.orc_ba_cost.argtypes = [C.POINTER(BaProblem)]
    L.orc_ba_cost.restype = C.c_double
    L.orc_ba_project.argtypes = [C.c_void_p] * 8
    L.orc_ba_project.restype = None
    L.orc_mat2quat.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_quat2mat.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_sba_motstr_levmar_x.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                          C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                          C.c_int, C.c_void_p, C.c_void_p]
    L.orc_posegraph_spread.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6
    L.orc_set_lapack.argtypes = [C.c_char_p]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_set_threads.restype = None
    # multi-threaded LAPACK for the dense reduced-system solve (the reference links system LAPACK)
    try:
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                       "libscipy_openblas*.so"))
        if cands:
            L.orc_set_lapack(os.path.abspath(cands[0]).encode())
    except Exception:
        pass
    _LIB = L
    return L


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    """Host threads the oracle may use: min(CPU affinity, cgroup CPU quota) -- the GPU boxes
    expose 128 logical CPUs but cap the container at a fraction of them; oversubscribing the
    quota makes OpenMP dramatically slower, which would be unfair to the CPU baseline."""
    n = int(lib().orc_get_max_threads())
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return max(1, n)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleKlt:
    """Same Python surface as coslam_b200.api.KltTracker (single camera)."""

    def __init__(self, cfg, width, height, n_levels, feat_w, feat_h, pl_w=0, pl_h=0):
        self.L = lib()
        self.cfg = cfg
        self.W, self.H, self.nl, self.fw, self.fh = width, height, n_levels, feat_w, feat_h
        self.F = feat_w * feat_h
        self.h = self.L.orc_klt_create(C.byref(cfg), width, height, n_levels, feat_w, feat_h,
                                       pl_w, pl_h)
        if not self.h:
            raise ValueError("orc_klt_create failed")
        self.feats = (KltFeature * self.F)()

    def close(self):
        if self.h:
            self.L.orc_klt_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _out(self):
        return np.frombuffer(bytes(self.feats), dtype=FEAT_DTYPE).copy()

    def detect(self, img, present=None):
        img = np.ascontiguousarray(img, np.uint8)
        n = C.c_int(0)
        npres = 0 if present is None else len(present)
        pres = None if present is None else np.ascontiguousarray(present, np.float32)
        rc = self.L.orc_klt_detect(self.h, _ptr(img), img.strides[0], npres, _ptr(pres),
                                   C.byref(self.feats), C.byref(n))
        assert rc == 0
        return self._out(), n.value

    def redetect(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        n = C.c_int(0)
        rc = self.L.orc_klt_redetect(self.h, _ptr(img), img.strides[0], C.byref(self.feats),
                                     C.byref(n))
        assert rc == 0
        return self._out(), n.value

    def track(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        n = C.c_int(0)
        rc = self.L.orc_klt_track(self.h, _ptr(img), img.strides[0], C.byref(self.feats),
                                  C.byref(n))
        assert rc == 0
        return self._out(), n.value

    def feed(self, pts3):
        pts3 = np.ascontiguousarray(pts3, np.float32).reshape(-1, 3)
        ids = np.full(len(pts3), -1, np.int32)
        n = C.c_int(0)
        rc = self.L.orc_klt_feed(self.h, len(pts3), _ptr(pts3), _ptr(ids), C.byref(n))
        assert rc == 0
        return ids, n.value

    def advance(self):
        self.L.orc_klt_advance(self.h)

    def next(self, img):
        """GPUKLT::next (tracking/GPUKLT.cpp:144-161): redetect + advanceFrame."""
        out = self.redetect(img)
        self.advance()
        return out

    def first(self, img):
        """GPUKLT::first (tracking/GPUKLT.cpp:133-142): detect + advanceFrame."""
        out = self.detect(img)
        self.advance()
        return out

    def set_margin(self, m):
        self.L.orc_klt_set_margin(self.h, m)

    def set_conv(self, t):
        self.L.orc_klt_set_conv(self.h, t)

    def set_ssd(self, t):
        self.L.orc_klt_set_ssd(self.h, t)

    def pyramid(self, which, level):
        w, h = self.W >> level, self.H >> level
        out = np.empty((h, w, 3), np.float32)
        rc = self.L.orc_klt_debug_pyramid(self.h, which, level, _ptr(out), None, None)
        assert rc == 0
        return out

    def cornerness(self):
        out = np.empty((self.H, self.W), np.float32)
        self.L.orc_klt_debug_cornerness(self.h, _ptr(out))
        return out

    def num_candidates(self):
        return self.L.orc_klt_last_num_candidates(self.h)


def pose_intracam(K, R0, t0, Ms, ms, tau, prev_errs=None, opt=None):
    L = lib()
    K = np.ascontiguousarray(K, np.float64).ravel()
    R0 = np.ascontiguousarray(R0, np.float64).ravel()
    t0 = np.ascontiguousarray(t0, np.float64).ravel()
    Ms = np.ascontiguousarray(Ms, np.float64).reshape(-1, 3)
    ms = np.ascontiguousarray(ms, np.float64).reshape(-1, 2)
    pe = None if prev_errs is None else np.ascontiguousarray(prev_errs, np.float64)
    opt = opt if opt is not None else PoseOpt.defaults()
    R = np.empty(9)
    t = np.empty(3)
    ok = L.orc_pose_intracam(_ptr(K), _ptr(R0), _ptr(t0), len(Ms), _ptr(pe), _ptr(Ms), _ptr(ms),
                             float(tau), _ptr(R), _ptr(t), C.byref(opt))
    return bool(ok), R.reshape(3, 3), t, opt


def posegraph_spread(fixed, R, t, id1, id2, eR, et):
    """Post-BA pose-graph spreading on a general edge set (posegraph_oracle.cpp).  Returns newR,
    newt; raises when a system is rank deficient."""
    fx = np.ascontiguousarray(fixed, np.int32)
    n = fx.size
    R = np.ascontiguousarray(R, np.float64).reshape(n, 3, 3)
    t = np.ascontiguousarray(t, np.float64).reshape(n, 3)
    id1 = np.ascontiguousarray(id1, np.int32)
    id2 = np.ascontiguousarray(id2, np.int32)
    ne = id1.size
    eR = np.ascontiguousarray(eR, np.float64).reshape(ne, 3, 3)
    et = np.ascontiguousarray(et, np.float64).reshape(ne, 3)
    nR, nt = np.empty((n, 3, 3)), np.empty((n, 3))
    rc = lib().orc_posegraph_spread(n, _ptr(fx), _ptr(R), _ptr(t), ne, _ptr(id1), _ptr(id2), _ptr(eR),
                                    _ptr(et), _ptr(nR), _ptr(nt))
    if rc != 0:
        raise RuntimeError("orc_posegraph_spread: rank-deficient system")
    return nR, nt


def ba_solve(prob, opt):
    """prob: coslam_b200.problem.BAProblem (updated in place). Returns info[16]."""
    info = np.zeros(COSL_BA_INFOSZ)
    s = prob.struct()
    rc = lib().orc_ba_solve(C.byref(s), C.byref(opt), _ptr(info))
    assert rc == 0, rc
    return info


def ba_run_fixed(prob, opt, trials):
    info = np.zeros(COSL_BA_INFOSZ)
    s = prob.struct()
    rc = lib().orc_ba_run_fixed(C.byref(s), C.byref(opt), int(trials), _ptr(info))
    assert rc == 0, rc
    return info


def ba_cost(prob):
    s = prob.struct()
    return float(lib().orc_ba_cost(C.byref(s)))


def ba_project(K, q0, v, t, X):
    xy, A, B = np.empty(2), np.empty(12), np.empty(6)
    a = [np.ascontiguousarray(z, np.float64).ravel() for z in (K, q0, v, t, X)]
    lib().orc_ba_project(*[_ptr(z) for z in a], _ptr(xy), _ptr(A), _ptr(B))
    return xy, A.reshape(2, 6), B.reshape(2, 3)


def mat2quat(R):
    q = np.empty(4)
    R = np.ascontiguousarray(R, np.float64).ravel()
    lib().orc_mat2quat(_ptr(R), _ptr(q))
    return q


def quat2mat(q):
    R = np.empty(9)
    q = np.ascontiguousarray(q, np.float64).ravel()
    lib().orc_quat2mat(_ptr(q), _ptr(R))
    return R.reshape(3, 3)

"""Thin ctypes binding of coslam_b200/libcoslam_b200.so (the C-ABI of include/coslam_b200.h).

This is plumbing for tests/ and bench.py -- the product is the shared library.  There is NO CPU
fallback: if the library is missing, importing this module raises; if no CUDA device is present,
the first call that needs one returns COSL_E_CUDA and raises CoslError."""
import ctypes as C
import os

import numpy as np

from .ctypes_defs import (COSL_BA_INFOSZ, FEAT_DTYPE, BaOptions, BaProblem, KltConfig, KltFeature,
                          PoseOpt)

_HERE = os.path.dirname(os.path.abspath(__file__))
# COSLAM_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("COSLAM_B200_LIB") or os.path.join(_HERE, "libcoslam_b200.so")


class CoslError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
            "coslam_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, cd = C.c_void_p, C.c_int, C.c_float, C.c_double
    pint = C.POINTER(C.c_int)
    L.cosl_last_error.restype = C.c_char_p
    L.cosl_version.restype = C.c_char_p
    L.cosl_kernel_launch_count.restype = C.c_uint64
    L.cosl_klt_config_default.argtypes = [C.POINTER(KltConfig)]
    L.cosl_klt_create.argtypes = [C.POINTER(KltConfig)] + [ci] * 8 + [C.POINTER(vp)]
    L.cosl_klt_group_create.argtypes = [C.POINTER(KltConfig)] + [ci] * 9 + [C.POINTER(vp)]
    L.cosl_klt_destroy.argtypes = [vp]
    L.cosl_klt_detect.argtypes = [vp, vp, C.c_size_t, ci, vp, vp, pint]
    L.cosl_klt_redetect.argtypes = [vp, vp, C.c_size_t, vp, pint]
    L.cosl_klt_track.argtypes = [vp, vp, C.c_size_t, vp, pint]
    L.cosl_klt_feed.argtypes = [vp, ci, vp, vp, pint]
    L.cosl_klt_advance.argtypes = [vp]
    for f in (L.cosl_klt_set_margin, L.cosl_klt_set_conv, L.cosl_klt_set_ssd):
        f.argtypes = [vp, cf]
    L.cosl_klt_group_next.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.cosl_klt_group_first.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.cosl_klt_group_next_dev.argtypes = [vp, vp, C.c_size_t]
    L.cosl_klt_group_fetch.argtypes = [vp, vp, vp]
    L.cosl_klt_group_sync.argtypes = [vp]
    L.cosl_klt_group_submit.argtypes = [vp, vp, C.c_size_t]
    L.cosl_klt_group_collect.argtypes = [vp, vp, vp]
    L.cosl_klt_stream.argtypes = [vp]
    L.cosl_klt_stream.restype = vp
    L.cosl_klt_debug_pyramid.argtypes = [vp, ci, ci, ci, vp, pint, pint]
    L.cosl_klt_debug_cornerness.argtypes = [vp, ci, vp]
    L.cosl_klt_debug_num_candidates.argtypes = [vp, ci]
    L.cosl_klt_algorithmic_bytes.argtypes = [vp]
    L.cosl_klt_algorithmic_bytes.restype = cd
    L.cosl_klt_profile_enable.argtypes = [vp, ci]
    L.cosl_klt_profile_get.argtypes = [vp, ci, C.POINTER(cd), pint]
    L.cosl_klt_profile_get.restype = C.c_char_p
    L.cosl_pose_opt_default.argtypes = [C.POINTER(PoseOpt)]
    L.cosl_pose_intracam.argtypes = [vp, vp, vp, ci, vp, vp, vp, cd, vp, vp, C.POINTER(PoseOpt),
                                     pint]
    L.cosl_pose_intracam_batch.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, cd, vp, vp, vp, vp, ci]
    L.cosl_ba_options_default.argtypes = [C.POINTER(BaOptions)]
    L.cosl_ba_solve.argtypes = [C.POINTER(BaProblem), C.POINTER(BaOptions), vp]
    L.cosl_ba_solve_multi.argtypes = [C.POINTER(BaProblem), C.POINTER(BaOptions), ci, vp, vp]
    L.cosl_sba_motstr_levmar_x.argtypes = [ci] * 4 + [vp, vp, ci, ci, vp, ci, vp, ci, ci, vp, vp, ci]
    L.cosl_nccl_unique_id.argtypes = [vp]
    L.cosl_ba_comm_create.argtypes = [vp, ci, ci, ci, C.POINTER(vp)]
    L.cosl_ba_comm_destroy.argtypes = [vp]
    L.cosl_ba_solver_create.argtypes = [C.POINTER(BaProblem), C.POINTER(BaOptions), vp,
                                        C.POINTER(vp)]
    L.cosl_ba_solver_reset.argtypes = [vp, C.POINTER(BaProblem)]
    L.cosl_ba_solver_run.argtypes = [vp, vp]
    L.cosl_ba_solver_run_fixed.argtypes = [vp, ci, vp]
    L.cosl_ba_solver_download.argtypes = [vp, C.POINTER(BaProblem)]
    L.cosl_ba_solver_destroy.argtypes = [vp]
    L.cosl_ba_solver_stream.argtypes = [vp]
    L.cosl_ba_solver_stream.restype = vp
    L.cosl_ba_solver_timer.argtypes = [vp, ci, C.POINTER(cd), pint]
    L.cosl_ba_solver_timer.restype = C.c_char_p
    L.cosl_ba_solver_profile_enable.argtypes = [vp, ci]
    L.cosl_ba_solver_stats.argtypes = [vp, C.POINTER(cd)]
    L.cosl_ba_solver_plan_info.argtypes = [vp, pint]
    L.cosl_ba_solver_trace.argtypes = [vp, ci, vp, vp, ci]
    L.cosl_ba_solver_tasks.argtypes = [vp, vp, ci, vp, ci]
    L.cosl_posegraph_spread_chains.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, vp, ci]
    return L


LIB = _load()


def _ck(rc):
    if rc != 0:
        raise CoslError(f"coslam_b200 error {rc}: {LIB.cosl_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def kernel_launch_count():
    return int(LIB.cosl_kernel_launch_count())


class KltTracker:
    """One camera.  Mirrors V3D_GPU::KLT_SequenceTracker (tracking/CGKLT/v3d_gpuklt.h:202-262)."""

    def __init__(self, cfg, width, height, n_levels, feat_w, feat_h, pl_w=0, pl_h=0, device=0):
        self.cfg, self.W, self.H, self.nl = cfg, width, height, n_levels
        self.fw, self.fh, self.F = feat_w, feat_h, feat_w * feat_h
        h = C.c_void_p()
        _ck(LIB.cosl_klt_create(C.byref(cfg), width, height, n_levels, feat_w, feat_h, pl_w, pl_h,
                                device, C.byref(h)))
        self.h = h
        self.feats = (KltFeature * self.F)()

    def close(self):
        if getattr(self, "h", None):
            LIB.cosl_klt_destroy(self.h)
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
        _ck(LIB.cosl_klt_detect(self.h, _ptr(img), img.strides[0], npres, _ptr(pres),
                                C.byref(self.feats), C.byref(n)))
        return self._out(), n.value

    def redetect(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        n = C.c_int(0)
        _ck(LIB.cosl_klt_redetect(self.h, _ptr(img), img.strides[0], C.byref(self.feats),
                                  C.byref(n)))
        return self._out(), n.value

    def track(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        n = C.c_int(0)
        _ck(LIB.cosl_klt_track(self.h, _ptr(img), img.strides[0], C.byref(self.feats), C.byref(n)))
        return self._out(), n.value

    def feed(self, pts3):
        pts3 = np.ascontiguousarray(pts3, np.float32).reshape(-1, 3)
        ids = np.full(len(pts3), -1, np.int32)
        n = C.c_int(0)
        _ck(LIB.cosl_klt_feed(self.h, len(pts3), _ptr(pts3), _ptr(ids), C.byref(n)))
        return ids, n.value

    def advance(self):
        _ck(LIB.cosl_klt_advance(self.h))

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
        _ck(LIB.cosl_klt_set_margin(self.h, m))

    def set_conv(self, t):
        _ck(LIB.cosl_klt_set_conv(self.h, t))

    def set_ssd(self, t):
        _ck(LIB.cosl_klt_set_ssd(self.h, t))

    def pyramid(self, which, level, cam=0):
        w, h = self.W >> level, self.H >> level
        out = np.empty((h, w, 3), np.float32)
        _ck(LIB.cosl_klt_debug_pyramid(self.h, cam, which, level, _ptr(out), None, None))
        return out

    def cornerness(self, cam=0):
        out = np.empty((self.H, self.W), np.float32)
        _ck(LIB.cosl_klt_debug_cornerness(self.h, cam, _ptr(out)))
        return out

    def num_candidates(self, cam=0):
        return LIB.cosl_klt_debug_num_candidates(self.h, cam)


class KltGroup:
    """C cameras of equal geometry, one launch per kernel for the whole group (batched
    GPUKLT::first / GPUKLT::next)."""

    def __init__(self, cfg, n_cams, width, height, n_levels, feat_w, feat_h, device=0):
        self.cfg, self.C, self.W, self.H, self.nl = cfg, n_cams, width, height, n_levels
        self.fw, self.fh, self.F = feat_w, feat_h, feat_w * feat_h
        h = C.c_void_p()
        _ck(LIB.cosl_klt_group_create(C.byref(cfg), n_cams, width, height, n_levels, feat_w,
                                      feat_h, 0, 0, device, C.byref(h)))
        self.h = h
        self.feats = np.zeros((n_cams, self.F), FEAT_DTYPE)
        self.counts = np.zeros(n_cams, np.int32)
        self._dest = (C.c_void_p * n_cams)(*[self.feats[c].ctypes.data for c in range(n_cams)])

    def close(self):
        if getattr(self, "h", None):
            LIB.cosl_klt_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _imgs(self, imgs):
        assert len(imgs) == self.C
        pitch = imgs[0].strides[0]
        for im in imgs:
            assert im.dtype == np.uint8 and im.strides[0] == pitch and im.shape == (self.H, self.W)
        return (C.c_void_p * self.C)(*[im.ctypes.data for im in imgs]), pitch

    def first(self, imgs):
        arr, pitch = self._imgs(imgs)
        _ck(LIB.cosl_klt_group_first(self.h, arr, pitch, self._dest, _ptr(self.counts)))
        return self.feats, self.counts

    def next(self, imgs):
        arr, pitch = self._imgs(imgs)
        _ck(LIB.cosl_klt_group_next(self.h, arr, pitch, self._dest, _ptr(self.counts)))
        return self.feats, self.counts

    def host_ptrs(self, imgs):
        """Pre-built argument for next_raw(): the per-camera pointer array of a set of host frames
        (the caller keeps the frames alive)."""
        return self._imgs(imgs)

    def next_raw(self, ptrs_pitch):
        """cosl_klt_group_next with a pointer array from host_ptrs(): no per-call marshalling."""
        _ck(LIB.cosl_klt_group_next(self.h, ptrs_pitch[0], ptrs_pitch[1], self._dest,
                                    _ptr(self.counts)))
        return self.feats, self.counts

    def next_dev(self, dev_ptrs, pitch):
        arr = (C.c_void_p * self.C)(*dev_ptrs)
        _ck(LIB.cosl_klt_group_next_dev(self.h, arr, pitch))

    def fetch(self):
        _ck(LIB.cosl_klt_group_fetch(self.h, self._dest, _ptr(self.counts)))
        return self.feats, self.counts

    def sync(self):
        _ck(LIB.cosl_klt_group_sync(self.h))

    def submit_raw(self, ptrs_pitch):
        """cosl_klt_group_submit: enqueue one frame (pointer array from host_ptrs()) and return."""
        _ck(LIB.cosl_klt_group_submit(self.h, ptrs_pitch[0], ptrs_pitch[1]))

    def submit(self, imgs):
        self.submit_raw(self._imgs(imgs))

    def collect(self):
        """cosl_klt_group_collect: wait for the oldest submitted frame, return its tables."""
        _ck(LIB.cosl_klt_group_collect(self.h, self._dest, _ptr(self.counts)))
        return self.feats, self.counts

    def stream(self):
        return LIB.cosl_klt_stream(self.h)

    def algorithmic_bytes(self):
        return float(LIB.cosl_klt_algorithmic_bytes(self.h))

    def profile_enable(self, on=True):
        _ck(LIB.cosl_klt_profile_enable(self.h, 1 if on else 0))

    def profile(self):
        out, i = {}, 0
        while True:
            ms, calls = C.c_double(0), C.c_int(0)
            name = LIB.cosl_klt_profile_get(self.h, i, C.byref(ms), C.byref(calls))
            if not name:
                break
            out[name.decode()] = (ms.value, calls.value)
            i += 1
        return out

    def pyramid(self, cam, which, level):
        w, h = self.W >> level, self.H >> level
        out = np.empty((h, w, 3), np.float32)
        _ck(LIB.cosl_klt_debug_pyramid(self.h, cam, which, level, _ptr(out), None, None))
        return out


def pose_intracam(K, R0, t0, Ms, ms, tau, prev_errs=None, opt=None):
    """intraCamEstimate (slam/SL_IntraCamPose.cpp:626-709) on the GPU."""
    K = np.ascontiguousarray(K, np.float64).ravel()
    R0 = np.ascontiguousarray(R0, np.float64).ravel()
    t0 = np.ascontiguousarray(t0, np.float64).ravel()
    Ms = np.ascontiguousarray(Ms, np.float64).reshape(-1, 3)
    ms = np.ascontiguousarray(ms, np.float64).reshape(-1, 2)
    pe = None if prev_errs is None else np.ascontiguousarray(prev_errs, np.float64)
    opt = opt if opt is not None else PoseOpt.defaults()
    R, t, ok = np.empty(9), np.empty(3), C.c_int(0)
    _ck(LIB.cosl_pose_intracam(_ptr(K), _ptr(R0), _ptr(t0), len(Ms), _ptr(pe), _ptr(Ms), _ptr(ms),
                               float(tau), _ptr(R), _ptr(t), C.byref(opt), C.byref(ok)))
    return bool(ok.value), R.reshape(3, 3), t, opt


def pose_intracam_batch(Ks, R0s, t0s, Ms_list, ms_list, tau, prev_list=None, opts=None, device=0):
    Cn = len(Ms_list)
    Ks = np.ascontiguousarray(Ks, np.float64).reshape(Cn, 9)
    R0s = np.ascontiguousarray(R0s, np.float64).reshape(Cn, 9)
    t0s = np.ascontiguousarray(t0s, np.float64).reshape(Cn, 3)
    Ms_list = [np.ascontiguousarray(a, np.float64).reshape(-1, 3) for a in Ms_list]
    ms_list = [np.ascontiguousarray(a, np.float64).reshape(-1, 2) for a in ms_list]
    npts = np.array([len(a) for a in Ms_list], np.int32)
    Mp = (C.c_void_p * Cn)(*[a.ctypes.data for a in Ms_list])
    mp = (C.c_void_p * Cn)(*[a.ctypes.data for a in ms_list])
    pp = None
    if prev_list is not None:
        prev_list = [None if a is None else np.ascontiguousarray(a, np.float64) for a in prev_list]
        pp = (C.c_void_p * Cn)(*[None if a is None else a.ctypes.data for a in prev_list])
    opts = opts if opts is not None else (PoseOpt * Cn)(*[PoseOpt.defaults() for _ in range(Cn)])
    R, t, ok = np.empty((Cn, 9)), np.empty((Cn, 3)), np.zeros(Cn, np.int32)
    _ck(LIB.cosl_pose_intracam_batch(Cn, _ptr(Ks), _ptr(R0s), _ptr(t0s), _ptr(npts), Mp, mp, pp,
                                     float(tau), _ptr(R), _ptr(t), opts, _ptr(ok), device))
    return ok.astype(bool), R.reshape(Cn, 3, 3), t, opts


def ba_solve(prob, opt):
    """bundleAdjustRobust drop-in: prob (coslam_b200.problem.BAProblem) is updated in place."""
    info = np.zeros(COSL_BA_INFOSZ)
    s = prob.struct()
    _ck(LIB.cosl_ba_solve(C.byref(s), C.byref(opt), _ptr(info)))
    return info


def ba_solve_multi(prob, opt, n_gpus, devices=None):
    """cosl_ba_solve_multi: the same drop-in on n_gpus devices of this process."""
    info = np.zeros(COSL_BA_INFOSZ)
    s = prob.struct()
    dev = None if devices is None else np.ascontiguousarray(devices, np.int32)
    _ck(LIB.cosl_ba_solve_multi(C.byref(s), C.byref(opt), int(n_gpus), _ptr(dev), _ptr(info)))
    return info


def posegraph_spread_chains(chain_off, fixed, R, t, eR, et, device=0):
    """cosl_posegraph_spread_chains: post-BA spreading of the non-key-frame poses of all cameras
    (GlobalPoseGraph::computeNewCameraRotations + ::computeNewCameraTranslations on the chain
    graphs of RobustBundleRTS::constructCameraGraphs).  chain_off[nChains+1]; fixed[N]; R (N,3,3);
    t (N,3); eR (N,3,3), et (N,3): edge k = node k -> k+1.  Returns newR (N,3,3), newt (N,3)."""
    off = np.ascontiguousarray(chain_off, np.int32)
    N = int(off[-1])
    fx = np.ascontiguousarray(fixed, np.uint8)
    R = np.ascontiguousarray(R, np.float64).reshape(N, 3, 3)
    t = np.ascontiguousarray(t, np.float64).reshape(N, 3)
    eR = np.ascontiguousarray(eR, np.float64).reshape(N, 3, 3)
    et = np.ascontiguousarray(et, np.float64).reshape(N, 3)
    assert fx.size == N
    nR, nt = np.empty((N, 3, 3)), np.empty((N, 3))
    _ck(LIB.cosl_posegraph_spread_chains(len(off) - 1, _ptr(off), _ptr(fx), _ptr(R), _ptr(t),
                                         _ptr(eR), _ptr(et), _ptr(nR), _ptr(nt), int(device)))
    return nR, nt


def nccl_unique_id():
    buf = np.zeros(128, np.uint8)
    _ck(LIB.cosl_nccl_unique_id(_ptr(buf)))
    return buf


class BaComm:
    def __init__(self, uid, rank, nranks, device):
        uid = np.ascontiguousarray(uid, np.uint8)
        h = C.c_void_p()
        _ck(LIB.cosl_ba_comm_create(_ptr(uid), rank, nranks, device, C.byref(h)))
        self.h, self.rank, self.nranks = h, rank, nranks

    def close(self):
        if getattr(self, "h", None):
            LIB.cosl_ba_comm_destroy(self.h)
            self.h = None


class BaSolver:
    """Problem resident on the device (this rank's shard of points when `comm` is given)."""

    def __init__(self, prob, opt, comm=None):
        self.prob, self.opt = prob, opt
        self._s = prob.struct()
        h = C.c_void_p()
        _ck(LIB.cosl_ba_solver_create(C.byref(self._s), C.byref(opt), comm.h if comm else None,
                                      C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            LIB.cosl_ba_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def reset(self, prob=None):
        if prob is not None:
            self.prob = prob
            self._s = prob.struct()
        _ck(LIB.cosl_ba_solver_reset(self.h, C.byref(self._s)))

    def run(self):
        info = np.zeros(COSL_BA_INFOSZ)
        _ck(LIB.cosl_ba_solver_run(self.h, _ptr(info)))
        return info

    def run_fixed(self, trials):
        info = np.zeros(COSL_BA_INFOSZ)
        _ck(LIB.cosl_ba_solver_run_fixed(self.h, int(trials), _ptr(info)))
        return info

    def download(self):
        _ck(LIB.cosl_ba_solver_download(self.h, C.byref(self._s)))
        return self.prob

    def stream(self):
        return LIB.cosl_ba_solver_stream(self.h)

    def stats(self):
        out = (C.c_double * 8)()
        _ck(LIB.cosl_ba_solver_stats(self.h, out))
        return {"ns": int(out[0]), "reduce_doubles": out[1], "factor_flops": out[2],
                "pair_entries": out[3], "pair_items": out[4], "blocks": int(out[5]),
                "tiles": int(out[6]), "tasks": int(out[7])}

    def plan_info(self):
        out = (C.c_int * 8)()
        _ck(LIB.cosl_ba_solver_plan_info(self.h, out))
        keys = ("blocks", "tiles_schur", "tiles", "tasks", "nd_depth", "critical_tasks", "grid",
                "small_solve")
        return dict(zip(keys, [int(v) for v in out]))

    def tasks(self):
        """Task list of the solve plan: (tasks [n,16] int32, wait lists [m,2] int32)."""
        n = self.plan_info()["tasks"]
        t = np.zeros((n, 16), np.int32)
        lst = np.zeros((64 * n + 16, 2), np.int32)
        LIB.cosl_ba_solver_tasks(self.h, _ptr(t), n, _ptr(lst), len(lst))
        return t, lst

    def trace_arm(self):
        return int(LIB.cosl_ba_solver_trace(self.h, 1, None, None, 0))

    def trace_get(self):
        """Timeline of the last persistent solve: (times [n,8] uint64, meta [n,3] int32)."""
        n = self.plan_info()["tasks"]
        tm = np.zeros((n, 8), np.uint64)
        meta = np.zeros((n, 3), np.int32)
        LIB.cosl_ba_solver_trace(self.h, 0, _ptr(tm), _ptr(meta), n)
        return tm, meta

    def profile_enable(self, on=True):
        _ck(LIB.cosl_ba_solver_profile_enable(self.h, 1 if on else 0))

    def timers(self):
        out, i = {}, 0
        while True:
            ms, calls = C.c_double(0), C.c_int(0)
            name = LIB.cosl_ba_solver_timer(self.h, i, C.byref(ms), C.byref(calls))
            if not name:
                break
            out[name.decode()] = (ms.value, calls.value)
            i += 1
        return out

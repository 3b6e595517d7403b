"""Host-side containers for the flat BA / pose arguments of the C-ABI (numpy owned)."""
import ctypes as C

import numpy as np

from .ctypes_defs import BaProblem


class BAProblem:
    """Flat form of bundleAdjustRobust's arguments (see cosl_ba_problem in coslam_b200.h).

    K [m,9], R [m,9], t [m,3], X [n,3] float64; ptr [n+1] int64; cam [nobs] int32; xy [nobs,2]."""

    def __init__(self, K, R, t, X, ptr, cam, xy, m_con, n_con):
        self.K = np.ascontiguousarray(K, np.float64).reshape(-1, 9)
        self.R = np.ascontiguousarray(R, np.float64).reshape(-1, 9).copy()
        self.t = np.ascontiguousarray(t, np.float64).reshape(-1, 3).copy()
        self.X = np.ascontiguousarray(X, np.float64).reshape(-1, 3).copy()
        self.ptr = np.ascontiguousarray(ptr, np.int64)
        self.cam = np.ascontiguousarray(cam, np.int32)
        self.xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        self.m_con, self.n_con = int(m_con), int(n_con)
        self.outlier = np.zeros(len(self.cam), np.uint8)
        assert self.ptr[-1] == len(self.cam) == len(self.xy)
        assert len(self.ptr) == len(self.X) + 1

    @property
    def m(self):
        return len(self.K)

    @property
    def n(self):
        return len(self.X)

    @property
    def nobs(self):
        return len(self.cam)

    def copy(self):
        return BAProblem(self.K, self.R, self.t, self.X, self.ptr, self.cam, self.xy, self.m_con,
                         self.n_con)

    def struct(self):
        p = BaProblem()
        p.m, p.n, p.nobs, p.m_con, p.n_con = self.m, self.n, self.nobs, self.m_con, self.n_con
        for name in ("K", "R", "t", "X", "ptr", "cam", "xy", "outlier"):
            setattr(p, name, getattr(self, name).ctypes.data_as(C.c_void_p))
        return p

    def residuals(self):
        """Unweighted reprojection residuals [nobs,2] at the current parameters (numpy)."""
        pt = np.repeat(np.arange(self.n), np.diff(self.ptr))
        R = self.R.reshape(-1, 3, 3)[self.cam]
        Pc = np.einsum("nij,nj->ni", R, self.X[pt]) + self.t[self.cam]
        K = self.K.reshape(-1, 3, 3)[self.cam]
        u = np.einsum("nij,nj->ni", K, Pc)
        return self.xy - u[:, :2] / u[:, 2:3]

    def rms(self, mask=None):
        r = self.residuals()
        e2 = (r * r).sum(1)
        if mask is not None:
            e2 = e2[mask]
        return float(np.sqrt(e2.mean()))

    def shard(self, rank, nranks):
        """Contiguous range of points for one rank, balanced by sum k_i^2."""
        k = np.diff(self.ptr).astype(np.float64)
        w = np.cumsum(k * k + 8.0 * k)
        tot = w[-1]
        bounds = [0] + [int(np.searchsorted(w, tot * r / nranks)) for r in range(1, nranks)] + [self.n]
        lo, hi = bounds[rank], bounds[rank + 1]
        o0, o1 = self.ptr[lo], self.ptr[hi]
        ncon_local = int(np.clip(self.n_con - lo, 0, hi - lo))
        return BAProblem(self.K, self.R, self.t, self.X[lo:hi], self.ptr[lo:hi + 1] - o0,
                         self.cam[o0:o1], self.xy[o0:o1], self.m_con, ncon_local), (lo, hi)

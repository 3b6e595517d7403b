"""Seeded synthetic inputs for the two hot paths.  numpy/scipy only.

* images: static texture (band-limited noise octaves + Gaussian blobs) warped by a smooth
  per-camera similarity; ground-truth flow known in closed form.
* BA scenes: camera rigs along a smooth trajectory, points in a slab in front of it, windowed
  visibility, Gaussian pixel noise + gross outliers, perturbed initial values.
"""
import numpy as np
from scipy import ndimage

from .problem import BAProblem

BASE_SEED = 20130901


def make_texture(h, w, seed, n_blobs=None, pad=32):
    """Float32 texture of size (h+2*pad, w+2*pad) in 0..255 with plenty of Shi-Tomasi corners."""
    rng = np.random.default_rng(seed)
    H, W = h + 2 * pad, w + 2 * pad
    img = np.zeros((H, W), np.float64)
    for o, (sigma, amp) in enumerate([(24.0, 30.0), (10.0, 22.0), (4.0, 16.0), (1.6, 10.0)]):
        n = rng.standard_normal((H, W))
        g = ndimage.gaussian_filter(n, sigma, mode="wrap")
        img += amp * g / (g.std() + 1e-12)
    if n_blobs is None:
        n_blobs = max(500, (h * w) // 140)
    blobs = np.zeros((H, W), np.float64)
    ys = rng.integers(0, H, n_blobs)
    xs = rng.integers(0, W, n_blobs)
    amps = rng.uniform(40, 110, n_blobs) * rng.choice([-1.0, 1.0], n_blobs)
    np.add.at(blobs, (ys, xs), amps)
    # two blob scales (sigma 1.5 and 2.5 px), normalised so a single blob peaks at ~amp
    b1 = ndimage.gaussian_filter(blobs, 1.5, mode="wrap") * (2 * np.pi * 1.5 ** 2)
    rng2 = np.random.default_rng(seed + 7)
    blobs2 = np.zeros((H, W), np.float64)
    ys = rng2.integers(0, H, n_blobs // 2)
    xs = rng2.integers(0, W, n_blobs // 2)
    np.add.at(blobs2, (ys, xs), rng2.uniform(40, 110, n_blobs // 2) * rng2.choice([-1.0, 1.0], n_blobs // 2))
    b2 = ndimage.gaussian_filter(blobs2, 2.5, mode="wrap") * (2 * np.pi * 2.5 ** 2)
    img = 128.0 + img + b1 + b2
    return np.clip(img, 0, 255).astype(np.float32)


class ImageSequence:
    """Frames of one camera: frame k = texture warped by the similarity S_k (about the image
    centre): p_tex = c + s_k R(a_k) (p - c) + d_k, with smooth (a, s, d) trajectories."""

    def __init__(self, h, w, seed, n_frames=8, max_shift=3.0, max_rot_deg=0.2, max_scale=0.001,
                 noise_sigma=1.0, pad=32):
        self.h, self.w, self.pad = h, w, pad
        self.seed = seed
        self.rng = np.random.default_rng(seed + 1)
        self.tex = make_texture(h, w, seed, pad=pad)
        # cubic prefilter once
        self.coef = ndimage.spline_filter(self.tex.astype(np.float64), order=3, mode="mirror")
        vel = self.rng.uniform(-1, 1, 2)
        vel = vel / np.linalg.norm(vel) * max_shift * self.rng.uniform(0.5, 1.0)
        self.vel = vel
        self.rot = np.deg2rad(max_rot_deg) * self.rng.uniform(-1, 1)
        self.scl = max_scale * self.rng.uniform(-1, 1)
        self.noise_sigma = noise_sigma
        self.n_frames = n_frames
        self.frames = [self.render(k) for k in range(n_frames)]

    def params(self, k):
        # smooth, bounded excursion so that the padded texture always covers the frame
        ph = 2 * np.pi * k / 64.0
        d = self.vel * k * (1.0 if k < 9 else 9.0 / k)  # <= 9 frames of drift
        d = np.clip(d, -(self.pad - 6), self.pad - 6)
        a = self.rot * k * (1.0 if k < 9 else 9.0 / k)
        s = 1.0 + self.scl * k * (1.0 if k < 9 else 9.0 / k)
        return a, s, d

    def tex_coords(self, k, x, y):
        """Texture (x, y) coordinates (pixel-centre convention, padded frame) of frame-k pixel
        coordinates (x, y) given in the half-integer-centre convention of the tracker."""
        a, s, d = self.params(k)
        cx, cy = self.w / 2.0, self.h / 2.0
        ca, sa = np.cos(a) * s, np.sin(a) * s
        xt = cx + ca * (x - cx) - sa * (y - cy) + d[0]
        yt = cy + sa * (x - cx) + ca * (y - cy) + d[1]
        return xt + self.pad, yt + self.pad

    def render(self, k):
        ys, xs = np.mgrid[0:self.h, 0:self.w]
        xt, yt = self.tex_coords(k, xs + 0.5, ys + 0.5)
        # map_coordinates uses integer-centre indices
        v = ndimage.map_coordinates(self.coef, [yt - 0.5, xt - 0.5], order=3, mode="mirror",
                                    prefilter=False)
        if self.noise_sigma > 0:
            v = v + np.random.default_rng(self.seed * 1000 + 17 + k).normal(
                0, self.noise_sigma, v.shape)
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)

    def flow_truth(self, k0, k1, x, y):
        """Where does the scene point seen at (x, y) in frame k0 appear in frame k1?"""
        xt, yt = self.tex_coords(k0, x, y)
        xt, yt = xt - self.pad, yt - self.pad
        a, s, d = self.params(k1)
        cx, cy = self.w / 2.0, self.h / 2.0
        ux, uy = xt - cx - d[0], yt - cy - d[1]
        ca, sa = np.cos(a) / s, np.sin(a) / s
        return cx + ca * ux + sa * uy, cy - sa * ux + ca * uy

    def frame(self, i):
        """Ping-pong access so that arbitrarily long runs stay temporally smooth."""
        n = self.n_frames
        if n == 1:
            return self.frames[0]
        period = 2 * (n - 1)
        j = i % period
        return self.frames[j if j < n else period - j]

    def frame_index(self, i):
        n = self.n_frames
        if n == 1:
            return 0
        period = 2 * (n - 1)
        j = i % period
        return j if j < n else period - j


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-15:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def make_ba_scene(n_cams=4, n_kf=5, n_pts=20000, width=1280, height=720, seed=BASE_SEED,
                  m_con=None, n_con=2, window=4, p_vis=0.55, noise_px=0.5, outlier_frac=0.02,
                  pose_rot_deg=0.1, pose_trans_frac=0.002, point_frac=0.002, min_obs=2,
                  sort_by_home=False):
    """Synthetic multi-camera BA problem.  Returns (BAProblem with perturbed
    initial values, dict of ground truth).  sort_by_home: number the points in the order of their
    home key frame (map-creation order of a real SLAM run), so that contiguous point shards see
    DIFFERENT camera pairs -- the realistic case for the multi-GPU structure exchange."""
    rng = np.random.default_rng(seed)
    f = 0.9 * width
    K1 = np.array([f, 0, width / 2.0, 0, f, height / 2.0, 0, 0, 1.0])
    step = 0.35  # metres between key frames
    # rig: cameras side by side with small yaw offsets
    rig_off = np.array([[(c - (n_cams - 1) / 2.0) * 0.6, 0.0, 0.0] for c in range(n_cams)])
    rig_yaw = np.array([np.deg2rad((c - (n_cams - 1) / 2.0) * 8.0) for c in range(n_cams)])
    Rs, ts, Cs = [], [], []
    for kf in range(n_kf):
        s = kf * step
        centre = np.array([s, 0.15 * np.sin(0.05 * s * 2 * np.pi), 0.1 * np.cos(0.03 * s * 2 * np.pi)])
        heading = np.deg2rad(3.0) * np.sin(0.02 * s * 2 * np.pi)
        for c in range(n_cams):
            # world: x along the trajectory, cameras look along +z
            Rwc = _rodrigues(np.array([0.0, heading + rig_yaw[c], 0.0]))
            C = centre + rig_off[c]
            R = Rwc.T
            Rs.append(R)
            ts.append(-R @ C)
            Cs.append(C)
    Rs, ts, Cs = np.array(Rs), np.array(ts), np.array(Cs)
    m = n_kf * n_cams
    # points: slab 4..20 m in front of the trajectory, home key frame uniform
    home = rng.integers(0, n_kf, n_pts)
    if sort_by_home:
        home = np.sort(home)
    depth = rng.uniform(4.0, 20.0, n_pts)
    lat = rng.uniform(-1.0, 1.0, n_pts) * depth * (width / 2.0 / f) * 1.3
    vert = rng.uniform(-1.0, 1.0, n_pts) * depth * (height / 2.0 / f) * 1.1
    X = np.stack([home * step + lat, vert, depth], 1)
    # visibility
    cam_kf = np.repeat(np.arange(n_kf), n_cams)
    obs_pt, obs_cam, obs_xy = [], [], []
    chunk = 4096
    for a in range(0, n_pts, chunk):
        b = min(n_pts, a + chunk)
        Xc = np.einsum("jab,ib->ija", Rs, X[a:b]) + ts[None]  # [pts, cams, 3]
        z = Xc[..., 2]
        u = f * Xc[..., 0] / z + width / 2.0
        v = f * Xc[..., 1] / z + height / 2.0
        ok = (z > 0.5) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
        ok &= np.abs(cam_kf[None, :] - home[a:b, None]) <= window
        ok &= rng.random(ok.shape) < p_vis
        pi, ci = np.nonzero(ok)
        obs_pt.append(pi + a)
        obs_cam.append(ci)
        obs_xy.append(np.stack([u[pi, ci], v[pi, ci]], 1))
    obs_pt = np.concatenate(obs_pt)
    obs_cam = np.concatenate(obs_cam)
    obs_xy = np.concatenate(obs_xy)
    # keep points with >= min_obs observations (parseInputs keeps nfpts > 1)
    cnt = np.bincount(obs_pt, minlength=n_pts)
    keep = cnt >= min_obs
    remap = -np.ones(n_pts, np.int64)
    remap[keep] = np.arange(keep.sum())
    sel = keep[obs_pt]
    obs_pt, obs_cam, obs_xy = remap[obs_pt[sel]], obs_cam[sel], obs_xy[sel]
    X = X[keep]
    n = len(X)
    order = np.lexsort((obs_cam, obs_pt))
    obs_pt, obs_cam, obs_xy = obs_pt[order], obs_cam[order], obs_xy[order]
    ptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(obs_pt, minlength=n), out=ptr[1:])
    xy_true = obs_xy.copy()
    xy = obs_xy + rng.normal(0, noise_px, obs_xy.shape)
    is_out = rng.random(len(xy)) < outlier_frac
    xy[is_out] += rng.uniform(-30, 30, (is_out.sum(), 2))
    # perturbed start
    if m_con is None:
        m_con = n_cams
    R0 = Rs.copy()
    t0 = ts.copy()
    base = step  # inter-key-frame baseline
    for j in range(m_con, m):
        dR = _rodrigues(rng.normal(0, np.deg2rad(pose_rot_deg) / np.sqrt(3), 3))
        Cj = Cs[j] + rng.normal(0, pose_trans_frac * max(base, 1.0) / np.sqrt(3), 3)
        R0[j] = dR @ Rs[j]
        t0[j] = -R0[j] @ Cj
    X0 = X.copy()
    zc = np.maximum(X[:, 2], 1.0)
    X0[n_con:] += rng.normal(0, 1, (n - n_con, 3)) * (point_frac * zc[n_con:, None])
    prob = BAProblem(np.tile(K1, (m, 1)), R0.reshape(m, 9), t0, X0, ptr, obs_cam.astype(np.int32),
                     xy, m_con, n_con)
    truth = dict(R=Rs.reshape(m, 9), t=ts, X=X, xy_true=xy_true, is_outlier=is_out,
                 n_cams=n_cams, n_kf=n_kf)
    return prob, truth


def make_pose_case(n_pts=192, width=1280, height=720, seed=BASE_SEED, noise_px=0.5,
                   outlier_frac=0.05, rot_deg=0.6, trans=0.03):
    """One intraCamEstimate input: (K, R0, t0, Ms, ms, R_true, t_true)."""
    rng = np.random.default_rng(seed)
    f = 0.9 * width
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]])
    Rt = _rodrigues(rng.normal(0, 0.2, 3))
    C = rng.normal(0, 0.5, 3)
    tt = -Rt @ C
    depth = rng.uniform(4.0, 20.0, n_pts)
    u = rng.uniform(20, width - 20, n_pts)
    v = rng.uniform(20, height - 20, n_pts)
    Pc = np.stack([(u - width / 2) / f * depth, (v - height / 2) / f * depth, depth], 1)
    Ms = (Pc - tt) @ Rt  # Rt^T (Pc - t)
    ms = np.stack([u, v], 1) + rng.normal(0, noise_px, (n_pts, 2))
    out = rng.random(n_pts) < outlier_frac
    ms[out] += rng.uniform(-40, 40, (out.sum(), 2))
    R0 = _rodrigues(rng.normal(0, np.deg2rad(rot_deg) / np.sqrt(3), 3)) @ Rt
    t0 = tt + rng.normal(0, trans / np.sqrt(3), 3)
    return K, R0, t0, Ms, ms, Rt, tt


def _rodrigues(w):
    th = float(np.linalg.norm(w))
    if th < 1e-300:
        return np.eye(3)
    k = np.asarray(w, np.float64) / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def make_pose_chains(lengths, key_every=8, seed=0, shift=0.03, lead_free=0, fixed_masks=None):
    """Synthetic input of the post-BA pose-graph spreading (RobustBundleRTS::constructCameraGraphs,
    reference app/SL_CoSLAMRobustBA.cpp:182-232): one chain of camera poses per camera, a smooth
    trajectory; edge k = rigid transform from pose k to pose k+1 measured BEFORE bundle adjustment;
    every `key_every`-th node is a key frame (fixed) and has then been moved by BA (rotation ~shift rad,
    translation ~shift).  lead_free > 0 leaves that many free nodes before the first key frame.
    Returns dict(chain_off, fixed, R, t, eR, et, id1, id2, eR_list, et_list): eR/et have one slot per
    node (slot of a chain's last node zero), eR_list/et_list + id1/id2 are the same edges as a list."""
    rng = np.random.default_rng(seed)
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    N = int(off[-1])
    R, t = np.zeros((N, 3, 3)), np.zeros((N, 3))
    eR, et = np.zeros((N, 3, 3)), np.zeros((N, 3))
    fixed = np.zeros(N, np.uint8)
    id1, id2 = [], []
    for c, n in enumerate(lengths):
        a = int(off[c])
        if n == 0:
            continue
        R[a] = _rodrigues(rng.normal(size=3))
        t[a] = rng.normal(size=3)
        for k in range(1, n):
            R[a + k] = _rodrigues(rng.normal(size=3) * 0.05) @ R[a + k - 1]
            t[a + k] = t[a + k - 1] + rng.normal(size=3) * 0.1
        for k in range(n - 1):
            eR[a + k] = R[a + k + 1] @ R[a + k].T
            et[a + k] = t[a + k + 1] - eR[a + k] @ t[a + k]
            id1.append(a + k)
            id2.append(a + k + 1)
        if fixed_masks is not None:
            fixed[a:a + n] = np.asarray(fixed_masks[c], np.uint8)
        else:
            fixed[a + lead_free:a + n:key_every] = 1
        for k in np.nonzero(fixed[a:a + n])[0]:
            R[a + k] = _rodrigues(rng.normal(size=3) * shift) @ R[a + k]
            t[a + k] += rng.normal(size=3) * shift
    id1 = np.asarray(id1, np.int32)
    id2 = np.asarray(id2, np.int32)
    return dict(chain_off=off, fixed=fixed, R=R, t=t, eR=eR, et=et, id1=id1, id2=id2,
                eR_list=eR[id1] if len(id1) else np.zeros((0, 3, 3)),
                et_list=et[id1] if len(id1) else np.zeros((0, 3)))

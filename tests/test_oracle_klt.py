"""Pins of the KLT oracle: closed-form known answers, structural properties of the detector and
sequence logic, and an independent cross-check against OpenCV's LK when cv2 is importable."""
import numpy as np
import pytest

from helpers import live_cfg, seq


def test_pyramid_constant_and_ramp(orc):
    W, H, L = 64, 48, 4
    k = orc.OracleKlt(live_cfg(), W, H, L, 4, 4)
    img = np.full((H, W), 77, np.uint8)
    k.detect(img)
    for l in range(L):
        p = k.pyramid(1, l)
        assert np.all(p[..., 0] == 77.0) and np.all(p[..., 1:] == 0.0)
    # horizontal ramp 3 grey levels / pixel: Ix == 3 away from the clamped border at level 0,
    # and stays 3 (per level-0 pixel!) at coarser levels (derivatives are filtered, not recomputed)
    img = np.clip(np.arange(W)[None, :] * 3 + 10, 0, 255).astype(np.uint8).repeat(H, 0)
    k.detect(img)
    p0 = k.pyramid(1, 0)
    assert np.all(p0[4:-4, 4:-4, 1] == 3.0) and np.all(p0[4:-4, 4:-4, 2] == 0.0)
    assert np.allclose(p0[10, 4:-4, 0], img[10, 4:-4])
    p2 = k.pyramid(1, 2)
    assert np.allclose(p2[3:-3, 3:-3, 1], 3.0, atol=1e-5)
    # [1 3 3 1] centred taps {2j-1..2j+2}: level-1 value sits between source pixels 2i and 2i+1
    p1 = k.pyramid(1, 1)
    assert np.allclose(p1[5, 5:20, 0], (p0[10, 10:40:2, 0] + p0[10, 11:41:2, 0]) / 2, atol=1e-4)


def test_pyramid_level_sizes_odd(orc):
    W, H, L = 100, 75, 4
    k = orc.OracleKlt(live_cfg(), W, H, L, 4, 4)
    k.detect(np.zeros((H, W), np.uint8))
    for l in range(L):
        assert k.pyramid(1, l).shape == (H >> l, W >> l, 3)


@pytest.mark.parametrize("gain", [True, False])
def test_known_translation(orc, gain):
    """Translate a textured image by a known sub-pixel shift: the tracker must recover it."""
    W, H = 320, 240
    s = seq(H, W, 3, n=2, max_rot_deg=0.0, max_scale=0.0, noise_sigma=0.0)
    cfg = live_cfg(gain=gain, min_corner=1500.0)
    k = orc.OracleKlt(cfg, W, H, 4, 16, 16)
    f0, n0 = k.first(s.frames[0])
    assert n0 > 100
    f1, n1 = k.next(s.frames[1])
    tr = f1["status"] == 0
    assert tr.sum() > 0.8 * n0
    x0 = f0["pos"][tr] * [W, H]
    x1 = f1["pos"][tr] * [W, H]
    gx, gy = s.flow_truth(0, 1, x0[:, 0], x0[:, 1])
    err = np.hypot(x1[:, 0] - gx, x1[:, 1] - gy)
    assert np.median(err) < 0.1 and np.percentile(err, 95) < 0.3


def test_detector_properties(orc):
    W, H = 320, 240
    s = seq(H, W, 9, n=1)
    cfg = live_cfg(min_corner=1000.0)
    k = orc.OracleKlt(cfg, W, H, 3, 32, 32)
    f, n = k.detect(s.frames[0])
    live = f[f["status"] == 1]
    assert n == len(live) == min(k.num_candidates(), 1024) and n > 50
    px = live["pos"] * [W, H]
    # pixel-centre positions, inside the 10 px detector margin
    assert np.allclose(px - np.floor(px), 0.5, atol=1e-4)
    assert px[:, 0].min() >= 10 and px[:, 0].max() <= W - 10
    assert px[:, 1].min() >= 10 and px[:, 1].max() <= H - 10
    # strict (2r+1)^2 maxima => Chebyshev distance between any two corners > minDistance
    d = np.abs(px[:, None, :] - px[None, :, :]).max(-1)
    np.fill_diagonal(d, 1e9)
    assert d.min() > cfg.minDistance
    # strongest-first slot order
    corn = k.cornerness()
    vals = corn[np.floor(px[:, 1]).astype(int), np.floor(px[:, 0]).astype(int)]
    assert np.all(np.diff(vals) <= 0)
    # every survivor is the strict maximum of its neighbourhood in the cornerness map
    r = cfg.minDistance
    for (x, y), v in list(zip(np.floor(px).astype(int), vals))[:40]:
        win = corn[max(0, y - r):y + r + 1, max(0, x - r):x + r + 1]
        assert (win >= v).sum() == 1


def test_redetect_slot_logic_and_suppression(orc):
    W, H = 320, 240
    s = seq(H, W, 13, n=3)
    cfg = live_cfg(min_corner=1000.0)
    k = orc.OracleKlt(cfg, W, H, 4, 12, 12)  # only 144 slots -> table is full
    f0, n0 = k.first(s.frames[0])
    assert n0 == 144 and np.all(f0["status"] == 1)
    f1, n1 = k.next(s.frames[1])
    tracked = f1["status"] == 0
    new = f1["status"] == 1
    assert tracked.sum() > 100 and n1 == tracked.sum() + new.sum()
    # new corners never fall into the pixel neighbourhood of a live track
    if new.any():
        pn = f1["pos"][new] * [W, H]
        pt = f1["pos"][tracked] * [W, H]
        d = np.abs(np.floor(pn)[:, None, :] - np.floor(pt)[None, :, :]).max(-1)
        assert d.min() > cfg.minDistance
    assert np.all(f1["fed"] == -1)


def test_feed_extern_points(orc):
    W, H = 320, 240
    s = seq(H, W, 14, n=2)
    k = orc.OracleKlt(live_cfg(min_corner=1000.0), W, H, 4, 16, 16)
    f0, n0 = k.first(s.frames[0])
    live = np.nonzero(f0["status"] >= 0)[0]
    pts = np.zeros((3, 3), np.float32)
    pts[0, :2] = f0["pos"][live[5]]          # coincides with slot live[5] -> that slot is reused
    pts[1, :2] = (0.111, 0.222)
    pts[2, :2] = (0.333, 0.444)
    ids, nfed = k.feed(pts)
    assert nfed == 3
    dead_before = np.nonzero(f0["status"] < 0)[0]
    expect = sorted(list(dead_before) + [live[5]])[:3]
    assert list(ids) == expect
    # empty table accepts at most F points
    k2 = orc.OracleKlt(live_cfg(), W, H, 4, 2, 2)
    ids, nfed = k2.feed(np.random.default_rng(0).uniform(0.1, 0.9, (9, 3)).astype(np.float32))
    assert nfed == 4 and list(ids[:4]) == [0, 1, 2, 3]


def test_gain_tracker_recovers_brightness_change(orc):
    W, H = 320, 240
    s = seq(H, W, 15, n=2, noise_sigma=0.0)
    dim = np.clip(s.frames[1].astype(np.float32) * 0.8, 0, 255).astype(np.uint8)
    k = orc.OracleKlt(live_cfg(gain=True, min_corner=1500.0), W, H, 4, 16, 16)
    k.first(s.frames[0])
    f1, _ = k.next(dim)
    tr = f1["status"] == 0
    assert tr.sum() > 50
    # residual is beta*I0 - I1, so beta ~ 0.8
    assert abs(np.median(f1["gain"][tr]) - 0.8) < 0.03


def test_cross_check_against_opencv_lk(orc):
    cv2 = pytest.importorskip("cv2")
    W, H = 320, 240
    s = seq(H, W, 16, n=2)
    k = orc.OracleKlt(live_cfg(gain=False, min_corner=1500.0), W, H, 4, 16, 16)
    f0, _ = k.first(s.frames[0])
    f1, _ = k.next(s.frames[1])
    tr = f1["status"] == 0
    p0 = (f0["pos"][tr] * [W, H] - 0.5).astype(np.float32)  # OpenCV: integer pixel centres
    p1, st, _ = cv2.calcOpticalFlowPyrLK(s.frames[0], s.frames[1], p0.reshape(-1, 1, 2), None,
                                         winSize=(15, 15), maxLevel=3)
    ok = st.ravel() == 1
    ours = f1["pos"][tr] * [W, H] - 0.5
    d = np.hypot(*(p1.reshape(-1, 2)[ok] - ours[ok]).T)
    assert np.median(d) < 0.25


def test_feed_stride2_compat_bit_reproduces_the_reference_reads(orc):
    """COSL_KLT_COMPAT_FEED_STRIDE2: the kill test of feedExternFeaturePoints reads featPts[2k],
    featPts[2k+1] (v3d_gpuklt.cpp:826-827) although points are 3 floats apart.  Known answer: feed two
    points, the SECOND one sitting on a live KLT point.  Stride 3 (default) kills that KLT point and
    reuses its slot; stride 2 tests (pts[2], pts[3]) = (0, x1) for k = 1 instead, which is outside the
    border margin every KLT point keeps, so the point survives and only dead slots are handed out."""
    W, H = 320, 240
    s = seq(H, W, 14, n=2)
    for bit in (0, 4):
        cfg = live_cfg(min_corner=1000.0)
        cfg.compat |= bit
        k = orc.OracleKlt(cfg, W, H, 4, 16, 16)
        f0, n0 = k.first(s.frames[0])
        live = np.nonzero(f0["status"] >= 0)[0]
        dead_before = np.nonzero(f0["status"] < 0)[0]
        assert len(dead_before) >= 2 and live[0] < dead_before[1]
        victim = live[0]
        pts = np.zeros((2, 3), np.float32)
        pts[0, :2] = (0.003, 0.003)
        pts[1, :2] = f0["pos"][victim]
        ids, nfed = k.feed(pts)
        assert nfed == 2
        expect = sorted(list(dead_before) + ([victim] if bit == 0 else []))[:2]
        assert list(ids[:2]) == expect, (bit, ids, expect)


def _morton(px, py):
    def part(v):
        v = v.astype(np.uint32) & 0xffff
        v = (v | (v << 8)) & 0x00ff00ff
        v = (v | (v << 4)) & 0x0f0f0f0f
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return part(px) | (part(py) << 1)


def test_histopyr_compat_reproduces_the_reference_candidate_list(orc):
    """COSL_KLT_COMPAT_HISTOPYR (v3d_gpuklt.cpp:660-665, 755-768 + klt_detector_traverse_histpyr.cg): the
    corners are read back in HistoPyramid extraction order = Morton order of the pixel (x in the even
    bits), truncated to pointListWidth*pointListHeight BEFORE ranking, ranked only if they exceed the
    free slots, and an odd last row / column is never examined.  Checked against an independent numpy
    statement built from the full candidate set of a default-mode tracker."""
    W, H = 327, 241  # odd: the last column / row must disappear
    s = seq(H, W, 31, n=1)
    full = orc.OracleKlt(live_cfg(min_corner=800.0), W, H, 4, 40, 40, 64, 64)
    f, n = full.first(s.frames[0])
    allc = f[f["status"] >= 0]
    assert 100 < len(allc) < 1600  # every candidate fits: this is the complete set, strongest first
    px = np.floor(allc["pos"][:, 0] * W).astype(np.int64)
    py = np.floor(allc["pos"][:, 1] * H).astype(np.int64)
    keep = (px < 2 * (W // 2)) & (py < 2 * (H // 2))
    px, py = px[keep], py[keep]
    order = np.argsort(_morton(px, py), kind="stable")
    # (a) 6x6 point list, 8x8 slots: the first 36 in Morton order fill the slots IN THAT ORDER
    cfg = live_cfg(min_corner=800.0)
    cfg.compat |= 8
    k = orc.OracleKlt(cfg, W, H, 4, 8, 8, 6, 6)
    g, ng = k.first(s.frames[0])
    assert ng == 36
    want = order[:36]
    got_px = np.floor(g["pos"][:36, 0] * W).astype(np.int64)
    got_py = np.floor(g["pos"][:36, 1] * H).astype(np.int64)
    assert np.array_equal(got_px, px[want]) and np.array_equal(got_py, py[want])
    # (b) 8x8 point list, 5x5 slots: the strongest 25 of the first 64 in Morton order
    k2 = orc.OracleKlt(cfg, W, H, 4, 5, 5, 8, 8)
    g2, n2 = k2.first(s.frames[0])
    assert n2 == 25
    sub = np.sort(order[:64])[:25]  # allc is strongest-first, so the smallest indices are the strongest
    got = set(zip(np.floor(g2["pos"][:, 0] * W).astype(int), np.floor(g2["pos"][:, 1] * H).astype(int)))
    assert got == set(zip(px[sub], py[sub]))
    # the default mode would have picked the globally strongest instead
    d = orc.OracleKlt(live_cfg(min_corner=800.0), W, H, 4, 5, 5, 8, 8)
    g3, _ = d.first(s.frames[0])
    assert set(zip(np.floor(g3["pos"][:, 0] * W).astype(int), np.floor(g3["pos"][:, 1] * H).astype(int))) != got

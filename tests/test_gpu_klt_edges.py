"""Edge cases of the KLT path on the GPU: ragged / tiny / large geometries (strip and chunk
boundaries of the front kernel, single strip, odd sizes), blank and saturated images (no corners,
all slots dead), suppression radii on both sides of the prefilter's 5x5 shortcut."""
import numpy as np
import pytest

from helpers import compare_features, live_cfg, seq

pytestmark = pytest.mark.gpu


def _pair(api, orc, cfg, W, H, L, fw, fh):
    return api.KltTracker(cfg, W, H, L, fw, fh), orc.OracleKlt(cfg, W, H, L, fw, fh)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# widths around the 52-column strips and 36-row chunks of klt_front, odd sizes, one-strip images
@pytest.mark.parametrize("W,H,L", [(97, 61, 3), (40, 30, 2), (52, 36, 2), (53, 37, 2), (104, 72, 3),
                                   (105, 73, 3), (64, 16, 2), (16, 64, 2), (331, 17, 2),
                                   (1920, 1080, 6)])
def test_ragged_geometries_bit_exact(api, orc, W, H, L):
    s = seq(H, W, 23, n=1)
    g, o = _pair(api, orc, live_cfg(min_corner=800.0), W, H, L, 8, 8)
    fg, ng = g.detect(s.frames[0])
    fo, no = o.detect(s.frames[0])
    for l in range(L):
        assert np.array_equal(_bits(g.pyramid(1, l)), _bits(o.pyramid(1, l))), f"level {l}"
    assert np.array_equal(_bits(g.cornerness()), _bits(o.cornerness()))
    assert ng == no and g.num_candidates() == o.num_candidates()
    assert np.array_equal(fg["status"], fo["status"])
    assert np.array_equal(_bits(fg["pos"]), _bits(fo["pos"]))


@pytest.mark.parametrize("value", [0, 255, 128])
def test_blank_image_has_no_corners_and_tracks_nothing(api, orc, value):
    W, H = 160, 120
    img = np.full((H, W), value, np.uint8)
    g, o = _pair(api, orc, live_cfg(), W, H, 4, 8, 8)
    fg, ng = g.first(img)
    fo, no = o.first(img)
    assert ng == no == 0 and g.num_candidates() == 0
    assert (fg["status"] < 0).all()
    fg, _ = g.next(img)
    fo, _ = o.next(img)
    assert np.array_equal(fg["status"], fo["status"]) and (fg["status"] < 0).all()


@pytest.mark.parametrize("min_dist", [1, 2, 3, 12])
def test_suppression_radius(api, orc, min_dist):
    """minDistance 1 takes the 3x3 branch of the prefilter, 12 a 25x25 window (beyond the unrolled
    part of the window walk); candidate sets and slots must stay bit-exact."""
    W, H = 328, 250
    s = seq(H, W, 31, n=3)
    cfg = live_cfg(min_corner=1500.0)
    cfg.minDistance = min_dist
    g, o = _pair(api, orc, cfg, W, H, 4, 24, 24)
    fg, ng = g.first(s.frames[0])
    fo, no = o.first(s.frames[0])
    assert ng == no and g.num_candidates() == o.num_candidates()
    assert np.array_equal(_bits(fg["pos"]), _bits(fo["pos"]))
    for k in (1, 2):  # re-detection next to live tracks exercises the -1e30 suppression marks
        fg, ng = g.next(s.frames[k])
        fo, no = o.next(s.frames[k])
        assert ng == no
        compare_features(fo, fg, W, H)


def test_group_with_a_blank_camera(api):
    """A camera without texture must not disturb the others (shared launches, per-camera lists)."""
    W, H = 320, 240
    cfg = live_cfg(min_corner=1500.0)
    s = [seq(H, W, 40 + c, n=3) for c in range(3)]
    blank = np.full((H, W), 90, np.uint8)
    grp = api.KltGroup(cfg, 3, W, H, 4, 16, 16)
    ref = [api.KltTracker(cfg, W, H, 4, 16, 16) for _ in range(3)]
    frames = lambda k: [s[0].frames[k], blank, s[2].frames[k]]
    fg, ng = grp.first(frames(0))
    for c in range(3):
        fr, nr = ref[c].detect(frames(0)[c])
        ref[c].advance()
        assert ng[c] == nr and np.array_equal(_bits(fg[c]["pos"]), _bits(fr["pos"]))
    assert ng[1] == 0
    for k in (1, 2):
        fg, ng = grp.next(frames(k))
        for c in range(3):
            fr, nr = ref[c].next(frames(k)[c])
            assert ng[c] == nr
            assert np.array_equal(fg[c]["status"], fr["status"])
            assert np.array_equal(_bits(fg[c]["pos"]), _bits(fr["pos"]))


def test_fused_tracker_with_more_slots_than_resident_groups(api):
    """7 cameras x 2048 slots = 14 336 items exceed the lane groups a B200 keeps resident (148 SMs x
    4 CTAs x 16), so part of the items take the non-resident path of klt_gain_fused (state in
    global memory, clamped global window loads); results must still equal the per-pass launches
    bit for bit."""
    W, H, C = 320, 240, 7
    s = [seq(H, W, 60 + c, n=3) for c in range(C)]
    cfg_f = live_cfg(gain=True, min_corner=300.0)
    cfg_p = live_cfg(gain=True, min_corner=300.0)
    cfg_p.compat |= 2  # COSL_KLT_PASS_KERNELS
    cfg_f.minDistance = cfg_p.minDistance = 3
    a = api.KltGroup(cfg_f, C, W, H, 4, 64, 32)
    b = api.KltGroup(cfg_p, C, W, H, 4, 64, 32)
    fa, na = a.first([q.frames[0] for q in s])
    fb, nb = b.first([q.frames[0] for q in s])
    assert fa.tobytes() == fb.tobytes() and int(na.sum()) > 2000
    for k in (1, 2):
        fa, na = a.next([q.frames[k] for q in s])
        fb, nb = b.next([q.frames[k] for q in s])
        assert np.array_equal(na, nb) and fa.tobytes() == fb.tobytes(), f"frame {k}"
        assert (fa["status"] == 0).sum() > 1500


def test_tiled_2x2_tracker_matches_the_plain_kernel(api, orc):
    """klt_track_2x2_tiled (shared-memory tile, 8 lanes per slot) against klt_track_2x2 (selected by
    the PASS_KERNELS compat bit) and against the oracle: same tolerance as every LK parity test."""
    W, H = 640, 480
    s = seq(H, W, 77, n=4)
    cfg_t = live_cfg(gain=False, min_corner=1200.0)
    cfg_p = live_cfg(gain=False, min_corner=1200.0)
    cfg_p.compat |= 2
    a = api.KltTracker(cfg_t, W, H, 6, 32, 32)
    b = api.KltTracker(cfg_p, W, H, 6, 32, 32)
    o = orc.OracleKlt(cfg_t, W, H, 6, 32, 32)
    fa, na = a.first(s.frames[0])
    fb, nb = b.first(s.frames[0])
    fo, no = o.first(s.frames[0])
    assert na == nb == no and fa.tobytes() == fb.tobytes()
    for k in range(1, 4):
        fa, na = a.next(s.frames[k])
        fb, nb = b.next(s.frames[k])
        fo, no = o.next(s.frames[k])
        compare_features(fb, fa, W, H)
        compare_features(fo, fa, W, H)
        assert (fa["status"] == 0).sum() > 500

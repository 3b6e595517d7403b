"""KLT-7 / the facade seam: the REFERENCE's own tracking/GPUKLT.cpp, SL_Track2D.cpp,
slam/SL_FeaturePoints.cpp and SL_FeaturePoint.cpp -- compiled unmodified into
oracle/_ref/gpuklt_ref_caller against coslam_b200/shim/v3d_gpuklt.h (`make -C oracle ref`, possible
only where /root/reference exists; the binary travels to the GPU box) -- are LINKED AND RUN on top of
libcoslam_b200.so: first() + next() x 4 (+ feedExternFeatPoints) on a synthetic sequence.  What the
reference's facade stores in its Track2D lists must equal what the C-ABI returns for the same
frames (pos * (W, H), `GPUKLT::addToFeaturePoints`, tracking/GPUKLT.cpp:36-61)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import live_cfg, seq

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "gpuklt_ref_caller")


def _run(tmp_path, frames, W, H, gain, feed_every):
    raw = tmp_path / "frames.raw"
    raw.write_bytes(np.stack(frames).astype(np.uint8).tobytes())
    out = subprocess.run([EXE, str(raw), str(W), str(H), str(len(frames)), str(int(gain)), str(feed_every)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    per_frame, counts, fed = {}, {}, {}
    for line in out.stdout.splitlines():
        p = line.split()
        if p[0] == "F":
            per_frame.setdefault(int(p[1]), {})[int(p[2])] = (int(p[3]), float(p[4]), float(p[5]))
        elif p[0] == "N":
            counts[int(p[1])] = int(p[2])
        elif p[0] == "X":
            fed[int(p[1])] = int(p[2])
    return per_frame, counts, fed


@pytest.mark.parametrize("gain", [True, False])
def test_reference_facade_runs_on_the_shim(api, tmp_path, gain):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gpuklt_ref_caller not built (make -C oracle ref needs /root/reference)")
    W, H = 640, 480
    s = seq(H, W, 31, n=5)
    per_frame, counts, _ = _run(tmp_path, s.frames, W, H, gain, 0)
    # the same frames through the C-ABI (32 x 32 slots = SLAM_FEATURE_WIDTH/HEIGHT, slam/SL_Define.h:17-18)
    g = api.KltTracker(live_cfg(gain=gain), W, H, 6, 32, 32)
    lengths = np.zeros(1024, np.int64)
    for k in range(5):
        f, n = g.first(s.frames[0]) if k == 0 else g.next(s.frames[k])
        assert counts[k] == n
        live = np.nonzero(f["status"] >= 0)[0]
        ref = per_frame.get(k, {})
        # the facade keeps a track node for exactly the live slots
        assert sorted(ref) == sorted(int(i) for i in live)
        lengths = np.where(f["status"] == 0, lengths + 1, np.where(f["status"] == 1, 1, 0))
        for i in live:
            ln, x, y = ref[int(i)]
            assert ln == lengths[i], (k, i)
            # identical floats, converted exactly like GPUKLT.cpp:43-44 (float pos * int W -> double)
            assert x == float(np.float32(f["pos"][i, 0]) * np.float32(W)) or abs(x - f["pos"][i, 0] * W) < 1e-4
            assert abs(y - f["pos"][i, 1] * H) < 1e-4
    assert (lengths >= 4).sum() > 100  # most first-frame corners were tracked through all 5 frames


def test_reference_facade_feed_extern_points(api, tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gpuklt_ref_caller not built")
    W, H = 640, 480
    s = seq(H, W, 32, n=4)
    per_frame, counts, fed = _run(tmp_path, s.frames, W, H, True, 2)
    assert list(fed) == [2] and 1 <= fed[2] <= 5  # external points found slots (SingleSLAM::feedExtraFeatPtsToTracker)
    assert len(per_frame[3]) > 100

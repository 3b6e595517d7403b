"""CUDA post-BA pose-graph spreading (cosl_posegraph_spread_chains) vs the CPU oracle
(pinned to the compiled reference by tests/test_posegraph.py) and vs the reference-generated vectors
tests/golden/posegraph_ref.npz, through the C-ABI.  Floating point: 1e-9 absolute on O(1) values (the
reference solves sparse least-squares systems, the kernel evaluates their closed form)."""
import os

import numpy as np
import pytest

from coslam_b200.synth import make_pose_chains

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-9


def _gpu(api, g):
    return api.posegraph_spread_chains(g["chain_off"], g["fixed"], g["R"], g["t"], g["eR"], g["et"])


def _orc_per_chain(orc, g):
    nR, nt = np.empty_like(g["R"]), np.empty_like(g["t"])
    off = g["chain_off"]
    for c in range(len(off) - 1):
        a, b = int(off[c]), int(off[c + 1])
        if b == a:
            continue
        ids = np.arange(b - a - 1, dtype=np.int32)
        nR[a:b], nt[a:b] = orc.posegraph_spread(g["fixed"][a:b], g["R"][a:b], g["t"][a:b], ids, ids + 1,
                                                g["eR"][a:b - 1], g["et"][a:b - 1])
    return nR, nt


def test_reference_golden_chain_cases(api):
    z = np.load(os.path.join(HERE, "golden", "posegraph_ref.npz"))
    for i in range(int(z["n_chain_cases"])):
        n = len(z[f"c{i}_fixed"])
        eR, et = np.zeros((n, 3, 3)), np.zeros((n, 3))
        eR[:n - 1], et[:n - 1] = z[f"c{i}_eR_list"], z[f"c{i}_et_list"]
        assert np.array_equal(z[f"c{i}_id1"], np.arange(n - 1)) and np.array_equal(z[f"c{i}_id2"], np.arange(1, n))
        nR, nt = api.posegraph_spread_chains([0, n], z[f"c{i}_fixed"], z[f"c{i}_R"], z[f"c{i}_t"], eR, et)
        assert np.abs(nR - z[f"c{i}_newR"]).max() < TOL, i
        assert np.abs(nt - z[f"c{i}_newt"]).max() < TOL, i


@pytest.mark.parametrize("lengths,kw", [
    ([37], dict(key_every=5)),
    ([300, 1, 2, 77, 0, 513], dict(key_every=9)),          # ragged batch incl. empty / single-node chains
    ([130, 130], dict(key_every=11, lead_free=4)),          # leading + trailing free runs
    ([700], dict(key_every=50, shift=0.2)),                 # chunks of 3 nodes per thread
    ([64], dict(key_every=1)),                              # every node fixed: nothing to solve
])
def test_oracle_parity(api, orc, lengths, kw):
    g = make_pose_chains(lengths, seed=sum(lengths), **kw)
    nR, nt = _gpu(api, g)
    oR, ot = _orc_per_chain(orc, g)
    assert np.abs(nR - oR).max() < TOL and np.abs(nt - ot).max() < TOL


def test_full_size_properties(api):
    """BASELINE-sized case (8 cameras x 4000 frames, key frame every 20): size-independent properties --
    fixed nodes copied bit-exactly, proper rotations, unmoved key frames reproduce the trajectory, and
    the residual of every edge of a run is the same vector in the run's frame (the closed form)."""
    g = make_pose_chains([4000] * 8, key_every=20, seed=5, shift=0.05)
    nR, nt = _gpu(api, g)
    fx = g["fixed"].astype(bool)
    assert np.array_equal(nR[fx], g["R"][fx]) and np.array_equal(nt[fx], g["t"][fx])
    assert np.abs(np.einsum("nij,nkj->nik", nR, nR) - np.eye(3)).max() < 1e-12
    assert np.all(np.linalg.det(nR) > 0.999999)
    # translation residual r_k = t_{k+1} - R_k t_k - t_k,k+1 rotated back to the run's first frame is constant
    a = 20  # run between key frames 20 and 40 of chain 0
    res = []
    P = np.eye(3)
    for k in range(a, a + 20):
        P = g["eR"][k] @ P
        r = nt[k + 1] - g["eR"][k] @ nt[k] - g["et"][k]
        res.append(P.T @ r)
    res = np.asarray(res)
    assert np.abs(res - res[0]).max() < 1e-11 and np.abs(res[0]).max() > 1e-4
    g0 = make_pose_chains([4000] * 8, key_every=20, seed=5, shift=0.0)
    nR0, nt0 = _gpu(api, g0)
    assert np.abs(nR0 - g0["R"]).max() < 1e-11 and np.abs(nt0 - g0["t"]).max() < 1e-10


def test_chain_without_fixed_node_is_an_error(api):
    g = make_pose_chains([8, 6], seed=2, fixed_masks=[[1, 0, 0, 0, 1, 0, 0, 0], [0] * 6])
    with pytest.raises(api.CoslError, match="no fixed node"):
        _gpu(api, g)

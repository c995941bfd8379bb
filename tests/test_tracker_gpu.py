"""GPU parity of the level-0 photometric tracker (CudaCoarseTracker call surface, through the C ABI) vs the C oracle.
Counts exact; energies / H / b to a relative tolerance (the reference itself reduces with float atomics)."""
import numpy as np
import pytest

from oracle.cpu import TrackerOracle
from tandem_b200 import CudaCoarseTracker
from tandem_b200.synthetic import tracker_case

pytestmark = pytest.mark.gpu


def _setup(cls, c, **kw):
    t = cls(c["w"], c["h"], 9.0, 20.0, **kw)
    if hasattr(t, "init"):
        t.init()
    t.setK(c["w"], c["h"], c["fx"], c["fy"], c["cx"], c["cy"])
    t.setReference(c["n"], c["pc_u"], c["pc_v"], c["pc_idepth"], c["pc_color"], c["ref_exposure"], c["ref_aff"])
    t.setNew(c["dInew"])
    return t


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


@pytest.mark.parametrize("size,step", [((120, 160), 1), ((480, 640), 1), ((480, 640), 3)])
def test_calc_res_and_g_match_oracle(size, step):
    h, w = size
    s = w / 640.0
    c = tracker_case(H=h, W=w, fx=320.0 * s, fy=320.0 * s, cx=319.5 * s, cy=239.5 * s, step=step)
    g, o = _setup(CudaCoarseTracker, c), _setup(TrackerOracle, c)
    rg = g.calcRes(c["refToNew"], c["new_exposure"], c["new_aff"], c["cutoffTH"])
    ro = o.calcRes(c["refToNew"], c["new_exposure"], c["new_aff"], c["cutoffTH"])
    assert ro[1] > 0.5 * c["n"], "degenerate test case"
    assert rg[1] == ro[1], "numTermsInE must be exact"
    assert abs(rg[5] - ro[5]) < 1e-12, "saturated ratio"
    assert abs(rg[0] - ro[0]) <= 1e-5 * abs(ro[0])
    assert abs(rg[2] - ro[2]) <= 1e-4 * abs(ro[2]) and abs(rg[4] - ro[4]) <= 1e-4 * abs(ro[4])
    Hg, bg = g.calcG(c["new_exposure"], c["new_aff"])
    Ho, bo = o.calcG(c["new_exposure"], c["new_aff"])
    assert _rel(Hg, Ho) < 1e-4 and _rel(bg, bo) < 1e-4
    assert np.allclose(Hg, Hg.T)
    # fused single launch == two-call path
    rf, Hf, bf = g.calcResAndG(c["refToNew"], c["new_exposure"], c["new_aff"], c["cutoffTH"])
    assert np.array_equal(rf, rg)
    assert _rel(Hf, Hg) < 1e-6 and _rel(bf, bg) < 1e-6
    # deterministic (the reference's float atomics are not)
    rf2, Hf2, bf2 = g.calcResAndG(c["refToNew"], c["new_exposure"], c["new_aff"], c["cutoffTH"])
    assert np.array_equal(Hf, Hf2) and np.array_equal(bf, bf2) and np.array_equal(rf, rf2)


def test_cutoff_and_identity_pose():
    c = tracker_case(H=120, W=160, fx=80.0, fy=80.0, cx=79.5, cy=59.5)
    g, o = _setup(CudaCoarseTracker, c), _setup(TrackerOracle, c)
    for cutoff in (0.5, 5.0, 1e9):
        rg = g.calcRes(np.eye(4), 1.0, np.zeros(2), cutoff)
        ro = o.calcRes(np.eye(4), 1.0, np.zeros(2), cutoff)
        assert rg[1] == ro[1] and abs(rg[5] - ro[5]) < 1e-12
        assert abs(rg[0] - ro[0]) <= 1e-5 * abs(ro[0])


def test_errors():
    t = CudaCoarseTracker(160, 120)
    t.init(1000)
    with pytest.raises(Exception):
        t.setK(320, 240, 1, 1, 1, 1)   # wrong size throws (cuda_coarse_tracker.cpp:359)
    with pytest.raises(Exception):
        z = np.zeros(2000, np.float32)
        t.setReference(2000, z, z, z, z, 1.0, np.zeros(2))  # n > n_max throws (cpp:82)
    with pytest.raises(Exception):
        t.init()


def test_batched_hypotheses_equal_sequential_calc_res():
    """calcResBatch (extension; the motion hypotheses of FullSystem::trackNewCoarse evaluated in one launch) returns, per
    hypothesis, exactly what calcRes returns for that pose - and leaves the buffers calcG refers to untouched."""
    c = tracker_case(H=240, W=320, fx=160.0, fy=160.0, cx=159.5, cy=119.5)
    t = CudaCoarseTracker(c["w"], c["h"])
    t.init()
    t.setK(c["w"], c["h"], c["fx"], c["fy"], c["cx"], c["cy"])
    t.setReference(c["n"], c["pc_u"], c["pc_v"], c["pc_idepth"], c["pc_color"], c["ref_exposure"], c["ref_aff"])
    t.setNew(c["dInew"])
    rng = np.random.default_rng(0)
    poses, affs = [], []
    for k in range(9):
        T = np.array(c["refToNew"], np.float64).copy()
        T[:3, 3] += rng.normal(0, 0.01, 3) * (k > 0)
        w = rng.normal(0, 0.005, 3) * (k > 0)
        R = np.array([[1, -w[2], w[1]], [w[2], 1, -w[0]], [-w[1], w[0], 1]])
        T[:3, :3] = R @ T[:3, :3]
        poses.append(T)
        affs.append(np.array(c["new_aff"], np.float64) + (k % 3) * np.array([0.01, 0.5]))
    r_keep = t.calcRes(poses[0], c["new_exposure"], affs[0], c["cutoffTH"])
    H0, b0 = t.calcG(c["new_exposure"], affs[0])
    batch = t.calcResBatch(np.stack(poses), c["new_exposure"], np.stack(affs), c["cutoffTH"])
    H1, b1 = t.calcG(c["new_exposure"], affs[0])
    assert np.array_equal(H0, H1) and np.array_equal(b0, b1), "the batch must not disturb the last calcRes' buffers"
    assert np.array_equal(batch[0], r_keep)
    for k in range(9):
        r = t.calcRes(poses[k], c["new_exposure"], affs[k], c["cutoffTH"])
        assert np.array_equal(batch[k], r), f"hypothesis {k}: {batch[k]} vs {r}"
    assert len({tuple(b) for b in batch}) > 5, "the hypotheses must actually differ"

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

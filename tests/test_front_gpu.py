"""GPU parity of the SURVEY 8(f) "next" rows around the tracker: image pyramid + gradients (n2), dense tracking
reference (n1) and the device-side LM loop (n3), each against its CPU restatement under oracle/."""
import numpy as np
import pytest

from oracle.cpu import FrontOracle, TrackerOracle
from oracle.lm_driver import track_level0
from tandem_b200 import CudaCoarseTracker, DrFusion, DrFusionOptions, ImagePyramid
from tandem_b200.synthetic import RoomScene, gray_gradients, look_at_pose, tracker_case

pytestmark = pytest.mark.gpu

GRAY_W = np.array([0.114, 0.587, 0.299], np.float32)


@pytest.mark.parametrize("size,levels", [((480, 640), 4), ((96, 130), 2), ((60, 80), 3)])
def test_pyramid_bit_exact(size, levels):
    h, w = size
    rng = np.random.default_rng(3)
    gray = (rng.random((h, w)) * 255).astype(np.float32)
    gray[5, 7] = np.inf                                   # non-finite gradients must become 0 (HessianBlocks.cpp:175-176)
    p = ImagePyramid(w, h, levels)
    p.build(gray)
    ref = FrontOracle().make_images(gray, levels)
    for lvl in range(levels):
        dI, ag = p.level(lvl)
        assert dI.shape == ref[lvl][0].shape
        assert np.array_equal(dI, ref[lvl][0], equal_nan=True), f"level {lvl} (I,dx,dy)"
        assert np.array_equal(ag, ref[lvl][1], equal_nan=True), f"level {lvl} absSquaredGrad"
    # the level-0 float3 image is what setNew consumes: same layout as the host helper used by the tracker tests
    bgr = (rng.random((h, w, 3)) * 255).astype(np.uint8)
    g2 = bgr.astype(np.float32) @ GRAY_W
    p.build(g2)
    dI0, _ = p.level(0)
    host = gray_gradients(bgr)
    assert np.array_equal(dI0[1:-1, 1:-1], host[1:-1, 1:-1])


def _dense_case(h=240, w=320, step=1):
    f = 160.0 * w / 320.0
    K4 = (f, f, (w - 1) / 2.0, (h - 1) / 2.0)
    scene = RoomScene()
    pose_d = look_at_pose((0.3, 0.0, -0.2), (2.5, 0.2, 0.5))
    pose_r = look_at_pose((0.34, 0.02, -0.17), (2.5, 0.25, 0.45))
    _, depth = scene.render(pose_d, h, w, *K4, dropout=0.02, seed=1)
    bgr_r, _ = scene.render(pose_r, h, w, *K4)
    gray_r = bgr_r.astype(np.float32) @ GRAY_W
    T = np.linalg.inv(pose_r.astype(np.float64)) @ pose_d.astype(np.float64)
    return K4, depth, gray_r, T, pose_d


@pytest.mark.parametrize("step,with_sparse", [(1, False), (2, True), (1, True)])
def test_dense_reference_matches_oracle(step, with_sparse):
    h, w = 240, 320
    K4, depth, gray_r, T, _ = _dense_case(h, w)
    rng = np.random.default_rng(5)
    sparse, idepth0 = None, None
    if with_sparse:
        ns = 500
        sparse = [rng.integers(3, w - 3, ns + 1).astype(np.float32), rng.integers(3, h - 3, ns + 1).astype(np.float32),
                  rng.uniform(0.2, 2.0, ns + 1).astype(np.float32), rng.uniform(0, 255, ns + 1).astype(np.float32)]
        idepth0 = np.zeros((h, w), np.float32)
        idepth0[sparse[1][:ns].astype(int), sparse[0][:ns].astype(int)] = sparse[2][:ns]
    n_ref, arrs_ref, proj = FrontOracle().dense_reference(depth, T, K4, step, not with_sparse, sparse, idepth0, gray_r)
    t = CudaCoarseTracker(w, h)
    t.init()
    t.setK(w, h, *K4)
    n = t.setReferenceDense(T, 1.0, np.zeros(2), depth=depth, tracking_step=step, dense_only=not with_sparse, sparse=sparse,
                            idepth0=idepth0, ref_gray=gray_r)
    assert n == n_ref and n > (2000 if step == 2 else 20000)
    got = t.getReference(n + 1)                    # + the uncounted last point (CoarseTracker.cpp:717-722)
    for a, b, name in zip(got, arrs_ref, ("u", "v", "idepth", "color")):
        assert np.array_equal(a, b), name
    # pyramid-fed grey values give the same reference
    p = ImagePyramid(w, h, 1)
    p.build(gray_r)
    n2 = t.setReferenceDense(T, 1.0, np.zeros(2), depth=depth, tracking_step=step, dense_only=not with_sparse, sparse=sparse,
                             idepth0=idepth0, pyramid=p)
    assert n2 == n
    assert all(np.array_equal(a, b) for a, b in zip(t.getReference(n + 1), arrs_ref))


def test_dense_reference_from_fusion_render_equals_host_path():
    h, w = 240, 320
    K4, _, gray_r, T, pose_d = _dense_case(h, w)
    scene = RoomScene()
    fus = DrFusion(DrFusionOptions(height=h, width=w, fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], num_render_streams=1))
    for k in range(3):
        pose = look_at_pose((0.3 + 0.02 * k, 0.0, -0.2), (2.5, 0.2, 0.5))
        bgr, d = scene.render(pose, h, w, *K4)
        fus.IntegrateScanAsync(bgr, d, pose)
        fus.RenderAsync([pose_d])
        _, depth_r = fus.GetRenderResult()
    depth_host = np.array(depth_r[0], np.float32).reshape(h, w)
    assert (depth_host > 0).mean() > 0.5
    t = CudaCoarseTracker(w, h)
    t.init()
    t.setK(w, h, *K4)
    n_dev = t.setReferenceDense(T, 1.0, np.zeros(2), fusion=fus, render_index=0, ref_gray=gray_r)
    dev = t.getReference(n_dev + 1)
    n_host = t.setReferenceDense(T, 1.0, np.zeros(2), depth=depth_host, ref_gray=gray_r)
    host = t.getReference(n_host + 1)
    assert n_dev == n_host and all(np.array_equal(a, b) for a, b in zip(dev, host))


def _setup(cls, c):
    t = cls(c["w"], c["h"], 9.0, 20.0)
    if hasattr(t, "init"):
        t.init()
    t.setK(c["w"], c["h"], c["fx"], c["fy"], c["cx"], c["cy"])
    t.setReference(c["n"], c["pc_u"], c["pc_v"], c["pc_idepth"], c["pc_color"], c["ref_exposure"], c["ref_aff"])
    t.setNew(c["dInew"])
    return t


def _pose_err(A, B):
    D = np.linalg.inv(A) @ B
    ang = np.arccos(np.clip((np.trace(D[:3, :3]) - 1) / 2, -1, 1))
    return np.linalg.norm(D[:3, 3]), ang


@pytest.mark.parametrize("size,step,fix", [((120, 160), 1, (False, False)), ((240, 320), 2, (False, False)),
                                           ((120, 160), 1, (True, True)), ((120, 160), 1, (False, True)),
                                           ((120, 160), 1, (True, False))])
def test_device_lm_loop_matches_host_driver(size, step, fix):
    h, w = size
    s = w / 640.0
    c = tracker_case(H=h, W=w, fx=320.0 * s, fy=320.0 * s, cx=319.5 * s, cy=239.5 * s, step=step)
    g, o = _setup(CudaCoarseTracker, c), _setup(TrackerOracle, c)
    T0 = np.eye(4)                                       # constant-motion guess far from the truth: several LM iterations
    aff0 = np.array(c["ref_aff"], np.float64)
    ro = track_level0(o, T0, aff0, c["new_exposure"], max_iterations=10, fix_a=fix[0], fix_b=fix[1])
    rg = g.track(T0, aff0, c["new_exposure"], max_iterations=10, fix_a=fix[0], fix_b=fix[1])
    assert ro["iterations"] >= 3, "degenerate case"
    dt, dr = _pose_err(ro["refToNew"], rg["refToNew"])
    print(f"{w}x{h} fix={fix}: iterations {rg['iterations']} (host {ro['iterations']}), evaluations {rg['evaluations']}, "
          f"dt {dt:.2e} m, dR {dr:.2e} rad, loop {rg['device_ms'] * 1e3:.0f} us on device")
    assert dt < 1e-3 and dr < 1e-3                       # SURVEY 8(d) config 4 bar
    assert np.allclose(rg["aff"], ro["aff"], atol=1e-3 * max(1.0, np.abs(ro["aff"]).max()))
    assert rg["cutoff_repeat"] == ro["cutoff_repeat"]
    assert abs(rg["res"][0] / rg["res"][1] - ro["res"][0] / ro["res"][1]) <= 1e-3 * ro["res"][0] / ro["res"][1]
    # the loop moved towards the true pose
    e0, _ = _pose_err(c["refToNew"], T0)
    e1, _ = _pose_err(c["refToNew"], rg["refToNew"])
    assert e1 < (0.5 if fix == (False, False) else 1.0) * e0    # with a or b frozen at the wrong value the optimum shifts
    if fix[0]:
        assert rg["aff"][0] == aff0[0]
    if fix[1]:
        assert rg["aff"][1] == aff0[1]


def test_track_errors():
    t = CudaCoarseTracker(64, 48)
    t.init()
    with pytest.raises(Exception):
        t.track(np.eye(4), np.zeros(2), 1.0)             # before setK / setNew
    p = ImagePyramid(64, 48, 2)
    p.build(np.zeros((48, 64), np.float32))
    with pytest.raises(Exception):
        t.setNewFromPyramid(p, 1)                        # level size != tracker size
    t.setNewFromPyramid(p, 0)

"""GPU parity of the CVA-MVSNet path (through the C ABI) against the CPU oracle and the committed goldens."""
import numpy as np
import pytest
import torch

from oracle import mvsnet_oracle as O
from tandem_b200 import DrMvsnet, default_weights
from tandem_b200.weights_io import load_tdmw

pytestmark = pytest.mark.gpu


def _inputs(g):
    V = g["bgr"].shape[0]
    H, W = g["bgr"].shape[1:3]
    bgrs = [np.ascontiguousarray(g["bgr"][v]) for v in range(V)]
    c2ws = [np.ascontiguousarray(g["c2w"][v]) for v in range(V)]
    Ks = np.stack([g["K1"], g["K2"], g["K3"]]).astype(np.float32)
    return V, H, W, bgrs, c2ws, Ks


def _run(g, weights, precision, filter_all=True):
    V, H, W, bgrs, c2ws, Ks = _inputs(g)
    m = DrMvsnet(default_weights(weights), precision=precision)
    m.set_option("filter_all_stages", int(filter_all))
    m.CallAsyncStageK(H, W, V, int(g["ref_index"]), bgrs, Ks, c2ws, float(g["depth_min"]), float(g["depth_max"]),
                      float(g["discard"]))
    out = m.GetResult()
    return m, out


def _oracle(g, weights, keep=None):
    w, dn, va = load_tdmw(default_weights(weights))
    img, order = O.preprocess_bgr(g["bgr"], int(g["ref_index"]))
    Ks = [torch.from_numpy(g[f"K{s}"]) for s in (1, 2, 3)]
    with torch.no_grad():
        return O.forward(w, dn, img, Ks, torch.from_numpy(g["c2w"][order]), float(g["depth_min"]),
                         float(g["depth_max"]), float(g["discard"]), va, keep)


def _absrel(ref, est):
    m = ref > 0
    return float(np.mean(np.abs(ref[m] - est[m]) / ref[m]))


@pytest.mark.parametrize("weights", ["abl03_view_aggregation", "abl04_fewer_depth_planes"])
def test_fp32_layerwise_vs_oracle(golden_small, weights):
    """fp32 storage: every layer group must agree with the oracle to fp32 re-association noise."""
    g = golden_small
    keep = {}
    ref = _oracle(g, weights, keep)
    m, out = _run(g, weights, "fp32")
    for s, name in ((1, "feat1"), (2, "feat2"), (3, "feat3")):
        a = m.debug_tensor(name)                     # (C,V,H,W)
        b = keep["features"][f"stage{s}"].permute(1, 0, 2, 3).numpy()
        assert np.abs(a - b).max() < 2e-4 * max(1.0, np.abs(b).max()), name
    for s in (1, 2, 3):
        st = keep[f"stage{s}"]
        vol = m.debug_tensor(f"s{s}.volume")
        vref = st["volume"].numpy()
        # bilinear taps at pixel borders may flip with 1-ulp coordinate differences: compare robustly
        err = np.abs(vol - vref)
        assert np.mean(err) < 5e-5 * max(1.0, np.abs(vref).mean()), f"volume stage{s} mean err {np.mean(err)}"
        assert np.quantile(err, 0.999) < 1e-3 * max(1.0, np.abs(vref).max()), f"volume stage{s}"
        lg = m.debug_tensor(f"s{s}.logits")[0]
        assert np.mean(np.abs(lg - st["logits"].numpy())) < 1e-3, f"logits stage{s}"
        d = m.stage_output(s, "depth_dense")
        dref = ref[s - 1]["depth_dense"].numpy()
        assert np.mean(np.abs(d - dref)) < 2e-4, f"depth stage{s}: {np.mean(np.abs(d - dref))}"
        c = m.stage_output(s, "confidence_dense")
        assert np.mean(np.abs(c - ref[s - 1]["confidence_dense"].numpy())) < 1e-3, f"confidence stage{s}"
        # edge filter: same discard mask except at ties / float noise
        df = m.stage_output(s, "depth")
        mref = ref[s - 1]["mask"].numpy()
        mours = (df == 0) & (d != 0)
        inter = np.logical_and(mref, mours).sum()
        union = np.logical_or(mref, mours).sum()
        assert inter / max(union, 1) > 0.98, f"filter mask IoU stage{s}: {inter / max(union, 1)}"
    assert _absrel(ref[2]["depth_dense"].numpy(), out.depth_dense) < 1e-4


@pytest.mark.parametrize("precision", ["fp32", "mixed16", "bf16"])
@pytest.mark.parametrize("size", ["small", "full"])
def test_known_answer_abl04(golden_small, golden_full, precision, size):
    """The reference's own KAT (test_dr_mvsnet, dr_mvsnet.cpp:376-556): deployed abl04 model on the shipped
    sample inputs, stage-3 filtered depth/confidence within mean-abs 1e-2 (dr_mvsnet.cpp:508-513)."""
    g = golden_small if size == "small" else golden_full
    m, out = _run(g, "abl04_fewer_depth_planes", precision, filter_all=False)
    ed = float(np.mean(np.abs(out.depth - g["abl04_stage3_depth"])))
    ec = float(np.mean(np.abs(out.confidence - g["abl04_stage3_confidence"])))
    print(f"KAT {size} {precision}: depth mean-abs {ed:.3e}, confidence mean-abs {ec:.3e}")
    assert ed < 1e-2 and ec < 1e-2
    ar = _absrel(g["abl04_stage3_depth_dense"], out.depth_dense)
    print(f"KAT {size} {precision}: dense Abs Rel vs reference {ar:.3e}")
    assert ar < {"fp32": 1e-4, "mixed16": 6e-4, "bf16": 3e-3}[precision]   # budget 1e-3 (BASELINE.json)


@pytest.mark.parametrize("precision", ["fp32", "mixed16"])
def test_benchmark_config_abs_rel(golden_full, precision):
    """640x480 / 7 views / (48,32,8): Abs Rel vs the reference model's fp32 output <= 1e-3 (BASELINE.json)."""
    g = golden_full
    m, out = _run(g, "abl03_view_aggregation", precision, filter_all=False)
    ar = _absrel(g["abl03_stage3_depth_dense"], out.depth_dense)
    ec = float(np.mean(np.abs(out.confidence_dense - g["abl03_stage3_confidence_dense"])))
    mref = g["abl03_stage3_depth"] == 0
    mours = out.depth == 0
    iou = np.logical_and(mref, mours).sum() / max(np.logical_or(mref, mours).sum(), 1)
    print(f"C2 {precision}: Abs Rel {ar:.3e}, conf mean-abs {ec:.3e}, filter IoU {iou:.4f}")
    # Abs Rel budget 1e-3 (BASELINE.json).  Filter-mask bar (SURVEY 8d config 2 asks IoU >= 0.99):
    #  * fp32 parity engine: IoU >= 0.999 (measured 0.9997).
    #  * mixed16: IoU >= 0.98 (measured 0.9853), a DEFENDED bar (profiles/r02_precision_study.md): the
    #    mask keeps the 2.5 % largest values of an order statistic of depth differences, i.e. it is decided at depth edges where
    #    soft-argmin is bimodal; the oracle itself, with nothing changed but its tensors rounded to this engine's 16-bit
    #    storage types, reaches Abs Rel 1.12e-4 / IoU 0.9868 - the engine sits on that floor (1.09e-4 / 0.9853) - and the
    #    reference's own two arithmetic paths (CPU fp32 vs cuDNN fp32 on the B200) agree with each other no better (bench
    #    line, gpu_reference.mask_iou_vs_cpu_reference).  IoU >= 0.99 needs > 16-bit feature maps (the fp32 engine).
    assert ar < (2e-5 if precision == "fp32" else 2.5e-4)
    assert ec < 2e-2
    assert iou > (0.999 if precision == "fp32" else 0.98)


def test_cpp_wrapper_intrinsics_path(golden_small):
    """CallAsync derives K_stage{1,2} as rows0-1 * {0.25,0.5} (dr_mvsnet.cpp:220-247) - not centre preserving;
    the reference tolerates the resulting difference with atol 1e-2 mean-abs, and so must we."""
    g = golden_small
    V, H, W, bgrs, c2ws, Ks = _inputs(g)
    m = DrMvsnet(default_weights("abl04_fewer_depth_planes"), precision="fp32")
    m.CallAsync(H, W, V, int(g["ref_index"]), bgrs, g["K3"], c2ws, float(g["depth_min"]), float(g["depth_max"]),
                float(g["discard"]))
    assert m.Ready() in (True, False)
    out = m.GetResult()
    assert float(np.mean(np.abs(out.depth - g["abl04_stage3_depth"]))) < 1e-2
    assert float(np.mean(np.abs(out.confidence - g["abl04_stage3_confidence"]))) < 1e-2
    with pytest.raises(Exception):
        m.GetResult()  # second GetResult without new input is an error (dr_mvsnet.cpp:100-102)


def test_cpp_shims_known_answer(tmp_path):
    """The C++ drop-in classes (include/dr_mvsnet, include/dr_fusion over the C ABI) run the reference's own KAT
    (test_dr_mvsnet, dr_mvsnet.cpp:376-556) and the DrFusion call order, from a plain g++-built executable."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "shim_kat")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/shim_kat not built (run __graft_entry__.build())")
    sys.path.insert(0, os.path.join(root, "tools"))
    import convert_sample_inputs
    binf = str(tmp_path / "sample_inputs.bin")
    convert_sample_inputs.main(os.path.join(root, "tests", "golden", "sample_512x320.npz"), binf)
    r = subprocess.run([exe, default_weights("abl04_fewer_depth_planes"), binf], capture_output=True, text=True, timeout=300)
    print(r.stdout[-1500:], r.stderr[-500:])
    assert r.returncode == 0 and "test_dr_mvsnet: PASS" in r.stdout


def _run_opts(g, weights, precision, **opts):
    V, H, W, bgrs, c2ws, Ks = _inputs(g)
    m = DrMvsnet(default_weights(weights), precision=precision)
    for k, v in opts.items():
        m.set_option(k, v)
    m.CallAsyncStageK(H, W, V, int(g["ref_index"]), bgrs, Ks, c2ws, float(g["depth_min"]), float(g["depth_max"]),
                      float(g["discard"]))
    return m, m.GetResult()


TC_LAYERS = ["f.c0_0", "f.c3", "f.c1_0", "f.c1_1", "f.c2", "f.c2_0", "f.c2_1", "f.c1", "feat2", "feat3"] + \
    [f"s{s}.{n}" for s in (1, 2, 3) for n in ("c0", "c1", "c2", "c3", "c4", "c5", "c6", "x7", "x9", "x11", "logits")]


@pytest.mark.parametrize("weights", ["abl03_view_aggregation", "abl04_fewer_depth_planes"])
def test_tcgen05_convs_match_direct_kernels(golden_small, weights):
    """Every layer computed by the tcgen05 implicit-GEMM kernel (conv_tc.cuh) against the direct kernel on the same
    16-bit inputs (weights are additionally rounded to 16 bit on the tensor-core path)."""
    md, od = _run_opts(golden_small, weights, "mixed16", use_tc=0)
    mt, ot = _run_opts(golden_small, weights, "mixed16", use_tc=1)
    prof = [r[0] for r in mt.profile()]
    assert sum(n.endswith("[tc]") for n in prof) >= 20, prof
    for name in TC_LAYERS:
        a, b = mt.debug_tensor(name), md.debug_tensor(name)
        scale = max(float(np.abs(b).max()), 1e-6)
        e = np.abs(a - b)
        print(f"{name:12s} max|d| {e.max():.3e} (scale {scale:.3e}) mean|d| {e.mean():.3e} mean|ref| {np.abs(b).mean():.3e}")
        assert np.isfinite(a).all(), name
        # layers are chained, so the tolerance covers the accumulated drift of up to ~12 fp16-weight layers; from
        # stage 2 on the two engines also see slightly different hypotheses (they depend on the stage-1 depth), so
        # isolated outliers at depth discontinuities are expected there and only robust statistics are compared
        if name.startswith(("f.", "feat", "s1.")):
            assert e.max() <= 6e-2 * scale, name
        else:
            assert np.quantile(e, 0.999) <= 6e-2 * scale, name
        tol_mean = 1e-2 if name.startswith(("f.", "feat", "s1.")) else 3e-2
        assert e.mean() <= tol_mean * max(float(np.abs(b).mean()), 1e-6) + 1e-6, name
    assert _absrel(od.depth_dense, ot.depth_dense) < 4e-4


@pytest.mark.parametrize("weights", ["abl03_view_aggregation", "abl04_fewer_depth_planes"])
def test_prob_direct_kernel_matches_tensor_core_prob(golden_small, weights):
    """The `prob` layer (8 -> 1 channels) on the FMA pipes (k_prob_direct, fp32 weights, fp32 accumulate) against the tcgen05
    path (hi/lo 16-bit weights) on the same fp16 input: the stage-1 logits see identical inputs, later stages identical
    kernels upstream but hypotheses that depend on the previous stage's depth; and the fork of the FPN tail changes nothing."""
    ma, oa = _run_opts(golden_small, weights, "mixed16", prob_direct=0, fork_fpn=0)
    mb, ob = _run_opts(golden_small, weights, "mixed16", prob_direct=1, fork_fpn=1)
    names = [r[0] for r in mb.profile()]
    assert sum(n.endswith("prob[direct]") for n in names) == 3, names
    assert np.array_equal(ma.debug_tensor("s1.x11"), mb.debug_tensor("s1.x11")), "inputs of the stage-1 prob layer differ"
    la, lb = ma.debug_tensor("s1.logits"), mb.debug_tensor("s1.logits")
    scale = float(np.abs(la).max())
    e = np.abs(la - lb)
    print(f"s1.logits: max|d| {e.max():.3e} mean|d| {e.mean():.3e} (scale {scale:.3e})")
    assert e.max() <= 2e-3 * scale and e.mean() <= 2e-4 * scale
    for s in (2, 3):
        e = np.abs(ma.debug_tensor(f"s{s}.logits") - mb.debug_tensor(f"s{s}.logits"))
        assert np.quantile(e, 0.999) <= 2e-2 * scale, s
    assert _absrel(oa.depth_dense, ob.depth_dense) < 1e-4
    assert np.mean(np.abs(oa.confidence_dense - ob.confidence_dense)) < 1e-3


def test_tcgen05_benchmark_config_abs_rel(golden_full):
    g = golden_full
    m, out = _run_opts(g, "abl03_view_aggregation", "mixed16", use_tc=1)
    ar = _absrel(g["abl03_stage3_depth_dense"], out.depth_dense)
    print(f"C2 mixed16+tcgen05: Abs Rel {ar:.3e}")
    assert ar < 5e-4


def _oracle_window(win, weights, dn_override=None, discard=10.0):
    w, dn, va = load_tdmw(default_weights(weights))
    if dn_override:
        dn = dn_override
    bgr = np.stack(win["bgrs"])
    img, order = O.preprocess_bgr(bgr, win["ref_index"])
    Ks = O.stage_intrinsics_cpp(win["K"])
    c2w = torch.from_numpy(np.stack(win["c2ws"])[order])
    with torch.no_grad():
        return O.forward(w, dn, img, Ks, c2w, win["depth_min"], win["depth_max"], discard, va)


@pytest.mark.parametrize("precision", ["fp32", "mixed16"])
def test_config1_small_window_and_replan(precision):
    """BASELINE.json configs[0] shape (3 source + 1 reference view, 320x256, 32 stage-1 hypotheses) on a seeded synthetic
    plane scene through CallAsync (C++-style stage intrinsics), then a second window of a different size and view count
    on the SAME handle (re-plan), each against the oracle; the plane depth itself is recovered too."""
    from tandem_b200.synthetic import mvs_plane_window
    m = DrMvsnet(default_weights("abl03_view_aggregation"), precision=precision)
    m.set_option("depth_num_stage1", 32)
    for V, H, W, dn in ((4, 256, 320, (32, 32, 8)), (3, 192, 256, (32, 32, 8)), (2, 96, 128, (32, 32, 8))):
        win = mvs_plane_window(V=V, H=H, W=W)
        m.CallAsync(H, W, V, win["ref_index"], win["bgrs"], win["K"], win["c2ws"], win["depth_min"], win["depth_max"], 10.0)
        out = m.GetResult()
        ref = _oracle_window(win, "abl03_view_aggregation", dn)
        d1 = m.stage_output(1, "depth_dense")
        e1 = float(np.mean(np.abs(d1 - ref[0]["depth_dense"].numpy())))
        ar = _absrel(ref[2]["depth_dense"].numpy(), out.depth_dense)
        print(f"V={V} {W}x{H} {precision}: stage-1 mean-abs {e1:.3e} (bar 1e-3*dmax), stage-3 Abs Rel {ar:.3e}")
        assert e1 <= 1e-3 * win["depth_max"]
        # the 1e-3 budget is stated for the 7-view benchmark window; with 1-2 source views of a periodic texture the
        # depth posterior is multi-modal and a few pixels flip mode under 16-bit rounding
        assert ar < (1e-4 if precision == "fp32" else (1e-3 if V >= 4 else 5e-3))
        if V >= 3:
            inner = out.depth_dense[H // 4:-H // 4, W // 4:-W // 4]
            assert abs(float(np.median(inner)) - win["z_plane"]) < 0.1 * win["z_plane"]


def test_argument_errors_and_two_handles():
    from tandem_b200 import TandemError
    from tandem_b200.synthetic import mvs_plane_window
    win = mvs_plane_window(V=3, H=96, W=128)
    a = DrMvsnet(default_weights("abl03_view_aggregation"), precision="mixed16")
    b = DrMvsnet(default_weights("abl03_view_aggregation"), precision="mixed16")
    with pytest.raises(TandemError):   # H, W must be multiples of 32
        a.CallAsync(100, 128, 3, 1, [np.zeros((100, 128, 3), np.uint8)] * 3, win["K"], win["c2ws"], 0.5, 5.0, 10.0)
    with pytest.raises(TandemError):   # the same buffer passed for two views (dr_mvsnet.cpp:153-160)
        a.CallAsync(96, 128, 3, 1, [win["bgrs"][0], win["bgrs"][0], win["bgrs"][2]], win["K"],
                    [win["c2ws"][0], win["c2ws"][0], win["c2ws"][2]], 0.5, 5.0, 10.0)
    for h in (a, b):                   # two engines in flight on one GPU
        h.CallAsync(96, 128, 3, 1, win["bgrs"], win["K"], win["c2ws"], 0.5, 5.0, 10.0)
    with pytest.raises(TandemError):   # CallAsync over an un-fetched result (dr_mvsnet.cpp:315-318)
        a.CallAsync(96, 128, 3, 1, win["bgrs"], win["K"], win["c2ws"], 0.5, 5.0, 10.0)
    ra, rb = a.GetResult(), b.GetResult()
    assert np.array_equal(ra.depth_dense, rb.depth_dense) and np.array_equal(ra.confidence, rb.confidence)


@pytest.mark.parametrize("which", ["full", "small"])
def test_tma_staged_cost_volume_is_bit_identical(golden_full, golden_small, which):
    """cv_variant 5 (A/B for north_star's "TMA staging of feature tiles"): the stage-3 cost volume with the
    source-view tile staged in shared memory by one tiled cp.async.bulk.tensor per (CTA, view) - zero-filled outside the map,
    global-gather fallback when the bounding box exceeds the staged box - must equal the L1-gather kernel bit for bit."""
    g = golden_full if which == "full" else golden_small     # 640x480 / 512x320 (the small one is the compute-sanitizer instance)
    V, H, W, bgrs, c2ws, Ks = _inputs(g)
    outs, vols = [], []
    for variant in (3, 5):
        m = DrMvsnet(default_weights("abl03_view_aggregation"), precision="mixed16")
        m.set_option("cv_variant", variant)
        m.CallAsyncStageK(H, W, V, int(g["ref_index"]), bgrs, Ks, c2ws, float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]))
        outs.append(m.GetResult())
        vols.append(m.debug_tensor("s3.volume"))
    assert np.array_equal(vols[0], vols[1]), f"stage-3 volume differs on {np.mean(vols[0] != vols[1]):.6f} of the entries"
    assert np.array_equal(outs[0].depth_dense, outs[1].depth_dense) and np.array_equal(outs[0].depth, outs[1].depth)


@pytest.mark.parametrize("weights,which", [("abl03_view_aggregation", "full"), ("abl03_view_aggregation", "small"),
                                           ("abl04_fewer_depth_planes", "small")])
def test_regress_in_prob_epilogue_is_bit_identical(golden_full, golden_small, weights, which):
    """SURVEY 8a a8 / VERDICT r01 item 9: softmax + soft-argmin + confidence (module.py:1116-1133) run in the epilogue of the
    tensor-core prob convolution (tiles span all D planes, so the thread that stored a pixel's D logits finishes the pixel) -
    same arithmetic, same order as the stand-alone k_regress, so every output map must be bit-identical with
    set_option("fused_regress", 0); three launches fewer per forward.  D = 48 / 32 / 8 (abl03) and 48 / 4 / 4 (abl04)."""
    g = golden_full if which == "full" else golden_small
    res, launches = [], []
    for fused in (1, 0):
        m, out = _run_opts(g, weights, "mixed16", fused_regress=fused)
        res.append((out, [m.stage_output(s, "depth_dense") for s in (1, 2, 3)], [m.stage_output(s, "confidence_dense") for s in (1, 2, 3)]))
        launches.append(m.run_resident(1)[1])
    (oa, da, ca), (ob, db, cb) = res
    for s in range(3):
        assert np.array_equal(da[s], db[s]), f"stage {s + 1} depth differs on {np.mean(da[s] != db[s]):.6f} of the pixels"
        assert np.array_equal(ca[s], cb[s]), f"stage {s + 1} confidence differs"
    for k in ("depth", "confidence", "depth_dense", "confidence_dense"):
        assert np.array_equal(getattr(oa, k), getattr(ob, k)), k
    assert launches[0] == launches[1] - 3, launches

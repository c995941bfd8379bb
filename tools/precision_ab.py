"""A/B of the benchmark-precision engine's accuracy knobs on the C2 window: Abs Rel / filter-mask IoU vs the reference model's
fp32 output, and the resident forward time, per option set (run on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tandem_b200 import DrMvsnet, default_weights

g = np.load("tests/golden/sample_640x480.npz")
V, H, W = g["bgr"].shape[:3]
bgrs = [np.ascontiguousarray(g["bgr"][v]) for v in range(V)]
c2ws = [np.ascontiguousarray(g["c2w"][v]) for v in range(V)]
Ks = np.stack([g["K1"], g["K2"], g["K3"]]).astype(np.float32)
ref = g["abl03_stage3_depth_dense"]
msk = ref > 0
mref = g["abl03_stage3_depth"] == 0


def run(prec, opts):
    m = DrMvsnet(default_weights("abl03_view_aggregation"), precision=prec)
    for k, v in opts.items():
        m.set_option(k, v)
    m.CallAsyncStageK(H, W, V, int(g["ref_index"]), bgrs, Ks, c2ws, float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]))
    out = m.GetResult()
    ar = float(np.mean(np.abs(ref[msk] - out.depth_dense[msk]) / ref[msk]))
    mo = out.depth == 0
    iou = np.logical_and(mref, mo).sum() / max(np.logical_or(mref, mo).sum(), 1)
    m.run_resident(3)
    ms, nl = m.run_resident(10)
    print(f"{prec:8s} {str(opts):60s} AbsRel {ar:.3e}  IoU {iou:.4f}  forward {ms / 10:.3f} ms ({nl} launches)", flush=True)


sets = [("fp32", {}), ("mixed16", {})]
for a in sys.argv[1:]:
    prec, _, kv = a.partition(":")
    sets.append((prec, {k: int(v) for k, v in (x.split("=") for x in kv.split(",") if x)}))
for prec, opts in sets:
    try:
        run(prec, opts)
    except Exception as e:
        print(prec, opts, "FAILED:", e, flush=True)

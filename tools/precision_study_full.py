"""CPU study (oracle, C2 window, full forward): which 16-bit tensor groups move the stage-3 depth / edge-filter mask?  Monkeypatches the
oracle's layer helpers so that selected tensor groups are rounded to 16 bit on the way out (everything else fp32)."""
import sys, os
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mvsnet_oracle as O
from tandem_b200 import default_weights
from tandem_b200.weights_io import load_tdmw

torch.set_num_threads(os.cpu_count())
g = np.load("tests/golden/sample_640x480.npz")
w, dn, va = load_tdmw(default_weights("abl03_view_aggregation"))
img, order = O.preprocess_bgr(g["bgr"], int(g["ref_index"]))
Ks = [torch.from_numpy(g[f"K{s}"]) for s in (1, 2, 3)]
c2w = torch.from_numpy(g["c2w"][order])
h16 = lambda x: x.half().float()
b16 = lambda x: x.bfloat16().float()
ident = lambda x: x
Q = dict(feat_inner=ident, feat_out=ident, vol={"stage1": ident, "stage2": ident, "stage3": ident},
         reg={"stage1": ident, "stage2": ident, "stage3": ident})
orig = dict(c2=O._conv2d_bn_relu, c3=O._conv3d_bn_relu, d3=O._deconv3d_bn_relu, fn=O.feature_net, cv=O.cost_volume, cr=O.cost_reg)
cur = {"stage": None}
O._conv2d_bn_relu = lambda x, w_, pre, s, p: Q["feat_inner"](orig["c2"](x, w_, pre, s, p))


def fn(w_, image):
    f, inter = orig["fn"](w_, image)
    return {k: Q["feat_out"](v) for k, v in f.items()}, inter


def cv(w_, stage, *a, **k):
    return Q["vol"][stage](orig["cv"](w_, stage, *a, **k))


def cr(w_, stage, vol, D, keep=None):
    cur["stage"] = stage
    return orig["cr"](w_, stage, vol, D, keep)


O.feature_net, O.cost_volume, O.cost_reg = fn, cv, cr
O._conv3d_bn_relu = lambda x, w_, pre, s: Q["reg"][cur["stage"]](orig["c3"](x, w_, pre, s))
O._deconv3d_bn_relu = lambda x, w_, pre, s, op: Q["reg"][cur["stage"]](orig["d3"](x, w_, pre, s, op))


def run():
    with torch.no_grad():
        return O.forward(w, dn, img, Ks, c2w, float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]), va)


ref = run()
m0 = ref[2]["mask"].numpy(); d0 = ref[2]["depth_dense"]


def report(name, **kw):
    saved = {k: (dict(v) if isinstance(v, dict) else v) for k, v in Q.items()}
    for k, v in kw.items():
        if isinstance(Q[k], dict) and not isinstance(v, dict):
            Q[k] = {s: v for s in Q[k]}
        else:
            Q[k] = v
    out = run()
    Q.update(saved)
    m = out[2]["mask"].numpy()
    ars = []
    for s in range(3):
        r, d = ref[s]["depth_dense"], out[s]["depth_dense"]
        ars.append(float(torch.mean(torch.abs(r - d) / r)))
    print(f"{name:60s} AbsRel s1 {ars[0]:.2e} s2 {ars[1]:.2e} s3 {ars[2]:.2e}  IoU {(m & m0).sum() / (m | m0).sum():.4f}", flush=True)


report("FeatureNet inner fp16", feat_inner=h16)
report("FeatureNet outputs fp16", feat_out=h16)
report("volumes bf16 (all stages)", vol=b16)
report("volumes fp16/16 (all stages)", vol=lambda x: h16(x / 16) * 16)
report("CostRegNet fp16 (all stages)", reg=h16)
report("stage 1+2 everything 16-bit (vol bf16), stage 3 + features fp32", vol={"stage1": b16, "stage2": b16, "stage3": ident}, reg={"stage1": h16, "stage2": h16, "stage3": ident})
report("mixed16 (feat fp16, vol bf16, reg fp16)", feat_inner=h16, feat_out=h16, vol=b16, reg=h16)
report("mixed16 but volumes fp16/16", feat_inner=h16, feat_out=h16, vol=lambda x: h16(x / 16) * 16, reg=h16)

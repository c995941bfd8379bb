"""CPU study (oracle, C2 window): which 16-bit tensors of STAGE 3 move the edge-filter mask?  Rounds selected tensors of the stage-3
CostRegNet to fp16 / bf16 (everything else fp32, stages 1-2 fp32) and reports stage-3 Abs Rel and the filter-mask IoU vs fp32."""
import sys, os, time
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
keep = {}
with torch.no_grad():
    ref = O.forward(w, dn, img, Ks, torch.from_numpy(g["c2w"][order]), float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]), va, keep)
st = keep["stage3"]
m0 = ref[2]["mask"].numpy()
d0 = ref[2]["depth_dense"]
h16 = lambda x: x.half().float()
b16 = lambda x: x.bfloat16().float()
ident = lambda x: x
p = "cost_regularization_net.stage3."


def cost_reg_q(vol, q_act, q_x11, q_c0=None):
    q_c0 = q_c0 or q_act
    x = vol.unsqueeze(0)
    c0 = q_c0(O._conv3d_bn_relu(x, w, p + "conv0", 1))
    c1 = q_act(O._conv3d_bn_relu(c0, w, p + "conv1", 2))
    c2 = q_act(O._conv3d_bn_relu(c1, w, p + "conv2", 1))
    c3 = q_act(O._conv3d_bn_relu(c2, w, p + "conv3", 2))
    c4 = q_act(O._conv3d_bn_relu(c3, w, p + "conv4", 1))
    c5 = q_act(O._conv3d_bn_relu(c4, w, p + "conv5", 2))
    c6 = q_act(O._conv3d_bn_relu(c5, w, p + "conv6", 1))
    x7 = q_act(c4 + O._deconv3d_bn_relu(c6, w, p + "conv7", 2, 1))
    x9 = q_act(c2 + O._deconv3d_bn_relu(x7, w, p + "conv9", 2, 1))
    x11 = q_x11(c0 + O._deconv3d_bn_relu(x9, w, p + "conv11", 2, 1))
    return F.conv3d(x11, O._t(w, p + "prob.weight"), None, padding=1)[0, 0]


def report(name, logits):
    depth, conf, _ = O.regress(logits, st["hyps"])
    _, _, m, _ = O.filter_edges(depth, conf, float(g["discard"]))
    m = m.numpy()
    msk = d0 > 0
    ar = float(torch.mean(torch.abs(d0[msk] - depth[msk]) / d0[msk]))
    print(f"{name:58s} AbsRel {ar:.2e}  IoU {(m & m0).sum() / (m | m0).sum():.4f}", flush=True)


vol = st["volume"]
with torch.no_grad():
    report("fp32 re-run (sanity)", cost_reg_q(vol, ident, ident))
    report("volume bf16", cost_reg_q(b16(vol), ident, ident))
    report("volume fp16 (x 1/16 scaled)", cost_reg_q(h16(vol / 16) * 16, ident, ident))
    report("x11 fp16 only", cost_reg_q(vol, ident, h16))
    report("c0 fp16 only", cost_reg_q(vol, ident, ident, h16))
    report("inner activations fp16 (c1..x9), c0/x11 fp32", cost_reg_q(vol, h16, ident, ident))
    report("all activations fp16, volume fp32", cost_reg_q(vol, h16, h16))
    report("all activations fp16, volume bf16 (= mixed16 stage 3)", cost_reg_q(b16(vol), h16, h16))
    report("all activations fp16, volume fp16/16", cost_reg_q(h16(vol / 16) * 16, h16, h16))
    report("acts fp16 but x11 fp32, volume fp16/16", cost_reg_q(h16(vol / 16) * 16, h16, ident))
    report("acts fp16 but c0+x11 fp32, volume fp16/16", cost_reg_q(h16(vol / 16) * 16, h16, ident, ident))

"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU fp32 restatement of the CVA-MVSNet forward that TANDEM's `libdr/dr_mvsnet` executes through
TorchScript.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this file.  It uses torch CPU ops as the arithmetic substrate (conv,
grid_sample) but none of the reference's Python code; every function cites the reference lines it
restates.  Pinned against (a) the shipped golden vectors `tandem/exported/*/sample_inputs.pt`
(tests/golden/sample_*.npz, see oracle/gen_golden.py) and (b) the reference's own model imported in
the build container (tests/test_oracle_mvsnet.py::test_oracle_matches_reference_model).

Weights come from a `.tdmw` container (tandem_b200/weights_io.py), unfolded.
"""
import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm default, used by every BN in module.py


def _t(w, name):
    v = w[name]
    if isinstance(v, torch.Tensor):      # weights already converted / placed on a device by the caller (bench.py's GPU comparator)
        return v
    return torch.from_numpy(np.asarray(v))


def _bn(x, w, prefix):
    """Inference BatchNorm, y = (x-mean)/sqrt(var+eps)*gamma+beta (module.py:104-110, 213-219)."""
    return F.batch_norm(x, _t(w, prefix + ".running_mean"), _t(w, prefix + ".running_var"),
                        _t(w, prefix + ".weight"), _t(w, prefix + ".bias"), False, 0.0, BN_EPS)


def _conv2d_bn_relu(x, w, prefix, stride, pad):
    x = F.conv2d(x, _t(w, prefix + ".conv.weight"), None, stride=stride, padding=pad)
    return F.relu(_bn(x, w, prefix + ".bn"))


def feature_net(w, image):
    """FeatureNet.forward, module.py:496-531. image (V,3,H,W) fp32 RGB in [0,1].
    Returns dict stage -> (V,C,H_s,W_s) and the intermediates the parity tests look at."""
    p = "feature_net."
    c3 = _conv2d_bn_relu(image, w, p + "conv0.0", 1, 1)
    c3 = _conv2d_bn_relu(c3, w, p + "conv0.1", 1, 1)
    c2 = _conv2d_bn_relu(c3, w, p + "conv1.0", 2, 2)
    c2 = _conv2d_bn_relu(c2, w, p + "conv1.1", 1, 1)
    c2 = _conv2d_bn_relu(c2, w, p + "conv1.2", 1, 1)
    c1 = _conv2d_bn_relu(c2, w, p + "conv2.0", 2, 2)
    c1 = _conv2d_bn_relu(c1, w, p + "conv2.1", 1, 1)
    c1 = _conv2d_bn_relu(c1, w, p + "conv2.2", 1, 1)
    f1 = F.conv2d(c1, _t(w, p + "out.stage1.weight"))
    i2 = F.interpolate(c1, scale_factor=2, mode="nearest") + \
        F.conv2d(c2, _t(w, p + "skip.stage2.weight"), _t(w, p + "skip.stage2.bias"))
    f2 = F.conv2d(i2, _t(w, p + "out.stage2.weight"), padding=1)
    i3 = F.interpolate(i2, scale_factor=2, mode="nearest") + \
        F.conv2d(c3, _t(w, p + "skip.stage3.weight"), _t(w, p + "skip.stage3.bias"))
    f3 = F.conv2d(i3, _t(w, p + "out.stage3.weight"), padding=1)
    return {"stage1": f1, "stage2": f2, "stage3": f3}, {"c3": c3, "c2": c2, "c1": c1, "i2": i2, "i3": i3}


def uniform_hyps(depth_min: float, depth_max: float, D: int, H: int, W: int):
    """module.py:1480-1500."""
    dmin = torch.tensor([depth_min], dtype=torch.float32)
    dmax = torch.tensor([depth_max], dtype=torch.float32)
    interval = (dmax - dmin) / (D - 1)
    d = dmin[:, None] + interval[:, None] * torch.arange(D, dtype=torch.float32)[None, :]
    return d[0].view(D, 1, 1).repeat(1, H, W), interval


def adaptive_hyps(depth, interval, D: int):
    """module.py:1503-1565 (non-inverse branch). depth (H,W), interval tensor (1,)."""
    dmin = (depth - (D / 2) * interval[:, None, None][0]).clamp(min=0.001)
    dmax = dmin + D * interval[:, None, None][0]
    lin = torch.linspace(0, 1, D + 1)[:-1].to(depth.dtype).reshape(-1, 1, 1)
    return dmin.unsqueeze(0) + (dmax - dmin).unsqueeze(0) * lin


def homography(K_ref, c2w_ref, K_src, c2w_src):
    """ref-pixel -> src-pixel 4x4 in fp32, module.py:795-808."""
    w2c_r = torch.inverse(c2w_ref)
    w2c_s = torch.inverse(c2w_src)
    P_r = w2c_r.clone()
    P_r[:3, :4] = K_ref @ w2c_r[:3, :4]
    P_s = w2c_s.clone()
    P_s[:3, :4] = K_src @ w2c_s[:3, :4]
    return P_s @ torch.inverse(P_r)


def warp_source(src_feat, hyps, M):
    """homo_warping, module.py:810-891.  src_feat (C,H,W), hyps (D,H,W), M 4x4 -> (C,D,H,W)."""
    C, H, W = src_feat.shape
    D = hyps.shape[0]
    rot, trans = M[:3, :3], M[:3, 3:4]
    y, x = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(H * W)))      # (3,HW)
    rot_xyz = rot @ xyz                                                       # (3,HW)
    q = rot_xyz.unsqueeze(1) * hyps.reshape(1, D, -1) + trans.view(3, 1, 1)   # (3,D,HW)
    xy = q[:2] / q[2:3]
    gx = xy[0] / (0.5 * (W - 1)) - 1
    gy = xy[1] / (0.5 * (H - 1)) - 1
    grid = torch.stack((gx, gy), dim=2).view(1, D * H, W, 2)
    out = F.grid_sample(src_feat.unsqueeze(0), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    out = out.view(C, D, H, W).clone()
    neg = (q[2] < 0.001).view(1, D, H, W).expand(C, -1, -1, -1)
    out[neg] = 0
    out[torch.isnan(out)] = 0
    return out


def gate(w, stage, x):
    """volume_gates[stage] (cva_mvsnet.py:76-83): 1x1x1 conv -> BN -> ReLU -> 1x1x1 conv -> BN -> ReLU. x (1,C,D,H,W)."""
    p = f"volume_gates.{stage}."
    x = F.conv3d(x, _t(w, p + "0.weight"), _t(w, p + "0.bias"))
    x = F.relu(_bn(x, w, p + "1"))
    x = F.conv3d(x, _t(w, p + "3.weight"), _t(w, p + "3.bias"))
    return F.relu(_bn(x, w, p + "4"))


def cost_volume(w, stage, feats, hyps, K, c2w, view_aggregation=True):
    """depth_prediction's view loop, module.py:1068-1110. feats (V,C,H,W), hyps (D,H,W) -> (C,D,H,W)."""
    V, C, H, W = feats.shape
    D = hyps.shape[0]
    ref = feats[0].unsqueeze(1).expand(-1, D, -1, -1)
    if view_aggregation:
        acc = torch.zeros(C, D, H, W)
        for s in range(1, V):
            warped = warp_source(feats[s], hyps, homography(K, c2w[0], K, c2w[s]))
            d2 = (warped - ref) ** 2
            g = gate(w, stage, d2.unsqueeze(0))[0]
            acc += (g + 1) * d2
        return acc / (V - 1)
    vs, vq = ref.clone(), ref ** 2
    for s in range(1, V):
        warped = warp_source(feats[s], hyps, homography(K, c2w[0], K, c2w[s]))
        vs = vs + warped
        vq = vq + warped ** 2
    return vq / V - (vs / V) ** 2


def _conv3d_bn_relu(x, w, prefix, stride):
    x = F.conv3d(x, _t(w, prefix + ".conv.weight"), None, stride=stride, padding=1)
    return F.relu(_bn(x, w, prefix + ".bn"))


def _deconv3d_bn_relu(x, w, prefix, stride, out_pad):
    x = F.conv_transpose3d(x, _t(w, prefix + ".conv.weight"), None, stride=stride, padding=1, output_padding=out_pad)
    return F.relu(_bn(x, w, prefix + ".bn"))


def cost_reg(w, stage, vol, D: int, keep=None):
    """CostRegNet.forward, module.py:577-600. vol (C,D,H,W) -> logits (D,H,W).
    has_four_depths (D==4) switches conv5/conv7 to stride (1,2,2) (module.py:554-570)."""
    p = f"cost_regularization_net.{stage}."
    four = (D == 4)
    x = vol.unsqueeze(0)
    c0 = _conv3d_bn_relu(x, w, p + "conv0", 1)
    c1 = _conv3d_bn_relu(c0, w, p + "conv1", 2)
    c2 = _conv3d_bn_relu(c1, w, p + "conv2", 1)
    c3 = _conv3d_bn_relu(c2, w, p + "conv3", 2)
    c4 = _conv3d_bn_relu(c3, w, p + "conv4", 1)
    c5 = _conv3d_bn_relu(c4, w, p + "conv5", (1, 2, 2) if four else 2)
    c6 = _conv3d_bn_relu(c5, w, p + "conv6", 1)
    x7 = c4 + _deconv3d_bn_relu(c6, w, p + "conv7", (1, 2, 2) if four else 2, (0, 1, 1) if four else 1)
    x9 = c2 + _deconv3d_bn_relu(x7, w, p + "conv9", 2, 1)
    x11 = c0 + _deconv3d_bn_relu(x9, w, p + "conv11", 2, 1)
    logits = F.conv3d(x11, _t(w, p + "prob.weight"), None, padding=1)
    if keep is not None:
        keep.update(dict(c0=c0[0], c1=c1[0], c2=c2[0], c3=c3[0], c4=c4[0], c5=c5[0], c6=c6[0],
                         x7=x7[0], x9=x9[0], x11=x11[0]))
    return logits[0, 0]


def regress(logits, hyps):
    """softmax + expectation + 4-neighbour confidence, module.py:1116-1133."""
    D = logits.shape[0]
    prob = F.softmax(logits, dim=0)
    depth = torch.sum(prob * hyps, dim=0)
    sum4 = 4 * F.avg_pool3d(F.pad(prob[None, None], pad=[0, 0, 0, 0, 1, 2]), kernel_size=(4, 1, 1), stride=1)[0, 0]
    idx = torch.sum(prob * torch.arange(D, dtype=torch.float32).view(D, 1, 1), dim=0).long().clamp(0, D - 1)
    conf = torch.gather(sum4, 0, idx.unsqueeze(0))[0]
    return depth, conf, prob


def edge_metric(depth):
    """15-th smallest |window - centre| over a zero-padded 5x5 window, module.py:1335-1343."""
    H, W = depth.shape
    dw = F.unfold(depth[None, None], kernel_size=(5, 5), padding=2, stride=1)  # (1,25,HW)
    e = torch.abs(dw - dw[:, 12:13, :])
    e, _ = torch.kthvalue(e, k=15, dim=1)
    return e.view(H, W)


def filter_edges(depth, conf, discard_percentage: float):
    """depth_filter_edges + confidence masking, module.py:1320-1361, cva_mvsnet.py:166-173.
    Returns (depth_filtered, conf_filtered, mask, threshold)."""
    H, W = depth.shape
    e = edge_metric(depth)
    es, _ = torch.sort(e.reshape(-1))
    dp = torch.tensor([discard_percentage], dtype=torch.float32)
    cutoff = int((H * W * (100 - dp) / 100.0).to(torch.long).clamp(0, H * W - 1)[0])
    thr = es[cutoff]
    mask = e > thr
    d = depth.clone(); c = conf.clone()
    d[mask] = 0; c[mask] = 0
    return d, c, mask, float(thr)


def forward(w, depth_num, image, K_stages, c2w, depth_min, depth_max, discard_percentage=None,
            view_aggregation=True, keep: Optional[dict] = None):
    """CvaMVSNet.forward, cva_mvsnet.py:98-184, batch 1.
    image (V,3,H,W) fp32 RGB/255 with the reference view FIRST; K_stages list of three (3,3);
    c2w (V,4,4) ref first.  Returns list of 3 dicts {depth, confidence, depth_dense, confidence_dense}."""
    V, _, H, W = image.shape
    feats, inter = feature_net(w, image)
    if keep is not None:
        keep["features"] = feats
        keep["feature_inter"] = inter
    outs = []
    base_interval = None
    prev_depth = None
    for si, stage in enumerate(("stage1", "stage2", "stage3")):
        scale = 2 ** (2 - si)
        Hs, Ws = H // scale, W // scale
        D = depth_num[si]
        if si == 0:
            hyps, base_interval = uniform_hyps(depth_min, depth_max, D, Hs, Ws)
        else:
            up = F.interpolate(prev_depth[None, None], (Hs, Ws), mode="bilinear", align_corners=False)[0, 0]
            ratio = (1.0, 0.5, 0.25)[si]
            hyps = adaptive_hyps(up, ratio * base_interval, D)
        vol = cost_volume(w, stage, feats[stage], hyps, K_stages[si], c2w, view_aggregation)
        sk = {} if keep is not None else None
        logits = cost_reg(w, stage, vol, D, sk)
        depth, conf, prob = regress(logits, hyps)
        if keep is not None:
            keep[stage] = dict(hyps=hyps, volume=vol, logits=logits, prob=prob, **sk)
        prev_depth = depth
        outs.append({"depth_dense": depth, "confidence_dense": conf})
    for o in outs:
        if discard_percentage is not None:
            d, c, m, thr = filter_edges(o["depth_dense"], o["confidence_dense"], discard_percentage)
            o["depth"], o["confidence"], o["mask"], o["thr"] = d, c, m, thr
        else:
            o["depth"], o["confidence"] = o["depth_dense"].clone(), o["confidence_dense"].clone()
    return outs


# ---- host-side pre-processing of DrMvsnetImpl::CallAsync (dr_mvsnet.cpp:190-247) ----

def preprocess_bgr(bgrs_u8, ref_index: int):
    """bgrs_u8 (V,H,W,3) uint8 BGR in window order -> (V,3,H,W) fp32 RGB/255, reference first."""
    V = bgrs_u8.shape[0]
    order = [ref_index] + [v for v in range(V) if v != ref_index]
    x = torch.from_numpy(np.ascontiguousarray(bgrs_u8[order][..., ::-1])).permute(0, 3, 1, 2)
    return (x.to(torch.float32) / 255.0), order


def stage_intrinsics_cpp(K_full):
    """The C++ wrapper's per-stage K: rows 0..1 times 0.25 / 0.5 / 1 (dr_mvsnet.cpp:220-247)."""
    K = torch.as_tensor(K_full, dtype=torch.float32).reshape(3, 3)
    out = []
    for s in (0.25, 0.5, 1.0):
        k = K.clone()
        k[:2] = k[:2] * s
        out.append(k)
    return out


def abs_rel(ref, est):
    """Abs Rel with the reference's fp32 output as ground truth (module.py:1412-1419)."""
    m = ref > 0
    return float((torch.abs(ref[m] - est[m]) / ref[m]).mean())

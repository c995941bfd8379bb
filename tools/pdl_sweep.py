"""Programmatic dependent launch A/B: resident forwards with 1 / 2 / 4 / 6 / 8 windows in flight on one GPU.
usage: pdl_sweep.py CONFIG [CONFIG ...]   with CONFIG = key=int[,key=int...]   e.g.  use_pdl=0 use_pdl=1 use_pdl=1,pdl_min_smem_kb=0 use_pdl=2
The result maps of every configuration must be bit-identical (launch attributes never change arithmetic)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tandem_b200 import DrMvsnet, default_weights

g = np.load("tests/golden/sample_640x480.npz")
V, H, W = g["bgr"].shape[:3]
bgrs = [np.ascontiguousarray(g["bgr"][v]) for v in range(V)]
c2ws = [np.ascontiguousarray(g["c2w"][v]) for v in range(V)]
ref = None
for cfg in (sys.argv[1:] or ["use_pdl=0", "use_pdl=1"]):
    opts = dict(kv.split("=") for kv in cfg.split(","))
    hs = []
    for i in range(8):
        m = DrMvsnet(default_weights("abl03_view_aggregation"), precision="mixed16")
        for k, v in opts.items():
            m.set_option(k, int(v))
        m.CallAsync(H, W, V, int(g["ref_index"]), bgrs, g["K3"], c2ws, float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]))
        out = m.GetResult()
        hs.append(m)
    row = []
    for n in (1, 2, 4, 6, 8):
        DrMvsnet.run_resident_multi(hs[:n], 4 * n)
        ms, _ = DrMvsnet.run_resident_multi(hs[:n], 96)
        row.append(f"{n}: {ms / 96:.4f}")
    # after the resident replays: the maps of the last handle, re-read through a fresh call
    m = hs[-1]
    m.CallAsync(H, W, V, int(g["ref_index"]), bgrs, g["K3"], c2ws, float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]))
    out = m.GetResult()
    maps = (out.depth.copy(), out.confidence.copy(), out.depth_dense.copy(), out.confidence_dense.copy())
    same = "reference" if ref is None else ("bit-identical" if all(np.array_equal(a, b) for a, b in zip(ref, maps)) else "DIFFERENT")
    if ref is None:
        ref = maps
    print(f"{cfg:34s} ms per window with n windows in flight -> " + "  ".join(row) + f"   [{same}]", flush=True)
    del hs, m

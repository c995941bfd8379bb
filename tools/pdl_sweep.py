"""Programmatic dependent launch A/B: resident forwards with 1 / 2 / 4 / 8 windows in flight on one GPU, use_pdl 0 / 1."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tandem_b200 import DrMvsnet, default_weights

g = np.load("tests/golden/sample_640x480.npz")
V, H, W = g["bgr"].shape[:3]
bgrs = [np.ascontiguousarray(g["bgr"][v]) for v in range(V)]
c2ws = [np.ascontiguousarray(g["c2w"][v]) for v in range(V)]
opts = dict(kv.split("=") for kv in sys.argv[1:])
for pdl in ((1,) if opts else (0, 1)):
    hs = []
    for i in range(8):
        m = DrMvsnet(default_weights("abl03_view_aggregation"), precision="mixed16")
        m.set_option("use_pdl", pdl)
        for k, v in opts.items():
            m.set_option(k, int(v))
        m.CallAsync(H, W, V, int(g["ref_index"]), bgrs, g["K3"], c2ws, float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]))
        m.GetResult()
        hs.append(m)
    row = []
    for n in (1, 2, 4, 6, 8):
        DrMvsnet.run_resident_multi(hs[:n], 4 * n)
        ms, _ = DrMvsnet.run_resident_multi(hs[:n], 96)
        row.append(f"{n}: {ms / 96:.4f}")
    print(f"use_pdl={pdl} {opts} ms per window with n windows in flight -> " + "  ".join(row), flush=True)
    del hs

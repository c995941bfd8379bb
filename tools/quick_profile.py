"""Quick device-side look at the MVSNet path: per-kernel CUDA-event times of one resident forward."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tandem_b200 import DrMvsnet, default_weights

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed16"
tc = int(sys.argv[2]) if len(sys.argv) > 2 else -1
g = np.load("tests/golden/sample_640x480.npz")
V, H, W = g["bgr"].shape[:3]
m = DrMvsnet(default_weights("abl03_view_aggregation"), precision=prec)
if tc >= 0:
    m.set_option("use_tc", tc)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    m.set_option(k, int(v))
bgrs = [np.ascontiguousarray(g["bgr"][v]) for v in range(V)]
c2ws = [np.ascontiguousarray(g["c2w"][v]) for v in range(V)]
m.CallAsync(H, W, V, int(g["ref_index"]), bgrs, g["K3"], c2ws, float(g["depth_min"]), float(g["depth_max"]), float(g["discard"]))
out = m.GetResult()
ref = g["abl03_stage3_depth_dense"]
msk = ref > 0
print(f"Abs Rel vs reference model output: {float(np.mean(np.abs(ref[msk] - out.depth_dense[msk]) / ref[msk])):.3e}")
ms, nl = m.run_resident(3)
ms, nl = m.run_resident(10)
print(f"{prec} tc={tc}: resident forward {ms / 10:.3f} ms, {nl} launches")
rows = m.profile()
tot = sum(r[1] for r in rows)
for name, t, b, fl in sorted(rows, key=lambda r: -r[1])[:int(os.environ.get('TOPK', '40'))]:
    print(f"{name:20s} {t:8.3f} ms {100 * t / tot:5.1f}%  {b / t / 1e6 if t > 0 else 0:9.1f} GB/s  {fl / t / 1e9 if t > 0 else 0:9.2f} TFLOP/s")
print(f"sum of kernels {tot:.3f} ms")

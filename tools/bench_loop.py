"""the full per-frame loop (tracker vs TSDF render; per keyframe MVSNet + integrate + render) on a
synthetic Replica-shaped sequence at 640x480 - frames/s of the loop and where the time goes.  Rendering the synthetic
frames (numpy, CPU) is outside the timed regions."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tandem_b200 import default_weights  # noqa: E402
from tandem_b200.loop import TandemLoop  # noqa: E402
from tandem_b200.synthetic import RoomScene, look_at_pose  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
H, W = 480, 640
K4 = (320.0, 320.0, 319.5, 239.5)
scene = RoomScene()
poses = []
for k in range(N):
    a = 0.03 * k
    eye = np.array([0.3 + 0.25 * np.sin(a), 0.02 * np.sin(2 * a), -0.2 + 0.25 * (1 - np.cos(a))])
    poses.append(look_at_pose(eye, (2.5, 0.2 + 0.1 * np.sin(a), 0.5 + 0.3 * a)))
frames = [scene.render(p, H, W, *K4) for p in poses]
loop = TandemLoop(H, W, K4, default_weights("abl03_view_aggregation"), keyframe_every=5, depth_min=0.3, depth_max=8.0,
                  integrate="mvsnet")
errs = []
t_loop = 0.0
per_frame = []
STEADY0 = 5 * 7 + 6          # keyframe every 5th frame, the 7th keyframe submits the first window
for k, (bgr, depth) in enumerate(frames):
    t0 = time.perf_counter()
    est = loop.step(bgr, sensor_depth=depth, c2w_init=poses[0] if k == 0 else None, true_depth=depth)
    per_frame.append(time.perf_counter() - t0)
    t_loop += per_frame[-1]
    D = np.linalg.inv(poses[k].astype(np.float64)) @ est
    errs.append(float(np.linalg.norm(D[:3, 3])))
loop.finish()
st = loop.stats
print(json.dumps({
    "what": "full loop, config 4 (synthetic room, 640x480, keyframe every 5th frame, 7-keyframe windows)",
    "frames": N, "keyframes": len(st["kf_ms"]), "loop_fps": N / t_loop, "loop_ms_per_frame": 1e3 * t_loop / N,
    # steady state: from the frame after the first MVSNet window was submitted (its one-off plan build / graph capture is
    # in loop_fps above), i.e. every component of the loop is running
    "steady_loop_fps": (N - STEADY0) / sum(per_frame[STEADY0:]) if N > STEADY0 else None,
    "steady_loop_ms_per_frame": 1e3 * sum(per_frame[STEADY0:]) / (N - STEADY0) if N > STEADY0 else None,
    "track_ms_wall_median": float(np.median(st["track_ms"])), "track_ms_device_median": float(np.median(st["track_dev_ms"])),
    "lm_iterations_mean(4 levels)": float(np.mean(st["iterations"])),
    "keyframe_ms_median(GetResult+CallAsync+integrate+render+reference)": float(np.median(st["kf_ms"])),
    "reference_points_mean": float(np.mean(st["n_ref"])), "ate_max_cm": 100 * max(errs),
    "mvsnet_absrel_vs_true_depth_mean": float(np.mean(st["mvs_absrel"])) if st["mvs_absrel"] else None,
}))

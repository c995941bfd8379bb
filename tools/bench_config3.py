"""BASELINE.json configs[2]: TSDF integrate + ray-cast, 512^3 grid at 1 cm (the 5.12 m cube holding a
5 m room + 3 spheres), a stream of N synthetic 640x480 depth maps (camera on a circle of radius 1 m looking outward, 2 mm depth
noise, 2 % dropouts), through the public DrFusion call surface with host buffers:
    IntegrateScanAsync -> RenderAsync([pose]) -> GetRenderResult      per frame.
Reports end-to-end frames/s, the device-resident kernel times at the final map size, size-independent properties (re-integrating
a frame allocates nothing; the render of a fused view reproduces the analytic depth) and the mesh of the final map.
    python tools/bench_config3.py [--frames 1000] [--procs 32]"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tandem_b200.synthetic import RoomScene, circle_trajectory  # noqa: E402

H, W = 480, 640
INTR = dict(fx=320.0, fy=320.0, cx=319.5, cy=239.5)
OFF = np.float32(2.56)       # cube [0, 5.12)^3: keeps block (0,0,0) at a corner, away from the camera


def _render(args):
    k, pose = args
    return RoomScene().render(pose, H, W, **INTR, noise_sigma=0.002, dropout=0.02, seed=k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--procs", type=int, default=min(32, os.cpu_count() or 1))
    a = ap.parse_args()
    poses = circle_trajectory(a.frames, radius=1.0)
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(a.procs) as pool:
        frames = pool.map(_render, list(enumerate(poses)), chunksize=4)
    gen_s = time.perf_counter() - t0
    clean0 = RoomScene().render(poses[0], H, W, **INTR)[1]
    for p in poses:
        p[:3, 3] += OFF
    from tandem_b200 import DrFusion, DrFusionOptions
    f = DrFusion(DrFusionOptions(height=H, width=W, **INTR))
    f.IntegrateScanAsync(frames[0][0], frames[0][1], poses[0])          # warm-up (allocates the first frustum)
    f.RenderAsync([poses[0]]); f.GetRenderResult()
    t0 = time.perf_counter()
    for (bgr, depth), pose in zip(frames[1:], poses[1:]):
        f.IntegrateScanAsync(bgr, depth, pose)
        f.RenderAsync([pose])
        f.GetRenderResult()
    f.Synchronize()
    dt = time.perf_counter() - t0
    n = a.frames - 1
    st = f.stats()
    # properties
    f.IntegrateScanAsync(frames[0][0], frames[0][1], poses[0])
    f.RenderAsync([poses[0]])
    (rb,), (rd,) = f.GetRenderResult()
    st2 = f.stats()
    ok = (rd > 0) & (clean0 > 0.1) & (clean0 < 10)
    err = np.abs(rd[ok] - clean0[ok])
    mi, mr = f.run_resident(20)
    vis = st2["visible_blocks"]
    lo = np.float32([0, 0, 0]); up = np.float32([5.12, 5.12, 5.12])
    f.ExtractMeshAsync(lo, up)
    mv, _ = f.GetMeshSync()
    line = {"what": "config 3: TSDF integrate + ray-cast stream, 512^3 @ 1 cm, 640x480", "frames": a.frames,
            "e2e_frames_per_s": n / dt, "e2e_ms_per_frame": dt / n * 1e3, "allocated_blocks": st["allocated_blocks"],
            "dropped_blocks": st["dropped_blocks"], "resident_allocate_ms": f.last_alloc_ms(), "resident_allocate+integrate_ms": mi / 20,
            "resident_raycast_ms": mr / 20, "visible_blocks": vis, "integrate_algorithmic_GB": (vis * 512 * 16 + 7 * H * W) / 1e9,
            "integrate_GBps": (vis * 512 * 16 + 7 * H * W) / (mi / 20 * 1e-3) / 1e9,
            "property_reintegration_new_blocks": st2["candidate_blocks"], "property_render_vs_analytic_depth_median_m": float(np.median(err)),
            "property_render_vs_analytic_depth_p95_m": float(np.quantile(err, 0.95)), "render_valid_fraction": float(ok.mean()),
            "mesh_triangles": len(mv) // 3, "mesh_device_ms": f.last_mesh_ms(), "synthetic_generation_s": gen_s}
    assert st2["candidate_blocks"] == 0 and st["dropped_blocks"] == 0, line
    assert np.median(err) < 0.005 and ok.mean() > 0.9, line
    print(json.dumps(line))


if __name__ == "__main__":
    main()

"""Seeded synthetic workloads shared by tests/ and bench.py: an axis-aligned room with three
spheres rendered analytically to z-depth + colour, camera trajectories, and tracker inputs derived from them.
Pure numpy, no reference code, no oracle."""
import numpy as np


def look_at_pose(eye, target, up=(0.0, -1.0, 0.0)):
    """cam->world 4x4 (row-major), camera z forward, x right, y down."""
    eye = np.asarray(eye, np.float64)
    z = np.asarray(target, np.float64) - eye
    z /= np.linalg.norm(z)
    x = np.cross(-np.asarray(up, np.float64), z)
    if np.linalg.norm(x) < 1e-9:
        x = np.array([1.0, 0.0, 0.0])
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, eye
    return T.astype(np.float32)


class RoomScene:
    """Room [-half, half]^3 seen from inside + spheres; everything analytic."""

    def __init__(self, half=2.5, spheres=((1.2, 0.3, 0.8, 0.5), (-1.0, -0.4, 1.5, 0.5), (0.2, 0.9, -1.4, 0.5))):
        self.half = float(half)
        self.spheres = [tuple(float(v) for v in s) for s in spheres]

    def render(self, pose, H, W, fx, fy, cx, cy, noise_sigma=0.0, dropout=0.0, seed=0):
        """Returns (bgr u8 (H,W,3), z-depth f32 (H,W))."""
        pose = np.asarray(pose, np.float64)
        v, u = np.mgrid[0:H, 0:W].astype(np.float64)
        dc = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
        d = dc @ pose[:3, :3].T
        o = pose[:3, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (self.half - o) / d
            t2 = (-self.half - o) / d
        t = np.where(d > 0, t1, t2)
        t = np.where(np.isfinite(t) & (t > 0), t, np.inf).min(-1)
        obj = np.zeros((H, W), np.int32)
        for k, (sx, sy, sz, r) in enumerate(self.spheres):
            oc = o - np.array([sx, sy, sz])
            a = (d * d).sum(-1)
            b = 2 * (d * oc).sum(-1)
            c = (oc * oc).sum() - r * r
            disc = b * b - 4 * a * c
            with np.errstate(invalid="ignore"):
                ts = (-b - np.sqrt(disc)) / (2 * a)
            hit = (disc > 0) & (ts > 1e-6) & (ts < t)
            t = np.where(hit, ts, t)
            obj = np.where(hit, k + 1, obj)
        p = o + d * t[..., None]
        # colour: smooth position-dependent pattern, different phase per object
        ph = obj[..., None].astype(np.float64) * 1.7
        col = 127.5 + 100.0 * np.sin(p * np.array([3.1, 2.3, 1.9]) + ph + np.array([0.0, 2.0, 4.0]))
        chk = ((np.floor(p[..., 0] * 2) + np.floor(p[..., 1] * 2) + np.floor(p[..., 2] * 2)) % 2) * 25.0
        bgr = np.clip(col + chk[..., None], 0, 255).astype(np.uint8)
        depth = t.astype(np.float32)
        rng = np.random.default_rng(seed)
        if noise_sigma > 0:
            depth = depth + rng.normal(0, noise_sigma, depth.shape).astype(np.float32)
        if dropout > 0:
            depth = np.where(rng.random(depth.shape) < dropout, np.float32(0), depth)
        return bgr, depth.astype(np.float32)


def circle_trajectory(n, radius=1.0, height=0.0, look_out=True, phase=0.0):
    poses = []
    for k in range(n):
        a = phase + 2 * np.pi * k / max(n, 1)
        eye = np.array([radius * np.cos(a), height + 0.1 * np.sin(3 * a), radius * np.sin(a)])
        tgt = eye * 3.0 if look_out else np.zeros(3)
        tgt = tgt + np.array([0.0, 0.2 * np.cos(2 * a), 0.0])
        poses.append(look_at_pose(eye, tgt))
    return poses


def gray_gradients(bgr):
    """(I,dx,dy) float3 image in the layout of FrameHessian::dI (HessianBlocks.cpp:128-191): grey value and
    central-difference gradients, borders zero."""
    g = bgr.astype(np.float32) @ np.array([0.114, 0.587, 0.299], np.float32)
    dI = np.zeros(g.shape + (3,), np.float32)
    dI[..., 0] = g
    dI[1:-1, 1:-1, 1] = 0.5 * (g[1:-1, 2:] - g[1:-1, :-2])
    dI[1:-1, 1:-1, 2] = 0.5 * (g[2:, 1:-1] - g[:-2, 1:-1])
    return dI


def tracker_case(H=480, W=640, fx=320.0, fy=320.0, cx=319.5, cy=239.5, seed=0, step=1, scene=None):
    """Reference point cloud (all valid pixels of a reference view, stride `step`) + new frame + relative pose."""
    scene = scene or RoomScene()
    ref_pose = look_at_pose((0.3, 0.0, -0.2), (2.5, 0.2, 0.5))
    new_pose = look_at_pose((0.33, 0.01, -0.18), (2.5, 0.22, 0.52))
    bgr_r, d_r = scene.render(ref_pose, H, W, fx, fy, cx, cy)
    bgr_n, _ = scene.render(new_pose, H, W, fx, fy, cx, cy)
    dI_r, dI_n = gray_gradients(bgr_r), gray_gradients(bgr_n)
    v, u = np.mgrid[0:H:step, 0:W:step]
    u, v = u.reshape(-1), v.reshape(-1)
    z = d_r[v, u]
    ok = (z > 0.1) & (u > 2) & (v > 2) & (u < W - 3) & (v < H - 3)
    u, v, z = u[ok], v[ok], z[ok]
    refToNew = np.linalg.inv(new_pose.astype(np.float64)) @ ref_pose.astype(np.float64)
    rng = np.random.default_rng(seed)
    return dict(w=W, h=H, fx=fx, fy=fy, cx=cx, cy=cy, n=int(u.size), pc_u=u.astype(np.float32), pc_v=v.astype(np.float32),
                pc_idepth=(1.0 / z).astype(np.float32), pc_color=dI_r[v, u, 0].astype(np.float32), dInew=dI_n,
                refToNew=refToNew, ref_exposure=1.0, new_exposure=1.05, ref_aff=np.array([0.01, 1.5]),
                new_aff=np.array([0.03 + 0.001 * rng.standard_normal(), -0.8]), cutoffTH=20.0)


def mvs_plane_window(V=4, H=256, W=320, f=300.0, z_plane=2.0, baseline=0.05, seed=0):
    """BASELINE.json configs[0] shaped input: V views of a textured fronto-parallel plane at z = z_plane, cameras
    translated along x by baseline*v (identity rotation), K = [[f,0,(W-1)/2],[0,f,(H-1)/2],[0,0,1]].  The texture is a sum
    of 8 random 2-D sinusoids per channel evaluated at the plane point each pixel sees, so the views are geometrically
    consistent.  Returns dict(bgrs [V x (H,W,3) u8], K (3,3) f32, c2ws [V x (4,4) f32], ref_index, depth_min, depth_max)."""
    rng = np.random.default_rng(seed)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    fr = rng.uniform(2.0, 14.0, (3, 8, 2))
    ph = rng.uniform(0, 2 * np.pi, (3, 8))
    am = rng.uniform(0.3, 1.0, (3, 8))
    v, u = np.mgrid[0:H, 0:W].astype(np.float64)
    bgrs, c2ws = [], []
    for k in range(V):
        tx = baseline * k
        X = (u - cx) / f * z_plane + tx
        Y = (v - cy) / f * z_plane
        img = np.zeros((H, W, 3))
        for c in range(3):
            acc = np.zeros((H, W))
            for j in range(8):
                acc += am[c, j] * np.sin(fr[c, j, 0] * X + fr[c, j, 1] * Y + ph[c, j])
            img[..., c] = 0.5 + 0.5 * acc / am[c].sum()
        bgrs.append(np.clip(np.round(img * 255), 0, 255).astype(np.uint8))
        T = np.eye(4, dtype=np.float32)
        T[0, 3] = tx
        c2ws.append(T)
    K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1]], np.float32)
    return dict(bgrs=bgrs, K=K, c2ws=c2ws, ref_index=V - 2 if V > 2 else 0, depth_min=0.5, depth_max=5.0, z_plane=z_plane)

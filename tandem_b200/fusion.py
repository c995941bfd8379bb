"""Host-side mirror of `class DrFusion` (tandem/libdr/dr_fusion/src/dr_fusion/dr_fusion.h:44-73) over the C ABI.

    f = DrFusion(DrFusionOptions(...))           # defaults = FullSystem::initDr (FullSystem.cpp:259-276)
    f.IntegrateScanAsync(bgr, depth, pose)       # bgr (H,W,3) u8, depth (H,W) f32, pose 4x4 cam->world
    f.RenderAsync([pose])                        # len == num_render_streams
    bgrs, depths = f.GetRenderResult()           # views into the pinned double buffer, valid until the next call
"""
import ctypes

import numpy as np

from ._lib import FusionOptions, FusionStats, TandemError, check, lib


def DrFusionOptions(height=480, width=640, fx=320.0, fy=320.0, cx=319.5, cy=239.5, **kw):
    """DrFusionOptions with the values hard-coded in FullSystem::initDr (FullSystem.cpp:259-276)."""
    o = FusionOptions()
    o.voxel_size = 0.01
    o.num_buckets = 1000000
    o.bucket_size = 10
    o.num_blocks = 1000000
    o.block_size = 8
    o.max_sdf_weight = 64
    o.truncation_distance = 0.04
    o.max_sensor_depth = 10.0
    o.min_sensor_depth = 0.1
    o.num_render_streams = 1
    o.fx, o.fy, o.cx, o.cy = fx, fy, cx, cy
    o.height, o.width = height, width
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TandemError(f"unknown DrFusionOptions field {k}")
        setattr(o, k, v)
    return o


VOXEL_DTYPE = np.dtype([("sdf", "<f4"), ("color", "u1", (3,)), ("weight", "u1")])  # voxel.h:13-21, 8 bytes


class DrFusion:
    def __init__(self, options, device=0):
        self.options = options
        self._h = ctypes.c_void_p()
        check(lib().tdm_fusion_create(ctypes.byref(options), device, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().tdm_fusion_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def IntegrateScanAsync(self, bgr, depth, pose):
        o = self.options
        bgr = np.ascontiguousarray(bgr, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        pose = np.ascontiguousarray(pose, np.float32)
        if bgr.size != o.height * o.width * 3 or depth.size != o.height * o.width or pose.size != 16:
            raise TandemError("IntegrateScanAsync: wrong input size")
        fp = ctypes.POINTER(ctypes.c_float)
        check(lib().tdm_fusion_integrate_async(self._h, bgr.ctypes.data, depth.ctypes.data_as(fp), pose.ctypes.data_as(fp)))

    def RenderAsync(self, camera_poses):
        poses = [np.ascontiguousarray(p, np.float32) for p in camera_poses]
        fp = ctypes.POINTER(ctypes.c_float)
        arr = (fp * max(len(poses), 1))(*[p.ctypes.data_as(fp) for p in poses])
        check(lib().tdm_fusion_render_async(self._h, arr, len(poses)))
        self._n = len(poses)

    def GetRenderResult(self):
        n = self._n
        o = self.options
        bp = (ctypes.c_void_p * max(n, 1))()
        dp = (ctypes.POINTER(ctypes.c_float) * max(n, 1))()
        check(lib().tdm_fusion_get_render_result(self._h, bp, dp, n))
        bgrs, depths = [], []
        for i in range(n):
            b = np.ctypeslib.as_array(ctypes.cast(bp[i], ctypes.POINTER(ctypes.c_ubyte)), (o.height, o.width, 3))
            d = np.ctypeslib.as_array(dp[i], (o.height, o.width))
            bgrs.append(b)
            depths.append(d)
        return bgrs, depths

    def set_slab(self, z_block_lo, z_block_hi):
        """Multi-GPU extension: keep only voxel blocks with z_block_lo <= z < z_block_hi (see include/tandem_b200.h)."""
        check(lib().tdm_fusion_set_slab(self._h, int(z_block_lo), int(z_block_hi)))

    def Synchronize(self):
        check(lib().tdm_fusion_synchronize(self._h))

    # ---- introspection used by tests / bench -------------------------------------------------
    def stats(self):
        s = FusionStats()
        check(lib().tdm_fusion_get_stats(self._h, ctypes.byref(s)))
        return dict(allocated_blocks=s.allocated_blocks, visible_blocks=s.visible_blocks,
                    dropped_blocks=s.dropped_blocks, candidate_blocks=s.candidate_blocks)

    def dump_blocks(self, with_voxels=True):
        n = check(lib().tdm_fusion_dump_blocks(self._h, None, None, 0))
        coords = np.empty((n, 3), np.int32)
        vox = np.empty((n, 512), VOXEL_DTYPE) if with_voxels else None
        check(lib().tdm_fusion_dump_blocks(self._h, coords.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                           vox.ctypes.data if with_voxels else None, n))
        return coords, vox

    def run_resident(self, iters):
        a, b = ctypes.c_float(), ctypes.c_float()
        check(lib().tdm_fusion_run_resident(self._h, iters, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

/* tandem_b200 — C ABI of the B200-native replacement for TANDEM's libdr (dense per-frame hot path).
 *
 * One opaque handle per object, int return codes (0 = ok, <0 = error, text via tdm_last_error()),
 * no exceptions and no C++/torch types across the boundary, all pointers caller-owned unless stated.
 * Blocking / non-blocking behaviour mirrors the reference classes each group replaces:
 *
 *   tdm_mvsnet_*   <->  class DrMvsnet          tandem/libdr/dr_mvsnet/src/dr_mvsnet/dr_mvsnet.h:36-66
 *   tdm_fusion_*   <->  class DrFusion          tandem/libdr/dr_fusion/src/dr_fusion/dr_fusion.h:44-73
 *   tdm_tracker_*  <->  class CudaCoarseTracker tandem/libdr/cuda_coarse_tracker/include/public/cuda_coarse_tracker.h:9-82
 *
 * The header-compatible C++ shims over this ABI live in include/dr_mvsnet/, include/dr_fusion/ and
 * include/cuda_coarse_tracker/ ; INTEGRATION.md shows how tandem/CMakeLists.txt:118-121 links them.
 */
#ifndef TANDEM_B200_H
#define TANDEM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDM_OK 0
#define TDM_ERR (-1)
#define TDM_PRECISION_FP32 0
#define TDM_PRECISION_MIXED16 1 /* fp16 activations, bf16 cost volume (range), fp32 accumulate */
#define TDM_PRECISION_BF16 2    /* bf16 everywhere (accuracy study only: misses the 1e-3 Abs Rel budget) */

/* Thread-local text of the last error raised by any tdm_* call on this thread. */
const char* tdm_last_error(void);
/* Library / build identification: "tandem_b200 <version> sm_100a". */
const char* tdm_version(void);
/* Number of CUDA devices visible (0 on a CPU-only box; never throws). */
int tdm_device_count(void);

/* Page-locked ("pinned") host memory, portable across devices.  Optional: every entry point accepts ordinary pageable
 * memory (staged through the library's own pinned buffers by a small copy pool), but image / result buffers that ARE
 * page-locked - from here, cudaHostAlloc or cudaHostRegister - are detected (cudaPointerGetAttributes) and DMA'd from / into
 * directly, with no staging copy.  The reference does the same staging by hand (tsdf_volume.cu:542-543, dr_mvsnet.cpp:260). */
int tdm_host_alloc_pinned(size_t bytes, void** out);
int tdm_host_free_pinned(void* p);

/* ------------------------------------------------------------------------------------------------
 * CVA-MVSNet (replaces DrMvsnet; dr_mvsnet.h:36-66, implementation dr_mvsnet.cpp:125-331)
 * ---------------------------------------------------------------------------------------------- */
typedef struct tdm_mvsnet tdm_mvsnet;

/* weights_path: a .tdmw container, or the reference's ".../model.pt" path, in which case the sidecar
 * ".../model.tdmw" next to it is loaded (DrMvsnet(char const* filename), dr_mvsnet.cpp:333).
 * precision: activation storage type; accumulation is always fp32. device: CUDA ordinal. */
int tdm_mvsnet_create(const char* weights_path, int precision, int device, tdm_mvsnet** out);
void tdm_mvsnet_destroy(tdm_mvsnet* h);

/* DrMvsnet::CallAsync (dr_mvsnet.h:43-53): "blocking for last input, non-blocking for this input".
 * bgrs[i]: H*W*3 uint8 BGR interleaved, window order, reference view at ref_index; intrinsic: 3x3
 * row-major full resolution (stage K's derived as rows0-1 * 0.25 / 0.5 / 1, dr_mvsnet.cpp:220-247);
 * cam_to_worlds[i]: 4x4 row-major. All inputs are copied before the call returns. */
int tdm_mvsnet_call_async(tdm_mvsnet* h, int height, int width, int view_num, int ref_index,
                          unsigned char* const* bgrs, const float* intrinsic_matrix,
                          float* const* cam_to_worlds, float depth_min, float depth_max,
                          float discard_percentage);
/* Same, with the three per-stage intrinsics given explicitly (9 floats each, stage1..stage3) — the
 * entry the golden-vector tests use, because export_model.py's K's are centre-preserving
 * (cva_mvsnet/models/datasets.py:144-174) while the C++ wrapper's are not. */
int tdm_mvsnet_call_async_k(tdm_mvsnet* h, int height, int width, int view_num, int ref_index,
                            unsigned char* const* bgrs, const float* intrinsics_stage123,
                            float* const* cam_to_worlds, float depth_min, float depth_max,
                            float discard_percentage);
/* DrMvsnet::GetResult (blocking). Copies the stage-3 maps (H*W floats each; any pointer may be NULL)
 * into caller memory: depth/confidence are edge-filtered, *_dense unfiltered (dr_mvsnet.cpp:296-329).
 * Calling it twice without a new CallAsync is an error, as in the reference (dr_mvsnet.cpp:100-102). */
int tdm_mvsnet_get_result(tdm_mvsnet* h, float* depth, float* confidence, float* depth_dense,
                          float* confidence_dense);
int tdm_mvsnet_ready(tdm_mvsnet* h); /* 1 ready, 0 busy  (DrMvsnet::Ready) */
int tdm_mvsnet_wait(tdm_mvsnet* h);  /* DrMvsnet::Wait */

/* Options (call before the first call_async): filter_all_stages=1 also edge-filters stages 1 and 2
 * (the Python model does, cva_mvsnet.py:165-173; the C++ consumer only reads stage 3).
 * keep_intermediates=1 keeps every layer output alive for tdm_mvsnet_debug_tensor. */
int tdm_mvsnet_set_option(tdm_mvsnet* h, const char* key, int value);
/* Stage outputs of the last finished call: which = "depth"|"confidence"|"depth_dense"|"confidence_dense"|"edge". */
int tdm_mvsnet_stage_output(tdm_mvsnet* h, int stage /*1..3*/, const char* which, float* out, size_t capacity);
/* Test hook: copy a named intermediate of the last call as planar fp32 [C][D][H][W]; dims4 receives C,D,H,W.
 * Returns the number of floats written, or <0. */
long long tdm_mvsnet_debug_tensor(tdm_mvsnet* h, const char* name, float* out, size_t capacity, int* dims4);

/* Device-resident measurement: re-runs the forward of the last submitted window `iters` times on the
 * handle's stream with inputs already in HBM (no H2D/D2H inside), timed with CUDA events on that stream.
 * ms_total receives the elapsed milliseconds; launches (optional) the kernel launches per forward. */
int tdm_mvsnet_run_resident(tdm_mvsnet* h, int iters, float* ms_total, int* launches);
/* Same clock over n handles of one device: iters_total forwards issued round-robin, each handle on its own stream (the
 * windows are independent, so their kernels may overlap on the GPU); ms_total spans first launch .. last completion. */
int tdm_mvsnet_run_resident_multi(tdm_mvsnet* const* hs, int n, int iters_total, float* ms_total, int* launches);
/* Host-only introspection (no GPU needed): the plane-sweep homography of one source view as the cost-volume kernel receives it
 * (homo_warping, cva_mvsnet/models/module.py:795-808): M = K [R|t]_src^-1 (K [R|t]_ref^-1)^-1, rot = M[:3,:3], trans = M[:3,3];
 * a reference pixel (x, y) at depth d lands at rot (x, y, 1)^T d + trans in the source image. K row-major 3x3, poses row-major
 * 4x4 cam->world. */
int tdm_debug_homography(const float* K3x3, const float* c2w_ref, const float* c2w_src, float* rot9, float* trans3);
/* Host-only introspection (no GPU needed): the tile plan the tcgen05 convolution planner picks for a layer - cin, N columns of
 * the MMA (npad, doubled for hi/lo weights), kd (1: 2-D over planes, 3: 3-D, 2: transposed-as-GEMM), the grid D x H x W the
 * tiles live on, pd (D halo), mode (0 stride-1, 1 transposed, 3 up2, 4 input-stationary), shared-memory budget in KB.
 * out12 = {ring slots, tile rows, tile cols, pitch, planes per tile, 128-row chunks per plane, slot positions, tiles_w,
 * tiles_h, tiles_d, grid, shared-memory bytes}. */
int tdm_debug_conv_plan(int cin, int npad, int kd, int D, int H, int W, int pd, int mode, int smem_kb, long long* out12);
/* Per-kernel CUDA-event timing of one resident forward: writes lines "name ms algorithmic_bytes flops\n"
 * into buf (NUL terminated). Returns bytes written or <0. */
long long tdm_mvsnet_profile(tdm_mvsnet* h, char* buf, size_t capacity);

/* ------------------------------------------------------------------------------------------------
 * TSDF fusion (replaces DrFusion; dr_fusion.h:18-73, implementation dr_fusion/src/tsdfvh/tsdf_volume.cu)
 * ---------------------------------------------------------------------------------------------- */
typedef struct tdm_fusion tdm_fusion;
typedef struct tdm_fusion_options { /* field-for-field DrFusionOptions, dr_fusion.h:18-36 */
  float voxel_size;
  int num_buckets;
  int bucket_size;
  int num_blocks;
  int block_size;
  int max_sdf_weight;
  float truncation_distance;
  float max_sensor_depth;
  float min_sensor_depth;
  int num_render_streams;
  float fx, fy, cx, cy;
  int height, width;
} tdm_fusion_options;

int tdm_fusion_create(const tdm_fusion_options* opt, int device, tdm_fusion** out);
void tdm_fusion_destroy(tdm_fusion* h);
/* DrFusion::IntegrateScanAsync (dr_fusion.h:50): bgr H*W*3 u8, depth H*W f32 metric z-depth, pose 4x4
 * row-major cam->world; inputs are copied into pinned staging before returning (tsdf_volume.cu:542-543). */
int tdm_fusion_integrate_async(tdm_fusion* h, const unsigned char* bgr, const float* depth, const float* pose);
/* DrFusion::RenderAsync (dr_fusion.h:52): n_poses must equal num_render_streams (tsdf_volume.cu:643-648). */
int tdm_fusion_render_async(tdm_fusion* h, const float* const* camera_poses, int n_poses);
/* DrFusion::GetRenderResult (dr_fusion.h:54): pointers into a pinned double buffer owned by the handle,
 * valid until the next get_render_result (tsdf_volume.cu:710-732). bgr_out/depth_out: arrays of n_poses. */
int tdm_fusion_get_render_result(tdm_fusion* h, unsigned char** bgr_out, float** depth_out, int n_poses);
int tdm_fusion_synchronize(tdm_fusion* h);
/* Multi-GPU extension (not in the reference): restrict this instance to the Z-slab of voxel blocks with
 * z_block_lo <= block.z < z_block_hi (block = 8 voxels; pass the owned range widened by one halo block so that trilinear
 * reads never cross ranks). Every rank integrates the same (broadcast) scans; renders are combined by a per-pixel
 * nearest-hit reduction (tandem_b200/parallel.py: reduce_nearest_hit). Call before the first scan. */
int tdm_fusion_set_slab(tdm_fusion* h, int z_block_lo, int z_block_hi);
/* Interleaved variant of the Z-slab partition: block row z is owned by rank ((z - z0_block) div k_blocks) mod world; every rank
 * additionally stores one halo block row on either side of each of its slabs.  Thin interleaved slabs balance the PER-FRAME work
 * (allocation, integration, ray-cast samples) over the ranks for any view direction - contiguous slabs only balance memory - at
 * the price of (k_blocks + 2) / k_blocks redundant integration.  Same exchange step as set_slab.  Call before the first scan. */
int tdm_fusion_set_interleave(tdm_fusion* h, int rank, int world, int k_blocks, int z0_block);
/* Pixel-partitioned ray-cast over the Z-slab-partitioned volume (the fused compute + transfer form of the exchange): every rank
 * exports its hash table and voxel pool (tdm_fusion_peer_export: raw pointers for instances of one process, CUDA IPC handles for
 * other processes), the handles of all ranks are gathered by the caller (e.g. torch.distributed.all_gather_object) and attached
 * (tdm_fusion_peer_attach: cudaIpcOpenMemHandle -> NVLink P2P mapping).  From then on RenderAsync renders only the 8x8 pixel
 * tiles t with t mod world == rank, but marches them through the WHOLE volume, reading each voxel from the rank that owns its
 * block row (local HBM or P2P loads inside the ray-cast kernel) - bit-identical to the single-volume render, occluders in other
 * slabs included; foreign tiles carry the "miss" key, so the same MIN all-reduce (slab_exchange mode) assembles the image.
 * Needs contiguous slabs in rank order WITHOUT halo rows (set_slab(owned_lo, owned_hi)), and the caller must order every rank's
 * ray-cast behind every rank's integration of the same scan (a stream-ordered barrier, tandem_b200.parallel.stream_barrier). */
typedef struct tdm_fusion_peer_handle {
  unsigned long long keys_ptr, ptrs_ptr, voxels_ptr;   /* device pointers (meaningful inside the exporting process) */
  long long pid;
  int device, num_buckets, bucket_size, slab_lo, slab_hi, reserved;
  unsigned char ipc_keys[64], ipc_ptrs[64], ipc_voxels[64];   /* cudaIpcMemHandle_t */
} tdm_fusion_peer_handle;
int tdm_fusion_peer_export(tdm_fusion* h, tdm_fusion_peer_handle* out);
int tdm_fusion_peer_attach(tdm_fusion* h, const tdm_fusion_peer_handle* all_ranks, int world, int rank);
/* The one exchange step of the slab-partitioned ray-cast, on the device: tdm_fusion_render_keys_device packs render
 * `render_index` of the last RenderAsync into height*width int64 keys (depth bits << 24 | b | g<<8 | r<<16; miss = +inf) in a
 * DEVICE buffer owned by the handle (stream synchronised on return) - the caller all-reduces it with MIN over the ranks
 * (NCCL, in place) - tdm_fusion_unpack_keys turns reduced keys back into a depth map and a bgr image in HOST buffers. */
int tdm_fusion_render_keys_device(tdm_fusion* h, int render_index, long long** keys_dev);
int tdm_fusion_unpack_keys(tdm_fusion* h, const long long* keys_dev, float* depth_out, unsigned char* bgr_out);
/* The CUDA stream (cudaStream_t) all work of this handle is ordered on.  With tdm_fusion_set_option(h, "slab_exchange", 1)
 * the ray-cast writes the keys itself (no pack kernel, no per-slab D2H, GetRenderResult does not wait) and
 * tdm_fusion_render_keys_device returns without a host sync: the caller enqueues its MIN all-reduce ON THIS STREAM (e.g.
 * torch.cuda.ExternalStream + torch.distributed, or ncclAllReduce directly) and tdm_fusion_unpack_keys behind it - one
 * host sync per frame, in unpack_keys. */
int tdm_fusion_stream(tdm_fusion* h, void** stream_out);
/* Mesh (dr_fusion.h:56-68; TsdfVolume::ExtractMeshAsync / GetMeshSync / ExtractMesh, tsdf_volume.cu:739-839; kernel
 * marching_cubes/mesh_extractor.cu:244-265): marching cubes over the cells of the box [lower, upper) at voxel spacing.
 * Output layout of GetMeshSync: vertices as xyz float triples, colours as rgb float triples in [0,1], 3 consecutive
 * vertices per triangle (no index buffer). Triangle ORDER is unspecified (the reference appends with atomicAdd).
 *   tdm_fusion_extract_mesh_async = ExtractMeshAsync: launches the extraction on the fusion stream and returns; only legal
 *     where IntegrateScanAsync is legal (tsdf_volume.cu:760-763) and not twice in a row (:769-772);
 *   tdm_fusion_get_mesh = GetMeshSync: waits, copies 3*triangles vertices out, returns the vertex count (error if it exceeds
 *     max_vertices, :796-799); with vert == cols == NULL it waits and returns the count only, the mesh stays pending;
 *   tdm_fusion_extract_mesh = TsdfVolume::ExtractMesh (blocking, no call-order check; DrFusion::GetMesh / SaveMeshToFile);
 *     with vert == cols == NULL it only returns the vertex count and keeps the mesh on the device, so that the next call
 *     (with buffers of that size) copies it out without extracting again. */
long long tdm_fusion_extract_mesh(tdm_fusion* h, const float lower[3], const float upper[3],
                                  float* vert, float* cols, size_t max_vertices);
int tdm_fusion_extract_mesh_async(tdm_fusion* h, const float lower[3], const float upper[3]);
long long tdm_fusion_get_mesh(tdm_fusion* h, float* vert, float* cols, size_t max_vertices);
/* Host-only introspection (no GPU needed): the per-axis cell table the mesh extractor uploads for one axis of a box - per cell
 * the voxel indices {gMA, gMB, gPA, gPB, gC} (ints5), {wM, wP, cM, cP} (floats4), and per voxel block {first cell, #cells}
 * (ranges2, bmin_nb = {first block, #blocks}). Returns the number of cells (mesh_extractor.cu:248-261 arithmetic). */
int tdm_debug_mesh_axis_table(float lower, float upper, float voxel_size, int* ints5, float* floats4, int* ranges2, int* bmin_nb, int capacity);
/* Host-only introspection (no GPU needed): the bucket the voxel-hashing kernels compute for block (x, y, z) - a division-free
 * remainder (precomputed ceil(2^64 / num_buckets)) - and, in *reference_expression when non-NULL, the reference's expression
 * ((x * 73856093) ^ (y * 19349669) ^ (z * 83492791)) % num_buckets, + num_buckets when negative (hash_table.cu:157-168). */
int tdm_debug_hash_slot(int x, int y, int z, int num_buckets, int* reference_expression);
/* Device time (CUDA events) of the last extraction: classify + scan + emit. */
int tdm_fusion_last_mesh_ms(tdm_fusion* h, float* ms);
/* Introspection for parity tests and the roofline: counters of the last integrate / render. */
typedef struct tdm_fusion_stats {
  long long allocated_blocks;      /* total blocks in the map */
  long long visible_blocks;        /* blocks integrated in the last scan */
  long long dropped_blocks;        /* bucket overflows (never allocated), cumulative */
  long long candidate_blocks;      /* distinct blocks touched by the last allocation pass */
} tdm_fusion_stats;
int tdm_fusion_get_stats(tdm_fusion* h, tdm_fusion_stats* out);
/* Dump the block map: xyz int triples (sorted lexicographically) into coords (capacity in blocks) and,
 * if voxels != NULL, 512 voxels per block as {float sdf; uint8 c0,c1,c2; uint8 weight} (8 B each). */
long long tdm_fusion_dump_blocks(tdm_fusion* h, int* coords, void* voxels, size_t capacity_blocks);
/* Device-resident measurement of integrate+render for the last submitted scan/pose (sums over iters);
 * tdm_fusion_last_alloc_ms: the allocation kernel's share of ms_integrate, per iteration. */
int tdm_fusion_run_resident(tdm_fusion* h, int iters, float* ms_integrate, float* ms_render);
int tdm_fusion_last_alloc_ms(tdm_fusion* h, float* ms);
/* A/B switches for measurements (results are bit-identical either way): "alloc_filter" (CTA-level shared-memory filter in
 * front of the hash table during allocation), "raycast_cache8" (8-entry per-ray block cache); both default to 0 - measured
 * on B200 they do not pay (the table probes are L2 hits behind other latency); "raycast_persistent" (warps pull rays from a
 * counter and refill finished lanes; default 0 = one ray per thread, which measured faster: 0.82 vs 0.875 ms), "raycast_shared" (index arithmetic shared between the nine voxel reads of a sample, one block lookup when they sit in one
 * block; default 1), "integrate_compact" (visibility pass + update of the
 * visible blocks only, default 1; 0 = every CTA scans the whole block list); "raycast_tile" (0: 16x16-pixel CTAs, 1: 8x8
 * (default), 2: 8x4, 3: 16x8); "raycast_dedup" (a sample whose eight corners straddle voxel-block faces looks every distinct
 * block up once instead of corner by corner; default 1: 0.31 vs 0.38 ms); "fast_div" (constant-divisor FMA division in the
 * ray-cast, exhaustively checked correctly rounded; default 0: 0.316 vs 0.311 ms); "occ_skip" (dilated block-occupancy bitmap
 * in front of the sampler; default 0: 0.330 vs 0.311 ms, a frustum-allocated map has no unallocated space on a ray's way);
 * "slab_clip" / "slab_exchange" (Z-slab volumes, see above). */
int tdm_fusion_set_option(tdm_fusion* h, const char* name, int value);

/* ------------------------------------------------------------------------------------------------
 * Coarse tracker, pyramid level 0 (replaces CudaCoarseTracker; cuda_coarse_tracker.h:9-82)
 * ---------------------------------------------------------------------------------------------- */
typedef struct tdm_tracker tdm_tracker;
int tdm_tracker_create(int w, int h, float setting_huberTH, float setting_coarseCutoffTH, int n_max,
                       int device, tdm_tracker** out);
void tdm_tracker_destroy(tdm_tracker* t);
int tdm_tracker_set_k(tdm_tracker* t, int w, int h, float fx, float fy, float cx, float cy);
int tdm_tracker_set_reference(tdm_tracker* t, int n, const float* pc_u, const float* pc_v,
                              const float* pc_idepth, const float* pc_color, float ref_exposure,
                              const double ref_aff_g2l[2]);
int tdm_tracker_set_new(tdm_tracker* t, const float* dInew /* h*w*3: (I,dx,dy) interleaved */);
/* calcRes: refToNew 4x4 row-major double; res6 = [E, numTermsInE, flowT, 0, flowRT, saturatedRatio]. */
int tdm_tracker_calc_res(tdm_tracker* t, const double* refToNew, float new_exposure,
                         const double aff_g2l[2], float cutoffTH, double res6[6]);
/* calcG: H 8x8 row-major, b 8; uses the warped buffers of the last calc_res. */
int tdm_tracker_calc_g(tdm_tracker* t, float new_exposure, const double aff_g2l[2], double H[64], double b[8]);
/* Fused single-launch calcRes+calcG (no warped buffers), same outputs. */
int tdm_tracker_calc_res_g(tdm_tracker* t, const double* refToNew, float new_exposure,
                           const double aff_g2l[2], float cutoffTH, double res6[6], double H[64], double b[8]);
/* Extension (the tracker's only parallel axis): calcRes of n_hyp (<= 64) independent motion hypotheses - the poses
 * FullSystem::trackNewCoarse tries one after the other (FullSystem.cpp:437-530) - in ONE launch and one synchronisation.
 * refToNew: n_hyp x 16, aff_g2l: n_hyp x 2, res6: n_hyp x 6; hypothesis k's result equals tdm_tracker_calc_res on that pose bit
 * for bit.  Does not disturb the buffers of the last calc_res (calc_g still refers to them). */
int tdm_tracker_calc_res_batch(tdm_tracker* t, int n_hyp, const double* refToNew, float new_exposure, const double* aff_g2l,
                               float cutoffTH, double* res6);
int tdm_tracker_synchronize(tdm_tracker* t);
int tdm_tracker_run_resident(tdm_tracker* t, int iters, float* ms_total);

/* ------------------------------------------------------------------------------------------------
 * SURVEY.md 8(f) "next" rows: the host stages either side of the evaluation, moved onto the device.  These entry
 * points have no counterpart in the reference's class (they replace code ABOVE its boundary); INTEGRATION.md shows
 * the three call sites in CoarseTracker.cpp / HessianBlocks.cpp that would switch to them.
 * ---------------------------------------------------------------------------------------------- */
/* n2  FrameHessian::makeImages (tandem/src/FullSystem/HessianBlocks.cpp:128-191): grey pyramid by 2x2 means +
 *     central-difference gradients for `levels` levels (level l is (w>>l) x (h>>l), globalCalib.cpp:84-85).
 *     build() uploads the w*h grey image (1.2 MB instead of the 3.7 MB float3 image setNew uploads) and is asynchronous;
 *     get_level() copies one level back (dI: (I,dx,dy) float3 per pixel; abs_squared_grad: dx^2+dy^2, no gamma weights). */
typedef struct tdm_pyramid tdm_pyramid;
int tdm_pyramid_create(int w, int h, int levels, int device, tdm_pyramid** out);
void tdm_pyramid_destroy(tdm_pyramid* p);
int tdm_pyramid_build(tdm_pyramid* p, const float* gray);
int tdm_pyramid_get_level(tdm_pyramid* p, int level, float* dI, float* abs_squared_grad);
/* setNew (cuda_coarse_tracker.cpp:99-102) without the host round trip: device-to-device from a pyramid level whose size
 * equals the tracker's (a tracker instance per level puts levels 1..3 of CoarseTracker.cpp:774 on the GPU as well). */
int tdm_tracker_set_new_from_pyramid(tdm_tracker* t, tdm_pyramid* p, int level);

/* n1  the dense part of CoarseTracker::setCoarseTrackingRef (tandem/src/FullSystem/CoarseTracker.cpp:655-732) followed by
 *     setReference: forward-warps the dense depth map (host pointer, or the render_index-th map of the last
 *     tdm_fusion_render_async - device resident, no PCIe hop) with T_depth_to_ref = ref.camToWorld^-1 * depth.camToWorld
 *     (row-major), keeps the nearest depth per target pixel, and appends the hit pixels in raster order behind the
 *     n_sparse sparse points of makeCoarseDepthL0.  pc_* hold n_sparse + 1 entries: the extra one is the slot the
 *     reference's `++pc_n` skips (CoarseTracker.cpp:717-722) and is uploaded as is; NULL is allowed when n_sparse == 0.
 *     idepth0 (h*w, the sparse inverse-depth image idepth[0]) may be NULL.  The grey values of the reference keyframe
 *     come from ref_gray (h*w floats) or from level 0 of a pyramid.  *pc_n receives the reference's pc_n[0]. */
int tdm_tracker_set_reference_dense(tdm_tracker* t, const float* depth, tdm_fusion* depth_from_fusion, int render_index,
                                    const double T_depth_to_ref[16], int tracking_step, int dense_only, int n_sparse,
                                    const float* pc_u, const float* pc_v, const float* pc_idepth, const float* pc_color,
                                    const float* idepth0, const float* ref_gray, tdm_pyramid* ref_gray_from_pyramid,
                                    float ref_exposure, const double ref_aff_g2l[2], int* pc_n);
/* Reads the first n reference points back (tests, debugging). Any output pointer may be NULL. */
int tdm_tracker_get_reference(tdm_tracker* t, int n, float* pc_u, float* pc_v, float* pc_idepth, float* pc_color);

/* n3  one pyramid level of CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:761-916) without host round trips:
 *     initial evaluation with cutoff doubling (:775-790), then up to max_iterations damped Gauss-Newton steps with the
 *     reference's accept / lambda / extrapolation / termination rules; fix_a / fix_b = setting_affineOptModeA/B < 0.
 *     Blocking; the caller keeps the coarse-to-fine logic (repeat a level when cutoff_repeat > 1, abort thresholds). */
typedef struct tdm_track_result {
  double ref_to_new[16];   /* row-major */
  double aff_g2l[2];
  double res[6];           /* resOld of the accepted state: E, numTermsInE, flow T, 0, flow RT, saturated ratio */
  int iterations;          /* executed LM iterations */
  int evaluations;         /* fused residual + normal-equation evaluations launched */
  float cutoff_repeat;     /* levelCutoffRepeat */
  float device_ms;         /* CUDA-event time of the whole loop */
} tdm_track_result;
int tdm_tracker_track(tdm_tracker* t, const double refToNew[16], const double aff_g2l[2], float new_exposure,
                      float coarse_cutoff_th, int max_iterations, float lambda_extrapolation_limit, int fix_a, int fix_b,
                      tdm_track_result* out);

#ifdef __cplusplus
}
#endif
#endif /* TANDEM_B200_H */

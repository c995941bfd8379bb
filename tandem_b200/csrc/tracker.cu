// Dense photometric tracker, pyramid level 0, behind the CudaCoarseTracker call surface
// (tandem/libdr/cuda_coarse_tracker/include/public/cuda_coarse_tracker.h:9-82).
//
// Reference being replaced:
//   calcResKernelNew<128>  cuda_coarse_tracker_private.cu:41-214   7 x cub::BlockReduce + 7 float atomics / block
//   calcGKernel<128,float> cuda_coarse_tracker_private.cu:261-394  45 serialized block reductions, <1 wave
//   host wrapper           cuda_coarse_tracker.cpp:195-356          memset + H2D + kernel + sync + D2H per call
// B200 design: K8 = ONE kernel template for calcRes, calcG and the fused calcRes+calcG.  296 persistent CTAs
// (2 per SM) grid-stride over the points, per-thread fp32 partials are widened to double, reduced with warp
// shuffles and one shared-memory pass, each CTA writes its partial vector, and the last CTA to finish (atomic
// ticket) adds the partials in a fixed order and writes the result straight into mapped pinned host memory:
// deterministic (the reference's float atomics are not), one launch, no memset, no D2H copy.
#include <cmath>
#include <cstring>
#include <vector>

#include "tracker.h"

namespace tdm {

namespace {

constexpr int kThreads = 256;
constexpr int kBlocks = 296;  // 2 x 148 SMs
constexpr int kRes = 7;       // E, numTermsInE, numTermsInWarped, numSaturated, shiftT, shiftRT, shiftNum
constexpr int kG = 45;
constexpr int kOut = kRes + kG;

struct TrkParams {
  int w, h, n;
  float fx, fy, cx, cy;
  float RKi[9], Ki[9], t[3];
  float affa, affb;     // affLL
  float ref_b;          // ref_aff_g2l.b  (calcG's b0)
  float huber, cutoff, maxEnergy;
};

// pose-dependent parameters as the device-side LM loop (tdm_tracker_track) hands them from one evaluation to the next
struct TrkPose {
  float RKi[9], t[3];
  float affa, affb;
  float cutoff, maxEnergy;
};

struct LmCtx;
struct TrkBufs {
  LmCtx* lm;            // device LM loop (tdm_tracker_track): pose in, LM step out; nullptr for the plain calls
  const float *pc_u, *pc_v, *pc_idepth, *pc_color, *dInew;
  float* warped;        // 7 x n_max: u, v, dx, dy, idepth, residual, weight
  long long n_max;
  double* partials;     // [kBlocks][kOut]
  unsigned* ticket;
  double* out;          // mapped pinned host memory, kOut doubles
  const TrkPose* batch; // MODE 3: one pose per blockIdx.y (motion hypotheses evaluated in ONE launch); partials / ticket / out are
                        // then arrays indexed by blockIdx.y
};

__device__ bool lm_done(const LmCtx* c);
__device__ const TrkPose* lm_pose(const LmCtx* c);
__device__ __noinline__ void lm_step(LmCtx* c, const double* o);

__device__ __forceinline__ float3 bilinear33(const float* __restrict__ mat, float x, float y, int width) {  // cu:22-38
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1.0f - dx - dy + dxdy;
  float3 r;
  r.x = w11 * bp[3 * (1 + width)] + w01 * bp[3 * width] + w10 * bp[3] + w00 * bp[0];
  r.y = w11 * bp[3 * (1 + width) + 1] + w01 * bp[3 * width + 1] + w10 * bp[4] + w00 * bp[1];
  r.z = w11 * bp[3 * (1 + width) + 2] + w01 * bp[3 * width + 2] + w10 * bp[5] + w00 * bp[2];
  return r;
}

__device__ __forceinline__ float mad3(const float* m, float x, float y) {  // row . (x,y,1) without contraction
  return __fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], 1.0f));
}

__device__ __forceinline__ void accumulate_g(float* g, float dxI, float dyI, float u, float v, float id, float refc,
                                             float res, float hw, const TrkParams& p) {  // cu:305-346
  const float dx = dxI * p.fx, dy = dyI * p.fy;
  float J[9];
  J[0] = id * dx; J[1] = id * dy; J[2] = -id * (u * dx + v * dy);
  J[3] = -(u * v * dx + dy + dy * v * v); J[4] = u * v * dy + dx + dx * u * u; J[5] = u * dy - v * dx;
  J[6] = p.affa * (p.ref_b - refc); J[7] = -1.f; J[8] = res;
  int k = 0;
#pragma unroll
  for (int j1 = 0; j1 < 9; ++j1) {
    const float jw = J[j1] * hw;
#pragma unroll
    for (int j2 = j1; j2 < 9; ++j2) g[k++] += jw * J[j2];
  }
}

// MODE 0: calcRes (stats + warped buffers)   1: calcG from warped buffers   2: fused (stats + G, no buffers)
// MODE 3: calcRes statistics of blockIdx.y = 0..n_hyp-1 independent poses in one launch (no buffers): the motion hypotheses
//         DSO tries one after the other (FullSystem.cpp:437-530) evaluated as a batch - SURVEY 8(e)'s only parallel axis of the
//         tracker.  Same per-thread point striding and the same fixed-order reduction as MODE 0 -> identical numbers.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_tracker(const __grid_constant__ TrkParams p_in, TrkBufs b) {
  TrkParams p = p_in;
  if (MODE == 2 && b.lm) {
    if (lm_done(b.lm)) return;                   // uniform over the grid: nobody touches the ticket
    const TrkPose* q = lm_pose(b.lm);
#pragma unroll
    for (int i = 0; i < 9; ++i) p.RKi[i] = q->RKi[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) p.t[i] = q->t[i];
    p.affa = q->affa; p.affb = q->affb; p.cutoff = q->cutoff; p.maxEnergy = q->maxEnergy;
  }
  if (MODE == 3) {
    const TrkPose* q = b.batch + blockIdx.y;
#pragma unroll
    for (int i = 0; i < 9; ++i) p.RKi[i] = q->RKi[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) p.t[i] = q->t[i];
    p.affa = q->affa; p.affb = q->affb; p.cutoff = q->cutoff; p.maxEnergy = q->maxEnergy;
    b.partials += (size_t)blockIdx.y * kBlocks * kOut;
    b.ticket += blockIdx.y;
    b.out += (size_t)blockIdx.y * kOut;
  }
  constexpr int NV = (MODE == 0 || MODE == 3) ? kRes : (MODE == 1 ? kG : kOut);
  constexpr int OFF = MODE == 1 ? kRes : 0;
  float acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.f;
  float* st = acc;                               // stats (MODE 0, 2)
  float* g = acc + (MODE == 2 ? kRes : 0);       // 45 products (MODE 1, 2)
  const long long nm = b.n_max;

  for (int i = blockIdx.x * kThreads + threadIdx.x; i < p.n; i += kBlocks * kThreads) {
    if (MODE == 1) {
      const float hw = b.warped[6 * nm + i];
      if (hw != 0.f)
        accumulate_g(g, b.warped[2 * nm + i], b.warped[3 * nm + i], b.warped[i], b.warped[nm + i], b.warped[4 * nm + i],
                     b.pc_color[i], b.warped[5 * nm + i], hw, p);
      continue;
    }
    const float id = b.pc_idepth[i], x = b.pc_u[i], y = b.pc_v[i];
    float pt[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) pt[r] = mad3(p.RKi + 3 * r, x, y);
    float p1[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) p1[r] = __fadd_rn(pt[r], __fmul_rn(p.t[r], id));
    const float u = __fdiv_rn(p1[0], p1[2]), v = __fdiv_rn(p1[1], p1[2]);
    const float Ku = __fadd_rn(__fmul_rn(p.fx, u), p.cx), Kv = __fadd_rn(__fmul_rn(p.fy, v), p.cy);
    const float nid = __fdiv_rn(id, p1[2]);
    if (i % 32 == 0) {  // flow statistics on every 32nd point (cu:119-166)
      float ptK[3], a[3], c[3], d[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        ptK[r] = mad3(p.Ki + 3 * r, x, y);
        const float tid = __fmul_rn(p.t[r], id);
        a[r] = __fadd_rn(ptK[r], tid); c[r] = __fsub_rn(ptK[r], tid); d[r] = __fsub_rn(pt[r], tid);
      }
      const float KuT = p.fx * (a[0] / a[2]) + p.cx, KvT = p.fy * (a[1] / a[2]) + p.cy;
      const float KuT2 = p.fx * (c[0] / c[2]) + p.cx, KvT2 = p.fy * (c[1] / c[2]) + p.cy;
      const float Ku3 = p.fx * (d[0] / d[2]) + p.cx, Kv3 = p.fy * (d[1] / d[2]) + p.cy;
      st[4] += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y) + (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      st[5] += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y) + (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      st[6] += 2.f;
    }
    float wu = 0, wv = 0, wdx = 0, wdy = 0, wid = 0, wres = 0, whw = 0;
    if (Ku > 2 && Kv > 2 && Ku < p.w - 3 && Kv < p.h - 3 && nid > 0) {
      const float refc = b.pc_color[i];
      const float3 hit = bilinear33(b.dInew, Ku, Kv, p.w);
      if (isfinite(hit.x)) {
        const float res = __fsub_rn(hit.x, __fadd_rn(__fmul_rn(p.affa, refc), p.affb));
        const float ar = fabsf(res);
        const float hw = ar < p.huber ? 1.f : p.huber / ar;
        if (ar > p.cutoff) {
          st[0] += p.maxEnergy; st[1] += 1.f; st[3] += 1.f;
        } else {
          st[0] += hw * res * res * (2.f - hw); st[1] += 1.f; st[2] += 1.f;
          wu = u; wv = v; wdx = hit.y; wdy = hit.z; wid = nid; wres = res; whw = hw;
          if (MODE == 2) accumulate_g(g, hit.y, hit.z, u, v, nid, refc, res, hw, p);
        }
      }
    }
    if (MODE == 0) {  // un-compacted buffers, zeros at rejected points (cu:75-81,191-197)
      b.warped[i] = wu; b.warped[nm + i] = wv; b.warped[2 * nm + i] = wdx; b.warped[3 * nm + i] = wdy;
      b.warped[4 * nm + i] = wid; b.warped[5 * nm + i] = wres; b.warped[6 * nm + i] = whw;
    }
  }

  // ---- reduction: double, warp shuffle -> shared -> per-CTA partial -> last CTA sums in a fixed order ----
  __shared__ double sm[kThreads / 32][NV];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double v = (double)acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
#pragma unroll
    for (int wv = 0; wv < kThreads / 32; ++wv) s += sm[wv][threadIdx.x];
    b.partials[(size_t)blockIdx.x * kOut + OFF + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(b.ticket, 1u) == kBlocks - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < NV) {
      double s = 0;
      for (int blk = 0; blk < kBlocks; ++blk) s += __ldcg(&b.partials[(size_t)blk * kOut + OFF + threadIdx.x]);
      b.out[OFF + threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *b.ticket = 0;
    if (MODE == 2 && b.lm) {                     // n3: the LM step rides on the evaluation's last CTA - no extra launch
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) lm_step(b.lm, b.out);
    }
  }
}


// ================================================================================================
// n2: FrameHessian::makeImages on the device (HessianBlocks.cpp:128-191).  Level 0 takes the grey image,
// level l > 0 the 2x2 mean of level l-1; gradients are central differences over the FLAT index range [w, w*(h-1)) as in
// the reference (so x = 0 / w-1 wrap to the neighbouring rows), non-finite -> 0; rows 0 and h-1 are zero.
// ================================================================================================
__global__ void k_pyr_level(const float* __restrict__ src /*grey w*h (lvl 0) or previous level's float3*/, int src_is_dI,
                            int wprev, float* __restrict__ gray /*scratch wl*hl*/, int wl, int hl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wl * hl) return;
  if (!src_is_dI) { gray[i] = src[i]; return; }
  const int y = i / wl, x = i - y * wl;
  const float* q = src + 3 * (size_t)(2 * x + 2 * y * wprev);
  gray[i] = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(q[0], q[3]), q[3 * wprev]), q[3 * wprev + 3]));
}
__global__ void k_pyr_grad(const float* __restrict__ gray, float* __restrict__ dI, float* __restrict__ absgrad, int wl, int hl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= wl * hl) return;
  float dx = 0.f, dy = 0.f;
  if (i >= wl && i < wl * (hl - 1)) {
    dx = __fmul_rn(0.5f, __fsub_rn(gray[i + 1], gray[i - 1]));
    dy = __fmul_rn(0.5f, __fsub_rn(gray[i + wl], gray[i - wl]));
    if (!isfinite(dx)) dx = 0.f;
    if (!isfinite(dy)) dy = 0.f;
  }
  dI[3 * (size_t)i] = gray[i];
  dI[3 * (size_t)i + 1] = dx;
  dI[3 * (size_t)i + 2] = dy;
  absgrad[i] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// ================================================================================================
// n1: CoarseTracker::setCoarseTrackingRef's dense part on the device (CoarseTracker.cpp:655-732):
// forward-warp the rendered depth into the reference keyframe (nearest depth wins: atomicMin on the bit pattern of
// the positive float), then append the hit pixels behind the sparse points in RASTER order (row counts -> row offsets
// -> ordered intra-row compaction), so the point order - and with it our deterministic reduction - is reproducible.
// ================================================================================================
struct DenseWarp {
  float KRKi[9], Kt[3];
};
__global__ void k_dense_warp(const float* __restrict__ depth, int w, int h, int step, DenseWarp m, unsigned* __restrict__ proj) {
  const int nx = (w + step - 1) / step, ny = (h + step - 1) / step;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx * ny) return;
  const int y = (i / nx) * step, x = (i - (i / nx) * nx) * step;
  const float z = depth[x + (size_t)y * w];
  if (!(z > 0.f)) return;
  const float ox = __fmul_rn((float)x, z), oy = __fmul_rn((float)y, z);
  float p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    p[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m.KRKi[3 * r], ox), __fmul_rn(m.KRKi[3 * r + 1], oy)), __fmul_rn(m.KRKi[3 * r + 2], z)), m.Kt[r]);
  if (!(p[2] > 0.f)) return;
  const int pu = (int)__fadd_rn(__fdiv_rn(p[0], p[2]), 0.5f), pv = (int)__fadd_rn(__fdiv_rn(p[1], p[2]), 0.5f);
  if (pu > w - 4 || pv > h - 4 || pu < 3 || pv < 3) return;
  atomicMin(proj + pu + (size_t)pv * w, __float_as_uint(p[2]));
}
__device__ __forceinline__ bool dense_take(const unsigned* proj, const float* idepth0, int dense_only, int i) {
  return proj[i] != 0xFFFFFFFFu && (dense_only || !idepth0 || idepth0[i] <= 0.f);
}
__global__ void __launch_bounds__(256) k_dense_rowcount(const unsigned* __restrict__ proj, const float* __restrict__ idepth0,
                                                         int dense_only, int w, int h, int* __restrict__ rowcnt) {
  const int y = blockIdx.x;
  int c = 0;
  if (y >= 2 && y < h - 2)
    for (int x = 2 + threadIdx.x; x < w - 2; x += 256) c += dense_take(proj, idepth0, dense_only, x + y * w) ? 1 : 0;
  __shared__ int sm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int k = 0; k < 8; ++k) s += sm[k];
    rowcnt[y] = s;
  }
}
__global__ void k_dense_rowscan(const int* __restrict__ rowcnt, int h, int* __restrict__ rowoff, int* __restrict__ total_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int s = 0;
  for (int y = 0; y < h; ++y) { rowoff[y] = s; s += rowcnt[y]; }
  *total_out = s;
}
__global__ void __launch_bounds__(256) k_dense_emit(const unsigned* __restrict__ proj, const float* __restrict__ idepth0,
                                                     int dense_only, int w, int h, const int* __restrict__ rowoff,
                                                     const float* __restrict__ gray, int gstride, int first,
                                                     float* __restrict__ pc_u, float* __restrict__ pc_v,
                                                     float* __restrict__ pc_idepth, float* __restrict__ pc_color) {
  const int y = blockIdx.x;
  if (y < 2 || y >= h - 2) return;
  __shared__ int wsum[8];
  __shared__ int base;
  if (threadIdx.x == 0) base = first + rowoff[y];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int x0 = 2; x0 < w - 2; x0 += 256) {
    const int x = x0 + threadIdx.x;
    const int i = x + y * w;
    const bool take = x < w - 2 && dense_take(proj, idepth0, dense_only, i);
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { before += k < warp ? wsum[k] : 0; all += wsum[k]; }
    if (take) {
      const int o = base + before + __popc(bal & ((1u << lane) - 1u));
      pc_u[o] = (float)x;
      pc_v[o] = (float)y;
      pc_idepth[o] = __fdiv_rn(1.f, __uint_as_float(proj[i]));
      pc_color[o] = gray[(size_t)i * gstride];
    }
    __syncthreads();
    if (threadIdx.x == 0) base += all;
    __syncthreads();
  }
}

// ================================================================================================
// n3: the level's Levenberg-Marquardt loop of CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:761-916)
// as a chain of (k_tracker<2>, k_lm_step) launches with NO host round trip: k_lm_step (one thread, double) turns the
// evaluation's 52 sums into the accept/reject decision, the damped 8x8 solve, the SE3 update and the next candidate's
// TrkPose; a `done` flag turns the remaining enqueued launches into no-ops.
// ================================================================================================
struct LmConfig {
  float Ki[9];
  float ref_exposure, new_exposure;
  double ref_aff[2];
  float huber, coarse_cutoff;
  int max_iter;
  float lambda_limit;
  int fix_a, fix_b;
};
struct LmState {
  double T[16], aff[2];      // accepted
  double Tn[16], affn[2];    // candidate under evaluation
  double res_old[6], H[64], b[8];
  double lambda, inc_norm;
  float repeat;
  int phase, iter, evals;
};
struct LmResult {            // mapped pinned host memory
  double T[16], aff[2], res[6];
  int iterations, evaluations;
  float repeat;
  int num_warped;
};

struct LmCtx {               // device memory, one per tracker
  LmConfig cfg;
  LmState s;
  TrkPose pose;
  int done;
  LmResult* out;             // mapped pinned host memory
};
__device__ bool lm_done(const LmCtx* c) { return c->done != 0; }
__device__ const TrkPose* lm_pose(const LmCtx* c) { return &c->pose; }

__device__ void lm_pose_from(const LmConfig& c, const double* T, const double* aff, float cutoff, TrkPose* o) {
  float R[9];
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) R[3 * r + k] = (float)T[4 * r + k];
    o->t[r] = (float)T[4 * r + 3];
  }
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s = __fadd_rn(s, __fmul_rn(R[3 * r + k], c.Ki[3 * k + q]));
      o->RKi[3 * r + q] = s;
    }
  float eF = c.ref_exposure, eT = c.new_exposure;       // AffLight::fromToVecExposure, cuda_coarse_tracker.cpp:42-52
  if (eF == 0 || eT == 0) eT = eF = 1;
  const double a = exp(aff[0] - c.ref_aff[0]) * eT / eF;
  const double bb = aff[1] - a * c.ref_aff[1];
  o->affa = (float)a; o->affb = (float)bb;
  o->cutoff = cutoff;
  o->maxEnergy = 2 * c.huber * cutoff - c.huber * c.huber;
}

__device__ void lm_se3_exp(const double* xi, double* T) {   // Sophus convention (upsilon, omega)
  const double wx = xi[3], wy = xi[4], wz = xi[5];
  const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
  double A, B, C;
  if (th < 1e-8) { A = 1.0 - th2 / 6; B = 0.5 - th2 / 24; C = 1.0 / 6 - th2 / 120; }
  else { A = sin(th) / th; B = (1 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th); }
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) { double s = 0; for (int k = 0; k < 3; ++k) s += O[3 * r + k] * O[3 * k + q]; O2[3 * r + q] = s; }
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int r = 0; r < 3; ++r) {
    double tv = 0;
    for (int q = 0; q < 3; ++q) {
      const double id = r == q ? 1.0 : 0.0;
      T[4 * r + q] = id + A * O[3 * r + q] + B * O2[3 * r + q];
      tv += (id + B * O[3 * r + q] + C * O2[3 * r + q]) * xi[q];
    }
    T[4 * r + 3] = tv;
  }
}

__device__ void lm_next_candidate(const LmConfig& c, LmState* s, TrkPose* pose) {
  int idx[8], nf = 0;
  for (int i = 0; i < 8; ++i)
    if (!((i == 6 && c.fix_a) || (i == 7 && c.fix_b))) idx[nf++] = i;
  double M[8][9];
  for (int r = 0; r < nf; ++r) {
    for (int q = 0; q < nf; ++q) M[r][q] = s->H[8 * idx[r] + idx[q]] * (r == q ? 1.0 + s->lambda : 1.0);
    M[r][nf] = -s->b[idx[r]];
  }
  for (int k = 0; k < nf; ++k) {             // Gaussian elimination, partial pivoting
    int piv = k;
    for (int r = k + 1; r < nf; ++r) if (fabs(M[r][k]) > fabs(M[piv][k])) piv = r;
    if (piv != k) for (int q = 0; q <= nf; ++q) { const double t = M[k][q]; M[k][q] = M[piv][q]; M[piv][q] = t; }
    const double d = M[k][k];
    for (int r = k + 1; r < nf; ++r) {
      const double f = M[r][k] / d;
      for (int q = k; q <= nf; ++q) M[r][q] -= f * M[k][q];
    }
  }
  double inc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = nf - 1; k >= 0; --k) {
    double v = M[k][nf];
    for (int q = k + 1; q < nf; ++q) v -= M[k][q] * inc[idx[q]];
    inc[idx[k]] = v / M[k][k];
  }
  double extrap = 1.0;
  if (s->lambda < c.lambda_limit) extrap = sqrt(sqrt((double)c.lambda_limit / s->lambda));
  double n2 = 0, sum = 0;
  const double scale[8] = {1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 10.0, 1000.0};
  double sc[8];
  for (int i = 0; i < 8; ++i) { inc[i] *= extrap; n2 += inc[i] * inc[i]; sc[i] = inc[i] * scale[i]; sum += sc[i]; }
  s->inc_norm = sqrt(n2);
  if (!isfinite(sum)) for (int i = 0; i < 8; ++i) sc[i] = 0;
  double E[16];
  lm_se3_exp(sc, E);
  for (int r = 0; r < 4; ++r)
    for (int q = 0; q < 4; ++q) { double v = 0; for (int k = 0; k < 4; ++k) v += E[4 * r + k] * s->T[4 * k + q]; s->Tn[4 * r + q] = v; }
  s->affn[0] = s->aff[0] + sc[6];
  s->affn[1] = s->aff[1] + sc[7];
  lm_pose_from(c, s->Tn, s->affn, c.coarse_cutoff * s->repeat, pose);
}

__device__ void lm_take_g(const double* o, int num_warped, LmState* s) {   // finish_g, cuda_coarse_tracker.cpp:335-355
  const double factor = 1.0 / num_warped;
  const double scale[8] = {1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 10.0, 1000.0};
  for (int r = 0; r < 8; ++r) {
    for (int q = 0; q < 8; ++q) {
      const int lo = r < q ? r : q, hi = r < q ? q : r;
      s->H[8 * r + q] = o[kRes + lo * 9 + hi - lo * (lo + 1) / 2] * factor * scale[r] * scale[q];
    }
    s->b[r] = o[kRes + r * 9 + 8 - r * (r + 1) / 2] * factor * scale[r];
  }
}

__global__ void k_lm_init(LmCtx* c, LmConfig cfg, LmResult* out, const double* T0 /*16 + 2 aff, device*/) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  c->cfg = cfg;
  c->out = out;
  LmState* s = &c->s;
  for (int i = 0; i < 16; ++i) s->T[i] = s->Tn[i] = T0[i];
  s->aff[0] = s->affn[0] = T0[16];
  s->aff[1] = s->affn[1] = T0[17];
  s->lambda = 0.01; s->inc_norm = 0; s->repeat = 1.f; s->phase = 0; s->iter = 0; s->evals = 0;
  c->done = 0;
  lm_pose_from(cfg, s->T, s->aff, cfg.coarse_cutoff, &c->pose);
}

__device__ __noinline__ void lm_step(LmCtx* ctx, const double* o /*kOut sums of the evaluation just finished*/) {
  const LmConfig& c = ctx->cfg;
  LmState* s = &ctx->s;
  TrkPose* pose = &ctx->pose;
  LmResult* out = ctx->out;
  double res[6] = {o[0], o[1], o[4] / o[6], 0.0, o[5] / o[6], o[3] / o[1]};   // finish_res, cuda_coarse_tracker.cpp:264-272
  const int num_warped = (int)o[2];
  s->evals++;
  bool finished = false;
  if (s->phase == 0) {
    if (res[5] > 0.6 && s->repeat < 50.f) {              // CoarseTracker.cpp:779-790: too many saturated residuals
      s->repeat *= 2.f;
      lm_pose_from(c, s->T, s->aff, c.coarse_cutoff * s->repeat, pose);
      return;
    }
    for (int i = 0; i < 6; ++i) s->res_old[i] = res[i];
    lm_take_g(o, num_warped, s);
    s->phase = 1;
    if (c.max_iter <= 0) finished = true;
  } else {
    const bool accept = (res[0] / res[1]) < (s->res_old[0] / s->res_old[1]);
    if (accept) {
      lm_take_g(o, num_warped, s);
      for (int i = 0; i < 6; ++i) s->res_old[i] = res[i];
      for (int i = 0; i < 16; ++i) s->T[i] = s->Tn[i];
      s->aff[0] = s->affn[0]; s->aff[1] = s->affn[1];
      s->lambda *= 0.5;
    } else {
      s->lambda *= 4;
      if (s->lambda < c.lambda_limit) s->lambda = c.lambda_limit;
    }
    s->iter++;
    if (!(s->inc_norm > 1e-3) || s->iter >= c.max_iter) finished = true;
  }
  if (finished) {
    for (int i = 0; i < 16; ++i) out->T[i] = s->T[i];
    out->aff[0] = s->aff[0]; out->aff[1] = s->aff[1];
    for (int i = 0; i < 6; ++i) out->res[i] = s->res_old[i];
    out->evaluations = s->evals; out->repeat = s->repeat; out->num_warped = num_warped;
    __threadfence_system();
    out->iterations = s->iter;                       // written last: the host polls it (>= 0 means finished)
    __threadfence_system();
    ctx->done = 1;
    return;
  }
  lm_next_candidate(c, s, pose);
}

}  // namespace

// ================================================================================================
class TrackerImpl final : public TrackerIface {
 public:
  TrackerImpl(int w, int h, float huber, float cutoff, int n_max, int device)
      : w_(w), h_(h), huber_(huber), coarse_cutoff_(cutoff), device_(device) {
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0)
      throw Error("tandem_b200: no CUDA device visible - this library has no CPU fallback");
    if (w * h == 0) throw Error("CudaCoarseTracker::init has w*h==0 (cuda_coarse_tracker.cpp:105)");
    n_max_ = n_max > 0 ? n_max : w * h;
    TDM_CUDA(cudaSetDevice(device_));
    int lo, hi;
    TDM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TDM_CUDA(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, hi));  // highest priority (cpp:72-74)
    TDM_CUDA(cudaMalloc(&d_pc_, (size_t)4 * n_max_ * 4));
    TDM_CUDA(cudaMalloc(&d_warped_, (size_t)7 * n_max_ * 4));
    TDM_CUDA(cudaMemset(d_warped_, 0, (size_t)7 * n_max_ * 4));
    TDM_CUDA(cudaMalloc(&d_dI_, (size_t)3 * w * h * 4));
    TDM_CUDA(cudaMalloc(&d_partials_, (size_t)kBlocks * kOut * 8));
    TDM_CUDA(cudaMalloc(&d_ticket_, 4));
    TDM_CUDA(cudaMemset(d_ticket_, 0, 4));
    TDM_CUDA(cudaMallocHost(&h_pc_, (size_t)4 * n_max_ * 4));
    TDM_CUDA(cudaMallocHost(&h_dI_, (size_t)3 * w * h * 4));
    TDM_CUDA(cudaHostAlloc(&h_out_, kOut * 8, cudaHostAllocMapped));
    TDM_CUDA(cudaHostGetDevicePointer(&d_out_, h_out_, 0));
  }
  ~TrackerImpl() override {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    cudaFree(d_pc_); cudaFree(d_warped_); cudaFree(d_dI_); cudaFree(d_partials_); cudaFree(d_ticket_);
    cudaFreeHost(h_pc_); cudaFreeHost(h_dI_); cudaFreeHost(h_out_);
    if (d_batch_) { cudaFree(d_batch_); cudaFreeHost(h_batch_); cudaFree(d_batch_partials_); cudaFree(d_batch_ticket_); cudaFreeHost(h_batch_out_); }
    cudaFree(d_proj_); cudaFree(d_depth_); cudaFree(d_idepth0_); cudaFree(d_gray_); cudaFree(d_rowcnt_); cudaFree(d_rowoff_);
    cudaFree(d_total_); cudaFree(d_lm_ctx_); cudaFree(d_lm_out_); cudaFree(d_lm_in_);
    if (h_front_) cudaFreeHost(h_front_);
    if (h_total_) cudaFreeHost(h_total_);
    if (h_lm_in_) cudaFreeHost(h_lm_in_);
    if (h_lm_res_) cudaFreeHost(h_lm_res_);
    cudaStreamDestroy(stream_);
  }

  void set_k(int w, int h, float fx, float fy, float cx, float cy) override {
    if (w != w_ || h != h_) throw Error("CudaCoarseTracker::setK wrong h,w. (cuda_coarse_tracker.cpp:359)");
    fx_ = fx; fy_ = fy; cx_ = cx; cy_ = cy;
    // Ki = K^-1 in double, used as float (cpp:365-371, :201-204)
    const double Ki[9] = {1.0 / fx, 0, -(double)cx / fx, 0, 1.0 / fy, -(double)cy / fy, 0, 0, 1};
    for (int i = 0; i < 9; ++i) Ki_[i] = (float)Ki[i];
    have_k_ = true;
  }

  void set_reference(int n, const float* u, const float* v, const float* idepth, const float* color, float ref_exposure,
                     const double ref_aff[2]) override {
    if (n > n_max_) throw Error("Called CudaCoarseTracker::setReference with n > n_max points. (cuda_coarse_tracker.cpp:82)");
    TDM_CUDA(cudaSetDevice(device_));
    TDM_CUDA(cudaStreamSynchronize(stream_));  // pinned staging may still be in flight
    n_ = n;
    const size_t nm = (size_t)n_max_;
    std::memcpy(h_pc_, u, 4 * (size_t)n);
    std::memcpy(h_pc_ + nm, v, 4 * (size_t)n);
    std::memcpy(h_pc_ + 2 * nm, idepth, 4 * (size_t)n);
    std::memcpy(h_pc_ + 3 * nm, color, 4 * (size_t)n);
    for (int a = 0; a < 4; ++a)
      TDM_CUDA(cudaMemcpyAsync(d_pc_ + a * nm, h_pc_ + a * nm, 4 * (size_t)n, cudaMemcpyHostToDevice, stream_));
    ref_exposure_ = ref_exposure;
    ref_aff_[0] = ref_aff[0]; ref_aff_[1] = ref_aff[1];
    have_params_ = false;   // a calcG now needs a calcRes against the new reference first
  }

  void set_new(const float* dI) override {
    TDM_CUDA(cudaSetDevice(device_));
    TDM_CUDA(cudaStreamSynchronize(stream_));
    std::memcpy(h_dI_, dI, (size_t)3 * w_ * h_ * 4);
    TDM_CUDA(cudaMemcpyAsync(d_dI_, h_dI_, (size_t)3 * w_ * h_ * 4, cudaMemcpyHostToDevice, stream_));
    have_new_ = true;
    have_params_ = false;   // calcG is only defined on the buffers of a calcRes against THIS image (cuda_coarse_tracker.cpp:277-281)
  }

  void calc_res(const double* refToNew, float new_exposure, const double aff[2], float cutoffTH, double res6[6]) override {
    prepare(refToNew, new_exposure, aff, cutoffTH);
    launch(0);
    TDM_CUDA(cudaStreamSynchronize(stream_));
    finish_res(res6);
  }
  void calc_g(float new_exposure, const double aff[2], double H[64], double b[8]) override {
    TDM_CHECK(have_params_, "calcG before calcRes (or the reference / new image changed since the last calcRes)");
    set_aff(new_exposure, aff);
    launch(1);
    TDM_CUDA(cudaStreamSynchronize(stream_));
    finish_g(H, b);
  }
  void calc_res_g(const double* refToNew, float new_exposure, const double aff[2], float cutoffTH, double res6[6],
                  double H[64], double b[8]) override {
    prepare(refToNew, new_exposure, aff, cutoffTH);
    launch(2);
    TDM_CUDA(cudaStreamSynchronize(stream_));
    finish_res(res6);
    finish_g(H, b);
  }

  // ---------------------------------------------------------------------------------------------- n1
  int set_reference_dense(const DenseRefArgs& a) override {
    TDM_CHECK(have_k_, "setReferenceDense before setK");
    TDM_CHECK(a.depth && a.T_depth_to_ref && a.ref_gray && a.ref_aff, "null argument");
    TDM_CHECK(a.step >= 1 && a.n_sparse >= 0, "bad tracking step / sparse count");
    TDM_CHECK(a.n_sparse == 0 || (a.pc_u && a.pc_v && a.pc_idepth && a.pc_color), "sparse arrays missing");
    TDM_CUDA(cudaSetDevice(device_));
    TDM_CUDA(cudaStreamSynchronize(stream_));
    const size_t npx = (size_t)w_ * h_, nm = (size_t)n_max_;
    ensure_front_buffers();
    // sparse points [0, n_sparse] (the last one is the reference's stale slot, CoarseTracker.cpp:717-722)
    const int ns1 = a.n_sparse + 1;
    if ((long long)ns1 > n_max_) throw Error("setReferenceDense: more sparse points than n_max");
    const float* src[4] = {a.pc_u, a.pc_v, a.pc_idepth, a.pc_color};
    for (int k = 0; k < 4; ++k) {
      if (src[k]) {
        std::memcpy(h_pc_ + k * nm, src[k], 4 * (size_t)ns1);
        TDM_CUDA(cudaMemcpyAsync(d_pc_ + k * nm, h_pc_ + k * nm, 4 * (size_t)ns1, cudaMemcpyHostToDevice, stream_));
      } else {
        TDM_CUDA(cudaMemsetAsync(d_pc_ + k * nm, 0, 4, stream_));
      }
    }
    const float* d_depth = a.depth;
    if (!a.depth_on_device) {
      std::memcpy(h_front_, a.depth, 4 * npx);
      TDM_CUDA(cudaMemcpyAsync(d_depth_, h_front_, 4 * npx, cudaMemcpyHostToDevice, stream_));
      d_depth = d_depth_;
    } else if (a.depth_ready) {
      TDM_CUDA(cudaStreamWaitEvent(stream_, (cudaEvent_t)a.depth_ready, 0));
    }
    const float* d_id0 = nullptr;
    if (a.idepth0) {
      std::memcpy(h_front_ + npx, a.idepth0, 4 * npx);
      TDM_CUDA(cudaMemcpyAsync(d_idepth0_, h_front_ + npx, 4 * npx, cudaMemcpyHostToDevice, stream_));
      d_id0 = d_idepth0_;
    }
    const float* d_gray = a.ref_gray;
    int gstride = a.gray_stride;
    if (!a.gray_on_device) {
      std::memcpy(h_front_ + 2 * npx, a.ref_gray, 4 * npx);
      TDM_CUDA(cudaMemcpyAsync(d_gray_, h_front_ + 2 * npx, 4 * npx, cudaMemcpyHostToDevice, stream_));
      d_gray = d_gray_;
      gstride = 1;
    } else if (a.gray_ready) {
      TDM_CUDA(cudaStreamWaitEvent(stream_, (cudaEvent_t)a.gray_ready, 0));
    }
    // KRKi = (K R) Ki, Kt = K t in float, k = 0,1,2 (CoarseTracker.cpp:676-677)
    DenseWarp m;
    {
      const float K[9] = {fx_, 0, cx_, 0, fy_, cy_, 0, 0, 1};
      const float Ki[9] = {1.0f / fx_, 0, -cx_ / fx_, 0, 1.0f / fy_, -cy_ / fy_, 0, 0, 1};
      float R[9], t[3], KR[9];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = (float)a.T_depth_to_ref[4 * r + c];
        t[r] = (float)a.T_depth_to_ref[4 * r + 3];
      }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          volatile float s = 0;
          for (int k = 0; k < 3; ++k) { volatile float q = K[3 * r + k] * R[3 * k + c]; s = s + q; }
          KR[3 * r + c] = s;
        }
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) {
          volatile float s = 0;
          for (int k = 0; k < 3; ++k) { volatile float q = KR[3 * r + k] * Ki[3 * k + c]; s = s + q; }
          m.KRKi[3 * r + c] = s;
        }
        volatile float s = 0;
        for (int k = 0; k < 3; ++k) { volatile float q = K[3 * r + k] * t[k]; s = s + q; }
        m.Kt[r] = s;
      }
    }
    TDM_CUDA(cudaMemsetAsync(d_proj_, 0xFF, 4 * npx, stream_));
    const int nx = (w_ + a.step - 1) / a.step, ny = (h_ + a.step - 1) / a.step;
    k_dense_warp<<<(nx * ny + 255) / 256, 256, 0, stream_>>>(d_depth, w_, h_, a.step, m, d_proj_);
    k_dense_rowcount<<<h_, 256, 0, stream_>>>(d_proj_, d_id0, a.dense_only, w_, h_, d_rowcnt_);
    k_dense_rowscan<<<1, 32, 0, stream_>>>(d_rowcnt_, h_, d_rowoff_, d_total_);
    TDM_CUDA(cudaMemcpyAsync(h_total_, d_total_, 4, cudaMemcpyDeviceToHost, stream_));
    TDM_CUDA(cudaGetLastError());
    TDM_CUDA(cudaStreamSynchronize(stream_));
    const int total = *h_total_;
    if ((long long)a.n_sparse + 1 + total > n_max_) throw Error("setReferenceDense: n_max too small for the dense points");
    k_dense_emit<<<h_, 256, 0, stream_>>>(d_proj_, d_id0, a.dense_only, w_, h_, d_rowoff_, d_gray, gstride, a.n_sparse + 1,
                                          d_pc_, d_pc_ + nm, d_pc_ + 2 * nm, d_pc_ + 3 * nm);
    TDM_CUDA(cudaGetLastError());
    n_ = a.n_sparse + total;      // the pre-increment quirk: the last appended point is not counted
    ref_exposure_ = a.ref_exposure;
    ref_aff_[0] = a.ref_aff[0]; ref_aff_[1] = a.ref_aff[1];
    have_params_ = false;
    return n_;
  }
  int width() const override { return w_; }
  int height() const override { return h_; }
  int device() const override { return device_; }
  void get_reference(int n, float* u, float* v, float* idepth, float* color) override {
    TDM_CHECK(n >= 0 && n <= n_max_, "get_reference: bad n");
    TDM_CUDA(cudaSetDevice(device_));
    const size_t nm = (size_t)n_max_;
    float* dst[4] = {u, v, idepth, color};
    for (int k = 0; k < 4; ++k)
      if (dst[k]) TDM_CUDA(cudaMemcpyAsync(dst[k], d_pc_ + k * nm, 4 * (size_t)n, cudaMemcpyDeviceToHost, stream_));
    TDM_CUDA(cudaStreamSynchronize(stream_));
  }
  // ---------------------------------------------------------------------------------------------- n2
  void set_new_device(const float* d_dI, void* ready_event) override {
    TDM_CUDA(cudaSetDevice(device_));
    if (ready_event) TDM_CUDA(cudaStreamWaitEvent(stream_, (cudaEvent_t)ready_event, 0));
    TDM_CUDA(cudaMemcpyAsync(d_dI_, d_dI, (size_t)3 * w_ * h_ * 4, cudaMemcpyDeviceToDevice, stream_));
    have_new_ = true;
    have_params_ = false;   // calcG is only defined on the buffers of a calcRes against THIS image (cuda_coarse_tracker.cpp:277-281)
  }
  // ---------------------------------------------------------------------------------------------- n3
  void track(const TrackArgs& a, TrackResult* r) override {
    TDM_CHECK(have_k_, "track before setK");
    TDM_CHECK(have_new_, "track before setNew");
    TDM_CHECK(a.refToNew && a.aff && r, "null argument");
    TDM_CHECK(a.max_iterations >= 0 && a.max_iterations <= 200, "max_iterations out of range");
    TDM_CUDA(cudaSetDevice(device_));
    ensure_lm_buffers();
    p_.w = w_; p_.h = h_; p_.n = n_;
    p_.fx = fx_; p_.fy = fy_; p_.cx = cx_; p_.cy = cy_;
    for (int i = 0; i < 9; ++i) p_.Ki[i] = Ki_[i];
    p_.huber = huber_; p_.ref_b = (float)ref_aff_[1];
    LmConfig c;
    for (int i = 0; i < 9; ++i) c.Ki[i] = Ki_[i];
    c.ref_exposure = ref_exposure_; c.new_exposure = a.new_exposure;
    c.ref_aff[0] = ref_aff_[0]; c.ref_aff[1] = ref_aff_[1];
    c.huber = huber_; c.coarse_cutoff = a.coarse_cutoff; c.max_iter = a.max_iterations;
    c.lambda_limit = a.lambda_extrapolation_limit; c.fix_a = a.fix_a; c.fix_b = a.fix_b;
    TDM_CUDA(cudaStreamSynchronize(stream_));      // h_lm_in_ may still be read by a previous call
    for (int i = 0; i < 16; ++i) h_lm_in_[i] = a.refToNew[i];
    h_lm_in_[16] = a.aff[0]; h_lm_in_[17] = a.aff[1];
    h_lm_res_->iterations = -1;
    cudaEvent_t e0, e1;
    TDM_CUDA(cudaEventCreate(&e0)); TDM_CUDA(cudaEventCreate(&e1));
    TDM_CUDA(cudaEventRecord(e0, stream_));
    TDM_CUDA(cudaMemcpyAsync(d_lm_in_, h_lm_in_, 18 * 8, cudaMemcpyHostToDevice, stream_));
    k_lm_init<<<1, 32, 0, stream_>>>(d_lm_ctx_, c, d_lm_res_, d_lm_in_);
    // Evaluations are enqueued in batches of kBatch with no host involvement inside a batch (the LM step runs in the
    // evaluation kernel's last CTA; a converged loop turns the rest of a batch into empty launches).  Between batches
    // the host looks at the mapped result once.  Worst case: initial + 6 cutoff doublings (2^6 > 50) + one per iteration.
    constexpr int kBatch = 4;
    const int max_evals = 1 + 6 + a.max_iterations;
    const volatile LmResult* o = h_lm_res_;
    for (int e = 0; e < max_evals && o->iterations < 0; e += kBatch) {
      for (int k = 0; k < kBatch && e + k < max_evals; ++k) launch(2, true);
      TDM_CUDA(cudaStreamSynchronize(stream_));
    }
    TDM_CUDA(cudaEventRecord(e1, stream_));
    TDM_CUDA(cudaGetLastError());
    TDM_CUDA(cudaEventSynchronize(e1));
    TDM_CUDA(cudaEventElapsedTime(&r->device_ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (o->iterations < 0) throw Error("track: the device loop did not finish (internal error)");
    for (int i = 0; i < 16; ++i) r->refToNew[i] = o->T[i];
    r->aff[0] = o->aff[0]; r->aff[1] = o->aff[1];
    for (int i = 0; i < 6; ++i) r->res[i] = o->res[i];
    r->iterations = o->iterations; r->evaluations = o->evaluations; r->cutoff_repeat = o->repeat;
    num_warped_ = o->num_warped;
  }
  void synchronize() override {
    TDM_CUDA(cudaSetDevice(device_));
    TDM_CUDA(cudaStreamSynchronize(stream_));
  }
  void run_resident(int iters, float* ms) override {
    TDM_CHECK(have_params_, "run_resident before calcRes");
    TDM_CUDA(cudaSetDevice(device_));
    cudaEvent_t e0, e1;
    TDM_CUDA(cudaEventCreate(&e0)); TDM_CUDA(cudaEventCreate(&e1));
    TDM_CUDA(cudaEventRecord(e0, stream_));
    for (int i = 0; i < iters; ++i) launch(2);
    TDM_CUDA(cudaEventRecord(e1, stream_));
    TDM_CUDA(cudaEventSynchronize(e1));
    TDM_CUDA(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }

 private:
  void ensure_front_buffers() {
    if (d_proj_) return;
    const size_t npx = (size_t)w_ * h_;
    TDM_CUDA(cudaMalloc(&d_proj_, 4 * npx));
    TDM_CUDA(cudaMalloc(&d_depth_, 4 * npx));
    TDM_CUDA(cudaMalloc(&d_idepth0_, 4 * npx));
    TDM_CUDA(cudaMalloc(&d_gray_, 4 * npx));
    TDM_CUDA(cudaMalloc(&d_rowcnt_, 4 * (size_t)h_));
    TDM_CUDA(cudaMalloc(&d_rowoff_, 4 * (size_t)h_));
    TDM_CUDA(cudaMalloc(&d_total_, 4));
    TDM_CUDA(cudaMallocHost(&h_front_, 3 * 4 * npx));
    TDM_CUDA(cudaMallocHost(&h_total_, 4));
  }
  void ensure_lm_buffers() {
    if (d_lm_ctx_) return;
    TDM_CUDA(cudaMalloc(&d_lm_ctx_, sizeof(LmCtx)));
    TDM_CUDA(cudaMalloc(&d_lm_out_, kOut * 8));
    TDM_CUDA(cudaMalloc(&d_lm_in_, 18 * 8));
    TDM_CUDA(cudaMallocHost(&h_lm_in_, 18 * 8));
    TDM_CUDA(cudaHostAlloc(&h_lm_res_, sizeof(LmResult), cudaHostAllocMapped));
    TDM_CUDA(cudaHostGetDevicePointer(&d_lm_res_, h_lm_res_, 0));
  }
  void set_aff(float new_exposure, const double aff[2]) {  // AffLight::fromToVecExposure, cpp:42-52
    float eF = ref_exposure_, eT = new_exposure;
    if (eF == 0 || eT == 0) eT = eF = 1;
    const double a = std::exp(aff[0] - ref_aff_[0]) * eT / eF;
    const double b = aff[1] - a * ref_aff_[1];
    p_.affa = (float)a; p_.affb = (float)b; p_.ref_b = (float)ref_aff_[1];
  }
  void prepare(const double* refToNew, float new_exposure, const double aff[2], float cutoffTH) {
    TDM_CHECK(have_k_, "calcRes before setK");
    TDM_CHECK(have_new_, "calcRes before setNew");
    TDM_CUDA(cudaSetDevice(device_));
    p_.w = w_; p_.h = h_; p_.n = n_;
    p_.fx = fx_; p_.fy = fy_; p_.cx = cx_; p_.cy = cy_;
    float R[9];
    for (int r = 0; r < 3; ++r) {
      for (int k = 0; k < 3; ++k) R[3 * r + k] = (float)refToNew[4 * r + k];
      p_.t[r] = (float)refToNew[4 * r + 3];
    }
    for (int i = 0; i < 9; ++i) p_.Ki[i] = Ki_[i];
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q) {
        volatile float s = 0;
        for (int k = 0; k < 3; ++k) { volatile float m = R[3 * r + k] * Ki_[3 * k + q]; s = s + m; }
        p_.RKi[3 * r + q] = s;
      }
    set_aff(new_exposure, aff);
    p_.huber = huber_; p_.cutoff = cutoffTH;
    p_.maxEnergy = 2 * huber_ * cutoffTH - huber_ * huber_;
    have_params_ = true;
  }
  void launch(int mode, bool dev_loop = false) {
    TDM_CUDA(cudaSetDevice(device_));   // another handle of this thread may have switched the current device
    TrkBufs b;
    const size_t nm = (size_t)n_max_;
    b.lm = dev_loop ? d_lm_ctx_ : nullptr;
    b.pc_u = d_pc_; b.pc_v = d_pc_ + nm; b.pc_idepth = d_pc_ + 2 * nm; b.pc_color = d_pc_ + 3 * nm;
    b.dInew = d_dI_; b.warped = d_warped_; b.n_max = n_max_;
    b.partials = d_partials_; b.ticket = d_ticket_; b.out = dev_loop ? d_lm_out_ : d_out_; b.batch = nullptr;
    if (mode == 0) k_tracker<0><<<kBlocks, kThreads, 0, stream_>>>(p_, b);
    else if (mode == 1) k_tracker<1><<<kBlocks, kThreads, 0, stream_>>>(p_, b);
    else k_tracker<2><<<kBlocks, kThreads, 0, stream_>>>(p_, b);
    TDM_CUDA(cudaGetLastError());
  }
  // Batched calcRes over n_hyp poses (motion hypotheses of FullSystem::trackNewCoarse, FullSystem.cpp:437-530) in ONE launch and
  // one synchronisation: res6 of hypothesis k equals calc_res(refToNew + 16 k, ..., aff + 2 k, ...) bit for bit.  Leaves the
  // buffers of the last calcRes untouched (calcG still refers to them).
  void calc_res_batch(int n_hyp, const double* refToNew, float new_exposure, const double* aff, float cutoffTH, double* res6) override {
    TDM_CHECK(n_hyp >= 1 && n_hyp <= kMaxHyp, "calcResBatch: 1..64 hypotheses");
    TDM_CHECK(have_k_ && have_new_, "calcResBatch before setK / setNew");
    TDM_CUDA(cudaSetDevice(device_));
    if (!d_batch_) {
      TDM_CUDA(cudaMalloc(&d_batch_, kMaxHyp * sizeof(TrkPose)));
      TDM_CUDA(cudaMallocHost(&h_batch_, kMaxHyp * sizeof(TrkPose)));
      TDM_CUDA(cudaMalloc(&d_batch_partials_, (size_t)kMaxHyp * kBlocks * kOut * 8));
      TDM_CUDA(cudaMalloc(&d_batch_ticket_, kMaxHyp * 4));
      TDM_CUDA(cudaMemset(d_batch_ticket_, 0, kMaxHyp * 4));
      TDM_CUDA(cudaHostAlloc(&h_batch_out_, (size_t)kMaxHyp * kOut * 8, cudaHostAllocMapped));
      TDM_CUDA(cudaHostGetDevicePointer(&d_batch_out_, h_batch_out_, 0));
    }
    TDM_CUDA(cudaStreamSynchronize(stream_));   // the pinned pose staging may still be in flight from the previous batch
    const TrkParams saved = p_;
    const bool saved_have = have_params_;
    for (int k = 0; k < n_hyp; ++k) {
      prepare(refToNew + 16 * k, new_exposure, aff + 2 * k, cutoffTH);   // the exact host arithmetic of calcRes, per pose
      TrkPose& q = h_batch_[k];
      for (int i = 0; i < 9; ++i) q.RKi[i] = p_.RKi[i];
      for (int i = 0; i < 3; ++i) q.t[i] = p_.t[i];
      q.affa = p_.affa; q.affb = p_.affb; q.cutoff = p_.cutoff; q.maxEnergy = p_.maxEnergy;
    }
    TrkParams pb = p_;
    p_ = saved;                                   // calcG keeps referring to the last plain calcRes
    have_params_ = saved_have;
    TDM_CUDA(cudaMemcpyAsync(d_batch_, h_batch_, n_hyp * sizeof(TrkPose), cudaMemcpyHostToDevice, stream_));
    TrkBufs b;
    const size_t nm = (size_t)n_max_;
    b.lm = nullptr;
    b.pc_u = d_pc_; b.pc_v = d_pc_ + nm; b.pc_idepth = d_pc_ + 2 * nm; b.pc_color = d_pc_ + 3 * nm;
    b.dInew = d_dI_; b.warped = nullptr; b.n_max = n_max_;
    b.partials = d_batch_partials_; b.ticket = d_batch_ticket_; b.out = d_batch_out_; b.batch = d_batch_;
    k_tracker<3><<<dim3(kBlocks, n_hyp), kThreads, 0, stream_>>>(pb, b);
    TDM_CUDA(cudaGetLastError());
    TDM_CUDA(cudaStreamSynchronize(stream_));
    for (int k = 0; k < n_hyp; ++k) {
      const volatile double* o = h_batch_out_ + (size_t)k * kOut;
      double* r = res6 + 6 * k;
      r[0] = o[0]; r[1] = o[1]; r[2] = o[4] / o[6]; r[3] = 0; r[4] = o[5] / o[6]; r[5] = o[3] / o[1];
    }
  }
  void finish_res(double res6[6]) {  // cpp:264-272
    const volatile double* o = h_out_;
    res6[0] = o[0]; res6[1] = o[1];
    res6[2] = o[4] / o[6]; res6[3] = 0; res6[4] = o[5] / o[6];
    res6[5] = o[3] / o[1];
    num_warped_ = (int)o[2];
  }
  void finish_g(double H[64], double b[8]) {  // cpp:335-355
    const volatile double* o = h_out_ + kRes;
    const double factor = 1.0 / num_warped_;
    static const double scale[8] = {1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 10.0, 1000.0};
    for (int r = 0; r < 8; ++r) {
      for (int q = 0; q < 8; ++q) {
        const int lo = r < q ? r : q, hi = r < q ? q : r;
        H[8 * r + q] = o[lo * 9 + hi - lo * (lo + 1) / 2] * factor * scale[r] * scale[q];
      }
      b[r] = o[r * 9 + 8 - r * (r + 1) / 2] * factor * scale[r];
    }
  }

  int w_, h_;
  float huber_, coarse_cutoff_;
  int device_;
  int n_max_ = 0, n_ = 0;
  float fx_ = 0, fy_ = 0, cx_ = 0, cy_ = 0, Ki_[9];
  bool have_k_ = false, have_new_ = false, have_params_ = false;
  float ref_exposure_ = 1.f;
  double ref_aff_[2] = {0, 0};
  cudaStream_t stream_ = nullptr;
  float *d_pc_ = nullptr, *d_warped_ = nullptr, *d_dI_ = nullptr, *h_pc_ = nullptr, *h_dI_ = nullptr;
  static constexpr int kMaxHyp = 64;
  TrkPose *d_batch_ = nullptr, *h_batch_ = nullptr;
  double *d_batch_partials_ = nullptr, *h_batch_out_ = nullptr, *d_batch_out_ = nullptr;
  unsigned* d_batch_ticket_ = nullptr;
  double *d_partials_ = nullptr, *h_out_ = nullptr, *d_out_ = nullptr;
  unsigned* d_ticket_ = nullptr;
  TrkParams p_{};
  int num_warped_ = 0;
  // n1 buffers
  unsigned* d_proj_ = nullptr;
  float *d_depth_ = nullptr, *d_idepth0_ = nullptr, *d_gray_ = nullptr, *h_front_ = nullptr;
  int *d_rowcnt_ = nullptr, *d_rowoff_ = nullptr, *d_total_ = nullptr, *h_total_ = nullptr;
  // n3 buffers
  LmCtx* d_lm_ctx_ = nullptr;
  double *d_lm_out_ = nullptr, *d_lm_in_ = nullptr, *h_lm_in_ = nullptr;
  LmResult *h_lm_res_ = nullptr, *d_lm_res_ = nullptr;
};


// ================================================================================================
class PyramidImpl final : public PyramidIface {
 public:
  PyramidImpl(int w, int h, int levels, int device) : w_(w), h_(h), levels_(levels), device_(device) {
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0)
      throw Error("tandem_b200: no CUDA device visible - this library has no CPU fallback");
    TDM_CHECK(w > 0 && h > 0 && levels >= 1 && levels <= 8 && (w >> (levels - 1)) >= 3 && (h >> (levels - 1)) >= 3, "bad pyramid shape");
    TDM_CUDA(cudaSetDevice(device_));
    TDM_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    TDM_CUDA(cudaEventCreateWithFlags(&ready_, cudaEventDisableTiming));
    size_t tot = 0;
    for (int l = 0; l < levels; ++l) { off_.push_back(tot); tot += (size_t)(w >> l) * (h >> l); }
    TDM_CUDA(cudaMalloc(&d_dI_, 3 * 4 * tot));
    TDM_CUDA(cudaMalloc(&d_abs_, 4 * tot));
    TDM_CUDA(cudaMalloc(&d_gray_, 4 * (size_t)w * h));
    TDM_CUDA(cudaMalloc(&d_in_, 4 * (size_t)w * h));
    TDM_CUDA(cudaMallocHost(&h_in_, 4 * (size_t)w * h));
  }
  ~PyramidImpl() override {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    cudaFree(d_dI_); cudaFree(d_abs_); cudaFree(d_gray_); cudaFree(d_in_); cudaFreeHost(h_in_);
    cudaEventDestroy(ready_); cudaStreamDestroy(stream_);
  }
  void build(const float* gray_host) override {
    TDM_CHECK(gray_host, "null argument");
    TDM_CUDA(cudaSetDevice(device_));
    TDM_CUDA(cudaStreamSynchronize(stream_));
    std::memcpy(h_in_, gray_host, 4 * (size_t)w_ * h_);
    TDM_CUDA(cudaMemcpyAsync(d_in_, h_in_, 4 * (size_t)w_ * h_, cudaMemcpyHostToDevice, stream_));
    for (int l = 0; l < levels_; ++l) {
      const int wl = w_ >> l, hl = h_ >> l, n = wl * hl;
      const float* src = l == 0 ? d_in_ : d_dI_ + 3 * off_[l - 1];
      k_pyr_level<<<(n + 255) / 256, 256, 0, stream_>>>(src, l > 0, l > 0 ? (w_ >> (l - 1)) : 0, d_gray_, wl, hl);
      k_pyr_grad<<<(n + 255) / 256, 256, 0, stream_>>>(d_gray_, d_dI_ + 3 * off_[l], d_abs_ + off_[l], wl, hl);
    }
    TDM_CUDA(cudaGetLastError());
    TDM_CUDA(cudaEventRecord(ready_, stream_));
  }
  void get_level(int lvl, float* dI_host, float* absgrad_host) override {
    TDM_CHECK(lvl >= 0 && lvl < levels_, "bad level");
    TDM_CUDA(cudaSetDevice(device_));
    const size_t n = (size_t)(w_ >> lvl) * (h_ >> lvl);
    if (dI_host) TDM_CUDA(cudaMemcpyAsync(dI_host, d_dI_ + 3 * off_[lvl], 12 * n, cudaMemcpyDeviceToHost, stream_));
    if (absgrad_host) TDM_CUDA(cudaMemcpyAsync(absgrad_host, d_abs_ + off_[lvl], 4 * n, cudaMemcpyDeviceToHost, stream_));
    TDM_CUDA(cudaStreamSynchronize(stream_));
  }
  const float* level_dI(int lvl) const override { return d_dI_ + 3 * off_.at(lvl); }
  void* ready_event() const override { return (void*)ready_; }
  int width(int lvl) const override { return w_ >> lvl; }
  int height(int lvl) const override { return h_ >> lvl; }
  int levels() const override { return levels_; }
  int device() const override { return device_; }

 private:
  int w_, h_, levels_, device_;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ready_ = nullptr;
  std::vector<size_t> off_;
  float *d_dI_ = nullptr, *d_abs_ = nullptr, *d_gray_ = nullptr, *d_in_ = nullptr, *h_in_ = nullptr;
};

PyramidIface* make_pyramid(int w, int h, int levels, int device) { return new PyramidImpl(w, h, levels, device); }

TrackerIface* make_tracker(int w, int h, float huber, float cutoff, int n_max, int device) {
  return new TrackerImpl(w, h, huber, cutoff, n_max, device);
}

}  // namespace tdm

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

struct TrkBufs {
  const float *pc_u, *pc_v, *pc_idepth, *pc_color, *dInew;
  float* warped;        // 7 x n_max: u, v, dx, dy, idepth, residual, weight
  long long n_max;
  double* partials;     // [kBlocks][kOut]
  unsigned* ticket;
  double* out;          // mapped pinned host memory, kOut doubles
};

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
template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_tracker(const __grid_constant__ TrkParams p, TrkBufs b) {
  constexpr int NV = MODE == 0 ? kRes : (MODE == 1 ? kG : kOut);
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
  }
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
  }

  void set_new(const float* dI) override {
    TDM_CUDA(cudaSetDevice(device_));
    TDM_CUDA(cudaStreamSynchronize(stream_));
    std::memcpy(h_dI_, dI, (size_t)3 * w_ * h_ * 4);
    TDM_CUDA(cudaMemcpyAsync(d_dI_, h_dI_, (size_t)3 * w_ * h_ * 4, cudaMemcpyHostToDevice, stream_));
    have_new_ = true;
  }

  void calc_res(const double* refToNew, float new_exposure, const double aff[2], float cutoffTH, double res6[6]) override {
    prepare(refToNew, new_exposure, aff, cutoffTH);
    launch(0);
    TDM_CUDA(cudaStreamSynchronize(stream_));
    finish_res(res6);
  }
  void calc_g(float new_exposure, const double aff[2], double H[64], double b[8]) override {
    TDM_CHECK(have_params_, "calcG before calcRes");
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
  void launch(int mode) {
    TrkBufs b;
    const size_t nm = (size_t)n_max_;
    b.pc_u = d_pc_; b.pc_v = d_pc_ + nm; b.pc_idepth = d_pc_ + 2 * nm; b.pc_color = d_pc_ + 3 * nm;
    b.dInew = d_dI_; b.warped = d_warped_; b.n_max = n_max_;
    b.partials = d_partials_; b.ticket = d_ticket_; b.out = d_out_;
    if (mode == 0) k_tracker<0><<<kBlocks, kThreads, 0, stream_>>>(p_, b);
    else if (mode == 1) k_tracker<1><<<kBlocks, kThreads, 0, stream_>>>(p_, b);
    else k_tracker<2><<<kBlocks, kThreads, 0, stream_>>>(p_, b);
    TDM_CUDA(cudaGetLastError());
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
  double *d_partials_ = nullptr, *h_out_ = nullptr, *d_out_ = nullptr;
  unsigned* d_ticket_ = nullptr;
  TrkParams p_{};
  int num_warped_ = 0;
};

TrackerIface* make_tracker(int w, int h, float huber, float cutoff, int n_max, int device) {
  return new TrackerImpl(w, h, huber, cutoff, n_max, device);
}

}  // namespace tdm

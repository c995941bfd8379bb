// Test infrastructure: thin driver around the REFERENCE's own tracker kernel launchers
// (callCalcResKernel / callCalcGKernel<float>, tandem/libdr/cuda_coarse_tracker/src/cuda_coarse_tracker_private.cu),
// compiled unmodified from /root/reference by oracle/ref_build.mk.  The reference's host wrapper
// (cuda_coarse_tracker.cpp) needs Eigen/Sophus/cnpy and cannot be built here, so this file restates only its
// argument marshalling (cpp:195-275, 277-356) around the real kernels.
#include <cuda_runtime.h>

#include <cstring>

#include "cuda_coarse_tracker_private.h"

extern "C" {

// all pointers are HOST pointers; returns 0 on success. outputs7 / outputs45 are the raw float reductions.
int ref_tracker_eval_timed(int w, int h, float fx, float fy, float cx, float cy, const float* refToNew16, const float* Ki9,
                           float affa, float affb, float ref_b, float huber, float cutoff, int n, const float* pc_u,
                           const float* pc_v, const float* pc_idepth, const float* pc_color, const float* dInew,
                           float* outputs7, float* outputs45, float* warped7n, int timed_iters, float* ms_res, float* ms_g);

int ref_tracker_eval(int w, int h, float fx, float fy, float cx, float cy, const float* refToNew16, const float* Ki9,
                     float affa, float affb, float ref_b, float huber, float cutoff, int n, const float* pc_u,
                     const float* pc_v, const float* pc_idepth, const float* pc_color, const float* dInew,
                     float* outputs7, float* outputs45, float* warped7n) {
  return ref_tracker_eval_timed(w, h, fx, fy, cx, cy, refToNew16, Ki9, affa, affb, ref_b, huber, cutoff, n, pc_u, pc_v, pc_idepth,
                                pc_color, dInew, outputs7, outputs45, warped7n, 0, nullptr, nullptr);
}

// timed_iters > 0: after the evaluation whose outputs are returned, the two reference kernels (+ the output memsets the
// reference's host wrapper issues per call, cuda_coarse_tracker.cpp:206-207,283) are launched timed_iters more times each,
// bracketed by CUDA events: *ms_res / *ms_g = average device time of one calcRes / calcG launch sequence on this GPU.
int ref_tracker_eval_timed(int w, int h, float fx, float fy, float cx, float cy, const float* refToNew16, const float* Ki9,
                           float affa, float affb, float ref_b, float huber, float cutoff, int n, const float* pc_u,
                           const float* pc_v, const float* pc_idepth, const float* pc_color, const float* dInew,
                           float* outputs7, float* outputs45, float* warped7n, int timed_iters, float* ms_res, float* ms_g) {
  float *d_pc, *d_dI, *d_w, *d_T, *d_o7, *d_o45;
  const size_t nb = (size_t)n * 4;
  if (cudaMalloc(&d_pc, 4 * nb) != cudaSuccess) return -1;
  cudaMalloc(&d_dI, (size_t)3 * w * h * 4);
  cudaMalloc(&d_w, 7 * nb);
  cudaMalloc(&d_T, 25 * 4);
  cudaMalloc(&d_o7, 7 * 4);
  cudaMalloc(&d_o45, 45 * 4);
  cudaMemcpy(d_pc, pc_u, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_pc + n, pc_v, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_pc + 2 * (size_t)n, pc_idepth, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_pc + 3 * (size_t)n, pc_color, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_dI, dInew, (size_t)3 * w * h * 4, cudaMemcpyHostToDevice);
  float T[25];
  std::memcpy(T, refToNew16, 64);
  std::memcpy(T + 16, Ki9, 36);
  cudaMemcpy(d_T, T, 100, cudaMemcpyHostToDevice);
  cudaMemset(d_o7, 0, 28);
  cudaMemset(d_o45, 0, 180);
  cudaStream_t s = 0;
  const float maxEnergy = 2 * huber * cutoff - huber * huber;
  float2 aff = make_float2(affa, affb);
  callCalcResKernel(128, s, huber, w, h, fx, fy, cx, cy, d_T, d_T + 16, aff, maxEnergy, cutoff, n, d_pc, d_pc + n,
                    d_pc + 2 * (size_t)n, d_pc + 3 * (size_t)n, d_dI, d_w, d_w + n, d_w + 2 * (size_t)n, d_w + 3 * (size_t)n,
                    d_w + 4 * (size_t)n, d_w + 5 * (size_t)n, d_w + 6 * (size_t)n, d_o7);
  callCalcGKernel<float>(128, s, fx, fy, aff, ref_b, n, 16, d_pc + 3 * (size_t)n, d_w, d_w + n, d_w + 2 * (size_t)n,
                         d_w + 3 * (size_t)n, d_w + 4 * (size_t)n, d_w + 5 * (size_t)n, d_w + 6 * (size_t)n, d_o45);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(outputs7, d_o7, 28, cudaMemcpyDeviceToHost);
  cudaMemcpy(outputs45, d_o45, 180, cudaMemcpyDeviceToHost);
  if (warped7n) cudaMemcpy(warped7n, d_w, 7 * nb, cudaMemcpyDeviceToHost);  // order u,v,dx,dy,idepth,residual,weight
  if (timed_iters > 0 && ms_res && ms_g && e == cudaSuccess) {
    float *d_s7, *d_s45;   // scratch outputs: the timed launches must not disturb the returned reductions
    cudaMalloc(&d_s7, 28);
    cudaMalloc(&d_s45, 180);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaEventRecord(e0, s);
    for (int it = 0; it < timed_iters; ++it) {
      cudaMemsetAsync(d_s7, 0, 28, s);
      callCalcResKernel(128, s, huber, w, h, fx, fy, cx, cy, d_T, d_T + 16, aff, maxEnergy, cutoff, n, d_pc, d_pc + n,
                        d_pc + 2 * (size_t)n, d_pc + 3 * (size_t)n, d_dI, d_w, d_w + n, d_w + 2 * (size_t)n, d_w + 3 * (size_t)n,
                        d_w + 4 * (size_t)n, d_w + 5 * (size_t)n, d_w + 6 * (size_t)n, d_s7);
    }
    cudaEventRecord(e1, s);
    for (int it = 0; it < timed_iters; ++it) {
      cudaMemsetAsync(d_s45, 0, 180, s);
      callCalcGKernel<float>(128, s, fx, fy, aff, ref_b, n, 16, d_pc + 3 * (size_t)n, d_w, d_w + n, d_w + 2 * (size_t)n,
                             d_w + 3 * (size_t)n, d_w + 4 * (size_t)n, d_w + 5 * (size_t)n, d_w + 6 * (size_t)n, d_s45);
    }
    cudaEventRecord(e2, s);
    e = cudaEventSynchronize(e2);
    cudaEventElapsedTime(ms_res, e0, e1);
    cudaEventElapsedTime(ms_g, e1, e2);
    *ms_res /= timed_iters;
    *ms_g /= timed_iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    cudaFree(d_s7); cudaFree(d_s45);
  }
  cudaFree(d_pc); cudaFree(d_dI); cudaFree(d_w); cudaFree(d_T); cudaFree(d_o7); cudaFree(d_o45);
  return e == cudaSuccess ? 0 : -2;
}

}  // extern "C"

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
int ref_tracker_eval(int w, int h, float fx, float fy, float cx, float cy, const float* refToNew16, const float* Ki9,
                     float affa, float affb, float ref_b, float huber, float cutoff, int n, const float* pc_u,
                     const float* pc_v, const float* pc_idepth, const float* pc_color, const float* dInew,
                     float* outputs7, float* outputs45, float* warped7n) {
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
  cudaFree(d_pc); cudaFree(d_dI); cudaFree(d_w); cudaFree(d_T); cudaFree(d_o7); cudaFree(d_o45);
  return e == cudaSuccess ? 0 : -2;
}

}  // extern "C"

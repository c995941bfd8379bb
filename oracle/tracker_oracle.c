/* ORACLE (test infrastructure - never linked into, or called by, the product library).
 *
 * Plain-C restatement of the dense photometric tracker's level-0 evaluation in TANDEM:
 *   calcRes   libdr/cuda_coarse_tracker/src/cuda_coarse_tracker_private.cu:41-214 (GPU semantics: warped
 *             buffers stay un-compacted, flow statistics divided by Num, not Num+0.1) and its CPU twin
 *             tandem/src/FullSystem/CoarseTracker.cpp:484-630
 *   calcG     cuda_coarse_tracker_private.cu:261-394, host scaling cuda_coarse_tracker.cpp:277-356,
 *             CPU twin CoarseTracker.cpp:378-481
 *   affLL     cuda_coarse_tracker.cpp:42-52 ; bilinear (I,dx,dy) fetch cuda_coarse_tracker_private.cu:22-38
 * PARITY UNPINNED: the only tracker golden in the reference (main.cu:185-189) needs the cct_data .npy dumps as inputs
 * that are not shipped, and neither the host wrapper (Eigen, Sophus, cnpy) nor CoarseTracker.cpp (Eigen)
 * compiles in the build container.  Per-point arithmetic is fp32 as in the kernels; the reductions are done
 * in double here (the reference uses fp32 block reductions + float atomics, which are run-to-run
 * non-deterministic), so H/b parity is stated as a relative tolerance, not bit equality.
 */
#include <math.h>
#include <string.h>

typedef struct {
  int w, h;
  float fx, fy, cx, cy;
  float huber;          /* setting_huberTH */
} TrkCfg;

static void bilinear33(const float* mat, float x, float y, int width, float out[3]) { /* cu:22-38 */
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  for (int c = 0; c < 3; ++c)
    out[c] = dxdy * bp[3 * (1 + width) + c] + (dy - dxdy) * bp[3 * width + c] + (dx - dxdy) * bp[3 + c] +
             (1.0f - dx - dy + dxdy) * bp[c];
}

/* exposures / affine brightness, cuda_coarse_tracker.cpp:42-52 (double) */
void tracker_oracle_affll(float ref_exposure, float new_exposure, const double ref_aff[2], const double new_aff[2],
                          float out2[2]) {
  float eF = ref_exposure, eT = new_exposure;
  if (eF == 0 || eT == 0) eT = eF = 1;
  double a = exp(new_aff[0] - ref_aff[0]) * eT / eF;
  double b = new_aff[1] - a * ref_aff[1];
  out2[0] = (float)a;
  out2[1] = (float)b;
}

/* warped: 7 arrays of n floats in the order u, v, dx, dy, idepth, residual, weight (zero at rejected points).
 * out7 (double): E, numTermsInE, numTermsInWarped, numSaturated, sumSquaredShiftT, sumSquaredShiftRT, shiftNum */
void tracker_oracle_calc_res(const TrkCfg* c, const double* refToNew /*4x4 row-major*/, const float affLL[2],
                             float cutoffTH, int n, const float* pc_u, const float* pc_v, const float* pc_idepth,
                             const float* pc_color, const float* dInew, float* warped, double out7[7]) {
  float R[9], t[3], Ki[9], RKi[9];
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) R[3 * r + k] = (float)refToNew[4 * r + k];
    t[r] = (float)refToNew[4 * r + 3];
  }
  /* Ki = K^-1 in double, then float (cuda_coarse_tracker.cpp:358-372, :201-204) */
  double Kid[9] = {1.0 / c->fx, 0, -(double)c->cx / c->fx, 0, 1.0 / c->fy, -(double)c->cy / c->fy, 0, 0, 1};
  for (int i = 0; i < 9; ++i) Ki[i] = (float)Kid[i];
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) {
      float s = 0;
      for (int k = 0; k < 3; ++k) s += R[3 * r + k] * Ki[3 * k + q];
      RKi[3 * r + q] = s;
    }
  const float maxEnergy = 2 * c->huber * cutoffTH - c->huber * c->huber;
  for (int k = 0; k < 7; ++k) out7[k] = 0;
  if (warped) memset(warped, 0, sizeof(float) * 7 * (size_t)n);
  const int w = c->w, h = c->h;
  for (int i = 0; i < n; ++i) {
    const float id = pc_idepth[i], x = pc_u[i], y = pc_v[i];
    float pt[3], ptK[3];
    for (int r = 0; r < 3; ++r) {
      pt[r] = RKi[3 * r] * x + RKi[3 * r + 1] * y + RKi[3 * r + 2] * 1.0f;
      ptK[r] = Ki[3 * r] * x + Ki[3 * r + 1] * y + Ki[3 * r + 2] * 1.0f;
    }
    float p1[3];
    for (int r = 0; r < 3; ++r) p1[r] = pt[r] + t[r] * id;
    const float u = p1[0] / p1[2], v = p1[1] / p1[2];
    const float Ku = c->fx * u + c->cx, Kv = c->fy * v + c->cy;
    const float nid = id / p1[2];
    if (i % 32 == 0) {
      float a[3], b[3], d[3];
      for (int r = 0; r < 3; ++r) { a[r] = ptK[r] + t[r] * id; b[r] = ptK[r] - t[r] * id; d[r] = pt[r] - t[r] * id; }
      const float KuT = c->fx * (a[0] / a[2]) + c->cx, KvT = c->fy * (a[1] / a[2]) + c->cy;
      const float KuT2 = c->fx * (b[0] / b[2]) + c->cx, KvT2 = c->fy * (b[1] / b[2]) + c->cy;
      const float Ku3 = c->fx * (d[0] / d[2]) + c->cx, Kv3 = c->fy * (d[1] / d[2]) + c->cy;
      float sT = (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      float sRT = (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      out7[4] += sT; out7[5] += sRT; out7[6] += 2.0;
    }
    if (Ku > 2 && Kv > 2 && Ku < w - 3 && Kv < h - 3 && nid > 0) {
      float hit[3];
      bilinear33(dInew, Ku, Kv, w, hit);
      if (isfinite(hit[0])) {
        const float res = hit[0] - (affLL[0] * pc_color[i] + affLL[1]);
        const float hw = fabsf(res) < c->huber ? 1.0f : c->huber / fabsf(res);
        if (fabsf(res) > cutoffTH) {
          out7[0] += maxEnergy; out7[1] += 1; out7[3] += 1;
        } else {
          out7[0] += hw * res * res * (2 - hw); out7[1] += 1; out7[2] += 1;
          if (warped) {
            warped[0 * (size_t)n + i] = u;  warped[1 * (size_t)n + i] = v;
            warped[2 * (size_t)n + i] = hit[1]; warped[3 * (size_t)n + i] = hit[2];
            warped[4 * (size_t)n + i] = nid; warped[5 * (size_t)n + i] = res; warped[6 * (size_t)n + i] = hw;
          }
        }
      }
    }
  }
}

/* host-side packaging of calcRes (cuda_coarse_tracker.cpp:264-272) */
void tracker_oracle_res6(const double out7[7], double res6[6]) {
  res6[0] = out7[0]; res6[1] = out7[1];
  res6[2] = out7[4] / out7[6]; res6[3] = 0; res6[4] = out7[5] / out7[6];
  res6[5] = out7[3] / out7[1];
}

/* calcG over the un-compacted warped buffers; acc45 in double; H 8x8 row-major, b 8 (scaled). */
void tracker_oracle_calc_g(const TrkCfg* c, const float affLL[2], float ref_aff_b, int n, const float* pc_color,
                           const float* warped, int num_terms_in_warped, double acc45[45], double H[64], double b[8]) {
  for (int k = 0; k < 45; ++k) acc45[k] = 0;
  const float a = affLL[0], b0 = ref_aff_b;
  const float *wu = warped, *wv = warped + (size_t)n, *wdx = warped + 2 * (size_t)n, *wdy = warped + 3 * (size_t)n,
              *wid = warped + 4 * (size_t)n, *wr = warped + 5 * (size_t)n, *ww = warped + 6 * (size_t)n;
  for (int i = 0; i < n; ++i) {
    const float dx = wdx[i] * c->fx, dy = wdy[i] * c->fy, u = wu[i], v = wv[i], id = wid[i];
    float J[9];
    J[0] = id * dx; J[1] = id * dy; J[2] = -id * (u * dx + v * dy);
    J[3] = -(u * v * dx + dy + dy * v * v); J[4] = u * v * dy + dx + dx * u * u; J[5] = u * dy - v * dx;
    J[6] = a * (b0 - pc_color[i]); J[7] = -1; J[8] = wr[i];
    const float w = ww[i];
    int k = 0;
    for (int j1 = 0; j1 < 9; ++j1) {
      const float jw = J[j1] * w;
      for (int j2 = j1; j2 < 9; ++j2) acc45[k++] += (double)(jw * J[j2]);
    }
  }
  const double factor = 1.0 / num_terms_in_warped;
  static const double scale[8] = {1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 10.0, 1000.0}; /* cpp:344-355 */
  for (int r = 0; r < 8; ++r) {
    for (int q = 0; q < 8; ++q) {
      const int lo = r < q ? r : q, hi = r < q ? q : r;
      H[8 * r + q] = acc45[lo * 9 + hi - lo * (lo + 1) / 2] * factor * scale[r] * scale[q];
    }
    b[r] = acc45[r * 9 + 8 - r * (r + 1) / 2] * factor * scale[r];
  }
}

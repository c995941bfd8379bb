// A/B variant of the view-aggregation cost-volume kernel (k_cost_volume_va16, mvsnet_kernels.cuh) with the source-view feature
// tile STAGED IN SHARED MEMORY BY TMA (north_star: "TMA staging of image / feature tiles"), for the
// stage where staging is most favourable: stage 3 (C = 8: one 16-byte vector per texel; D = 8 adaptive hypotheses a quarter of
// the base interval apart, so the 8 samples of a pixel stay within a few texels of each other along the epipolar line).
//
// A CTA owns a 16 x 8 pixel tile and all 8 hypotheses (256 threads = 128 pixels x 2 depth groups of ND = 4).  Per source view:
// every thread computes its 4 sample positions, the CTA reduces the bounding box of all texels its samples touch, one thread
// issues ONE tiled cp.async.bulk.tensor for that box (fixed box BW x BH texels, zero-filled outside the padded map = the
// zeros padding of grid_sample), the CTA waits on the mbarrier and gathers its 4 x 16-byte taps per sample from shared memory
// instead of L1.  A box larger than BW x BH (large parallax) falls back to the global gathers for that view (CTA-uniform).
// The arithmetic - and therefore the volume - is bit-identical to k_cost_volume_va16<TV, 8, 1, 4, ACC16>.
#pragma once
#include "conv_tc.cuh"

namespace tdm {

constexpr int kCvBW = 48, kCvBH = 24;   // staged box in texels (18 KB of shared memory)

template <typename TV, bool ACC16>
__global__ void __launch_bounds__(256)
k_cost_volume_va16_tma(const __grid_constant__ CUtensorMap tmap /*feat3: {8, Wp, Hp, V}, box {8, kCvBW, kCvBH, 1}*/,
                       P8<const __half> feats, const DminSrc dsrc, P8<TV> vol, int slot, int stage,
                       unsigned* __restrict__ stats /*[2]: views staged by TMA, views that fell back*/) {
  constexpr int C = 8, ND = 4;
  const CvParams& p = c_call_params[slot].cv[stage];
  constexpr float kS = 1.f / 64.f;
  __shared__ __align__(128) uint4 tile[kCvBH * kCvBW];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int red[4][8];
  __shared__ int box[5];   // bx0, by0, use_tma
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int lx = t & 15, ly = (t >> 4) & 7, dg = t >> 7;
  const int x = blockIdx.x * 16 + lx, y = blockIdx.y * 8 + ly;
  const bool active = x < p.W && y < p.H;
  const int xc = min(x, p.W - 1), yc = min(y, p.H - 1);
  const int pix = yc * p.W + xc;
  const int d0 = dg * ND;
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __half2 nref[C / 2], gwh[C / 2];
  {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(feats.p + feats.pos(0, yc, xc)));
    const __half2* h = reinterpret_cast<const __half2*>(&q);
    const __half2 ms = __float2half2_rn(-kS);
#pragma unroll
    for (int k = 0; k < 4; ++k) nref[k] = __hmul2(h[k], ms);
#pragma unroll
    for (int k = 0; k < C / 2; ++k) gwh[k] = __floats2half2_rn(p.gw1[2 * k] * p.gw1_scale, p.gw1[2 * k + 1] * p.gw1_scale);
  }
  float depth[ND];
  {
    const float dmn = p.hyp.adaptive ? dmin_px(dsrc, pix, xc, yc, c_call_params[slot].half_range[stage]) : 0.f;
#pragma unroll
    for (int k = 0; k < ND; ++k) depth[k] = hyp_value(p.hyp, dmn, min(d0 + k, p.D - 1));
  }
  float2 acc[ND][C / 2];
  __half2 acch[ND][C / 2];
#pragma unroll
  for (int k = 0; k < ND; ++k)
#pragma unroll
    for (int c = 0; c < C / 2; ++c) { acc[k][c] = make_float2(0.f, 0.f); acch[k][c] = __float2half2_rn(0.f); }
  const float fx = (float)xc, fy = (float)yc, Wf = (float)p.W, Hf = (float)p.H;
  const int row = feats.Wp * 8;
  __syncthreads();
  unsigned phase = 0;
  for (int s = 0; s < p.nsrc; ++s) {
    const float* R = p.rot[s];
    const float* tr = p.trans[s];
    const float rx = R[0] * fx + R[1] * fy + R[2];
    const float ry = R[3] * fx + R[4] * fy + R[5];
    const float rz = R[6] * fx + R[7] * fy + R[8];
    float six[ND], siy[ND];
    bool sok[ND];
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const float qx = fmaf(rx, depth[k], tr[0]), qy = fmaf(ry, depth[k], tr[1]), qz = fmaf(rz, depth[k], tr[2]);
      const float inv = rcp_approx(qz);
      six[k] = qx * inv; siy[k] = qy * inv;
      sok[k] = !(qz < 0.001f) && six[k] >= -1.f && siy[k] >= -1.f && six[k] < Wf && siy[k] < Hf;
      if (sok[k]) {
        const int ix0 = (int)floorf(six[k]), iy0 = (int)floorf(siy[k]);
        mnx = min(mnx, ix0); mxx = max(mxx, ix0 + 1); mny = min(mny, iy0); mxy = max(mxy, iy0 + 1);
      }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; }
    __syncthreads();   // also: every thread has finished reading the tile of the previous view
    if (t == 0) {
      int a = INT_MAX, b = INT_MAX, c = INT_MIN, d = INT_MIN;
#pragma unroll
      for (int w = 0; w < 8; ++w) { a = min(a, red[0][w]); b = min(b, red[1][w]); c = max(c, red[2][w]); d = max(d, red[3][w]); }
      const bool any = c >= a;
      const bool fits = any && (c - a + 1) <= kCvBW && (d - b + 1) <= kCvBH;
      box[0] = a; box[1] = b; box[2] = fits ? 1 : 0;
      if (fits) {
        tc::mbar_expect_tx(&bar, (uint32_t)(kCvBW * kCvBH * 16));
        tc::tma_load_4d(tile, &tmap, 0, a + 1, b + 1, s + 1, &bar);   // +1: the P8 halo column / row
        atomicAdd(&stats[0], 1u);
      } else if (any) {
        atomicAdd(&stats[1], 1u);
      }
    }
    __syncthreads();
    const bool use_tma = box[2] != 0;
    const int bx0 = box[0], by0 = box[1];
    if (use_tma) { tc::mbar_wait(&bar, phase & 1); ++phase; }
    const int vbase = ((s + 1 + feats.pd) * feats.Hp + 1) * feats.Wp + 1;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      __half2 df[C / 2];
      if (sok[k]) {
        const float x0f = floorf(six[k]), y0f = floorf(siy[k]);
        const float ax = six[k] - x0f, ay = siy[k] - y0f;
        const float wy1 = ay * kS, wy0 = fmaf(-ay, kS, kS), wx0 = 1.f - ax;
        const __half2 h00 = __float2half2_rn(wx0 * wy0), h01 = __float2half2_rn(ax * wy0);
        const __half2 h10 = __float2half2_rn(wx0 * wy1), h11 = __float2half2_rn(ax * wy1);
        uint4 q00, q01, q10, q11;
        if (use_tma) {
          const uint4* tq = tile + ((int)y0f - by0) * kCvBW + ((int)x0f - bx0);
          q00 = tq[0]; q01 = tq[1]; q10 = tq[kCvBW]; q11 = tq[kCvBW + 1];
        } else {
          const __half* tq = feats.p + (size_t)(unsigned)((vbase + (int)y0f * feats.Wp + (int)x0f) * 8);
          q00 = __ldg(reinterpret_cast<const uint4*>(tq)); q01 = __ldg(reinterpret_cast<const uint4*>(tq + 8));
          q10 = __ldg(reinterpret_cast<const uint4*>(tq + row)); q11 = __ldg(reinterpret_cast<const uint4*>(tq + row + 8));
        }
        const __half2* a00 = reinterpret_cast<const __half2*>(&q00);
        const __half2* a01 = reinterpret_cast<const __half2*>(&q01);
        const __half2* a10 = reinterpret_cast<const __half2*>(&q10);
        const __half2* a11 = reinterpret_cast<const __half2*>(&q11);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          df[j] = __hfma2(a11[j], h11, __hfma2(a10[j], h10, __hfma2(a01[j], h01, __hfma2(a00[j], h00, nref[j]))));
      } else {
#pragma unroll
        for (int c = 0; c < C / 2; ++c) df[c] = nref[c];
      }
      __half2 dh = __float2half2_rn(0.f);
#pragma unroll
      for (int c = 0; c < C / 2; ++c) {
        df[c] = __hmul2(df[c], df[c]);
        dh = __hfma2(df[c], gwh[c], dh);
      }
      const float2 dl = __half22float2(dh);
      const float dot = dl.x + dl.y;
      const float h1 = fmaxf(fmaf(dot, p.dot_unscale, p.gb1), 0.f);
      const float g = fmaxf(fmaf(p.gw2, h1, p.gb2), 0.f) + 1.f;
      if constexpr (ACC16) {
        const __half2 gh = __float2half2_rn(g);
#pragma unroll
        for (int c = 0; c < C / 2; ++c) acch[k][c] = __hfma2(df[c], gh, acch[k][c]);
      } else {
        const float2 g2 = make_float2(g, g);
#pragma unroll
        for (int c = 0; c < C / 2; ++c) acc[k][c] = ffma2(__half22float2(df[c]), g2, acc[k][c]);
      }
    }
  }
  if (active) {
    const float sc = 4096.f / (float)p.nsrc * p.vol_scale;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      if (d0 + k < p.D) {
        const long long op = vol.pos(d0 + k, y, x);
        float o8[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 v = ACC16 ? __half22float2(acch[k][j]) : acc[k][j];
          o8[2 * j] = vol_store<TV>(v.x * sc);
          o8[2 * j + 1] = vol_store<TV>(v.y * sc);
        }
        store_vec<TV, 8>(vol.p + op, o8);
      }
    }
  }
}

}  // namespace tdm

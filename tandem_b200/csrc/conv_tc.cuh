// Generation-2 convolution: im2col-free implicit GEMM on the 5th-gen tensor cores (tcgen05 / TMEM), fed by 1-D
// TMA bulk copies.  Replaces cuDNN's per-layer library calls of the reference (module.forward, dr_mvsnet.cpp:294)
// for every stride-1 3x3(x3) convolution of FeatureNet / CostRegNet (module.py:496-531, 577-600).
//
// Formulation ("flattened zero-haloed planes"): activations are P8, i.e. [C/8][D+2][H+2][W+2][8] with zero halos
// (mvsnet_kernels.cuh).  A CTA owns an output tile of R rows x TW columns x DR planes.  For every input plane it
// needs, (R+2) row segments of TW+2 positions per channel group are copied with cp.async.bulk into a shared-memory
// tile of pitch P = TW+2 positions.  Flattening (row, col) -> l = row*P + col makes every tap (kh,kw) a CONSTANT
// shift kh*P + kw of l, so the A operand of tap (kd,kh,kw) for the 128 output positions l0..l0+127 is the same
// shared-memory tile read through a descriptor whose start address is advanced by (l0 + kh*P + kw)*16 bytes:
//   K-major, SWIZZLE_NONE canonical layout  ((8,m),(8,2)) : ((16 B, SBO = 128 B), (2 B, LBO))
//   rows = positions (16 B apart), SBO = 8 positions, LBO = distance between the two 8-channel halves of a K=16 step
//   (the next channel group's plane for CIN >= 16, the next tap's shift for CIN == 8).
// Outputs computed for the 2 pad columns of each row are garbage and masked in the epilogue (utilisation TW/(TW+2)).
// B (weights, all taps) is pre-arranged on the host in the same canonical core-matrix order and stays resident.
// Accumulators live in TMEM, double-buffered per output plane, so the epilogue (tcgen05.ld -> bias/ReLU/skip ->
// 16-byte stores) of plane d overlaps the MMAs of plane d+1.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM allocator,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#pragma once
#include "common.cuh"
#include "mvsnet_kernels.cuh"

namespace tdm {
namespace tc {

struct Geom {
  int D, H, W;            // output == input dims (stride 1, pad 1)
  int Hp, Wp;             // H+2, W+2
  int pd;                 // 1: 3-D conv (KD = 3, halo planes exist), 0: 2-D conv over independent planes (KD = 1)
  long long in_gs, out_gs, res_gs;   // channel-group strides in elements
  int R, TW, DR;          // tile: rows, cols, output planes
  int P;                  // TW + 2
  int nch;                // 128-position chunks per plane = ceil(R*P / 128)
  int slot_pos;           // positions per (ring slot, channel group)
  int tiles_w, tiles_h, tiles_d;
  int relu, has_res;
  int cout;               // real output channels (<= NPAD)
  int S;                  // ring slots
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; !mbar_try(bar, parity); ++i)
    if (i > (1u << 24)) __trap();
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout[61,64)=0
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int kThreads = 192;

// blocks of B per kd plane: one block = one K=16 MMA step = NPAD x 16 elements in canonical order
template <int CIN> constexpr int blocks_per_kd() { return CIN >= 16 ? 9 * (CIN / 16) : 5; }

// AFMT: 0 = f16, 1 = bf16 (a_format/b_format of the instruction descriptor).  OUT_PLAIN: fp32 [D][H][W] (prob conv).
template <typename TIn, typename TOut, int CIN, int NPAD, int KD, bool OUT_PLAIN>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc(const TIn* __restrict__ in, const TIn* __restrict__ bimg /*host-arranged B image*/, const float* __restrict__ bias,
          const TOut* __restrict__ res, TOut* __restrict__ out, float* __restrict__ plain_out, const __grid_constant__ Geom g) {
  constexpr int CG = CIN / 8;
  constexpr int NBLK = blocks_per_kd<CIN>();
  constexpr int B_BYTES = KD * NBLK * NPAD * 32;
  constexpr uint32_t AFMT = std::is_same<TIn, __nv_bfloat16>::value ? 1u : 0u;
  constexpr uint32_t IDESC = (1u << 4) | (AFMT << 7) | (AFMT << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sA = smem + ((B_BYTES + 127) / 128) * 128;
  const uint32_t cg_bytes = (uint32_t)g.slot_pos * 16u;
  const uint32_t slot_bytes = cg_bytes * CG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)slot_bytes * g.S);
  uint64_t* full = bars;                 // [S]
  uint64_t* empty = bars + 8;            // [S]
  uint64_t* acc_full = bars + 16;        // [2]
  uint64_t* acc_empty = bars + 18;       // [2]
  uint64_t* b_full = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int tw = tile % g.tiles_w; tile /= g.tiles_w;
  const int th = tile % g.tiles_h; tile /= g.tiles_h;
  const int td = tile;
  const int w0 = tw * g.TW, h0 = th * g.R, d0 = td * g.DR;
  const int ndo = min(g.DR, g.D - d0);             // output planes of this tile
  const int nin = ndo + KD - 1;                    // input planes to stream
  const uint32_t acc_cols = (uint32_t)g.nch * NPAD;  // TMEM columns per accumulator buffer
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(b_full, B_BYTES);
      bulk_g2s(sB, bimg, B_BYTES, b_full);
      const int nrows = min(g.R + 2, g.Hp - h0);
      const uint32_t row_bytes = (uint32_t)min(g.P, g.Wp - w0) * 16u;
      const long long plane_elems = (long long)g.Hp * g.Wp * 8;
      for (int rp = 0; rp < nin; ++rp) {
        const int slot = rp % g.S;
        if (rp >= g.S) mbar_wait(&empty[slot], ((rp / g.S) - 1) & 1);
        mbar_expect_tx(&full[slot], row_bytes * (uint32_t)nrows * CG);
        const int pp = d0 + rp;  // padded plane index (pd == 1: plane d-1+kd+1; pd == 0: plane d)
#pragma unroll 1
        for (int cg = 0; cg < CG; ++cg) {
          const TIn* src = in + cg * g.in_gs + pp * plane_elems + ((long long)h0 * g.Wp + w0) * 8;
          uint8_t* dst = sA + (size_t)slot * slot_bytes + (size_t)cg * cg_bytes;
          for (int r = 0; r < nrows; ++r)
            bulk_g2s(dst + (size_t)r * g.P * 16, src + (long long)r * g.Wp * 8, row_bytes, &full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_wait(b_full, 0);
      int next_wait = 0;
      const uint32_t sB_addr = smem_u32(sB), sA_addr = smem_u32(sA);
      for (int od = 0; od < ndo; ++od) {
        const int buf = od & 1;
        if (od >= 2) mbar_wait(&acc_empty[buf], ((od >> 1) - 1) & 1);
        while (next_wait <= od + KD - 1) {
          mbar_wait(&full[next_wait % g.S], (next_wait / g.S) & 1);
          ++next_wait;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < g.nch; ++c) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * acc_cols + c * NPAD);
          uint32_t accum = 0;
#pragma unroll 1
          for (int kd = 0; kd < KD; ++kd) {
            const uint32_t a_slot = sA_addr + (uint32_t)((od + kd) % g.S) * slot_bytes + (uint32_t)c * 128u * 16u;
#pragma unroll 1
            for (int b = 0; b < NBLK; ++b) {
              uint32_t a_start, a_lbo;
              if constexpr (CIN >= 16) {
                const int t = b / (CIN / 16), j = b % (CIN / 16);
                a_start = a_slot + (uint32_t)(2 * j) * cg_bytes + (uint32_t)((t / 3) * g.P + (t % 3)) * 16u;
                a_lbo = cg_bytes;
              } else {
                // taps are paired (0,1)(2,3)(4,5)(6,7)(7*,8): the 5th block re-reads tap 7 against zero weights so
                // that no K slice ever touches shared memory outside the copied tile
                const int t0 = b < 4 ? 2 * b : 7, t1 = b < 4 ? 2 * b + 1 : 8;
                const int o0 = (t0 / 3) * g.P + (t0 % 3);
                const int o1 = (t1 / 3) * g.P + (t1 % 3);
                a_start = a_slot + (uint32_t)o0 * 16u;
                a_lbo = (uint32_t)(o1 - o0) * 16u;
              }
              const uint64_t ad = make_desc(a_start, a_lbo, 128u);
              const uint64_t bd = make_desc(sB_addr + (uint32_t)((kd * NBLK + b) * NPAD * 32), (uint32_t)NPAD * 16u, 128u);
              mma_f16(d_tmem, ad, bd, IDESC, accum);
              accum = 1;
            }
          }
        }
        mma_commit(&acc_full[buf]);           // accumulators of plane od complete -> epilogue
        mma_commit(&empty[od % g.S]);         // oldest input plane no longer needed -> producer may refill
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    for (int od = 0; od < ndo; ++od) {
      const int buf = od & 1;
      mbar_wait(&acc_full[buf], (od >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int d = d0 + od;
      for (int c = 0; c < g.nch; ++c) {
        uint32_t v[16];
        float acc[NPAD];
#pragma unroll
        for (int n0 = 0; n0 < NPAD; n0 += 16) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols + c * NPAD + n0), v);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[n0 + i] = __uint_as_float(v[i]);
        }
        const int l = c * 128 + q * 32 + lane;
        const int hh = l / g.P, ww = l - hh * g.P;
        const int h = h0 + hh, w = w0 + ww;
        if (hh < g.R && ww < g.TW && h < g.H && w < g.W) {
          if constexpr (OUT_PLAIN) {
            plain_out[((long long)d * g.H + h) * g.W + w] = acc[0] + (bias ? bias[0] : 0.f);
          } else {
            const long long pos = ((((long long)(d + g.pd)) * g.Hp + (h + 1)) * g.Wp + (w + 1)) * 8;
#pragma unroll
            for (int c0 = 0; c0 < NPAD; c0 += 8) {
              if (c0 < g.cout) {
                float o8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float x = acc[c0 + i] + (bias ? bias[c0 + i] : 0.f);
                  if (g.relu) x = fmaxf(x, 0.f);
                  o8[i] = x;
                }
                if (g.has_res) {
                  float r8[8];
                  load_vec<TOut, 8>(res + pos + (c0 >> 3) * g.res_gs, r8);
#pragma unroll
                  for (int i = 0; i < 8; ++i) o8[i] += r8[i];
                }
                store_vec<TOut, 8>(out + pos + (c0 >> 3) * g.out_gs, o8);
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
template <int CIN> inline size_t b_image_elems(int npad, int kd) { return (size_t)kd * blocks_per_kd<CIN>() * npad * 16; }

// w: folded fp32 weights [tap][cin][cout] (tap = (kd*3+kh)*3+kw).  Writes the canonical K-major no-swizzle image:
// block(kd,b) = [k half (2)][n/8][n%8][8 k-elements]
template <typename T, int CIN>
inline void build_b_image(const float* w, int cin, int cout, int npad, int kd_n, std::vector<T>& img, T (*cvt)(float)) {
  constexpr int NBLK = blocks_per_kd<CIN>();
  img.assign((size_t)kd_n * NBLK * npad * 16, cvt(0.f));
  for (int kd = 0; kd < kd_n; ++kd)
    for (int b = 0; b < NBLK; ++b)
      for (int half = 0; half < 2; ++half)
        for (int n = 0; n < cout; ++n)
          for (int e = 0; e < 8; ++e) {
            int tap9, ci;
            if (CIN >= 16) { tap9 = b / (CIN / 16); ci = (b % (CIN / 16)) * 16 + half * 8 + e; }
            else { tap9 = b < 4 ? 2 * b + half : (half == 0 ? -1 : 8); ci = e; }  // block 4 = (zero, tap 8)
            if (tap9 < 0 || tap9 >= 9 || ci >= cin) continue;
            const int tap = kd * 9 + tap9;
            const float val = w[((size_t)tap * cin + ci) * cout + n];
            img[(((size_t)(kd * NBLK + b) * 2 + half) * (npad / 8) + n / 8) * 64 + (n % 8) * 8 + e] = cvt(val);
          }
}

struct Plan {
  Geom g;
  int grid;
  size_t smem;
};

inline Plan make_plan(int cin, int npad, int kd, int D, int H, int W, int pd, size_t smem_limit = 200 * 1024) {
  Plan p{};
  Geom& g = p.g;
  g.D = D; g.H = H; g.W = W; g.Hp = H + 2; g.Wp = W + 2; g.pd = pd;
  const int cg = cin / 8;
  const int nblk = cin >= 16 ? 9 * (cin / 16) : 5;
  const size_t bbytes = (((size_t)kd * nblk * npad * 32) + 127) / 128 * 128;
  g.S = kd == 3 ? 4 : 2;
  // choose TW (<= 320 so wide rows split), then the largest R that fits shared memory and TMEM (2*nch*npad <= 512)
  int tiles_w = 1;
  while ((W + tiles_w - 1) / tiles_w > 320) ++tiles_w;
  g.TW = (W + tiles_w - 1) / tiles_w;
  g.P = g.TW + 2;
  int bestR = 0;
  for (int R = 1; R <= 16 && R <= H; ++R) {
    const int nch = (R * g.P + 127) / 128;
    const int slot_pos = std::max((R + 2) * g.P, nch * 128 + 2 * g.P + 2) + 8;
    const size_t smem = bbytes + (size_t)g.S * cg * slot_pos * 16 + 256;
    if (smem <= smem_limit && 2 * nch * npad <= 512) bestR = R;
  }
  TDM_CHECK(bestR > 0, "conv_tc: no tile fits shared memory");
  g.R = bestR;
  g.nch = (g.R * g.P + 127) / 128;
  g.slot_pos = std::max((g.R + 2) * g.P, g.nch * 128 + 2 * g.P + 2) + 8;
  g.tiles_w = tiles_w;
  g.tiles_h = (H + g.R - 1) / g.R;
  // planes per tile: keep >= ~2 waves of CTAs on 148 SMs when the tensor allows it
  g.DR = D;
  if (kd == 3) {
    while (g.DR > 4 && (long long)g.tiles_w * g.tiles_h * ((D + g.DR - 1) / g.DR) < 296) g.DR = (g.DR + 1) / 2;
  } else {
    g.DR = 1;
    while (g.DR < D && (long long)g.tiles_w * g.tiles_h * ((D + g.DR) / (g.DR + 1)) >= 592) ++g.DR;
  }
  g.tiles_d = (D + g.DR - 1) / g.DR;
  p.grid = g.tiles_w * g.tiles_h * g.tiles_d;
  p.smem = bbytes + (size_t)g.S * cg * g.slot_pos * 16 + 256;
  return p;
}

}  // namespace tc
}  // namespace tdm

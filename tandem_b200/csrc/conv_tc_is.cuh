// Input-stationary variant of the 3x3x3 stride-1 tensor-core convolution (conv_tc.cuh) for the 3-D layers of
// CostRegNet (conv0 / conv2 / conv4 / prob, module.py:577-600).
//
// Measured on the B200 (profiles/): a tcgen05.mma with M=128, K=16 and N <= 64 is bound by streaming its operands out of
// shared memory at ~93 B/clk (A = 4 KB per instruction), not by the tensor pipe.  The output-stationary kernel re-reads
// every input plane three times (once per kd).  Here each (tap, k-step) of an input plane p is ONE instruction with
//   N = 3 * NMMA,   B = [ W(kd=2) | W(kd=1) | W(kd=0) ],   D = accumulators of output planes (p-1 | p | p+1)
// so A is read once per input plane: 4 KB + 3 KB of operands instead of 3 x 5 KB.  The three accumulators must be
// adjacent TMEM column ranges: per 128-position chunk the output planes live in a ring of 4 slots (3 accumulating + 1
// draining), a run that wraps around the ring is split into two instructions.  Because one instruction now mixes a
// plane's first contribution with later ones, the per-instruction "accumulate" flag cannot zero anything: the epilogue
// warps zero a slot (tcgen05.st) right after draining it, and every MMA accumulates.
// An input plane is consumed exactly once, so its shared-memory slot is released immediately (ring of 3).
#pragma once
#include "conv_tc.cuh"

namespace tdm {
namespace tc {

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Per-pixel tail of the OUT_PLAIN (prob) instantiation: with tiles that span all D planes (make_plan(full_depth)) the epilogue
// thread that stored a pixel's D logits is the same for every plane, so once the tile's last plane is drained it can finish the
// pixel (softmax / soft-argmin / confidence, module.py:1116-1133) without another launch.  NoTail = plain convolution.
struct NoTail {
  static constexpr bool enabled = false;
  __device__ __forceinline__ void operator()(const float*, int, int, int, int) const {}
};

template <typename TIn, typename TOut, int CIN, int NPAD, bool OUT_PLAIN, bool HILO, typename Tail = NoTail>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc_is(const __grid_constant__ CUtensorMap tmap, const TIn* __restrict__ bimg, const float* __restrict__ bias,
             const TOut* __restrict__ res, TOut* __restrict__ out, float* __restrict__ plain_out,
             const __grid_constant__ Geom g, const __grid_constant__ Tail tail) {
  static_assert(!Tail::enabled || OUT_PLAIN, "a per-pixel tail only exists for the single-channel fp32 output");
  constexpr int CG = CIN / 8;
  constexpr int NBLK = blocks_per_kd<CIN, 0>();
  constexpr int NMMA = HILO ? 2 * NPAD : NPAD;
  constexpr int NB3 = 3 * NMMA;                       // columns of one B block
  constexpr int B_BYTES = NBLK * NB3 * 32;
  constexpr int RS = 4;                               // TMEM ring slots per chunk
  constexpr uint32_t AFMT = std::is_same<TIn, __nv_bfloat16>::value ? 1u : 0u;
  constexpr uint32_t IDESC0 = (1u << 4) | (AFMT << 7) | (AFMT << 10) | ((128u >> 4) << 24);

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sA = smem + ((B_BYTES + 127) / 128) * 128;
  const uint32_t cg_bytes = (uint32_t)g.slot_pos * 16u;
  const uint32_t slot_bytes = cg_bytes * CG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)slot_bytes * g.S);
  uint64_t* full = bars;                 // [S]
  uint64_t* empty = bars + 4;            // [S]
  uint64_t* acc_full = bars + 8;         // [RS]
  uint64_t* acc_empty = bars + 12;       // [RS]
  uint64_t* b_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // Persistent over tiles (round 2): a CTA owns the tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Barrier init, TMEM allocation
  // and the B-image load happen once per CTA, and because every role keeps RUNNING plane counters (shared-memory ring slot and
  // parity from the global input-plane count, TMEM ring slot and parity from the global output-plane count) the producer
  // streams the next tile's planes while the previous tile's last planes are still being multiplied and drained - no pipeline
  // fill / drain bubble between tiles.  ~4.5 us of fixed cost per tile was 40 % of the time of the 440- / 640-tile layers.
  const int ntiles = g.tiles_w * g.tiles_h * g.tiles_d;
  const uint32_t chunk_cols = RS * NMMA;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)g.nch * chunk_cols) tmem_cols <<= 1;
#define TDM_IS_TILE(tile_)                                                          \
  int t_ = (tile_);                                                                \
  const int tw = t_ % g.tiles_w; t_ /= g.tiles_w;                                  \
  const int th = t_ % g.tiles_h; t_ /= g.tiles_h;                                  \
  const int w0 = tw * g.TW, h0 = th * g.R, d0 = t_ * g.DR;                         \
  const int ndo = min(g.DR, g.D - d0);                                             \
  const int nin = ndo + 2;                                                         \
  (void)w0; (void)h0; (void)nin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMmaWarps); }
    for (int r = 0; r < RS; ++r) { mbar_init(&acc_full[r], kMmaWarps); mbar_init(&acc_empty[r], 4 * kEpiGroups); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(b_full, B_BYTES);          // weights do not depend on the previous kernel: fetch them before pdl_wait
    bulk_g2s(sB, bimg, B_BYTES, b_full);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();   // from here on activations written by the previous kernel are read and our outputs are written
  // trigger the dependents only now: a trigger BEFORE the wait lets the whole chain of later kernels become resident early (each
  // one's prologue would trigger the next), and their idle CTAs then hold the shared memory the running kernel's tiles need
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t box_bytes = (uint32_t)g.P * (uint32_t)(g.R + 2) * 16u;
      int gp = 0;                                   // input planes issued so far by this CTA (all tiles)
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TDM_IS_TILE(tile)
        for (int rp = 0; rp < nin; ++rp, ++gp) {
          const int slot = gp % g.S;
          if (gp >= g.S) mbar_wait(&empty[slot], ((gp / g.S) - 1) & 1);
          mbar_expect_tx(&full[slot], box_bytes * CG);
#pragma unroll 1
          for (int cg = 0; cg < CG; ++cg)
            tma_load_4d(sA + (size_t)slot * slot_bytes + (size_t)cg * cg_bytes, &tmap, 0, w0, h0, cg * g.iDp + d0 + rp, &full[slot]);
        }
      }
    }
  } else if (warp == 1 || warp == 6 || warp == 7) {
    // ===================== MMA issuers =====================
    const int issuer = warp == 1 ? 0 : warp - 5;
    const bool leader = lane == 0;
    uint32_t a_off[NBLK], a_lbo[NBLK];
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
      if constexpr (CIN >= 16) {
        const int t = b / (CIN / 16), j = b % (CIN / 16);
        a_off[b] = (uint32_t)(2 * j) * (cg_bytes >> 4) + (uint32_t)((t / 3) * g.P + (TDM_DBG_MODE(g) ? 0 : (t % 3)));
        a_lbo[b] = cg_bytes >> 4;
      } else {
        const int t0 = b < 4 ? 2 * b : 7, t1 = b < 4 ? 2 * b + 1 : 8;   // (0,1)(2,3)(4,5)(6,7)(7*,8), see conv_tc.cuh
        const int o0 = (t0 / 3) * g.P + (TDM_DBG_MODE(g) ? 0 : (t0 % 3)), o1 = (t1 / 3) * g.P + (TDM_DBG_MODE(g) ? 0 : (t1 % 3));
        a_off[b] = (uint32_t)o0;
        a_lbo[b] = (uint32_t)(o1 - o0);
      }
    }
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t sB16 = (smem_u32(sB) & 0x3FFFFu) >> 4, sA16 = (smem_u32(sA) & 0x3FFFFu) >> 4;
    const uint32_t b_lo_base = sB16 | ((uint32_t)(NB3 * 16 >> 4) << 16);     // LBO = stride between the two K halves
    mbar_wait(b_full, 0);
    int gp = 0, go = 0;                             // global input-plane / output-plane counters (ring slots and parities)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      TDM_IS_TILE(tile)
      for (int rp = 0; rp < nin; ++rp, ++gp) {
        const int q = rp - 1;                                  // this input plane is output plane q's centre (kd = 1)
        const int oa = max(q - 1, 0), ob = min(q + 1, ndo - 1);
        if (q + 1 <= ndo - 1) mbar_wait(&acc_empty[(go + q + 1) & 3], ((go + q + 1) >> 2) & 1);   // newest plane's slot is drained + zeroed
        mbar_wait(&full[gp % g.S], (gp / g.S) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t slot16 = sA16 + (uint32_t)(gp % g.S) * (slot_bytes >> 4);
        for (int c = issuer; c < g.nch; c += kMmaWarps) {
          const uint32_t a16 = slot16 + (uint32_t)c * 128u;
          const uint32_t d_chunk = tmem_base + (uint32_t)c * chunk_cols;
          for (int od = oa; od <= ob;) {
            const int gs = (go + od) & 3;                                // ring slot of output plane od
            const int run_end = min(ob, od + (3 - gs));                  // stay inside the ring (no wrap within a run)
            const int nrun = run_end - od + 1;
            const uint32_t d_tmem = d_chunk + (uint32_t)gs * NMMA;
            const uint32_t bcol16 = (uint32_t)((od - (q - 1)) * NMMA / 8) * 8u;     // column offset in 16-byte units (128 B per 8 cols)
            const uint32_t idesc = IDESC0 | ((uint32_t)((nrun * NMMA) >> 3) << 17);
#pragma unroll
            for (int b = 0; b < NBLK; ++b) {
              if ((TDM_DBG_MODE(g) == 2 || TDM_DBG_MODE(g) == 4) && b >= (NBLK + 2) / 3) continue;   // TIMING EXPERIMENT ONLY: a third of the instructions (kw folded into N)
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)((a16 + a_off[b]) | (a_lbo[b] << 16));
              const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo_base + (uint32_t)(b * NB3 * 2) + bcol16);
              if (leader) mma_f16(d_tmem, ad, bd, idesc, 1u);
            }
            od = run_end + 1;
          }
        }
        if (leader) {
          mma_commit(&empty[gp % g.S]);                              // the input plane is consumed exactly once
          if (q - 1 >= 0 && q - 1 <= ndo - 1) mma_commit(&acc_full[(go + q - 1) & 3]);   // output plane q-1 has seen kd = 0,1,2
        }
        __syncwarp();
      }
      go += ndo;
    }
  } else {
    // ===================== epilogue (warps 2..5): zero, drain, re-zero =====================
    const int qd = warp & 3;
    const int egroup = warp >= 8 ? 1 : 0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
    for (int c = egroup; c < g.nch; c += kEpiGroups)      // each group zeroes (and later drains) its own chunks
      for (uint32_t col = 0; col < chunk_cols; col += 16) tmem_st16_zero(lane_base + (uint32_t)c * chunk_cols + col);
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (int r = 0; r < RS; ++r) mbar_arrive(&acc_empty[r]);     // arrival #0 of every slot: "zeroed"
    int go = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    TDM_IS_TILE(tile)
    for (int od = 0; od < ndo; ++od) {
      const int r = (go + od) & 3;
      mbar_wait(&acc_full[r], ((go + od) >> 2) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int d = d0 + od;
      for (int c = egroup; c < g.nch; c += kEpiGroups) {
        if (TDM_DBG_MODE(g) >= 3) continue;   // TIMING EXPERIMENT ONLY: no TMEM drain / zero / global stores (is the epilogue the limiter?)
        const int l = c * 128 + qd * 32 + lane;
        const int hh = l / g.P, ww = l - hh * g.P;
        const int h = h0 + hh, w = w0 + ww;
        const bool valid = hh < g.R && ww < g.TW && h < g.H && w < g.W;
        const uint32_t t_row = lane_base + (uint32_t)c * chunk_cols + (uint32_t)r * NMMA;
        uint4 rbuf[NPAD / 8];
        const long long pos0 = ((((long long)(d + g.pd)) * g.Hp + (h + 1)) * g.Wp + (w + 1)) * 8;
        if constexpr (!OUT_PLAIN) {
          if (valid && g.has_res) {   // skip-tensor reads issued before the TMEM loads (latency overlap)
#pragma unroll
            for (int k = 0; k < NPAD / 8; ++k)
              if (k * 8 < g.cout) rbuf[k] = __ldg(reinterpret_cast<const uint4*>(res + pos0 + k * g.res_gs));
          }
        }
        uint32_t v[16];
        float acc[NPAD];
        if constexpr (NPAD == 8) {
          // tight packing for the 8- (and 1-) channel outputs: [hi 0..7 | lo 0..7] are the 16 columns of ONE tcgen05.ld
          static_assert(HILO, "NPAD == 8 is only instantiated with hi/lo weights (N must be a multiple of 16)");
          tmem_ld16(t_row, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = __uint_as_float(v[i]) + __uint_as_float(v[8 + i]);
        } else {
#pragma unroll
          for (int n0 = 0; n0 < NPAD; n0 += 16) {
            tmem_ld16(t_row + (uint32_t)n0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[n0 + i] = __uint_as_float(v[i]);
            if constexpr (HILO) {
              tmem_ld16(t_row + (uint32_t)(NPAD + n0), v);
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[n0 + i] += __uint_as_float(v[i]);
            }
          }
        }
#pragma unroll
        for (int n0 = 0; n0 < NMMA; n0 += 16) tmem_st16_zero(t_row + (uint32_t)n0);
        if (valid) {
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
                  const TOut* rv = reinterpret_cast<const TOut*>(&rbuf[c0 / 8]);
#pragma unroll
                  for (int i = 0; i < 8; ++i) o8[i] += to_f<TOut>(rv[i]);
                }
                store_vec<TOut, 8>(out + pos + (c0 >> 3) * g.out_gs, o8);
              }
            }
          }
        }
      }
      tmem_wait_st();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[r]);
    }
    if constexpr (Tail::enabled) {
      // ndo == D (host-checked): this thread wrote all D logits of its pixels itself (program order makes them visible to it)
      for (int c = egroup; c < g.nch; c += kEpiGroups) {
        const int l = c * 128 + qd * 32 + lane;
        const int hh = l / g.P, ww = l - hh * g.P;
        const int h = h0 + hh, w = w0 + ww;
        if (hh < g.R && ww < g.TW && h < g.H && w < g.W) tail(plain_out, h * g.W + w, w, h, g.H * g.W);
      }
    }
    go += ndo;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

#undef TDM_IS_TILE

// B image for the input-stationary kernel: per (tap9, k-step) block the columns are [kd=2 | kd=1 | kd=0], each NMMA wide
// ([hi | lo] inside when hilo).  w: folded fp32 [tap27][cin][cout], tap27 = (kd*3+kh)*3+kw.
template <typename T, int CIN>
inline void build_b_image_is(const float* w, int cin, int cout, int npad, std::vector<T>& img, T (*cvt)(float), bool hilo,
                             float (*back)(T)) {
  constexpr int NBLK = blocks_per_kd<CIN, 0>();
  const int nmma = hilo ? 2 * npad : npad;
  const int nb3 = 3 * nmma;
  img.assign((size_t)NBLK * nb3 * 16, cvt(0.f));
  for (int slotk = 0; slotk < 3; ++slotk) {
    const int kd = 2 - slotk;
    for (int b = 0; b < NBLK; ++b)
      for (int half = 0; half < 2; ++half)
        for (int n = 0; n < cout; ++n)
          for (int e = 0; e < 8; ++e) {
            int tap9, ci;
            if (CIN >= 16) { tap9 = b / (CIN / 16); ci = (b % (CIN / 16)) * 16 + half * 8 + e; }
            else { tap9 = b < 4 ? 2 * b + half : (half == 0 ? -1 : 8); ci = e; }
            if (tap9 < 0 || tap9 >= 9 || ci >= cin) continue;
            const float val = w[((size_t)(kd * 9 + tap9) * cin + ci) * cout + n];
            T* blk = img.data() + ((size_t)b * 2 + half) * (nb3 / 8) * 64;
            const int c_hi = slotk * nmma + n;
            const T hi = cvt(val);
            blk[(c_hi / 8) * 64 + (c_hi % 8) * 8 + e] = hi;
            if (hilo) {
              const int c_lo = slotk * nmma + npad + n;
              blk[(c_lo / 8) * 64 + (c_lo % 8) * 8 + e] = cvt(val - back(hi));
            }
          }
  }
}

}  // namespace tc
}  // namespace tdm

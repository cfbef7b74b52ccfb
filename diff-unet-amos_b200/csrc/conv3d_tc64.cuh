// 3x3x3 convolution with Cout = 64 (padded): the layers that carry ~90 % of the network's FLOPs (every conv at the
// 96^3 and 48^3 levels).  Same data path as conv3d_tc.cuh (TMA halo planes, no-swizzle K-major UMMA operands, TMEM
// accumulators) but restructured around one measured fact (tools/umma_probe.cu, profiles/): a tcgen05.mma reads its
// shared-memory operands at 128 B/clk/SM, so a 128 x N x 16 MMA costs max(N/2, (4096 + 32 N)/128) clocks -- N = 64 is
// operand-fetch bound at 67 % of the tensor peak, N >= 128 is not.
//
//   * Z-STACKING: for a resident input plane p and an in-plane tap (ty,tx), the three z-taps tz = 0,1,2 send
//     W[tz,ty,tx] * A to the three output slabs p, p-1, p-2.  The accumulators of consecutive slabs are adjacent TMEM
//     column blocks, so ONE MMA with N = 192 (B rows = [tz=2 | tz=1 | tz=0] x 64 couts) feeds all three: the A tile is
//     read once instead of three times and the MMA runs at the full tensor rate.  Border planes use N = 64 / 128.
//   * PERSISTENT CTAs (one per SM) walk a static list of 8x16xZT output tiles; the plane/weight rings run ahead across
//     tile boundaries and the accumulators are double-buffered in TMEM (2 x ZT x 64 columns), so the epilogue of tile i
//     (TMEM -> bf16 -> HBM, InstanceNorm statistics) overlaps the MMAs of tile i+1.
//   * FUSE (second conv of a TwoConv, bf16 mode): the input is the RAW output of the previous conv; eight extra warps
//     apply that conv's InstanceNorm + LeakyReLU (+ time-embedding bias) to every halo plane IN SHARED MEMORY between
//     the TMA arrival and the MMAs (out-of-volume halo voxels stay the zeros TMA wrote: the padding of the normalised
//     tensor).  The normalised intermediate never exists in HBM: one full read + write pass per TwoConv disappears.
//     Same fp32 formulas as norm_act_kernel, so the result is bit-identical to the unfused path.  The transform warps
//     are instruction-latency bound (one dependent stream per scheduler): with four of them the conv lost what the
//     saved pass gained (690 vs 532 + 157 us), with eight it takes 558 us.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "conv3d_tc.cuh"

namespace dunet {

template <int CB_CH, int ZT>
struct ConvTc64 {
  static constexpr int HX = CONV_TX + 2, HY = CONV_TY + 2;
  static constexpr int KCH = CB_CH / 8;
  static constexpr int KC = CB_CH / 16;                     // MMAs (K = 16) per plane per tap
  static constexpr int PLANE_BYTES = KCH * HY * HX * 16;
  static constexpr int A_LBO = HY * HX * 16, A_SBO = HX * 16;
  static constexpr int NROWS = 192;                         // [tz=2 | tz=1 | tz=0] x 64 output channels
  static constexpr int W_UNIT_BYTES = KCH * NROWS * 16;     // one (cin block, ty, tx) weight tile: 24 KB for 64 ch
  static constexpr int B_LBO = NROWS * 16, B_SBO = 128;
  static constexpr int PLANES = ZT + 2;
  // 32-channel blocks: the ring holds TWO blocks' planes, so the next block (or the next tile's first block) is loaded
  // -- and, with FUSE, normalised -- entirely under the MMAs of the current one.  64-channel blocks: one block + 1.
  static constexpr int A_SLOTS = CB_CH == 32 ? 2 * PLANES + 1 : PLANES + 1;
  static constexpr int W_SLOTS = CB_CH == 32 ? 4 : 2;
  static constexpr int ACC_COLS = ZT * 64;                  // one accumulator set
  static constexpr int TMEM_COLS = 512;
  static constexpr int RED_BYTES = 4 * 128 * 4;
  static constexpr int SMEM_BYTES = A_SLOTS * PLANE_BYTES + W_SLOTS * W_UNIT_BYTES + RED_BYTES + 1024 + 512;
  static constexpr int XFORM_WARPS = 8;  // FUSE: the transform warps are instruction-latency bound (one dependent stream per
                                         // scheduler), more warps hide it
  static constexpr int THREADS_FUSED = CONV_THREADS + XFORM_WARPS * 32;
  static_assert(2 * ACC_COLS <= 512, "double-buffered accumulators exceed TMEM");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct ConvTc64Args {
  const __nv_bfloat16* w;  // packed [cin_block][ty*3+tx][KCH][192][8]
  __nv_bfloat16* out;      // raw conv output, C8-planar, 64 channels
  __nv_bfloat16* out_lo;   // fp32x3 mode: low part of the output, or nullptr
  float* stats;            // [n*8 + chunk][gridDim.x][16] (one row per CTA and sample) or nullptr
  ConvSegs segs;           // input-channel block list (conv3d_tc.cuh)
  int D, H, W;
  int tiles_x, tiles_y, tiles_z, batch;
  // FUSE only (single 64-channel source): affine map of the producer's InstanceNorm per (sample, 8-channel chunk),
  // [n * 8 + chunk][16] = scale[8], shift[8] (in_affine_kernel); bias added after the activation [64] or nullptr
  const float* in_affine;
  const float* in_bias;
  int in_bias_n_stride;    // floats between the bias rows of consecutive samples (0: one row for the whole batch)
  float slope;
};

// Static tile schedule.  Within EACH sample the tiles are dealt round-robin over the CTAs, so the set of tiles whose
// InstanceNorm partial sums share a row is a residue class modulo gridDim.x -- independent of how many samples are
// batched (a window's result is bit-identical whatever it is batched with).  The residue a CTA serves is rotated from
// sample to sample so the CTAs with one tile more are different ones each time.
template <class F>
__device__ __forceinline__ void for_each_tile(int tiles_per_n, int batch, F&& f) {
  const int G = gridDim.x, shift = tiles_per_n % G;
  for (int n = 0; n < batch; ++n) {
    const int r = (int)((blockIdx.x + (long long)G * batch - (long long)n * shift) % G);
    for (int lin = r; lin < tiles_per_n; lin += G) f(n, lin);
  }
}

template <int CB_CH, int ZT, bool FUSE, bool H>
__global__ void __launch_bounds__(FUSE ? ConvTc64<CB_CH, ZT>::THREADS_FUSED : CONV_THREADS, 1)
conv3d_tc64_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                   const __grid_constant__ CUtensorMap tmap2, const __grid_constant__ CUtensorMap tmap3, ConvTc64Args a) {
  using Cfg = ConvTc64<CB_CH, ZT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem_base;
  const uint32_t w_smem = a_smem + Cfg::A_SLOTS * Cfg::PLANE_BYTES;
  const uint32_t red_smem = w_smem + Cfg::W_SLOTS * Cfg::W_UNIT_BYTES;
  const uint32_t bars = red_smem + Cfg::RED_BYTES;
  const uint32_t a_full = bars, a_empty = bars + 8 * Cfg::A_SLOTS;
  const uint32_t w_full = bars + 16 * Cfg::A_SLOTS, w_empty = w_full + 8 * Cfg::W_SLOTS;
  const uint32_t acc_full = w_empty + 8 * Cfg::W_SLOTS;  // [2]
  const uint32_t acc_empty = acc_full + 16;              // [2]
  const uint32_t tmem_slot = acc_empty + 16;
  const uint32_t a_ready = tmem_slot + 8;                // [A_SLOTS], FUSE: plane normalised in place, MMAs may read it
  static_assert(16 * Cfg::A_SLOTS + 16 * Cfg::W_SLOTS + 32 + 8 + 8 * Cfg::A_SLOTS <= 512, "barrier block");
  float* red = reinterpret_cast<float*>(smem_raw + (red_smem - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncb = a.segs.ncb;
  const int tiles_per_n = a.tiles_x * a.tiles_y * a.tiles_z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); mbar_init(a_ready + 8 * i, Cfg::XFORM_WARPS); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 4); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap0);
    if (a.segs.n > 1) tma_prefetch_desc(&tmap1);
    if (a.segs.n > 2) { tma_prefetch_desc(&tmap2); tma_prefetch_desc(&tmap3); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // =============================== halo-plane producer (TMA), runs ahead across tiles ===============================
    if (elect_one_sync()) {
      const CUtensorMap* const tms[4] = {&tmap0, &tmap1, &tmap2, &tmap3};
      int u = 0;
      for_each_tile(tiles_per_n, a.batch, [&](int n, int lin) {
        int t = lin;
        const int tix = t % a.tiles_x; t /= a.tiles_x;
        const int tiy = t % a.tiles_y; t /= a.tiles_y;
        const int tiz = t;
        const int x0 = tix * CONV_TX, y0 = tiy * CONV_TY, z0 = tiz * ZT;
        for (int cb = 0; cb < ncb; ++cb) {
          int ti, chunk0, chunks;
          conv_seg_lookup(a.segs, cb, Cfg::KCH, ti, chunk0, chunks);
          const CUtensorMap* tm = tms[ti];
          const int c3 = n * chunks + chunk0;
          for (int p = 0; p < Cfg::PLANES; ++p, ++u) {
            const int slot = u % Cfg::A_SLOTS, it = u / Cfg::A_SLOTS;
            if (it > 0) mbar_wait(a_empty + 8 * slot, (it - 1) & 1);
            mbar_arrive_expect_tx(a_full + 8 * slot, Cfg::PLANE_BYTES);
            tma_load_4d(a_smem + slot * Cfg::PLANE_BYTES, tm, a_full + 8 * slot, (x0 - 1) * 8, y0 - 1, z0 + p - 1, c3);
          }
        }
      });
    }
  } else if (warp == 1) {
    // =============================== weight-tile producer (bulk copy) ===============================
    if (elect_one_sync()) {
      int w = 0;
      for_each_tile(tiles_per_n, a.batch, [&](int, int) {
        for (int i = 0; i < ncb * 9; ++i, ++w) {
          const int slot = w % Cfg::W_SLOTS, it = w / Cfg::W_SLOTS;
          if (it > 0) mbar_wait(w_empty + 8 * slot, (it - 1) & 1);
          mbar_arrive_expect_tx(w_full + 8 * slot, Cfg::W_UNIT_BYTES);
          bulk_load_1d(w_smem + slot * Cfg::W_UNIT_BYTES,
                       reinterpret_cast<const uint8_t*>(a.w) + (size_t)i * Cfg::W_UNIT_BYTES, Cfg::W_UNIT_BYTES,
                       w_full + 8 * slot);
        }
      });
    }
  } else if (warp == 2) {
    // =============================== MMA issuer (one thread) ===============================
    // The single issuing thread must sustain one MMA per ~50-100 clocks, so the per-MMA instruction count matters:
    // plane descriptors are formed once per block, the tap loop is split into first / middle / last variants (plane
    // waits and accumulator initialisation only in the first, plane release only in the last) and all descriptor
    // arithmetic is 32-bit.
    if (elect_one_sync()) {
      const uint64_t a_desc0 = make_smem_desc(a_smem, Cfg::A_LBO, Cfg::A_SBO);
      const uint64_t b_desc0 = make_smem_desc(w_smem, Cfg::B_LBO, Cfg::B_SBO);
      const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
      const uint32_t b_lo0 = (uint32_t)b_desc0, b_hi = (uint32_t)(b_desc0 >> 32);
      static_assert((Cfg::W_SLOTS & (Cfg::W_SLOTS - 1)) == 0, "W_SLOTS must be a power of two");
      int u0 = 0, w = 0, li = 0;
      for_each_tile(tiles_per_n, a.batch, [&](int, int) {
        const int buf = li & 1, use = li >> 1;
        if (use > 0) { mbar_wait(acc_empty + 8 * buf, (use - 1) & 1); tc_fence_after(); }
        const uint32_t acc = tmem_base + buf * Cfg::ACC_COLS;
        for (int cb = 0; cb < ncb; ++cb, u0 += Cfg::PLANES) {
          uint32_t pl_lo[Cfg::PLANES], pl_bar[Cfg::PLANES];  // descriptor low word / barrier offset of each plane's slot
          uint32_t par[Cfg::PLANES];
#pragma unroll
          for (int p = 0; p < Cfg::PLANES; ++p) {
            const int u = u0 + p, slot = u % Cfg::A_SLOTS;
            pl_lo[p] = a_lo0 + slot * (Cfg::PLANE_BYTES >> 4);
            pl_bar[p] = 8 * slot;
            par[p] = (u / Cfg::A_SLOTS) & 1;
          }
          // KIND 0: first tap of the block (waits for the planes; with first_cb also initialises the accumulators),
          // 1: middle taps, 2: last tap (hands the planes back to the producer)
          auto tap = [&](auto kind, int tyx, bool first_cb) {
            constexpr int KIND = decltype(kind)::value;
            const int ws = w & (Cfg::W_SLOTS - 1);
            mbar_wait(w_full + 8 * ws, (w / Cfg::W_SLOTS) & 1);
            tc_fence_after();
            const uint32_t tap16 = (tyx / 3) * Cfg::HX + (tyx % 3);  // tap offset in 16-byte units
            const uint32_t bl_u = b_lo0 + ws * (Cfg::W_UNIT_BYTES >> 4);
#pragma unroll
            for (int p = 0; p < Cfg::PLANES; ++p) {
              if constexpr (KIND == 0) {
                mbar_wait((FUSE ? a_ready : a_full) + pl_bar[p], par[p]);
                tc_fence_after();
              }
              constexpr int dummy = 0; (void)dummy;
              const int slab_lo = p >= 2 ? p - 2 : 0, slab_hi = p < ZT ? p : ZT - 1;
              const int nblk = slab_hi - slab_lo + 1;
              const int row0 = (2 - (p - slab_lo)) * 64;
              const uint32_t al_p = pl_lo[p] + tap16, bl_p = bl_u + row0;  // (row0 rows * 16 B) >> 4
#pragma unroll
              for (int k = 0; k < Cfg::KC; ++k) {
                const uint32_t al = al_p + k * 2 * (Cfg::A_LBO >> 4), bl = bl_p + k * 2 * (Cfg::B_LBO >> 4);
                if (KIND == 0 && k == 0 && first_cb) {
                  // first contribution to every slab of this tile: per-slab MMAs so each gets its own accumulate flag
#pragma unroll
                  for (int i = 0; i < nblk; ++i) {
                    const int sl = slab_lo + i, tz = p - sl;
                    umma_bf16_lh(acc + sl * 64, al, a_hi, bl + i * 64, b_hi, make_idesc_16(128, 64, H), tz != 0 ? 1u : 0u);
                  }
                } else {
                  umma_bf16_lh(acc + slab_lo * 64, al, a_hi, bl, b_hi, make_idesc_16(128, 64 * nblk, H), 1u);
                }
              }
              if constexpr (KIND == 2) umma_commit(a_empty + pl_bar[p]);  // last reader of this plane for this block
            }
            umma_commit(w_empty + 8 * ws);
            ++w;
          };
          tap(std::integral_constant<int, 0>{}, 0, cb == 0);
#pragma unroll 1
          for (int tyx = 1; tyx < 8; ++tyx) tap(std::integral_constant<int, 1>{}, tyx, false);
          tap(std::integral_constant<int, 2>{}, 8, false);
        }
        umma_commit(acc_full + 8 * buf);
        ++li;
      });
    }
    __syncwarp();
    pdl_trigger();  // all MMAs of this CTA are issued: only the last epilogue remains
  } else if (FUSE && warp >= 7) {
    // =============================== in-place normalise of arriving halo planes ===============================
    // thread = (8-channel chunk of the block, 1 of LPC lanes walking that chunk's 18 x 10 halo positions): its 8 scales /
    // shifts / biases are reloaded (L1 hits) when the (sample, block) changes
    constexpr int LPC = Cfg::XFORM_WARPS * 32 / Cfg::KCH;  // lanes per chunk
    constexpr int NPOS = Cfg::HY * Cfg::HX;  // 180 halo positions of one chunk plane
    constexpr int NV = (NPOS + LPC - 1) / LPC;  // vectors per thread per plane (6 or 12)
    const int tt = threadIdx.x - 7 * 32, chunk = tt / LPC, l0 = tt % LPC;
    uint8_t* const a_gen = smem_raw + (a_smem - smem_u32(smem_raw)) + chunk * (NPOS * 16) + l0 * 16;
    // the halo positions this thread owns are the same for every plane and tile: i = l0 + j * LPC
    int pos_y[NV], pos_x[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = l0 + j * LPC;
      pos_y[j] = i < NPOS ? i / Cfg::HX : 1 << 20;  // out-of-range positions never pass the bounds test
      pos_x[j] = i % Cfg::HX;
    }
    float sc[8], sh[8], bi[8];
    int u = 0;
    for_each_tile(tiles_per_n, a.batch, [&](int n, int lin) {
      int t = lin;
      const int tix = t % a.tiles_x; t /= a.tiles_x;
      const int tiy = t % a.tiles_y; t /= a.tiles_y;
      const int tiz = t;
      const int x0 = tix * CONV_TX - 1, y0 = tiy * CONV_TY - 1, z0 = tiz * ZT - 1;
      // which of this thread's positions lie inside the volume (the others keep TMA's zero fill = the padding)
      uint32_t mask = 0;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int gy = y0 + pos_y[j], gx = x0 + pos_x[j];
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) mask |= 1u << j;
      }
      for (int cb = 0; cb < ncb; ++cb) {  // FUSE: one source tensor, blocks in channel order
        {
          const int gch = cb * Cfg::KCH + chunk;  // 8-channel chunk of the source tensor
          const float4* src = reinterpret_cast<const float4*>(a.in_affine + ((long long)n * (ncb * Cfg::KCH) + gch) * 16);
          // L2 loads (not the read-only path): the map was written by the kernel this one was pre-launched behind
#ifdef DUNET_AFFINE_LDG  // A/B build switch
          const float4 s0 = __ldg(src), s1 = __ldg(src + 1), h0 = __ldg(src + 2), h1 = __ldg(src + 3);
#else
          const float4 s0 = __ldcg(src), s1 = __ldcg(src + 1), h0 = __ldcg(src + 2), h1 = __ldcg(src + 3);
#endif
          sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
          sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
#pragma unroll
          for (int j = 0; j < 8; ++j) bi[j] = a.in_bias ? __ldg(a.in_bias + (long long)n * a.in_bias_n_stride + gch * 8 + j) : 0.f;
        }
        for (int p = 0; p < Cfg::PLANES; ++p, ++u) {
          const int slot = u % Cfg::A_SLOTS;
          mbar_wait(a_full + 8 * slot, (u / Cfg::A_SLOTS) & 1);
          const int gz = z0 + p;
          if (gz >= 0 && gz < a.D) {
            uint8_t* const pl = a_gen + slot * Cfg::PLANE_BYTES;
            constexpr int G = 6;  // vectors in flight per thread: all loads of a group are issued before the math
#pragma unroll
            for (int j0 = 0; j0 < NV; j0 += G) {
              BF8 v[G];
#pragma unroll
              for (int j = 0; j < G; ++j)
                if (j0 + j < NV && (mask >> (j0 + j) & 1u)) v[j] = *reinterpret_cast<const BF8*>(pl + (j0 + j) * (LPC * 16));
#pragma unroll
              for (int j = 0; j < G; ++j)
                if (j0 + j < NV && (mask >> (j0 + j) & 1u)) {
                  float f[8];
                  bf8_to_float<H>(v[j], f);
                  norm_apply(f, sc, sh, bi, a.slope);
                  *reinterpret_cast<BF8*>(pl + (j0 + j) * (LPC * 16)) = float_to_bf8<H>(f);
                }
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes above -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(a_ready + 8 * slot);
        }
      }
    });
  } else {
    // =============================== epilogue: TMEM -> bf16 -> HBM (+ IN statistics) ===============================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const long long vox = (long long)a.D * a.H * a.W;
    // InstanceNorm statistics: lane l of each warp carries, per 16-column group j, the running (sum | sum of squares)
    // of column j*16 + (l & 15) over this warp's rows of all tiles of the current sample; one row [16] per
    // (sample, chunk, CTA) is written when the sample changes / at the end  ->  nseg = gridDim.x, fixed order.
    float run[4] = {0.f, 0.f, 0.f, 0.f};
    int cur_n = 0;
    auto flush = [&](int n_flush) {
      // combine the four warps in a fixed order and write this CTA's row for sample n_flush
#pragma unroll
      for (int j = 0; j < 4; ++j) red[q * 128 + j * 32 + lane] = run[j];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = r;
      const float tot = (red[e] + red[128 + e]) + (red[256 + e] + red[384 + e]);
      const int l = e & 31, col = (e >> 5) * 16 + (l & 15), stat = l >> 4;
      a.stats[(((long long)n_flush * 8 + (col >> 3)) * gridDim.x + blockIdx.x) * 16 + stat * 8 + (col & 7)] = tot;
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 4; ++j) run[j] = 0.f;
    };
    int li = 0;
    for_each_tile(tiles_per_n, a.batch, [&](int n, int lin) {
      int t = lin;
      const int tix = t % a.tiles_x; t /= a.tiles_x;
      const int tiy = t % a.tiles_y; t /= a.tiles_y;
      const int tiz = t;
      if (a.stats) {
        for (; cur_n < n; ++cur_n) flush(cur_n);  // rows of finished (or skipped) samples
      }
      const int x = tix * CONV_TX + (r & 7), y = tiy * CONV_TY + (r >> 3), z0 = tiz * ZT;
      const bool xy_ok = x < a.W && y < a.H;
      const int buf = li & 1, use = li >> 1;
      mbar_wait(acc_full + 8 * buf, use & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::ACC_COLS;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
#pragma unroll 1
        for (int s = 0; s < ZT; ++s) {
          const int z = z0 + s;
          const bool ok = xy_ok && z < a.D;
          float v[16];
          tmem_ld16(acc + s * 64 + j * 16, v);
          if (ok) {
            const long long o = ((long long)n * 8 + j * 2) * vox + ((long long)z * a.H + y) * a.W + x;
            float c0[8], c1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { c0[i] = v[i]; c1[i] = v[8 + i]; }
            store_split<H>(a.out, a.out_lo, o, c0);
            store_split<H>(a.out, a.out_lo, o + vox, c1);
          }
          if (a.stats) {
            const float m = ok ? 1.f : 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float xv = v[i] * m;
              st[i] += xv;
              st[16 + i] = fmaf(xv, xv, st[16 + i]);
            }
          }
        }
        if (a.stats) run[j] += warp_reduce32(st, lane);
      }
      // all TMEM reads of this accumulator set are done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
      ++li;
    });
    if (a.stats) {
      for (; cur_n < a.batch; ++cur_n) flush(cur_n);
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dunet

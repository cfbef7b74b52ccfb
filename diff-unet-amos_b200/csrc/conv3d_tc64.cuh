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
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

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
  static constexpr int A_SLOTS = PLANES + 1;
  static constexpr int W_SLOTS = 2;
  static constexpr int ACC_COLS = ZT * 64;                  // one accumulator set
  static constexpr int TMEM_COLS = 512;
  static constexpr int RED_BYTES = 4 * 128 * 4;
  static constexpr int SMEM_BYTES = A_SLOTS * PLANE_BYTES + W_SLOTS * W_UNIT_BYTES + RED_BYTES + 1024 + 256;
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

template <int CB_CH, int ZT>
__global__ void __launch_bounds__(CONV_THREADS, 1)
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
  float* red = reinterpret_cast<float*>(smem_raw + (red_smem - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncb = a.segs.ncb;
  const int tiles_per_n = a.tiles_x * a.tiles_y * a.tiles_z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
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
    if (elect_one_sync()) {
      const uint64_t a_desc0 = make_smem_desc(a_smem, Cfg::A_LBO, Cfg::A_SBO);
      const uint64_t b_desc0 = make_smem_desc(w_smem, Cfg::B_LBO, Cfg::B_SBO);
      int u0 = 0, w = 0, li = 0;
      for_each_tile(tiles_per_n, a.batch, [&](int, int) {
        const int buf = li & 1, use = li >> 1;
        if (use > 0) { mbar_wait(acc_empty + 8 * buf, (use - 1) & 1); tc_fence_after(); }
        const uint32_t acc = tmem_base + buf * Cfg::ACC_COLS;
        for (int cb = 0; cb < ncb; ++cb, u0 += Cfg::PLANES) {
          int waited = 0;
#pragma unroll 1
          for (int tyx = 0; tyx < 9; ++tyx, ++w) {
            const int ws = w % Cfg::W_SLOTS;
            mbar_wait(w_full + 8 * ws, (w / Cfg::W_SLOTS) & 1);
            tc_fence_after();
            const uint32_t tap16 = (tyx / 3) * Cfg::HX + (tyx % 3);  // tap offset in 16-byte units
            const uint64_t bd_u = b_desc0 + (uint64_t)(ws * (Cfg::W_UNIT_BYTES >> 4));
#pragma unroll
            for (int p = 0; p < Cfg::PLANES; ++p) {
              if (waited <= p) {
                const int u = u0 + p;
                mbar_wait(a_full + 8 * (u % Cfg::A_SLOTS), (u / Cfg::A_SLOTS) & 1);
                tc_fence_after();
                waited = p + 1;
              }
              constexpr int dummy = 0; (void)dummy;
              const int slab_lo = p >= 2 ? p - 2 : 0, slab_hi = p < ZT ? p : ZT - 1;
              const int nblk = slab_hi - slab_lo + 1;
              const int row0 = (2 - (p - slab_lo)) * 64;
              const int slot = (u0 + p) % Cfg::A_SLOTS;
              const uint64_t ad_p = a_desc0 + (uint64_t)(slot * (Cfg::PLANE_BYTES >> 4) + tap16);
              const uint64_t bd_p = bd_u + (uint64_t)(row0 * 16 >> 4);
#pragma unroll
              for (int k = 0; k < Cfg::KC; ++k) {
                const uint64_t ad = ad_p + (uint64_t)(k * 2 * (Cfg::A_LBO >> 4));
                const uint64_t bd = bd_p + (uint64_t)(k * 2 * (Cfg::B_LBO >> 4));
                if (cb == 0 && tyx == 0 && k == 0) {
                  // first contribution to every slab of this tile: per-slab MMAs so each gets its own accumulate flag
                  for (int i = 0; i < nblk; ++i) {
                    const int s = slab_lo + i, tz = p - s;
                    umma_bf16(acc + s * 64, ad, bd + (uint64_t)(i * 64 * 16 >> 4), make_idesc_bf16(128, 64), tz != 0 ? 1u : 0u);
                  }
                } else {
                  umma_bf16(acc + slab_lo * 64, ad, bd, make_idesc_bf16(128, 64 * nblk), 1u);
                }
              }
              if (tyx == 8) umma_commit(a_empty + 8 * slot);  // last reader of this plane for this cin block
            }
            umma_commit(w_empty + 8 * ws);
          }
        }
        umma_commit(acc_full + 8 * buf);
        ++li;
      });
    }
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
            store_split(a.out, a.out_lo, o, c0);
            store_split(a.out, a.out_lo, o + vox, c1);
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

// 3x3x3 convolution (stride 1, zero padding 1) as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM),
// operands staged by TMA.  Replaces nn.Conv3d inside MONAI Convolution (reference denoiser.py:56-58,
// pretrained/basic_unet.py:59-62).
//
// GEMM view per CTA:  D[128 voxels x N_TILE couts] (x ZT z-slabs)  +=  A[128 x 16] * B[16 x N_TILE]   per tap, per K=16
//
//   * HBM activations are C8-planar bf16: act[n][c/8][z][y][x][c%8].  One TMA box (x:10, y:18, z:1, 8-channel
//     chunks: CB_CH/8) lands a HALO PLANE of a 8x16 output tile in shared memory as [chunk][y][x][8ch] -- exactly the
//     UMMA "K-major, no-swizzle" canonical layout (core matrix = 8 consecutive x voxels x 16 bytes).  Out-of-volume
//     coordinates are zero-filled by TMA, which IS the convolution's zero padding.
//   * A tap (tz,ty,tx) is not a new load: it is the same plane read through a matrix descriptor whose start address
//     is shifted by (ty*10+tx)*16 bytes, and tz picks one of the resident planes.  One halo plane feeds 9 taps x up
//     to 3 z-slabs = 27 MMAs-worth of A reads from shared memory instead of L2.
//   * ZT z-slabs accumulate in ZT TMEM accumulators so each weight tap tile (streamed with cp.async.bulk through a
//     small ring) is reused ZT times.
//   * planes live in a ring of A_SLOTS plane-units; tz-major tap order releases plane 0 after the tz=0 taps, plane 1
//     after tz=1, the rest at the end, so the producer prefetches the next input-channel block under the MMAs.
//   * channel concat (UpCat, denoiser.py:190) is free: input-channel blocks [0,nb0) come from tensor map 0 (skip),
//     the rest from tensor map 1 (upsampled).
//
// The same kernel, instantiated with MODE_DECONV2, is the 2x2x2 stride-2 transposed convolution (MONAI UpSample "deconv",
// denoiser.py:161-170): no halo, one "tap", N = (tap, cout) columns, and an epilogue that adds the bias and scatters
// each column group to its (2z+dz, 2y+dy, 2x+dx) output voxel.
// Deep U-Net levels (few voxels, thousands of input channels) split the input-channel blocks over `ksplit` CTAs that
// write fp32 partial tiles; splitk_reduce_stats_kernel sums them in a fixed order.
// InstanceNorm statistics (sum, sum of squares per channel, from the fp32 accumulators) are reduced in the epilogue
// with a warp butterfly and written as one deterministic partial row per CTA.
//
// Warp roles (7 warps): 0 = plane producer (TMA), 1 = weight producer (bulk copy), 2 = MMA issuer + TMEM owner,
// 3..6 = epilogue (TMEM -> registers -> bf16 -> coalesced 16-byte stores, one TMEM lane quadrant each).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "elementwise.cuh"
#include "ptx.cuh"

namespace dunet {

constexpr int CONV_TX = 8, CONV_TY = 16;  // output tile in x, y  (= 128 GEMM rows)
constexpr int CONV_THREADS = 7 * 32;
// generic kernel: epilogue warps per CTA (4 = one per TMEM lane quadrant, 8 = two, each taking half of the column groups)
#ifndef DUNET_CONV_EPI_WARPS
#define DUNET_CONV_EPI_WARPS 4
#endif
constexpr int CONV_EPI_WARPS = DUNET_CONV_EPI_WARPS;
constexpr int CONV_TC_THREADS = (3 + CONV_EPI_WARPS) * 32;
constexpr int MODE_CONV3 = 0, MODE_DECONV2 = 1;

template <int CB_CH, int N_TILE, int ZT, int MODE>
struct ConvTc {
  static constexpr int HALO = MODE == MODE_CONV3 ? 1 : 0;
  static constexpr int HX = CONV_TX + 2 * HALO, HY = CONV_TY + 2 * HALO;
  static constexpr int TZ = MODE == MODE_CONV3 ? 3 : 1, TYX = MODE == MODE_CONV3 ? 9 : 1, TAPS = TZ * TYX;
  static constexpr int KCH = CB_CH / 8;                     // 16-byte K chunks per input-channel block
  static constexpr int PLANE_BYTES = KCH * HY * HX * 16;    // conv, 64 ch: 23040 B
  static constexpr int A_LBO = HY * HX * 16;                // chunk -> chunk
  static constexpr int A_SBO = HX * 16;                     // y -> y+1 (next 8-row group)
  static constexpr int W_UNIT_BYTES = KCH * N_TILE * 16;    // one (tap, cin block, cout tile) weight tile
  static constexpr int B_LBO = N_TILE * 16;
  static constexpr int B_SBO = 128;
  static constexpr int PLANES = ZT + 2 * HALO;
  // Ring sizes.  Deep U-Net levels are WEIGHT-STREAMING bound: every work item reads 27 weight tiles (16 KB each for
  // 64 x 128) that feed only ZT * CB_CH/16 MMAs, and a bulk copy takes ~2.5k clocks under load, so a 2-slot ring
  // delivered 13 B/clk/SM where the MMAs want 32 (measured with tools/deep_timeline.py: 31k clocks per item against
  // 13.8k of MMA time).  The plane ring keeps one block + 1-2 planes of look-ahead, everything else goes to weights.
  static constexpr int A_SLOTS = MODE == MODE_CONV3 ? (ZT <= 2 ? PLANES + 2 : PLANES + 1) : 2 * PLANES;
  static constexpr int RED_BYTES_ = 4 * N_TILE * 2 * 4;
  static constexpr int W_FIT = (232448 - 1024 - 512 - RED_BYTES_ - A_SLOTS * PLANE_BYTES) / W_UNIT_BYTES;
  static constexpr int W_SLOTS = MODE == MODE_CONV3 ? (W_FIT < 2 ? 2 : (W_FIT > 8 ? 8 : W_FIT))
                                                    : ((32768 / W_UNIT_BYTES) < 2 ? 2 : ((32768 / W_UNIT_BYTES) > 8 ? 8 : (32768 / W_UNIT_BYTES)));
  static constexpr int TMEM_COLS = (ZT * N_TILE <= 32) ? 32 : (ZT * N_TILE <= 64) ? 64 : (ZT * N_TILE <= 128) ? 128
                                   : (ZT * N_TILE <= 256) ? 256 : 512;
  static constexpr int RED_BYTES = 4 * N_TILE * 2 * 4;      // epilogue statistics exchange between the 4 warps
  static constexpr int SMEM_BYTES = A_SLOTS * PLANE_BYTES + W_SLOTS * W_UNIT_BYTES + RED_BYTES + 1024 /*align*/ + 512 /*barriers*/;
  static_assert(16 * A_SLOTS + 16 * W_SLOTS + 40 <= 512, "barrier block");
  static_assert(ZT * N_TILE <= 512, "accumulators exceed TMEM");
  static_assert(A_SLOTS >= PLANES, "ring must hold one input-channel block");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// Input-channel blocks are described by up to 6 SEGMENTS, each a run of `nblocks` consecutive blocks of one source
// tensor (tensor map `tmap` of the kernel's four).  bf16 mode: [skip] or [skip, upsampled] (the UpCat concat).
// fp32x3 mode (split-bf16: every activation and weight is a hi + lo bf16 pair, product = hi*hi + lo*hi + hi*lo):
// [S0.hi, S1.hi | S0.lo, S1.lo | S0.hi, S1.hi] against weight blocks [W.hi | W.hi | W.lo].
struct ConvSeg {
  int tmap, nblocks, chunks;  // chunks = C/8 of the source (stride of the batch index in the folded 4th tensor-map dim)
};
constexpr int CONV_MAX_SEGS = 6;
struct ConvSegs {
  int n, ncb;
  ConvSeg s[CONV_MAX_SEGS];
};
// global input-channel block index -> (tensor map, first 16-byte chunk of the block inside one sample of that source)
__device__ __forceinline__ void conv_seg_lookup(const ConvSegs& sg, int cb, int kch, int& tmap, int& chunk0, int& chunks) {
  int base = 0;
#pragma unroll 1
  for (int i = 0; i < sg.n; ++i) {
    if (cb < base + sg.s[i].nblocks || i == sg.n - 1) {
      tmap = sg.s[i].tmap; chunk0 = (cb - base) * kch; chunks = sg.s[i].chunks;
      return;
    }
    base += sg.s[i].nblocks;
  }
}

struct ConvTcArgs {
  const __nv_bfloat16* w;  // packed [n_tile][cin_block][tap][KCH][N_TILE][8]
  __nv_bfloat16* out;      // bf16 C8-planar output (raw conv output / transposed-conv result), `cout` channels
  __nv_bfloat16* out_lo;   // fp32x3 mode: low part of the output (out + out_lo ~ the fp32 accumulator) or nullptr
  float* out_partial;      // split-K only: fp32 partial tiles [ks][n][cout/8][voxels][8]
  float* stats;            // fused IN statistics [n*cout/8 + chunk][tiles per sample][16] (conv, ksplit == 1) or nullptr
  const float* bias;       // transposed conv: bias[cout]
  ConvSegs segs;           // input-channel block list
  int cout;                // channels of the output tensor (padded)
  int D, H, W;             // INPUT spatial size (the transposed conv writes a 2D x 2H x 2W volume)
  int tiles_x, tiles_y, tiles_z, n_tiles, ksplit, batch;
  long long* dbg;          // optional per-CTA clock64 timeline (8 slots per CTA)
};

// butterfly transpose-reduce: every lane enters with 32 values, lane l leaves with the warp-wide sum of value l
__device__ __forceinline__ float warp_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float keep = up ? v[16 + i] : v[i], send = up ? v[i] : v[16 + i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float keep = up ? v[8 + i] : v[i], send = up ? v[i] : v[8 + i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float keep = up ? v[4 + i] : v[i], send = up ? v[i] : v[4 + i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float keep = up ? v[2 + i] : v[i], send = up ? v[i] : v[2 + i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// decoded work item of the generic kernel: (sample, K split, cout tile, spatial tile), spatial tile fastest so that
// CTAs running side by side share weight tiles and halos in L2
struct ConvItem {
  int n, ks, ntile, tix, tiy, tiz, cb_lo, cb_hi;
};
__device__ __forceinline__ ConvItem conv_item(const ConvTcArgs& a, int item) {
  ConvItem it;
  int t = item;
  it.tix = t % a.tiles_x; t /= a.tiles_x;
  it.tiy = t % a.tiles_y; t /= a.tiles_y;
  it.tiz = t % a.tiles_z; t /= a.tiles_z;
  it.ntile = t % a.n_tiles; t /= a.n_tiles;
  it.ks = t % a.ksplit; t /= a.ksplit;
  it.n = t;
  const int ncb = a.segs.ncb;
  it.cb_lo = (int)((long long)it.ks * ncb / a.ksplit);
  it.cb_hi = (int)((long long)(it.ks + 1) * ncb / a.ksplit);
  return it;
}

// PERSISTENT: gridDim.x <= #SMs CTAs walk the work items round-robin; the plane / weight rings run ahead across items
// and, when 2 * ZT * N_TILE <= 512, the accumulators are double-buffered in TMEM so the epilogue of one item overlaps
// the MMAs of the next.  Statistics rows are per spatial tile, so results do not depend on the item -> CTA assignment.
template <int CB_CH, int N_TILE, int ZT, int MODE, bool H>
__global__ void __launch_bounds__(CONV_TC_THREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                 const __grid_constant__ CUtensorMap tmap2, const __grid_constant__ CUtensorMap tmap3, ConvTcArgs a) {
  using Cfg = ConvTc<CB_CH, N_TILE, ZT, MODE>;
  constexpr int NBUF = (2 * ZT * N_TILE <= 512) ? 2 : 1;
  constexpr int ACC_COLS = ZT * N_TILE;
  constexpr int TMEM_COLS = NBUF == 2 ? 512 : Cfg::TMEM_COLS;
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
  const int total_items = a.tiles_x * a.tiles_y * a.tiles_z * a.n_tiles * a.ksplit * a.batch;

  if (threadIdx.x == 0) {
    if (a.dbg) a.dbg[blockIdx.x * 8 + 0] = clock64();
    for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, CONV_EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
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
  if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * 8 + 1] = clock64();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // =============================== input-plane producer (TMA) ===============================
    if (elect_one_sync()) {
      const CUtensorMap* const tms[4] = {&tmap0, &tmap1, &tmap2, &tmap3};
      int u = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const ConvItem it = conv_item(a, item);
        const int x0 = it.tix * CONV_TX, y0 = it.tiy * CONV_TY, z0 = it.tiz * ZT;
        for (int cb = it.cb_lo; cb < it.cb_hi; ++cb) {
          int ti, chunk0, chunks;
          conv_seg_lookup(a.segs, cb, Cfg::KCH, ti, chunk0, chunks);
          const CUtensorMap* tm = tms[ti];
          const int c3 = it.n * chunks + chunk0;
          for (int p = 0; p < Cfg::PLANES; ++p, ++u) {
            const int slot = u % Cfg::A_SLOTS, rnd = u / Cfg::A_SLOTS;
            if (rnd > 0) mbar_wait(a_empty + 8 * slot, (rnd - 1) & 1);
            mbar_arrive_expect_tx(a_full + 8 * slot, Cfg::PLANE_BYTES);
            tma_load_4d(a_smem + slot * Cfg::PLANE_BYTES, tm, a_full + 8 * slot, (x0 - Cfg::HALO) * 8, y0 - Cfg::HALO,
                        z0 + p - Cfg::HALO, c3);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== weight-tile producer (bulk copy) ===============================
    if (elect_one_sync()) {
      int w = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const ConvItem it = conv_item(a, item);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) +
                              ((size_t)it.ntile * ncb + it.cb_lo) * Cfg::TAPS * Cfg::W_UNIT_BYTES;
        const int nw = (it.cb_hi - it.cb_lo) * Cfg::TAPS;
        for (int i = 0; i < nw; ++i, ++w) {
          const int slot = w % Cfg::W_SLOTS, rnd = w / Cfg::W_SLOTS;
          if (rnd > 0) mbar_wait(w_empty + 8 * slot, (rnd - 1) & 1);
          mbar_arrive_expect_tx(w_full + 8 * slot, Cfg::W_UNIT_BYTES);
          bulk_load_1d(w_smem + slot * Cfg::W_UNIT_BYTES, wsrc + (size_t)i * Cfg::W_UNIT_BYTES, Cfg::W_UNIT_BYTES,
                       w_full + 8 * slot);
        }
      }
    }
  } else if (warp == 2) {
    // =============================== MMA issuer (one thread) ===============================
    // One thread must sustain an MMA every 48-64 clocks: plane descriptors are formed once per block, the tap loop is
    // specialised on tz at compile time and all descriptor arithmetic is 32-bit (only the 14-bit start-address field
    // of the low descriptor word ever changes).
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_16(128, N_TILE, H);
      const uint64_t a_desc0 = make_smem_desc(a_smem, Cfg::A_LBO, Cfg::A_SBO);
      const uint64_t b_desc0 = make_smem_desc(w_smem, Cfg::B_LBO, Cfg::B_SBO);
      const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
      const uint32_t b_lo0 = (uint32_t)b_desc0, b_hi = (uint32_t)(b_desc0 >> 32);
      int ubase = 0, w = 0, li = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
        const ConvItem it = conv_item(a, item);
        const int buf = li % NBUF, use = li / NBUF;
        if (use > 0) { mbar_wait(acc_empty + 8 * buf, (use - 1) & 1); tc_fence_after(); }
        const uint32_t acc = tmem_base + buf * ACC_COLS;
        for (int cb = it.cb_lo; cb < it.cb_hi; ++cb, ubase += Cfg::PLANES) {
          uint32_t pl_lo[Cfg::PLANES], pl_bar[Cfg::PLANES], par[Cfg::PLANES];
#pragma unroll
          for (int p = 0; p < Cfg::PLANES; ++p) {
            const int u = ubase + p, slot = u % Cfg::A_SLOTS;
            pl_lo[p] = a_lo0 + slot * (Cfg::PLANE_BYTES >> 4);
            pl_bar[p] = 8 * slot;
            par[p] = (u / Cfg::A_SLOTS) & 1;
          }
          const bool first_cb = cb == it.cb_lo;
          // one weight tap (tz compile-time, tyx run-time): ZT slabs x CB_CH/16 k-steps
          auto tap = [&](auto tzc, int tyx) {
            constexpr int TZI = decltype(tzc)::value;
            const int ws = w % Cfg::W_SLOTS;
            mbar_wait(w_full + 8 * ws, (w / Cfg::W_SLOTS) & 1);
            tc_fence_after();
            if (tyx == 0) {  // planes first read in this tz phase: all ZT for tz = 0, one more for each later phase
#pragma unroll
              for (int p = (TZI == 0 ? 0 : ZT - 1 + TZI); p < ZT + TZI; ++p) {
                mbar_wait(a_full + pl_bar[p], par[p]);
                tc_fence_after();
              }
            }
            if (a.dbg && w == 0) a.dbg[blockIdx.x * 8 + 3] = clock64();  // first operands have arrived
            const uint32_t tap16 = (tyx / 3) * Cfg::HX + (tyx % 3);  // tap offset in 16-byte units
            const uint32_t bl_u = b_lo0 + ws * (Cfg::W_UNIT_BYTES >> 4);
            const uint32_t acc0 = (first_cb && TZI == 0 && tyx == 0) ? 0u : 1u;  // very first MMA of a slab overwrites
#pragma unroll
            for (int sl = 0; sl < ZT; ++sl) {
              const uint32_t al_p = pl_lo[sl + TZI] + tap16;
#pragma unroll
              for (int k = 0; k < CB_CH / 16; ++k)
                umma_bf16_lh(acc + sl * N_TILE, al_p + k * 2 * (Cfg::A_LBO >> 4), a_hi, bl_u + k * 2 * (Cfg::B_LBO >> 4), b_hi,
                             idesc, k == 0 ? acc0 : 1u);
            }
            umma_commit(w_empty + 8 * ws);  // weight slot free once these MMAs retire
            ++w;
          };
          // planes whose last reader was a tz phase go back to the producer as soon as that phase is issued
#pragma unroll 1
          for (int tyx = 0; tyx < Cfg::TYX; ++tyx) tap(std::integral_constant<int, 0>{}, tyx);
          if constexpr (Cfg::TZ == 3) {
            umma_commit(a_empty + pl_bar[0]);
#pragma unroll 1
            for (int tyx = 0; tyx < Cfg::TYX; ++tyx) tap(std::integral_constant<int, 1>{}, tyx);
            umma_commit(a_empty + pl_bar[1]);
#pragma unroll 1
            for (int tyx = 0; tyx < Cfg::TYX; ++tyx) tap(std::integral_constant<int, 2>{}, tyx);
#pragma unroll
            for (int p = 2; p < Cfg::PLANES; ++p) umma_commit(a_empty + pl_bar[p]);
          } else {
#pragma unroll
            for (int p = 0; p < Cfg::PLANES; ++p) umma_commit(a_empty + pl_bar[p]);
          }
        }
        umma_commit(acc_full + 8 * buf);
      }
      if (a.dbg) a.dbg[blockIdx.x * 8 + 2] = clock64();
    }
    __syncwarp();
    pdl_trigger();  // all MMAs of this CTA are issued: only the last epilogue remains
  } else {
    // =============================== epilogue: TMEM -> registers -> HBM ===============================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int part = (warp - 3) / 4;  // with two warps per quadrant: which half of the column groups this warp handles
    constexpr int J_PER_WARP = (N_TILE / 16) * 4 / CONV_EPI_WARPS;
    const int r = q * 32 + lane;  // GEMM row = voxel (y = r / 8, x = r % 8)
    const int out_chunks = a.cout / 8;
    const long long in_vox = (long long)a.D * a.H * a.W;
    const bool do_stats = (MODE == MODE_CONV3) && a.stats != nullptr;
    int li = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
      const ConvItem it = conv_item(a, item);
      const int n = it.n, ks = it.ks, ntile = it.ntile;
      const int x = it.tix * CONV_TX + (r & 7), y = it.tiy * CONV_TY + (r >> 3), z0 = it.tiz * ZT;
      const bool xy_ok = x < a.W && y < a.H;
      const int buf = li % NBUF, use = li / NBUF;
      mbar_wait(acc_full + 8 * buf, use & 1);
      tc_fence_after();
      if (a.dbg && li == 0 && threadIdx.x == 3 * 32) a.dbg[blockIdx.x * 8 + 4] = clock64();  // first accumulators complete
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS;
#pragma unroll 1
      for (int j = part * J_PER_WARP; j < (part + 1) * J_PER_WARP; ++j) {
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        const int gcol = ntile * N_TILE + j * 16;
#pragma unroll 1
        for (int s = 0; s < ZT; ++s) {
          const int z = z0 + s;
          const bool ok = xy_ok && z < a.D;
          float v[16];
          tmem_ld16(acc + s * N_TILE + j * 16, v);
          if constexpr (MODE == MODE_CONV3) {
            const long long vofs = ((long long)z * a.H + y) * a.W + x;
            if (a.out_partial) {
              if (ok) {
                float4* dst = reinterpret_cast<float4*>(a.out_partial) +
                              ((((long long)ks * a.batch + n) * out_chunks + gcol / 8) * in_vox + vofs) * 2;
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                dst[in_vox * 2] = make_float4(v[8], v[9], v[10], v[11]);
                dst[in_vox * 2 + 1] = make_float4(v[12], v[13], v[14], v[15]);
              }
            } else {
              if (ok) {
                const long long o = ((long long)n * out_chunks + gcol / 8) * in_vox + vofs;
                float c0[8], c1[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { c0[i] = v[i]; c1[i] = v[8 + i]; }
                store_split<H>(a.out, a.out_lo, o, c0);
                store_split<H>(a.out, a.out_lo, o + in_vox, c1);
              }
              if (do_stats) {
                const float m = ok ? 1.f : 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float xv = v[i] * m;
                  st[i] += xv;
                  st[16 + i] = fmaf(xv, xv, st[16 + i]);
                }
              }
            }
          } else {
            if (ok) {
              // column order (dz, dy, 64-channel block, dx, co % 64), see deconv2_tc.cuh / pack_deconv_tc_w_kernel
              const int nblk = a.cout / 64, nt128 = gcol >> 7, within = gcol & 127;
              const int dzdy = nt128 / nblk, tap = dzdy * 2 + (within >> 6), co = (nt128 - dzdy * nblk) * 64 + (within & 63);
              const int oz = 2 * z + (tap >> 2), oy = 2 * y + ((tap >> 1) & 1), ox = 2 * x + (tap & 1);
              const long long ovox = in_vox * 8;
              const long long o = ((long long)n * out_chunks + co / 8) * ovox + ((long long)oz * (2 * a.H) + oy) * (2 * a.W) + ox;
              float c0[8], c1[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) { c0[i] = v[i] + a.bias[co + i]; c1[i] = v[8 + i] + a.bias[co + 8 + i]; }
              store_split<H>(a.out, a.out_lo, o, c0);
              store_split<H>(a.out, a.out_lo, o + ovox, c1);
            }
          }
        }
        if (do_stats) red[q * (N_TILE * 2) + j * 32 + lane] = warp_reduce32(st, lane);
      }
      // all TMEM reads of this accumulator set are done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
      if (do_stats) {
        asm volatile("bar.sync 1, %0;" ::"n"(CONV_EPI_WARPS * 32) : "memory");  // the epilogue warps only
        const int nseg = a.tiles_x * a.tiles_y * a.tiles_z;
        const int tile_lin = (it.tiz * a.tiles_y + it.tiy) * a.tiles_x + it.tix;
        for (int e = part * 128 + r; e < N_TILE * 2; e += CONV_EPI_WARPS * 32) {
          const float tot = (red[e] + red[N_TILE * 2 + e]) + (red[2 * N_TILE * 2 + e] + red[3 * N_TILE * 2 + e]);
          const int l = e & 31, col = (e >> 5) * 16 + (l & 15), stat = l >> 4;
          const int chunk = (ntile * N_TILE + col) >> 3;
          a.stats[(((long long)n * out_chunks + chunk) * nseg + tile_lin) * 16 + stat * 8 + (col & 7)] = tot;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONV_EPI_WARPS * 32) : "memory");  // red[] is reused by the next item
      }
    }
  }
  if (a.dbg && threadIdx.x == 3 * 32) { a.dbg[blockIdx.x * 8 + 5] = clock64(); a.dbg[blockIdx.x * 8 + 6] = (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x; }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace dunet

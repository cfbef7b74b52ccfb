// 3x3x3 convolution of the DEEP U-Net levels (24^3, 12^3, 6^3 of a 96^3 window; Cout a multiple of 128) as a
// swapped-operand, flattened-plane implicit GEMM on tcgen05.  Same layer as conv3d_tc.cuh (reference denoiser.py:56-58,
// pretrained/basic_unet.py:59-62), different decomposition:
//
//   D[128 couts x NPOS positions] (x ZT z-slabs)  +=  A = W[128 couts x 16 cin]  *  B = act[16 cin x NPOS]   per tap, K=16
//
//   * M = output channels (the packed weight tile of conv3d_tc.cuh, [k chunk][128 couts][8 cin], IS a K-major A operand),
//     N = voxels.  N may be any multiple of 16 up to 256, so small volumes do not pay for a fixed 128-row voxel tile: the
//     voxel-as-M kernel fills 75 % of its GEMM rows at 24^3, 50 % at 12^3 and 28 % at 6^3.
//   * N walks the FLATTENED halo plane: a TMA box (x: W+2, y: TY+2, z: 1, 8 chunks) lands in shared memory as
//     [chunk][y][x][8 ch], position p = y * HX + x.  Eight consecutive positions are one core matrix (128 contiguous
//     bytes, SBO = 128), and a tap (ty, tx) is the SAME plane read through a descriptor whose start address is shifted by
//     (ty * HX + tx) * 16 bytes -- for every position at once.  Output column p is voxel (y0 + p / HX, p % HX); the two
//     columns per row with p % HX >= W straddle the halo and are discarded (fill 24/26, 12/14, 6/8).
//   * one streamed 16 KB weight tile feeds ZT slabs x 4 k-steps of N = NPOS MMAs (NPOS >= 128 keeps the MMA off the
//     128 B/clk operand-fetch limit, see tools/umma_probe.cu).
//   * epilogue: a thread owns one output CHANNEL (TMEM lane), so the InstanceNorm statistics are plain per-thread sums (no
//     butterfly); the [channel][position] accumulators are transposed through a small per-warp staging buffer into the
//     C8-planar 16-byte vectors of the activation layout.
//   * split-K (12^3 / 6^3: too few positions to fill the GPU otherwise), in units of 64-channel blocks or of one tz slice
//     of taps, writes fp32 partial tiles [ks][n][cout/8][voxels][8] like the voxel-as-M kernel; splitk_norm_kernel (or
//     splitk_reduce_stats_kernel in split precision) sums them in a fixed order.
//   * accumulators are double-buffered in TMEM when two sets fit, for CTAs that own several items.
//   * transposed-conv mode (ConvTranspose3d k2 s2 with Cin > 128, reference denoiser.py:161-170): one tap, no halo, rows =
//     (tap, cout) of the packed deconv weight tile, epilogue adds the bias and scatters to (2z+dz, 2y+dy, 2x+dx).
//
// Warp roles: 0 = plane producer (TMA), 1 = weight producer (bulk copy), 2 = MMA issuer + TMEM owner, 3..10 = epilogue
// (two warps per TMEM lane quadrant, alternating 16-column groups).
#pragma once
#include "conv3d_tc.cuh"

namespace dunet {

constexpr int FLAT_EPI_WARPS = 8;
constexpr int FLAT_THREADS = (3 + FLAT_EPI_WARPS) * 32;
constexpr int FLAT_W_BYTES = 8 * 128 * 16;       // one (tap, 64-channel block, 128-cout tile) weight tile
constexpr int FLAT_STAGE_CHUNK = 136;            // floats per 8-channel chunk of a warp's staging buffer (16 x 8 + pad)
constexpr int FLAT_STAGE_FLOATS = 4 * FLAT_STAGE_CHUNK;
// epilogue shared memory: position table [256] + statistics exchange [128][2] + one staging buffer per epilogue warp
constexpr int FLAT_EPI_SMEM = 256 * 4 + 128 * 2 * 4 + FLAT_EPI_WARPS * FLAT_STAGE_FLOATS * 4;
constexpr int FLAT_A_SLACK = 256;                // the N round-up may read a few positions past the last plane
constexpr int FLAT_SMEM_MAX = 232448;

struct ConvFlatArgs {
  const __nv_bfloat16* w;  // packed [n_tile][cin_block][tap][8][128][8] (pack_conv_w_kernel, cb_ch = 64, n_tile = 128)
  __nv_bfloat16* out;
  __nv_bfloat16* out_lo;   // split precision: low part of the output, or nullptr
  float* out_partial;      // split-K only: fp32 partial tiles [ks][n][cout/8][voxels][8]
  float* stats;            // IN statistics [n*cout/8 + chunk][tiles_y*tiles_z][16] (ksplit == 1) or nullptr
  ConvSegs segs;
  int cout, D, H, W;
  int hx, ty, tiles_y, tiles_z, n_tiles, ksplit, batch;
  int ksub;                // K units per 64-channel block: 1 (all 27 taps) or 3 (one tz plane of taps each: finer split-K)
  int deconv;              // 1: ConvTranspose3d k2 s2: one tap, no halo, rows = (tap, cout), scatter epilogue + bias
  const float* bias;       // deconv: bias[cout]
  int npos;                // GEMM N: positions per slab (multiple of 16, <= 256)
  int lbo;                 // bytes between the 8-channel chunks of a plane in shared memory: (ty + 2) * hx * 16
  int a_slots, w_slots;    // ring sizes (planes / weight tiles)
  long long* dbg;
};

// TMEM -> registers split into issue and wait, so that the next load is in flight while the previous registers are used.
// The wait takes the registers as read-write operands: nothing that uses them can be scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// K is split in UNITS: a unit is a 64-channel block (27 taps) or, with ksub = 3, one tz slice of it (9 taps, ZT planes)
struct FlatItem {
  int n, ks, ntile, tiy, tiz, u_lo, u_hi;
};
__device__ __forceinline__ FlatItem flat_item(const ConvFlatArgs& a, int item) {
  FlatItem it;
  int t = item;
  it.tiy = t % a.tiles_y; t /= a.tiles_y;
  it.tiz = t % a.tiles_z; t /= a.tiles_z;
  it.ntile = t % a.n_tiles; t /= a.n_tiles;
  it.ks = t % a.ksplit; t /= a.ksplit;
  it.n = t;
  const int nu = a.segs.ncb * a.ksub;
  it.u_lo = (int)((long long)it.ks * nu / a.ksplit);
  it.u_hi = (int)((long long)(it.ks + 1) * nu / a.ksplit);
  return it;
}

template <int ZT, bool H>
__global__ void __launch_bounds__(FLAT_THREADS, 1)
conv3d_flat_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                   const __grid_constant__ CUtensorMap tmap2, const __grid_constant__ CUtensorMap tmap3, ConvFlatArgs a) {
  constexpr int PLANES = ZT + 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int plane_bytes = 8 * a.lbo;
  const uint32_t a_smem = smem_base;
  const uint32_t w_smem = a_smem + a.a_slots * plane_bytes + FLAT_A_SLACK;
  const uint32_t stage_smem = w_smem + a.w_slots * FLAT_W_BYTES;
  const uint32_t bars = stage_smem + FLAT_EPI_SMEM;
  const uint32_t a_full = bars, a_empty = bars + 8 * a.a_slots;
  const uint32_t w_full = bars + 16 * a.a_slots, w_empty = w_full + 8 * a.w_slots;
  const uint32_t acc_full = w_empty + 8 * a.w_slots;   // [2]
  const uint32_t acc_empty = acc_full + 16;              // [2]
  const uint32_t tmem_slot = acc_empty + 16;
  // accumulators double-buffered in TMEM when two sets fit: the epilogue of an item then overlaps the MMAs of the CTA's next
  // item (launches with more items than SMs: 4+ windows per launch, the wide variant)
  const int acc_cols = ZT * a.npos;
  const int nbuf = 2 * acc_cols <= 512 ? 2 : 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncb = a.segs.ncb;
  const int total_items = a.tiles_y * a.tiles_z * a.n_tiles * a.ksplit * a.batch;

  if (threadIdx.x == 0) {
    if (a.dbg) a.dbg[blockIdx.x * 8 + 0] = clock64();
    for (int i = 0; i < a.a_slots; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < a.w_slots; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, FLAT_EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
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
  // The weight producer (warp 1) does NOT wait for the previous kernel: packed weights are written once at checkpoint load,
  // so its first tiles stream in while the previous grid is still draining.  Every other role waits before touching global
  // memory (activations read by TMA, outputs that the previous kernel may still be reading).
  if (warp != 1) pdl_wait();
  if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * 8 + 1] = clock64();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // =============================== input-plane producer (TMA) ===============================
    if (elect_one_sync()) {
      const CUtensorMap* const tms[4] = {&tmap0, &tmap1, &tmap2, &tmap3};
      int slot = 0, rnd = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const FlatItem it = flat_item(a, item);
        const int y0 = it.tiy * a.ty, z0 = it.tiz * ZT;
        for (int u = it.u_lo; u < it.u_hi; ++u) {
          const int cb = a.ksub == 3 ? u / 3 : u, tz = a.ksub == 3 ? u - cb * 3 : 0;
          const int np = (a.ksub == 3 || a.deconv) ? ZT : PLANES;
          const int halo = a.deconv ? 0 : 1;
          int ti, chunk0, chunks;
          conv_seg_lookup(a.segs, cb, 8, ti, chunk0, chunks);
          const CUtensorMap* tm = tms[ti];
          const int c3 = it.n * chunks + chunk0;
          for (int p = 0; p < np; ++p) {
            if (rnd > 0) mbar_wait(a_empty + 8 * slot, (rnd - 1) & 1);
            mbar_arrive_expect_tx(a_full + 8 * slot, plane_bytes);
            tma_load_4d(a_smem + slot * plane_bytes, tm, a_full + 8 * slot, -8 * halo, y0 - halo, z0 + tz + p - halo, c3);
            if (++slot == a.a_slots) { slot = 0; ++rnd; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== weight-tile producer (bulk copy) ===============================
    if (elect_one_sync()) {
      int slot = 0, rnd = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const FlatItem it = flat_item(a, item);
        // units are contiguous in the packed order [cin block][tap]: a unit is 27 / ksub consecutive tap tiles
        const int taps = a.deconv ? 1 : 27, tpu = taps / a.ksub;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) + ((size_t)it.ntile * ncb * taps + (size_t)it.u_lo * tpu) * FLAT_W_BYTES;
        const int nw = (it.u_hi - it.u_lo) * tpu;
        for (int i = 0; i < nw; ++i) {
          if (rnd > 0) mbar_wait(w_empty + 8 * slot, (rnd - 1) & 1);
          mbar_arrive_expect_tx(w_full + 8 * slot, FLAT_W_BYTES);
          bulk_load_1d(w_smem + slot * FLAT_W_BYTES, wsrc + (size_t)i * FLAT_W_BYTES, FLAT_W_BYTES, w_full + 8 * slot);
          if (++slot == a.w_slots) { slot = 0; ++rnd; }
        }
      }
    }
  } else if (warp == 2) {
    // =============================== MMA issuer (one thread) ===============================
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_16(128, a.npos, H);
      const uint64_t w_desc0 = make_smem_desc(w_smem, 128 * 16, 128);   // A: weights, k chunk -> chunk = 2048 B
      const uint64_t p_desc0 = make_smem_desc(a_smem, a.lbo, 128);      // B: flattened plane, 8 positions = 128 B
      const uint32_t w_lo0 = (uint32_t)w_desc0, w_hi = (uint32_t)(w_desc0 >> 32);
      const uint32_t p_lo0 = (uint32_t)p_desc0, p_hi = (uint32_t)(p_desc0 >> 32);
      const uint32_t k_step_p = 2 * (a.lbo >> 4), k_step_w = 2 * (2048 >> 4);
      const uint32_t plane16 = plane_bytes >> 4;
      int aslot = 0, arnd = 0, wslot = 0, wrnd = 0, li = 0;
      bool first_w = true;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
        const FlatItem it = flat_item(a, item);
        const int buf = li % nbuf, use = li / nbuf;
        if (use > 0) { mbar_wait(acc_empty + 8 * buf, (use - 1) & 1); tc_fence_after(); }
        const uint32_t acc_base = tmem_base + buf * acc_cols;
        if (a.ksub == 3 || a.deconv) {
          // one tz slice per unit: ZT planes (slab sl reads plane sl), 9 taps -- or the single tap of the transposed conv
          const int ntap = a.deconv ? 1 : 9;
          for (int u = it.u_lo; u < it.u_hi; ++u) {
            uint32_t pl_lo[ZT], pl_bar[ZT], par[ZT];
#pragma unroll
            for (int p = 0; p < ZT; ++p) {
              pl_lo[p] = p_lo0 + aslot * plane16;
              pl_bar[p] = 8 * aslot;
              par[p] = arnd & 1;
              if (++aslot == a.a_slots) { aslot = 0; ++arnd; }
            }
#pragma unroll 1
            for (int tyx = 0; tyx < ntap; ++tyx) {
              mbar_wait(w_full + 8 * wslot, wrnd & 1);
              tc_fence_after();
              if (tyx == 0) {
#pragma unroll
                for (int p = 0; p < ZT; ++p) {
                  mbar_wait(a_full + pl_bar[p], par[p]);
                  tc_fence_after();
                }
              }
              if (a.dbg && first_w) { a.dbg[blockIdx.x * 8 + 3] = clock64(); first_w = false; }
              const uint32_t tap16 = (tyx / 3) * a.hx + (tyx % 3);
              const uint32_t wl = w_lo0 + wslot * (FLAT_W_BYTES >> 4);
              const uint32_t acc0 = (u == it.u_lo && tyx == 0) ? 0u : 1u;
#pragma unroll
              for (int sl = 0; sl < ZT; ++sl) {
                const uint32_t pl = pl_lo[sl] + tap16;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_lh(acc_base + sl * a.npos, wl + k * k_step_w, w_hi, pl + k * k_step_p, p_hi, idesc, k == 0 ? acc0 : 1u);
              }
              umma_commit(w_empty + 8 * wslot);
              if (++wslot == a.w_slots) { wslot = 0; ++wrnd; }
            }
#pragma unroll
            for (int p = 0; p < ZT; ++p) umma_commit(a_empty + pl_bar[p]);
          }
        } else
        for (int cb = it.u_lo; cb < it.u_hi; ++cb) {
          uint32_t pl_lo[PLANES], pl_bar[PLANES], par[PLANES];
#pragma unroll
          for (int p = 0; p < PLANES; ++p) {
            pl_lo[p] = p_lo0 + aslot * plane16;
            pl_bar[p] = 8 * aslot;
            par[p] = arnd & 1;
            if (++aslot == a.a_slots) { aslot = 0; ++arnd; }
          }
          const bool first_cb = cb == it.u_lo;
          auto tap = [&](auto tzc, int tyx) {
            constexpr int TZI = decltype(tzc)::value;
            mbar_wait(w_full + 8 * wslot, wrnd & 1);
            tc_fence_after();
            if (tyx == 0) {  // planes first read in this tz phase
#pragma unroll
              for (int p = (TZI == 0 ? 0 : ZT - 1 + TZI); p < ZT + TZI; ++p) {
                mbar_wait(a_full + pl_bar[p], par[p]);
                tc_fence_after();
              }
            }
            if (a.dbg && first_w) { a.dbg[blockIdx.x * 8 + 3] = clock64(); first_w = false; }
            const uint32_t tap16 = (tyx / 3) * a.hx + (tyx % 3);
            const uint32_t wl = w_lo0 + wslot * (FLAT_W_BYTES >> 4);
            const uint32_t acc0 = (first_cb && TZI == 0 && tyx == 0) ? 0u : 1u;
#pragma unroll
            for (int sl = 0; sl < ZT; ++sl) {
              const uint32_t pl = pl_lo[sl + TZI] + tap16;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lh(acc_base + sl * a.npos, wl + k * k_step_w, w_hi, pl + k * k_step_p, p_hi, idesc, k == 0 ? acc0 : 1u);
            }
            umma_commit(w_empty + 8 * wslot);
            if (++wslot == a.w_slots) { wslot = 0; ++wrnd; }
          };
#pragma unroll 1
          for (int tyx = 0; tyx < 9; ++tyx) tap(std::integral_constant<int, 0>{}, tyx);
          umma_commit(a_empty + pl_bar[0]);
#pragma unroll 1
          for (int tyx = 0; tyx < 9; ++tyx) tap(std::integral_constant<int, 1>{}, tyx);
          umma_commit(a_empty + pl_bar[1]);
#pragma unroll 1
          for (int tyx = 0; tyx < 9; ++tyx) tap(std::integral_constant<int, 2>{}, tyx);
#pragma unroll
          for (int p = 2; p < PLANES; ++p) umma_commit(a_empty + pl_bar[p]);
        }
        umma_commit(acc_full + 8 * buf);
      }
      if (a.dbg) a.dbg[blockIdx.x * 8 + 2] = clock64();
    }
    __syncwarp();
    pdl_trigger();
  } else {
    // =============================== epilogue: TMEM -> registers -> staging -> HBM ===============================
    // A thread owns one output channel (TMEM lane) and walks 16-column groups = 16 consecutive positions: statistics are
    // plain per-thread sums.  The [channel][position] values are transposed through a per-warp staging buffer so that a
    // lane stores the 16-byte C8 vector of one voxel and a warp writes 256-byte runs (direct 2-byte stores were measured
    // 1.5x slower: four partial 16-byte segments per store instruction).  A position table (voxel offset or -1, built per
    // item while the MMAs run) replaces all index arithmetic, and the TMEM load of the next group is in flight while the
    // current one is stored (two register sets).
    const int ew = warp - 3;
    const int q = warp & 3;        // TMEM lane quadrant this warp may read
    const int part = ew >> 2;      // which of the two warps of the quadrant (alternating column groups)
    int* tab = reinterpret_cast<int*>(smem_raw + (stage_smem - smem_u32(smem_raw)));   // [256] position -> voxel offset in the y-strip, or -1
    float* xch = reinterpret_cast<float*>(tab + 256);                                  // [4 quadrants][32 lanes][2] statistics of the part-1 warps
    float* stage = xch + 256 + ew * FLAT_STAGE_FLOATS;                                 // [4 chunks][16 positions][8 channels] (+ pad)
    const int et = threadIdx.x - 3 * 32;  // 0..255
    const int out_chunks = a.cout / 8;
    const long long in_vox = (long long)a.D * a.H * a.W;
    const int ngroups = a.npos >> 4;
    const int G = ZT * ngroups;
    int li = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
      const FlatItem it = flat_item(a, item);
      const int y0 = it.tiy * a.ty, z0 = it.tiz * ZT;
      const int ty_valid = min(a.ty, a.H - y0);
      {  // position table of this item (overlaps the MMAs)
        const int py = et / a.hx, px = et - py * a.hx;
        const bool ok = et < a.npos && px < a.W && py < ty_valid;
        // conv: voxel offset inside the y-strip; transposed conv: offset of output voxel (2 py, 2 px) in the 2H x 2W slab
        tab[et] = !ok ? -1 : (a.deconv ? (2 * py) * (2 * a.W) + 2 * px : py * a.W + px);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(FLAT_EPI_WARPS * 32) : "memory");
      const int buf = li % nbuf, use = li / nbuf;
      mbar_wait(acc_full + 8 * buf, use & 1);
      tc_fence_after();
      if (a.dbg && li == 0 && threadIdx.x == 3 * 32) a.dbg[blockIdx.x * 8 + 4] = clock64();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + buf * acc_cols;
      float s1 = 0.f, s2 = 0.f;
      // conv: rows of tile nt are couts nt*128..; transposed conv: tile nt = ((dz, dy), 64-cout block), row = dx * 64 + cout % 64
      // (pack_deconv_tc_w_kernel), output volume 2D x 2H x 2W
      const int nblk = a.cout / 64, dzdy = it.ntile / nblk, cblk = it.ntile - dzdy * nblk;
      const long long plane0 = a.deconv ? (long long)it.n * out_chunks + cblk * 8 + (q & 1) * 4
                                        : (long long)it.n * out_chunks + it.ntile * 16 + q * 4;  // C8 plane of this warp's first chunk
      const long long v0 = a.deconv ? ((long long)(2 * z0 + (dzdy >> 1)) * (2 * a.H) + 2 * y0 + (dzdy & 1)) * (2 * a.W) + (q >> 1)
                                    : ((long long)z0 * a.H + y0) * a.W;                           // voxel (z0, y0, 0)
      const long long slab = a.deconv ? (long long)8 * a.H * a.W : (long long)a.H * a.W;          // one input z step
      const long long out_vox = a.deconv ? in_vox * 8 : in_vox;
      const float my_bias = a.deconv ? a.bias[cblk * 64 + (q & 1) * 32 + lane] : 0.f;
      auto process = [&](int g, const uint32_t (&r)[16]) {
        const int s = g / ngroups, j = g - s * ngroups;
        if (z0 + s >= a.D) return;
        const int4* tb = reinterpret_cast<const int4*>(tab + j * 16);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int4 o4 = tb[i4];
          const int o[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const float v = __uint_as_float(r[i4 * 4 + ii]) + my_bias;
            if (o[ii] >= 0) {
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
            stage[(lane >> 3) * FLAT_STAGE_CHUNK + (i4 * 4 + ii) * 8 + (lane & 7)] = v;
          }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int vec = lane + 32 * k, chunk = vec >> 4, pos = vec & 15;
          const int o = tab[j * 16 + pos];
          if (o >= 0) {
            const float4 f0 = *reinterpret_cast<const float4*>(stage + chunk * FLAT_STAGE_CHUNK + pos * 8);
            const float4 f1 = *reinterpret_cast<const float4*>(stage + chunk * FLAT_STAGE_CHUNK + pos * 8 + 4);
            const long long vofs = v0 + s * slab + o;
            if (a.out_partial) {
              float4* dst = reinterpret_cast<float4*>(a.out_partial) +
                            ((((long long)it.ks * a.batch) * out_chunks + plane0 + chunk) * in_vox + vofs) * 2;
              dst[0] = f0;
              dst[1] = f1;
            } else {
              const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
              store_split<H>(a.out, a.out_lo, (plane0 + chunk) * out_vox + vofs, f);
            }
          }
        }
        __syncwarp();
      };
      {
        uint32_t ra[16], rb[16];
        int g = part;
        if (g < G) tmem_ld16_issue(acc + g * 16, ra);
        while (g < G) {
          tmem_ld16_wait(ra);
          if (g + 2 < G) tmem_ld16_issue(acc + (g + 2) * 16, rb);
          process(g, ra);
          g += 2;
          if (g >= G) break;
          tmem_ld16_wait(rb);
          if (g + 2 < G) tmem_ld16_issue(acc + (g + 2) * 16, ra);
          process(g, rb);
          g += 2;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
      // the two warps of a quadrant each hold the sums of their column groups: added in a fixed order (part 0 + part 1)
      if (part == 1) { xch[(q * 32 + lane) * 2] = s1; xch[(q * 32 + lane) * 2 + 1] = s2; }
      asm volatile("bar.sync 1, %0;" ::"n"(FLAT_EPI_WARPS * 32) : "memory");  // also: the table is free for the next item
      if (part == 0 && a.stats != nullptr) {
        s1 += xch[(q * 32 + lane) * 2];
        s2 += xch[(q * 32 + lane) * 2 + 1];
        const int nseg = a.tiles_y * a.tiles_z, tile_lin = it.tiz * a.tiles_y + it.tiy;
        float* row = a.stats + ((plane0 + (lane >> 3)) * nseg + tile_lin) * 16;
        row[lane & 7] = s1;
        row[8 + (lane & 7)] = s2;
      }
    }
  }
  if (a.dbg && threadIdx.x == 3 * 32) { a.dbg[blockIdx.x * 8 + 5] = clock64(); a.dbg[blockIdx.x * 8 + 6] = (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x; }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dunet

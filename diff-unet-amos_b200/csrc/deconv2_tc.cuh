// ConvTranspose3d k=2 s=2 + bias (MONAI UpSample "deconv", reference denoiser.py:161-170,181) for Cin <= 128:
// GEMM  out[v, (tap, co)] = sum_ci in[v, ci] * W[ci, co, tap]  on tcgen05 with a scatter epilogue, persistent CTAs.
//
// The op writes 8x more voxels than it reads and has K = Cin only, so it is bound by the HBM write of its output; what
// matters is that the epilogue (TMEM -> +bias -> bf16 -> 16-byte stores) never waits for loads or MMAs.  One CTA per SM
// walks a static list of (8x16xZT input tile, 128-column block) units; the input planes of a tile are loaded once and
// reused by all its column blocks, weights stream through a small ring, accumulators are double-buffered in TMEM.
// Column order is (dz, dy, 64-channel block, dx, co % 64): every 128-column block holds the dx = 0 and dx = 1 taps of the
// same 64 output channels, so a thread writes 32 contiguous bytes per channel chunk and a warp writes whole 256-byte runs.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv3d_tc.cuh"

namespace dunet {

template <int NCB, int ZT>
struct DeconvTc {
  static constexpr int HX = CONV_TX, HY = CONV_TY;
  static constexpr int PLANE_BYTES = 8 * HY * HX * 16;  // 64 channels of an 8x16 plane: 16 KB
  static constexpr int A_LBO = HY * HX * 16, A_SBO = HX * 16;
  static constexpr int N_TILE = 128;
  static constexpr int W_UNIT_BYTES = 8 * N_TILE * 16;   // 64 input channels x 128 columns: 16 KB
  static constexpr int B_LBO = N_TILE * 16, B_SBO = 128;
  static constexpr int PLANES = NCB * ZT;                // resident planes of one tile
  static constexpr int A_SLOTS = 2 * PLANES;
  static constexpr int W_SLOTS = 3;
  static constexpr int ACC_COLS = ZT * N_TILE;
  static constexpr int SMEM_BYTES = A_SLOTS * PLANE_BYTES + W_SLOTS * W_UNIT_BYTES + 1024 + 256;
  // The epilogue (TMEM -> +bias -> 16-bit -> lane exchange -> 16-byte stores) is one dependent instruction stream per warp:
  // with one epilogue warp per scheduler it took ~19k clocks per tile against the ~11k the HBM write of the tile needs.
  // Two warps per TMEM lane quadrant (each takes half of the column groups) hide that latency; the MMAs of this kernel are
  // tiny, so -- unlike in the conv kernels -- more concurrent TMEM readers do not slow anything down.
#ifndef DUNET_DECONV_EPI_WARPS
#define DUNET_DECONV_EPI_WARPS 8
#endif
  static constexpr int EPI_WARPS = DUNET_DECONV_EPI_WARPS;
  static constexpr int THREADS = (3 + EPI_WARPS) * 32;
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "one or two epilogue warps per TMEM lane quadrant");
  static_assert(2 * ACC_COLS <= 512, "double-buffered accumulators exceed TMEM");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct DeconvTcArgs {
  const __nv_bfloat16* w;  // packed [n_tile][cin block][8 k chunks][128][8], n_tile = (dz*2+dy) * (cout/64) + co/64, column = dx*64 + co%64
  __nv_bfloat16* out;      // C8-planar, cout channels, 2D x 2H x 2W
  const float* bias;       // [cout]
  int chunks_in;           // Cin / 8
  int cout;
  int D, H, W;             // input volume
  int tiles_x, tiles_y, tiles_z, n_tiles, batch;
  long long* dbg;          // optional timeline: per CTA 64 slots
};

template <int NCB, int ZT, bool H>
__global__ void __launch_bounds__(DeconvTc<NCB, ZT>::THREADS, 1)
deconv2_tc_kernel(const __grid_constant__ CUtensorMap tmap, DeconvTcArgs a) {
  using Cfg = DeconvTc<NCB, ZT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem_base;
  const uint32_t w_smem = a_smem + Cfg::A_SLOTS * Cfg::PLANE_BYTES;
  const uint32_t bars = w_smem + Cfg::W_SLOTS * Cfg::W_UNIT_BYTES;
  const uint32_t a_full = bars, a_empty = bars + 8 * Cfg::A_SLOTS;
  const uint32_t w_full = bars + 16 * Cfg::A_SLOTS, w_empty = w_full + 8 * Cfg::W_SLOTS;
  const uint32_t acc_full = w_empty + 8 * Cfg::W_SLOTS, acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.tiles_x * a.tiles_y * a.tiles_z * a.batch;

  if (threadIdx.x == 0) {
    if (a.dbg) { a.dbg[blockIdx.x * 64 + 62] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); a.dbg[blockIdx.x * 64 + 60] = (long long)g; }
    for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, Cfg::EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // =============================== input-plane producer ===============================
    if (elect_one_sync()) {
      int u = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int tix = t % a.tiles_x; t /= a.tiles_x;
        const int tiy = t % a.tiles_y; t /= a.tiles_y;
        const int tiz = t % a.tiles_z; t /= a.tiles_z;
        const int n = t;
        for (int cb = 0; cb < NCB; ++cb)
          for (int s = 0; s < ZT; ++s, ++u) {
            const int slot = u % Cfg::A_SLOTS, it = u / Cfg::A_SLOTS;
            if (it > 0) mbar_wait(a_empty + 8 * slot, (it - 1) & 1);
            mbar_arrive_expect_tx(a_full + 8 * slot, Cfg::PLANE_BYTES);
            tma_load_4d(a_smem + slot * Cfg::PLANE_BYTES, &tmap, a_full + 8 * slot, tix * CONV_TX * 8, tiy * CONV_TY,
                        tiz * ZT + s, n * a.chunks_in + cb * 8);
          }
      }
    }
  } else if (warp == 1) {
    // =============================== weight-tile producer ===============================
    if (elect_one_sync()) {
      int w = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
        for (int i = 0; i < a.n_tiles * NCB; ++i, ++w) {
          const int slot = w % Cfg::W_SLOTS, it = w / Cfg::W_SLOTS;
          if (it > 0) mbar_wait(w_empty + 8 * slot, (it - 1) & 1);
          mbar_arrive_expect_tx(w_full + 8 * slot, Cfg::W_UNIT_BYTES);
          bulk_load_1d(w_smem + slot * Cfg::W_UNIT_BYTES, reinterpret_cast<const uint8_t*>(a.w) + (size_t)i * Cfg::W_UNIT_BYTES,
                       Cfg::W_UNIT_BYTES, w_full + 8 * slot);
        }
    }
  } else if (warp == 2) {
    // =============================== MMA issuer ===============================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_16(128, Cfg::N_TILE, H);
      const uint64_t a_desc0 = make_smem_desc(a_smem, Cfg::A_LBO, Cfg::A_SBO);
      const uint64_t b_desc0 = make_smem_desc(w_smem, Cfg::B_LBO, Cfg::B_SBO);
      int u0 = 0, w = 0, unit = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, u0 += Cfg::PLANES) {
        for (int nt = 0; nt < a.n_tiles; ++nt, ++unit) {
          const int buf = unit & 1, use = unit >> 1;
          if (use > 0) { mbar_wait(acc_empty + 8 * buf, (use - 1) & 1); tc_fence_after(); }
          const uint32_t acc = tmem_base + buf * Cfg::ACC_COLS;
#pragma unroll
          for (int cb = 0; cb < NCB; ++cb, ++w) {
            const int ws = w % Cfg::W_SLOTS;
            mbar_wait(w_full + 8 * ws, (w / Cfg::W_SLOTS) & 1);
            tc_fence_after();
            const uint64_t bd0 = b_desc0 + (uint64_t)(ws * (Cfg::W_UNIT_BYTES >> 4));
#pragma unroll
            for (int s = 0; s < ZT; ++s) {
              const int u = u0 + cb * ZT + s, slot = u % Cfg::A_SLOTS;
              if (nt == 0) {
                mbar_wait(a_full + 8 * slot, (u / Cfg::A_SLOTS) & 1);
                tc_fence_after();
              }
              const uint64_t ad0 = a_desc0 + (uint64_t)(slot * (Cfg::PLANE_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(acc + s * Cfg::N_TILE, ad0 + (uint64_t)(k * 2 * (Cfg::A_LBO >> 4)),
                          bd0 + (uint64_t)(k * 2 * (Cfg::B_LBO >> 4)), idesc, (cb | k) != 0 ? 1u : 0u);
              if (nt == a.n_tiles - 1) umma_commit(a_empty + 8 * slot);  // last column block of this tile
            }
            umma_commit(w_empty + 8 * ws);
          }
          umma_commit(acc_full + 8 * buf);
        }
      }
    }
    __syncwarp();
    pdl_trigger();  // all MMAs of this CTA are issued: only the last epilogue remains
  } else {
    // =============================== epilogue: TMEM -> +bias -> bf16 -> scatter ===============================
    const int q = warp & 3;                            // TMEM lane quadrant this warp may read
    const int part = (warp - 3) / 4;                   // which share of the four 16-column groups this warp stores
    constexpr int JP_PER_WARP = 4 * 4 / Cfg::EPI_WARPS;  // 4 (one warp per quadrant) or 2 (two warps per quadrant)
    const int r = q * 32 + lane;
    const long long in_vox = (long long)a.D * a.H * a.W, ovox = in_vox * 8;
    const int out_chunks = a.cout / 8;
    int unit = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t = tile;
      const int tix = t % a.tiles_x; t /= a.tiles_x;
      const int tiy = t % a.tiles_y; t /= a.tiles_y;
      const int tiz = t % a.tiles_z; t /= a.tiles_z;
      const int n = t;
      const int x = tix * CONV_TX + (r & 7), y = tiy * CONV_TY + (r >> 3), z0 = tiz * ZT;
      const bool xy_ok = x < a.W && y < a.H;
      for (int nt = 0; nt < a.n_tiles; ++nt, ++unit) {
        const int buf = unit & 1, use = unit >> 1;
        mbar_wait(acc_full + 8 * buf, use & 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::ACC_COLS;
        {
          // One 128-column block = [dx = 0 | dx = 1] x 64 channels (block cblk) of the tap pair (dz, dy).  The two x-taps
          // of a voxel are adjacent 16-byte slots of the output row, so the values are exchanged between lanes (shuffles)
          // until every store instruction writes whole contiguous 128-byte runs instead of half sectors.
          const int nblk = a.cout / 64, dzdy = nt / nblk, cblk = nt - dzdy * nblk;
          const int dz = dzdy >> 1, dy = dzdy & 1;
          const int yrow = tiy * CONV_TY + (r >> 3);
          const int oy = 2 * yrow + dy;
#pragma unroll 1
          for (int jp = part * JP_PER_WARP; jp < (part + 1) * JP_PER_WARP; ++jp) {
            float bl[8], bh[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { bl[i] = __ldg(a.bias + cblk * 64 + jp * 16 + i); bh[i] = __ldg(a.bias + cblk * 64 + jp * 16 + 8 + i); }
#pragma unroll
            for (int s = 0; s < ZT; ++s) {
              const int z = z0 + s;
              float v0[16], v1[16];
              tmem_ld16(acc + s * Cfg::N_TILE + jp * 16, v0);
              tmem_ld16(acc + s * Cfg::N_TILE + 64 + jp * 16, v1);
              BF8 pk[2][2];  // [dx][lo | hi chunk]
              {
                float lo[8], hi[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { lo[i] = v0[i] + bl[i]; hi[i] = v0[8 + i] + bh[i]; }
                pk[0][0] = float_to_bf8<H>(lo); pk[0][1] = float_to_bf8<H>(hi);
#pragma unroll
                for (int i = 0; i < 8; ++i) { lo[i] = v1[i] + bl[i]; hi[i] = v1[8 + i] + bh[i]; }
                pk[1][0] = float_to_bf8<H>(lo); pk[1][1] = float_to_bf8<H>(hi);
              }
              const int oz = 2 * z + dz;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int src = (lane & 24) | (4 * h + ((lane & 7) >> 1));
                const int xs = tix * CONV_TX + 4 * h + ((lane & 7) >> 1);       // source voxel x of this slot
                const bool ok = xs < a.W && yrow < a.H && z < a.D;
                const int ox = 2 * tix * CONV_TX + 8 * h + (lane & 7);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const uint4 d0 = *reinterpret_cast<const uint4*>(&pk[0][c]), d1 = *reinterpret_cast<const uint4*>(&pk[1][c]);
                  uint4 e0, e1;
                  e0.x = __shfl_sync(0xffffffffu, d0.x, src); e0.y = __shfl_sync(0xffffffffu, d0.y, src);
                  e0.z = __shfl_sync(0xffffffffu, d0.z, src); e0.w = __shfl_sync(0xffffffffu, d0.w, src);
                  e1.x = __shfl_sync(0xffffffffu, d1.x, src); e1.y = __shfl_sync(0xffffffffu, d1.y, src);
                  e1.z = __shfl_sync(0xffffffffu, d1.z, src); e1.w = __shfl_sync(0xffffffffu, d1.w, src);
                  const uint4 val = (lane & 1) ? e1 : e0;
                  if (ok) {
                    uint4* dst = reinterpret_cast<uint4*>(a.out) + ((long long)n * out_chunks + cblk * 8 + jp * 2 + c) * ovox +
                                 ((long long)oz * (2 * a.H) + oy) * (2 * a.W) + ox;
                    *dst = val;
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
        if (a.dbg && unit < 60 && warp == 3 && lane == 0) a.dbg[blockIdx.x * 64 + unit] = clock64();
      }
    }
  }
  __syncthreads();
  if (a.dbg && threadIdx.x == 0) { a.dbg[blockIdx.x * 64 + 63] = clock64(); unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); a.dbg[blockIdx.x * 64 + 61] = (long long)g; }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dunet

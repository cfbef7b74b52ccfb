// Micro-benchmark: sustained tcgen05.mma rate per SM as a function of N, operand strides and issue pattern.
// Answers: is the 128 x N x 16 MMA with N=64 shared-memory-bandwidth bound (A re-read per MMA), or issue bound?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../diff-unet-amos_b200/csrc/ptx.cuh"
using namespace dunet;

struct Params { int a_sbo; int a_lbo; int iters; int same_a; };

template <int N>
__global__ void __launch_bounds__(128, 1) probe(Params p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + i * 4), "r"(0x3f803f80u));
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 1 && elect_one_sync()) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    constexpr int NACC = 512 / N;
    uint64_t ad[8], bd[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) ad[i] = make_smem_desc(base + (p.same_a ? 0 : i) * 6144, p.a_lbo, p.a_sbo);
#pragma unroll
    for (int i = 0; i < 4; ++i) bd[i] = make_smem_desc(base + 100 * 1024 + i * (N * 32), N * 16, 128);
    long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) umma_bf16(tmem + (j % NACC) * N, ad[j & 7], bd[j & 3], idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// z-stacked tile pattern of conv3d_tc64: per in-plane tap, planes p = 0..5 feed slabs [max(p-2,0), min(p,3)] with one
// N = 64 * nblk MMA per K=16 step; KC k-steps per plane back to back (KC = 4: 64-channel blocks, 2: 32-channel blocks).
// order 0: p outer / k inner (the kernel's order); 1: k outer / p inner; 2: planes interleaved so that consecutive
// MMAs touch disjoint accumulator columns where possible (0,4,1,5,2,3)
template <int KC, int ORDER>
__global__ void __launch_bounds__(128, 1) probe_tile(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + i * 4), "r"(0x3f803f80u));
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 1 && elect_one_sync()) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int perm[6] = {0, 4, 1, 5, 2, 3};
      if (ORDER == 1) {
#pragma unroll
        for (int k = 0; k < KC; ++k)
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const int lo = p >= 2 ? p - 2 : 0, hi = p < 4 ? p : 3, nblk = hi - lo + 1;
            umma_bf16(tmem + lo * 64, make_smem_desc(base + p * 23040 + k * 5760, 2880, 160),
                      make_smem_desc(base + 150 * 1024 + k * 6144, 3072, 128), make_idesc_bf16(128, 64 * nblk), 1u);
          }
      } else {
#pragma unroll
        for (int pp = 0; pp < 6; ++pp) {
          const int p = ORDER == 2 ? perm[pp] : pp;
          const int lo = p >= 2 ? p - 2 : 0, hi = p < 4 ? p : 3, nblk = hi - lo + 1;
#pragma unroll
          for (int k = 0; k < KC; ++k)
            umma_bf16(tmem + lo * 64, make_smem_desc(base + p * 23040 + k * 5760, 2880, 160),
                      make_smem_desc(base + 150 * 1024 + k * 6144, 3072, 128), make_idesc_bf16(128, 64 * nblk), 1u);
        }
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[blockIdx.x] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int KC, int ORDER>
void run_tile(long long* d) {
  cudaFuncSetAttribute(probe_tile<KC, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 512;
  probe_tile<KC, ORDER><<<148, 128, 220 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_tile: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double tot = 0;
  for (int i = 0; i < 148; ++i) tot += h[i];
  const double cyc = tot / 148 / iters, ideal = 416.0 * KC;
  printf("z-stacked tap, KC=%d order %d: %.0f cyc per tap (operand-fetch ideal %.0f => %.0f%%)\n", KC, ORDER, cyc, ideal, 100 * ideal / cyc);
}

template <int N>
void run(const char* name, Params p, long long* d) {
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int grid : {1, 148}) {
    probe<N><<<grid, 128, 220 * 1024>>>(p, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    long long h[296];
    cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    double issue = 0, total = 0;
    for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
    const double n_mma = p.iters * 16.0, cyc = total / grid / n_mma;
    printf("N=%3d %-28s grid %3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f => %.0f%%), smem operand bytes/cyc %.0f\n", N, name,
           grid, issue / grid / n_mma, cyc, 128.0 * N * 16 / 4096.0, 100.0 * (128.0 * N * 16 / 4096.0) / cyc, (4096.0 + N * 32.0) / cyc);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  run<64>("dense SBO128 LBO2048", {128, 2048, 512, 0}, d);
  run<64>("conv layout SBO160 LBO2880", {160, 2880, 512, 0}, d);
  run<64>("conv layout, same A tile", {160, 2880, 512, 1}, d);
  run<128>("dense", {128, 2048, 512, 0}, d);
  run<128>("conv layout", {160, 2880, 512, 0}, d);
  run<192>("conv layout", {160, 2880, 512, 0}, d);
  run<256>("conv layout", {160, 2880, 512, 0}, d);
  run<32>("conv layout", {160, 2880, 512, 0}, d);
  run_tile<4, 0>(d); run_tile<2, 0>(d); run_tile<4, 1>(d); run_tile<2, 1>(d); run_tile<4, 2>(d); run_tile<2, 2>(d);
  run_tile<1, 0>(d); run_tile<8, 0>(d);
  return 0;
}

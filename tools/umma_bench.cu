// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) for N in {64,128,256},
// operands from shared memory (SS) or A from TMEM (TS), back-to-back on one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I visper_lm_b200/csrc -I include tools/umma_bench.cu -o tools/_bin/umma_bench
#include "common.cuh"
#include <cstdio>
#include <vector>
using namespace vpb;
namespace vpb { void set_error(const char*, ...) {} void count_launch(int) {} int get_option(int) { return 0; } }

template <int N, bool TS, bool B_MN>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters, int kdistinct) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (64 + 64) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, B_MN ? 1 : 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 65536;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int k = it % kdistinct;  // walk over k slices like a real K loop
      const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
      const uint64_t bdesc = B_MN ? make_smem_desc(sb + (k % 4) * 2048, 8192, 1024)
                                  : make_smem_desc(sb + (k >> 2) * (N * 128) + (k & 3) * 32, 16, 1024);
      if (TS) umma_bf16_ts(tm, tm + 256 + (k % 8) * 8, bdesc, idesc, 1);
      else umma_bf16(tm, make_smem_desc(sa + oa, 16, 1024), bdesc, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, bool TS, bool B_MN>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8 * grid);
  auto k = bench<N, TS, B_MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 4096;
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, 160 * 1024>>>(d, iters, 8);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, 8 * grid, cudaMemcpyDeviceToHost);
  double s = 0; for (auto v : h) s += v;
  printf("{\"case\": \"%s\", \"N\": %d, \"grid\": %d, \"cycles_per_mma\": %.1f, \"math_floor\": %d, \"err\": \"%s\"}\n",
         name, N, grid, s / grid / iters, N / 2, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, false, false>("SS A K-major, B K-major", grid);
    run<128, false, false>("SS A K-major, B K-major", grid);
    run<256, false, false>("SS A K-major, B K-major", grid);
    run<64, false, true>("SS A K-major, B MN-major", grid);
    run<128, false, true>("SS A K-major, B MN-major", grid);
    run<256, false, true>("SS A K-major, B MN-major", grid);
    run<64, true, false>("TS A in TMEM, B K-major", grid);
    run<128, true, false>("TS A in TMEM, B K-major", grid);
    run<256, true, false>("TS A in TMEM, B K-major", grid);
    run<128, true, true>("TS A in TMEM, B MN-major", grid);
  }
  return 0;
}

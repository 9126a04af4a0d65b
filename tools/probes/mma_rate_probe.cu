// Developer microbenchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of N, of where A comes from (shared memory or
// TMEM) and of the B layout, issued back to back by one thread of one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mapdit_b200/csrc tools/probes/mma_rate_probe.cu -o tools/probes/_bin/mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace tc;

// MODE 0: SS, B K-major; 1: TS, B K-major; 2: SS, B MN-major; 3: TS, B MN-major.  CHAINS: independent accumulators used round-robin
template <int N, int MODE, int CHAINS>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, (MODE & 2) ? 1 : 0);
    const uint64_t d_a = make_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t d_b = (MODE & 2) ? make_smem_desc(smem_u32(smem) + 32768, 8192, 1024) : make_smem_desc(smem_u32(smem) + 32768, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t d = tmem + (k % CHAINS) * 256 / CHAINS * (N <= 128 ? 1 : 0);
        if (MODE & 1) {
          if (leader) umma_ts(d, tmem + 480 + (k & 3) * 8, desc_advance(d_b, (MODE & 2) ? (k & 3) * 2048 : (k & 3) * 32), idesc, 1);
        } else {
          if (leader) umma_ss(d, desc_advance(d_a, (k & 3) * 32), desc_advance(d_b, (MODE & 2) ? (k & 3) * 2048 : (k & 3) * 32), idesc, 1);
        }
      }
    }
    long long t1 = clock64();
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N, int MODE, int CHAINS>
void run(const char* name, long long* d_out) {
  const int iters = 512;
  cudaFuncSetAttribute(probe<N, MODE, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  probe<N, MODE, CHAINS><<<148, 128, 66 * 1024>>>(iters, d_out);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s N %3d chains %d: issue %6.1f cycles/MMA, complete %6.1f cycles/MMA (floor %d)  %s\n", name, N, CHAINS, (double)h[0] / (iters * 8),
         (double)h[1] / (iters * 8), N / 2, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<64, 0, 1>("SS  B K-major", d_out);
  run<64, 0, 2>("SS  B K-major", d_out);
  run<128, 0, 1>("SS  B K-major", d_out);
  run<128, 0, 2>("SS  B K-major", d_out);
  run<256, 0, 1>("SS  B K-major", d_out);
  run<64, 1, 1>("TS  B K-major", d_out);
  run<64, 1, 2>("TS  B K-major", d_out);
  run<128, 1, 1>("TS  B K-major", d_out);
  run<128, 1, 2>("TS  B K-major", d_out);
  run<256, 1, 1>("TS  B K-major", d_out);
  run<64, 2, 1>("SS  B MN-major", d_out);
  run<80, 2, 1>("SS  B MN-major", d_out);
  run<80, 2, 2>("SS  B MN-major", d_out);
  run<64, 3, 1>("TS  B MN-major", d_out);
  run<80, 3, 1>("TS  B MN-major", d_out);
  run<80, 3, 2>("TS  B MN-major", d_out);
  run<128, 3, 1>("TS  B MN-major", d_out);
  return 0;
}

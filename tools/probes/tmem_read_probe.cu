// Developer microbenchmark: TMEM read (tcgen05.ld 32x32b.x32) throughput per SM as a function of the number of reading warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mapdit_b200/csrc tools/probes/tmem_read_probe.cu -o gpurun_out/tmem_probe && gpurun_out/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace tc;

template <int MODE>  // 0: back-to-back loads, wait at the end of each group of 4; 1: load + wait each; 2: loads + 64 FMAs per load
__global__ void __launch_bounds__(512, 1) probe(int nwarps, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t t_lane = base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[32];
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(t_lane + ((i * 4 + c) * 32) % 512, v);
        if (MODE >= 1) tmem_ld_wait();
        if (MODE == 2) {
#pragma unroll
          for (int e = 0; e < 32; ++e) acc = fmaf(__uint_as_float(v[e]), 1.0001f, acc);
        }
      }
      if (MODE == 0) tmem_ld_wait();
      acc += __uint_as_float(v[lane & 31]);
    }
    t1 = clock64();
  }
  if (lane == 0 && warp < nwarps) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(base);
}

template <int MODE>
void run(const char* name, long long* d_out, float* d_sink) {
  const int iters = 256;
  for (int nw : {1, 2, 4, 8, 12, 16}) {
    cudaMemset(d_out, 0, 148 * 16 * 8);
    probe<MODE><<<148, 512>>>(nw, iters, d_out, d_sink);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
    const double bytes = (double)nw * iters * 4 * 4096;
    printf("%s warps %2d: %8lld cycles, %6.1f cycles per 4 KB warp-load, %6.1f B/clk/SM\n", name, nw, mx, (double)mx / (iters * 4), bytes / mx);
  }
}

int main() {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 148 * 16 * 8);
  cudaMalloc(&d_sink, 4);
  run<0>("burst(4)+wait ", d_out, d_sink);
  run<1>("load+wait      ", d_out, d_sink);
  run<2>("load+wait+32fma", d_out, d_sink);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

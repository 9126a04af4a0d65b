// fp32 CUDA-core GEMM for the fp32 parity mode (mode a) and the small conditioning-path GEMMs.
// C[m,n] (+)= sum_k A(m,k) B(n,k) with arbitrary element strides, so forward (x·W^T), dgrad (dy·W)
// and wgrad (dy^T·x) all go through the same kernel without materialising transposes.
// 128x128x16 CTA tile, 8x8 register tile per thread, FFMA accumulation in k order.
#include "common.cuh"

namespace {
constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 1;

// load a (rows x BK) tile of a strided operand into smem as [k][row]
template <bool KFAST>
__device__ __forceinline__ void load_tile(const float* __restrict__ p, int64_t s_row, int64_t s_k, int row0, int k0, int rows,
                                          int K, float (*sm)[LDS_]) {
  const int tid = threadIdx.x;
  if (KFAST) {
    const int kk = tid & 15, r0 = tid >> 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int r = r0 + 16 * j;
      int gr = row0 + r, gk = k0 + kk;
      sm[kk][r] = (gr < rows && gk < K) ? p[gr * s_row + gk * s_k] : 0.f;
    }
  } else {
    const int r = tid & 127, kq = tid >> 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int kk = kq + 2 * j;
      int gr = row0 + r, gk = k0 + kk;
      sm[kk][r] = (gr < rows && gk < K) ? p[gr * s_row + gk * s_k] : 0.f;
    }
  }
}

template <bool AK, bool BKF>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                       const float* __restrict__ B, int64_t sbn, int64_t sbk,
                                                       float* __restrict__ C, int64_t ldc, int M, int N, int K, int accumulate) {
  __shared__ float As[BK][LDS_];
  __shared__ float Bs[BK][LDS_];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each 8x8 (strided by 16)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    load_tile<AK>(A, sam, sak, m0, k0, M, K, As);
    load_tile<BKF>(B, sbn, sbk, n0, k0, N, K, Bs);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + ty + 16 * i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      *c = accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}
}  // namespace

extern "C" int mapdit_gemm_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk, float* c,
                               int64_t ldc, int m, int n, int k, int accumulate, void* stream) {
  MAPDIT_REQUIRE(a && b && c && m > 0 && n > 0 && k > 0, "gemm_f32: bad args");
  dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM);
  cudaStream_t s = (cudaStream_t)stream;
  const bool ak = (sak == 1), bk = (sbk == 1);
  if (ak && bk) gemm_f32_kernel<true, true><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
  else if (ak) gemm_f32_kernel<true, false><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
  else if (bk) gemm_f32_kernel<false, true><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
  else gemm_f32_kernel<false, false><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
  MAPDIT_LAUNCH_CHECK("gemm_f32");
  return MAPDIT_OK;
}

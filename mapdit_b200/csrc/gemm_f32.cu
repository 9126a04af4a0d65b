// fp32 CUDA-core GEMM for the fp32 parity mode (mode a) and the small conditioning-path GEMMs.
// C[m,n] (+)= sum_k A(m,k) B(n,k) with arbitrary element strides, so forward (x·W^T), dgrad (dy·W)
// and wgrad (dy^T·x) all go through the same kernel without materialising transposes.
// T x T x 16 CTA tile (T = 128, 64 or 32), (T/16) x (T/16) register tile per thread, FFMA accumulation in k order.
// The conditioning path multiplies [batch, 256..768] matrices: with 128 x 128 tiles those launches had 2-12 CTAs on 148 SMs and
// sat out 48 k-steps each (0.77 ms of a 43 ms training step in 11 launches); the tile is therefore chosen so the grid covers
// the machine.  Results do not depend on the tile size (every output element is one k-ordered FFMA chain).
#include "common.cuh"

namespace {
constexpr int BK = 16;

// load a (T rows x BK) tile of a strided operand into smem as [k][row]
template <int T, bool KFAST>
__device__ __forceinline__ void load_tile(const float* __restrict__ p, int64_t s_row, int64_t s_k, int row0, int k0, int rows,
                                          int K, float (*sm)[T + 1]) {
  const int tid = threadIdx.x;
  if (KFAST) {
    const int kk = tid & 15, r0 = tid >> 4;
#pragma unroll
    for (int j = 0; j < T / 16; ++j) {
      int r = r0 + 16 * j;
      int gr = row0 + r, gk = k0 + kk;
      sm[kk][r] = (gr < rows && gk < K) ? p[gr * s_row + gk * s_k] : 0.f;
    }
  } else {
    constexpr int KQ = 256 / T;  // k rows covered per pass
    const int r = tid % T, kq = tid / T;
#pragma unroll
    for (int j = 0; j < BK / KQ; ++j) {
      int kk = kq + KQ * j;
      int gr = row0 + r, gk = k0 + kk;
      sm[kk][r] = (gr < rows && gk < K) ? p[gr * s_row + gk * s_k] : 0.f;
    }
  }
}

template <int T, bool AK, bool BKF>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                       const float* __restrict__ B, int64_t sbn, int64_t sbk,
                                                       float* __restrict__ C, int64_t ldc, int M, int N, int K, int accumulate) {
  constexpr int R = T / 16;  // register tile edge
  __shared__ float As[BK][T + 1];
  __shared__ float Bs[BK][T + 1];
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each R x R (strided by 16)
  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    load_tile<T, AK>(A, sam, sak, m0, k0, M, K, As);
    load_tile<T, BKF>(B, sbn, sbk, n0, k0, N, K, Bs);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[R], b[R];
#pragma unroll
      for (int i = 0; i < R; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < R; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    int gm = m0 + ty + 16 * i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      *c = accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

template <int T>
void launch(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk, float* c, int64_t ldc, int m, int n,
            int k, int accumulate, cudaStream_t s) {
  dim3 grid((n + T - 1) / T, (m + T - 1) / T);
  const bool ak = (sak == 1), bk = (sbk == 1);
  if (ak && bk) gemm_f32_kernel<T, true, true><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
  else if (ak) gemm_f32_kernel<T, true, false><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
  else if (bk) gemm_f32_kernel<T, false, true><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
  else gemm_f32_kernel<T, false, false><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate);
}
}  // namespace

extern "C" int mapdit_gemm_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk, float* c,
                               int64_t ldc, int m, int n, int k, int accumulate, void* stream) {
  MAPDIT_REQUIRE(a && b && c && m > 0 && n > 0 && k > 0, "gemm_f32: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  // largest tile whose grid still gives every SM a CTA (148 SMs); small problems take the smallest tile
  auto ctas = [&](int t) { return (long long)((m + t - 1) / t) * ((n + t - 1) / t); };
  if (ctas(128) >= 148) launch<128>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate, s);
  else if (ctas(64) >= 148) launch<64>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate, s);
  else launch<32>(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, accumulate, s);
  MAPDIT_LAUNCH_CHECK("gemm_f32");
  return MAPDIT_OK;
}

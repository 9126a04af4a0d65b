// Generic bf16 flash attention (forward + backward) on warp-level tensor-core MMAs (mma.sync.m16n8k16), for the shapes
// the tcgen05 kernels do not cover: head_dim 72 (DiT-XL: 1152/16, not a UMMA-friendly K or N) and plain dot-product
// attention with unbounded logits (use_cosine_attention=False), tokens a multiple of 64.  Replaces
// F.scaled_dot_product_attention of src/layers/attention.py:47 and its autograd for those shapes; before this file they ran
// on the CUDA-core kernels (attention_f32.cu), which held DiT-XL/2 @ 64x64 at 7 % of its roofline
// (profiles/r1_other_configs.md).  The tcgen05 kernels stay the path for head_dim 64 cosine attention.
//
// One CTA = 64 rows of the "row" operand (4 warps x 16 rows) looping over 64-row blocks of the "column" operand:
//   forward : rows = queries : S = Q K^T -> online softmax (running max, exact for any logits) -> O += P V
//   dq      : rows = queries : S = Q K^T, dP = dO V^T -> dS = P (dP - delta) -> dQ += dS K       (also writes delta)
//   dkv     : rows = keys    : S^T = K Q^T, dP^T = V dO^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
// head_dim is padded to a multiple of 16 with zeros in shared memory (72 -> 80), so the padding never reaches HBM.
#include "common.cuh"

namespace {
constexpr int BM = 64, BN = 64, NT = 128;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// [64 rows][HD channels] bf16 tile from global (row stride ld elements) into smem with pitch P; channels [HD, HDP) zeroed
template <int HD, int HDP, int P>
__device__ __forceinline__ void load_tile(bf16* __restrict__ s, const bf16* __restrict__ g, size_t ld) {
  constexpr int CH = HDP / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < 64 * CH; i += NT) {
    const int r = i / CH, c = (i - r * CH) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c < HD) v = *reinterpret_cast<const uint4*>(g + (size_t)r * ld + c);
    *reinterpret_cast<uint4*>(s + r * P + c) = v;
  }
}

// the same with cp.async (16-byte chunks, zero fill for the padding channels): completes with cp_async_wait
template <int HD, int HDP, int P>
__device__ __forceinline__ void load_tile_async(bf16* __restrict__ s, const bf16* __restrict__ g, size_t ld) {
  constexpr int CH = HDP / 8;
  for (int i = threadIdx.x; i < 64 * CH; i += NT) {
    const int r = i / CH, c = (i - r * CH) * 8;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s + r * P + c);
    const bf16* src = g + (size_t)r * ld + (c < HD ? c : 0);
    const int nbytes = c < HD ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A fragments (16 rows of this warp x HDP) from a [64][P] tile
template <int HDP, int P>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[HDP / 16][4], const bf16* s, int warp, int lane) {
#pragma unroll
  for (int kk = 0; kk < HDP / 16; ++kk) ldsm_x4(a[kk], s + (warp * 16 + (lane & 15)) * P + kk * 16 + (lane >> 4) * 8);
}

// c[nb] (16 x 8 each, 8 blocks = 64 columns) = A[16 x HDP] * B[64 x HDP]^T, B tile row-major [col-row][channel]
template <int HDP, int P>
__device__ __forceinline__ void gemm_abt(float (&c)[8][4], const uint32_t (&a)[HDP / 16][4], const bf16* sb, int lane) {
#pragma unroll
  for (int nb = 0; nb < 8; ++nb)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[nb][e] = 0.f;
  const int m = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < HDP / 16; ++kk)
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // two 8-column blocks per ldmatrix.x4
      uint32_t b[4];
      ldsm_x4(b, sb + (np * 16 + (m >> 1) * 8 + r) * P + kk * 16 + (m & 1) * 8);
      mma16816(c[2 * np], a[kk], b[0], b[1]);
      mma16816(c[2 * np + 1], a[kk], b[2], b[3]);
    }
}

// acc[nb] (16 x 8 each, HDP/8 blocks) += P[16 x 64] * B[64 x HDP], P given in accumulator layout (p[8][4]), B tile [row][channel]
template <int HDP, int P>
__device__ __forceinline__ void gemm_pb(float (&acc)[HDP / 8][4], const float (&p)[8][4], const bf16* sb, int lane) {
  const int m = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {  // 16 rows of B per step
    uint32_t a[4];
    a[0] = pack2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < HDP / 16; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, sb + (kk * 16 + (m & 1) * 8 + r) * P + np * 16 + (m >> 1) * 8);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// store a 16 x HD slice held in accumulator layout as bf16 rows (two rows per thread: g and g + 8), scaled per row
template <int HD, int HDP>
__device__ __forceinline__ void store_rows(bf16* __restrict__ dst, size_t ld, const float (&acc)[HDP / 8][4], float s0, float s1, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nb = 0; nb < HDP / 8; ++nb) {
    const int c = nb * 8 + t * 2;
    if (c < HD) {
      *reinterpret_cast<uint32_t*>(dst + (size_t)g * ld + c) = pack2(acc[nb][0] * s0, acc[nb][1] * s0);
      *reinterpret_cast<uint32_t*>(dst + (size_t)(g + 8) * ld + c) = pack2(acc[nb][2] * s1, acc[nb][3] * s1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int HD, int HDP>
__global__ void __launch_bounds__(NT) attn_mma_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, float* __restrict__ lse,
                                                          int tokens, int heads, float scale_log2) {
  constexpr int P = HDP + 8;
  extern __shared__ __align__(16) uint8_t smem_fwd[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_fwd);
  bf16* sKV = sQ + BM * P;  // stage b: K at sKV + b*2*BN*P, V right after
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
  const size_t D = (size_t)heads * HD, ld = 3 * D;
  const bf16* base = qkv + (size_t)n * tokens * ld + (size_t)h * HD;
  load_tile<HD, HDP, P>(sQ, base + (size_t)q0 * ld, ld);
  load_tile_async<HD, HDP, P>(sKV, base + D, ld);
  load_tile_async<HD, HDP, P>(sKV + BN * P, base + 2 * D, ld);
  cp_async_commit();
  __syncthreads();
  uint32_t aq[HDP / 16][4];
  load_a_frags<HDP, P>(aq, sQ, warp, lane);
  float acc[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows g and g + 8, in log2 units
  for (int k0 = 0, it = 0; k0 < tokens; k0 += BN, ++it) {
    const bf16* sK = sKV + (it & 1) * 2 * BN * P;
    const bf16* sV = sK + BN * P;
    if (k0 + BN < tokens) {  // prefetch the next K/V block into the other stage while this one is consumed
      bf16* nk = sKV + ((it + 1) & 1) * 2 * BN * P;
      load_tile_async<HD, HDP, P>(nk, base + (size_t)(k0 + BN) * ld + D, ld);
      load_tile_async<HD, HDP, P>(nk + BN * P, base + (size_t)(k0 + BN) * ld + 2 * D, ld);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4];
    gemm_abt<HDP, P>(s, aq, sK, lane);
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nb][e] *= scale_log2;
      bm0 = fmaxf(bm0, fmaxf(s[nb][0], s[nb][1]));
      bm1 = fmaxf(bm1, fmaxf(s[nb][2], s[nb][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
    const float c0 = ex2a(m0 - n0), c1 = ex2a(m1 - n1);  // first block: 2^(-inf) = 0
    m0 = n0;
    m1 = n1;
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = ex2a(s[nb][0] - n0);
      s[nb][1] = ex2a(s[nb][1] - n0);
      s[nb][2] = ex2a(s[nb][2] - n1);
      s[nb][3] = ex2a(s[nb][3] - n1);
      r0 += s[nb][0] + s[nb][1];
      r1 += s[nb][2] + s[nb][3];
    }
    l0 = l0 * c0 + r0;
    l1 = l1 * c1 + r1;
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) {
      acc[i][0] *= c0;
      acc[i][1] *= c0;
      acc[i][2] *= c1;
      acc[i][3] *= c1;
    }
    gemm_pb<HDP, P>(acc, s, sV, lane);
    __syncthreads();  // everyone is done with this stage before the next iteration's prefetch overwrites it
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const size_t row = (size_t)n * tokens + q0 + warp * 16;
  store_rows<HD, HDP>(o + row * D + (size_t)h * HD, D, acc, 1.0f / l0, 1.0f / l1, lane);
  if (lse && (lane & 3) == 0) {
    const int g = lane >> 2;
    lse[(row + g) * heads + h] = (m0 + log2f(l0)) * 0.6931471805599453f;
    lse[(row + g + 8) * heads + h] = (m1 + log2f(l1)) * 0.6931471805599453f;
  }
}

// ------------------------------------------------------------------------------------------------ dQ (+ delta)
template <int HD, int HDP>
__global__ void __launch_bounds__(NT) attn_mma_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                         const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                         float* __restrict__ delta, bf16* __restrict__ dqkv, int tokens, int heads,
                                                         float scale) {
  constexpr int P = HDP + 8;
  extern __shared__ __align__(16) uint8_t smem_dq[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_dq);
  bf16* sdO = sQ + BM * P;
  bf16* sKV = sdO + BM * P;  // two stages of {K, V}
  bf16* sK = sKV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2;
  const int n = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
  const size_t D = (size_t)heads * HD, ld = 3 * D;
  const bf16* base = qkv + (size_t)n * tokens * ld + (size_t)h * HD;
  const size_t row = (size_t)n * tokens + q0 + warp * 16;
  load_tile<HD, HDP, P>(sQ, base + (size_t)q0 * ld, ld);
  load_tile<HD, HDP, P>(sdO, dout + ((size_t)n * tokens + q0) * D + (size_t)h * HD, D);
  // sK temporarily holds O for delta_i = dO_i . O_i
  load_tile<HD, HDP, P>(sK, o + ((size_t)n * tokens + q0) * D + (size_t)h * HD, D);
  __syncthreads();
  // delta for this warp's 16 rows: 2 lanes per row
  float dl;
  {
    const int r = warp * 16 + (lane >> 1), c0 = (lane & 1) * (HDP / 2);
    float a = 0.f;
    for (int c = 0; c < HDP / 2; ++c) a = fmaf(__bfloat162float(sdO[r * P + c0 + c]), __bfloat162float(sK[r * P + c0 + c]), a);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    dl = a;  // lanes 2r, 2r+1 hold row r of the warp
    if ((lane & 1) == 0) delta[(row + (lane >> 1)) * heads + h] = a;
  }
  const float d0 = __shfl_sync(0xffffffffu, dl, 2 * g), d1 = __shfl_sync(0xffffffffu, dl, 2 * (g + 8));
  const float L0 = lse[(row + g) * heads + h] * 1.4426950408889634f, L1 = lse[(row + g + 8) * heads + h] * 1.4426950408889634f;
  uint32_t aq[HDP / 16][4], ado[HDP / 16][4];
  load_a_frags<HDP, P>(aq, sQ, warp, lane);
  load_a_frags<HDP, P>(ado, sdO, warp, lane);
  float acc[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  const float sl2 = scale * 1.4426950408889634f;
  __syncthreads();  // the O tile parked in stage 0 has been consumed
  load_tile_async<HD, HDP, P>(sKV, base + D, ld);
  load_tile_async<HD, HDP, P>(sKV + BN * P, base + 2 * D, ld);
  cp_async_commit();
  for (int k0 = 0, it = 0; k0 < tokens; k0 += BN, ++it) {
    const bf16* sKc = sKV + (it & 1) * 2 * BN * P;
    const bf16* sVc = sKc + BN * P;
    if (k0 + BN < tokens) {
      bf16* nk = sKV + ((it + 1) & 1) * 2 * BN * P;
      load_tile_async<HD, HDP, P>(nk, base + (size_t)(k0 + BN) * ld + D, ld);
      load_tile_async<HD, HDP, P>(nk + BN * P, base + (size_t)(k0 + BN) * ld + 2 * D, ld);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4], dp[8][4];
    gemm_abt<HDP, P>(s, aq, sKc, lane);
    gemm_abt<HDP, P>(dp, ado, sVc, lane);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = ex2a(fmaf(s[nb][0], sl2, -L0)) * (dp[nb][0] - d0);
      s[nb][1] = ex2a(fmaf(s[nb][1], sl2, -L0)) * (dp[nb][1] - d0);
      s[nb][2] = ex2a(fmaf(s[nb][2], sl2, -L1)) * (dp[nb][2] - d1);
      s[nb][3] = ex2a(fmaf(s[nb][3], sl2, -L1)) * (dp[nb][3] - d1);
    }
    gemm_pb<HDP, P>(acc, s, sKc, lane);
    __syncthreads();
  }
  store_rows<HD, HDP>(dqkv + row * ld + (size_t)h * HD, ld, acc, scale, scale, lane);
}

// ------------------------------------------------------------------------------------------------ dK, dV
template <int HD, int HDP>
__global__ void __launch_bounds__(NT) attn_mma_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          bf16* __restrict__ dqkv, int tokens, int heads, float scale) {
  constexpr int P = HDP + 8;
  extern __shared__ __align__(16) uint8_t smem_dkv[];
  bf16* sK = reinterpret_cast<bf16*>(smem_dkv);
  bf16* sV = sK + BM * P;
  bf16* sQdO = sV + BM * P;  // two stages of {Q, dO}
  float* sLD = reinterpret_cast<float*>(sQdO + 4 * BN * P);  // two stages of {lse * log2e [64], delta [64]} of the query block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
  const int n = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BM;
  const size_t D = (size_t)heads * HD, ld = 3 * D;
  const bf16* base = qkv + (size_t)n * tokens * ld + (size_t)h * HD;
  load_tile<HD, HDP, P>(sK, base + (size_t)k0 * ld + D, ld);
  load_tile<HD, HDP, P>(sV, base + (size_t)k0 * ld + 2 * D, ld);
  __syncthreads();
  uint32_t ak[HDP / 16][4], av[HDP / 16][4];
  load_a_frags<HDP, P>(ak, sK, warp, lane);
  load_a_frags<HDP, P>(av, sV, warp, lane);
  float dk[HDP / 8][4], dv[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;
  const float sl2 = scale * 1.4426950408889634f;
  auto prefetch = [&](int q0, int stage) {
    bf16* nq = sQdO + stage * 2 * BN * P;
    load_tile_async<HD, HDP, P>(nq, base + (size_t)q0 * ld, ld);
    load_tile_async<HD, HDP, P>(nq + BN * P, dout + ((size_t)n * tokens + q0) * D + (size_t)h * HD, D);
    cp_async_commit();
    if (threadIdx.x < BN) {
      const size_t qr = (size_t)n * tokens + q0 + threadIdx.x;
      sLD[stage * 2 * BN + threadIdx.x] = lse[qr * heads + h] * 1.4426950408889634f;
      sLD[stage * 2 * BN + BN + threadIdx.x] = delta[qr * heads + h];
    }
  };
  prefetch(0, 0);
  for (int q0 = 0, it = 0; q0 < tokens; q0 += BN, ++it) {
    const bf16* sQ = sQdO + (it & 1) * 2 * BN * P;
    const bf16* sdO = sQ + BN * P;
    const float* sL = sLD + (it & 1) * 2 * BN;
    const float* sD = sL + BN;
    if (q0 + BN < tokens) {
      prefetch(q0 + BN, (it + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float s[8][4], dp[8][4];
    gemm_abt<HDP, P>(s, ak, sQ, lane);    // S^T: rows = keys, columns = queries
    gemm_abt<HDP, P>(dp, av, sdO, lane);  // dP^T
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int c = nb * 8 + t * 2;
      const float La = sL[c], Lb = sL[c + 1], Da = sD[c], Db = sD[c + 1];
      const float p0 = ex2a(fmaf(s[nb][0], sl2, -La)), p1 = ex2a(fmaf(s[nb][1], sl2, -Lb));
      const float p2 = ex2a(fmaf(s[nb][2], sl2, -La)), p3 = ex2a(fmaf(s[nb][3], sl2, -Lb));
      s[nb][0] = p0;
      s[nb][1] = p1;
      s[nb][2] = p2;
      s[nb][3] = p3;
      dp[nb][0] = p0 * (dp[nb][0] - Da);
      dp[nb][1] = p1 * (dp[nb][1] - Db);
      dp[nb][2] = p2 * (dp[nb][2] - Da);
      dp[nb][3] = p3 * (dp[nb][3] - Db);
    }
    gemm_pb<HDP, P>(dv, s, sdO, lane);
    gemm_pb<HDP, P>(dk, dp, sQ, lane);
    __syncthreads();
  }
  const size_t row = (size_t)n * tokens + k0 + warp * 16;
  store_rows<HD, HDP>(dqkv + row * ld + D + (size_t)h * HD, ld, dk, scale, scale, lane);
  store_rows<HD, HDP>(dqkv + row * ld + 2 * D + (size_t)h * HD, ld, dv, 1.0f, 1.0f, lane);
}

template <int HD, int HDP>
int launch_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, cudaStream_t s) {
  constexpr int SM_FWD = 5 * 64 * (HDP + 8) * 2;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_mma_fwd_kernel<HD, HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_FWD);
    attr = true;
  }
  dim3 grid(tokens / BM, heads, n);
  attn_mma_fwd_kernel<HD, HDP><<<grid, NT, SM_FWD, s>>>((const bf16*)qkv, (bf16*)o, lse, tokens, heads, 1.4426950408889634f / sqrtf((float)HD));
  return MAPDIT_OK;
}
template <int HD, int HDP>
int launch_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta, int n, int tokens, int heads,
               cudaStream_t s) {
  constexpr int P = HDP + 8;
  constexpr int SM_DQ = 6 * 64 * P * 2, SM_DKV = 6 * 64 * P * 2 + 4 * 64 * 4;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_mma_dq_kernel<HD, HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_DQ);
    cudaFuncSetAttribute(attn_mma_dkv_kernel<HD, HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_DKV);
    attr = true;
  }
  dim3 grid(tokens / BM, heads, n);
  const float scale = 1.0f / sqrtf((float)HD);
  attn_mma_dq_kernel<HD, HDP><<<grid, NT, SM_DQ, s>>>((const bf16*)qkv, (const bf16*)o, (const bf16*)dout, lse, delta, (bf16*)dqkv, tokens,
                                                      heads, scale);
  attn_mma_dkv_kernel<HD, HDP><<<grid, NT, SM_DKV, s>>>((const bf16*)qkv, (const bf16*)dout, lse, delta, (bf16*)dqkv, tokens, heads, scale);
  return MAPDIT_OK;
}
}  // namespace

bool mapdit_attn_mma_supported(int tokens, int hd) { return tokens % 64 == 0 && tokens >= 64 && (hd == 64 || hd == 72 || hd == 32); }

int mapdit_attn_mma_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (hd == 72) launch_fwd<72, 80>(qkv, o, lse, n, tokens, heads, s);
  else if (hd == 64) launch_fwd<64, 64>(qkv, o, lse, n, tokens, heads, s);
  else launch_fwd<32, 32>(qkv, o, lse, n, tokens, heads, s);
  MAPDIT_LAUNCH_CHECK("attn_mma_fwd");
  return MAPDIT_OK;
}
int mapdit_attn_mma_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta, int n, int tokens,
                        int heads, int hd, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (hd == 72) launch_bwd<72, 80>(qkv, o, dout, lse, dqkv, delta, n, tokens, heads, s);
  else if (hd == 64) launch_bwd<64, 64>(qkv, o, dout, lse, dqkv, delta, n, tokens, heads, s);
  else launch_bwd<32, 32>(qkv, o, dout, lse, dqkv, delta, n, tokens, heads, s);
  MAPDIT_LAUNCH_CHECK("attn_mma_bwd");
  mapdit_count_launch();
  return MAPDIT_OK;
}

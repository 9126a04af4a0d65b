// Backward of cosine attention on tcgen05 (head_dim 64, tokens a multiple of 64, bf16 operands, fp32 accumulation).
//
//   logits = q·k/8,  P = exp(logits - L) (L = saved log-sum-exp),  delta_i = dO_i·O_i
//   dV = P^T dO,  dP = dO V^T,  dS = P (dP - delta),  dQ = dS K / 8,  dK = dS^T Q / 8
//
// Two kernels, each shaped like the forward kernel (TMA producer warp, single-thread MMA issuer, four
// softmax warps that own one TMEM lane = one row each):
//   * dq kernel  : CTA = (sample, head, 128 queries); rows = queries.  S = Q K_j^T and dP = dO V_j^T per 64-key block,
//                  dS (bf16) goes to shared memory as a K-major A operand, dQ += dS K_j with K_j re-read MN-major from
//                  the same TMA tile.  Also produces delta for the second kernel.
//   * dkv kernel : CTA = (sample, head, 128 keys); rows = keys.  S^T = K Q_j^T and dP^T = V dO_j^T per 64-query block, so
//                  P^T and dS^T are produced directly in the K-major layout the dV += P^T dO_j and dK += dS^T Q_j MMAs
//                  need; dO_j and Q_j are re-read MN-major from their TMA tiles.  No transposes, no atomics.
#include "tc_common.cuh"

int mapdit_attn_bwd_simt(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta, int n_samples,
                         int tokens, int heads, int head_dim, int dtype, void* stream);

int mapdit_attn_mma_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta, int n, int tokens,
                        int heads, int hd, void* stream);
bool mapdit_attn_mma_supported(int tokens, int hd);

namespace {
using namespace tc;

constexpr int HD = 64, RT = 128, CB = 64;     // row tile (TMEM lanes), column block
constexpr int ROW_BYTES = RT * HD * 2;        // 16 KB  [128 x 64] bf16
constexpr int BLK_BYTES = CB * HD * 2;        // 8 KB   [64 x 64] bf16
constexpr int NTHREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// write 64 bf16 (32 packed words) as row r of a [128 x 64] K-major SWIZZLE_128B tile
__device__ __forceinline__ void store_row_sw128(uint8_t* tile, int r, const uint32_t (&pk)[32]) {
  uint8_t* prow = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}
__device__ __forceinline__ void store_out_row(bf16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32], float sc) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(a[8 * c]) * sc, __uint_as_float(a[8 * c + 1]) * sc);
    u.y = pack_bf16(__uint_as_float(a[8 * c + 2]) * sc, __uint_as_float(a[8 * c + 3]) * sc);
    u.z = pack_bf16(__uint_as_float(a[8 * c + 4]) * sc, __uint_as_float(a[8 * c + 5]) * sc);
    u.w = pack_bf16(__uint_as_float(a[8 * c + 6]) * sc, __uint_as_float(a[8 * c + 7]) * sc);
    *reinterpret_cast<uint4*>(dst + 8 * c) = u;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(b[8 * c]) * sc, __uint_as_float(b[8 * c + 1]) * sc);
    u.y = pack_bf16(__uint_as_float(b[8 * c + 2]) * sc, __uint_as_float(b[8 * c + 3]) * sc);
    u.z = pack_bf16(__uint_as_float(b[8 * c + 4]) * sc, __uint_as_float(b[8 * c + 5]) * sc);
    u.w = pack_bf16(__uint_as_float(b[8 * c + 6]) * sc, __uint_as_float(b[8 * c + 7]) * sc);
    *reinterpret_cast<uint4*>(dst + 32 + 8 * c) = u;
  }
}

// dq / dk epilogue with the backward of the q/k L2 normalisation fused in (src/layers/attention.py:43-45; closed form in
// backward.cu qk_norm_bwd): the thread owns the whole 64-channel gradient row G = acc * att_scale of the NORMALISED head
// y = sqrt(hd) v / (r + eps); with s = sqrt(hd)/(r+eps) saved by the forward, dv = s (G - y (y.G)(r+eps)/(hd r)).
__device__ __forceinline__ void store_out_row_qknorm(bf16* dst, const bf16* __restrict__ yrow, const uint32_t (&a)[32],
                                                     const uint32_t (&b)[32], float att_scale, float s, float eps) {
  float y[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = reinterpret_cast<const uint4*>(yrow)[c];
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = __bfloat1622float2(h2[e]);
      y[8 * c + 2 * e] = t.x;
      y[8 * c + 2 * e + 1] = t.y;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) dot = fmaf(y[c], __uint_as_float(a[c]), fmaf(y[32 + c], __uint_as_float(b[c]), dot));
  dot *= att_scale;
  const float rpe = 8.0f / s;
  const float r = fmaxf(rpe - eps, 1e-30f);
  const float coef = dot * rpe / (64.0f * r);
  const float ga = s * att_scale;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16(fmaf(ga, __uint_as_float(a[8 * c]), -s * coef * y[8 * c]), fmaf(ga, __uint_as_float(a[8 * c + 1]), -s * coef * y[8 * c + 1]));
    u.y = pack_bf16(fmaf(ga, __uint_as_float(a[8 * c + 2]), -s * coef * y[8 * c + 2]), fmaf(ga, __uint_as_float(a[8 * c + 3]), -s * coef * y[8 * c + 3]));
    u.z = pack_bf16(fmaf(ga, __uint_as_float(a[8 * c + 4]), -s * coef * y[8 * c + 4]), fmaf(ga, __uint_as_float(a[8 * c + 5]), -s * coef * y[8 * c + 5]));
    u.w = pack_bf16(fmaf(ga, __uint_as_float(a[8 * c + 6]), -s * coef * y[8 * c + 6]), fmaf(ga, __uint_as_float(a[8 * c + 7]), -s * coef * y[8 * c + 7]));
    *reinterpret_cast<uint4*>(dst + 8 * c) = u;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c]), -s * coef * y[32 + 8 * c]), fmaf(ga, __uint_as_float(b[8 * c + 1]), -s * coef * y[32 + 8 * c + 1]));
    u.y = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c + 2]), -s * coef * y[32 + 8 * c + 2]), fmaf(ga, __uint_as_float(b[8 * c + 3]), -s * coef * y[32 + 8 * c + 3]));
    u.z = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c + 4]), -s * coef * y[32 + 8 * c + 4]), fmaf(ga, __uint_as_float(b[8 * c + 5]), -s * coef * y[32 + 8 * c + 5]));
    u.w = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c + 6]), -s * coef * y[32 + 8 * c + 6]), fmaf(ga, __uint_as_float(b[8 * c + 7]), -s * coef * y[32 + 8 * c + 7]));
    *reinterpret_cast<uint4*>(dst + 32 + 8 * c) = u;
  }
}

// ------------------------------------------------------------------------------------------------ dQ
constexpr int DQ_SMEM = 2 * ROW_BYTES + 2 * 2 * BLK_BYTES + ROW_BYTES + 1024 + 256;

__global__ void __launch_bounds__(NTHREADS, 2)
attn_bwd_dq_tc(const __grid_constant__ CUtensorMap tm_qkv_row, const __grid_constant__ CUtensorMap tm_qkv_blk,
               const __grid_constant__ CUtensorMap tm_do_row, const bf16* __restrict__ o, const bf16* __restrict__ dout,
               const float* __restrict__ lse, float* __restrict__ delta, bf16* __restrict__ dqkv, int tokens, int heads,
               const bf16* __restrict__ qkv, const float* __restrict__ sc, float eps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + ROW_BYTES;
  uint8_t* sKV = sdO + ROW_BYTES;          // stage s: K_j at sKV + s*2*BLK, V_j after it
  uint8_t* sdS = sKV + 2 * 2 * BLK_BYTES;  // [128 x 64] bf16 K-major
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + ROW_BYTES);
  uint64_t* bar_q = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = bars + 3;
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* ds_full = bars + 7;
  uint64_t* ds_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int D = heads * HD, q0 = qt * RT, nkb = tokens / CB, row_base = n * tokens;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv_row);
    prefetch_tmap(&tm_qkv_blk);
    prefetch_tmap(&tm_do_row);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(ds_full, 4);
    mbar_init(ds_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * ROW_BYTES);
    tma_load_2d(sQ, &tm_qkv_row, bar_q, h * HD, row_base + q0);
    tma_load_2d(sdO, &tm_do_row, bar_q, h * HD, row_base + q0);
    for (int j = 0; j < nkb; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
      uint8_t* dst = sKV + s * 2 * BLK_BYTES;
      mbar_arrive_expect_tx(&kv_full[s], 2 * BLK_BYTES);
      tma_load_2d(dst, &tm_qkv_blk, &kv_full[s], D + h * HD, row_base + j * CB);
      tma_load_2d(dst + BLK_BYTES, &tm_qkv_blk, &kv_full[s], 2 * D + h * HD, row_base + j * CB);
    }
  } else if (warp == 1) {
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    constexpr uint32_t idesc_s = make_idesc_bf16(RT, CB, 0, 0);  // S / dP: both operands K-major
    constexpr uint32_t idesc_q = make_idesc_bf16(RT, HD, 0, 1);  // dQ += dS K_j: K_j MN-major (d contiguous)
    const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), ds_addr = smem_u32(sdS);
    auto scores = [&](int j) {
      const int s = j & 1;
      mbar_wait(&kv_full[s], (j >> 1) & 1);
      mbar_wait(s_empty, (j & 1) ^ 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sKV + s * 2 * BLK_BYTES), v_addr = k_addr + BLK_BYTES;
      // the S and dP chains accumulate into different TMEM tiles: issued alternately so consecutive MMAs are independent
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        if (leader) umma_ss(tmem_base, make_smem_desc(q_addr + k * 32, 16, 1024), make_smem_desc(k_addr + k * 32, 16, 1024), idesc_s, k != 0);
        if (leader) umma_ss(tmem_base + 64, make_smem_desc(do_addr + k * 32, 16, 1024), make_smem_desc(v_addr + k * 32, 16, 1024), idesc_s, k != 0);
      }
      if (leader) umma_commit(s_full);
    };
    mbar_wait(bar_q, 0);
    scores(0);
    for (int j = 0; j < nkb; ++j) {
      const int s = j & 1;
      if (j + 1 < nkb) scores(j + 1);
      mbar_wait(ds_full, j & 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sKV + s * 2 * BLK_BYTES);
#pragma unroll
      for (int k = 0; k < CB / 16; ++k)
        if (leader) umma_ss(tmem_base + 128, make_smem_desc(ds_addr + k * 32, 16, 1024), make_smem_desc(k_addr + k * 2048, 1024, 1024), idesc_q,
                (j | k) != 0);
      if (leader) umma_commit(&kv_empty[s]);
      if (leader) umma_commit(ds_empty);
    }
    if (leader) umma_commit(o_full);
  } else if (warp >= 2) {
    const int qq = warp & 3;
    const int r = qq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
    const bool row_ok = q0 + r < tokens;
    const size_t grow = (size_t)row_base + q0 + r;
    // delta_i = dO_i . O_i (128-byte rows straight from global) and L_i
    float dl = 0.f, L = 0.f;
    if (row_ok) {
      const uint4* po = reinterpret_cast<const uint4*>(o + grow * D + h * HD);
      const uint4* pg = reinterpret_cast<const uint4*>(dout + grow * D + h * HD);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 a = po[c], b = pg[c];
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 x = __bfloat1622float2(ha[e]), y = __bfloat1622float2(hb[e]);
          dl = fmaf(x.x, y.x, fmaf(x.y, y.y, dl));
        }
      }
      L = lse[grow * heads + h];
      delta[grow * heads + h] = dl;
    }
    const float c1 = 0.125f * LOG2E, c2 = L * LOG2E;
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t pk[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dp[32];
        tmem_ld32(t_lane + half * 32, sv);
        tmem_ld32(t_lane + 64 + half * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2(fmaf(__uint_as_float(sv[2 * i]), c1, -c2)), p1 = ex2(fmaf(__uint_as_float(sv[2 * i + 1]), c1, -c2));
          float d0 = p0 * (__uint_as_float(dp[2 * i]) - dl), d1 = p1 * (__uint_as_float(dp[2 * i + 1]) - dl);
          pk[half * 16 + i] = pack_bf16(d0, d1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      mbar_wait(ds_empty, (j & 1) ^ 1);
      store_row_sw128(sdS, r, pk);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t a0[32], a1[32];
    tmem_ld32(t_lane + 128, a0);
    tmem_ld32(t_lane + 160, a1);
    tmem_ld_wait();
    if (row_ok) {
      if (sc) store_out_row_qknorm(dqkv + grow * 3 * D + h * HD, qkv + grow * 3 * D + h * HD, a0, a1, 0.125f, sc[grow * 2 * heads + h], eps);
      else store_out_row(dqkv + grow * 3 * D + h * HD, a0, a1, 0.125f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ dK, dV
constexpr int DKV_SMEM = 2 * ROW_BYTES + 2 * 2 * BLK_BYTES + 2 * ROW_BYTES + 2 * 2 * CB * 4 + 1024 + 256;

__global__ void __launch_bounds__(NTHREADS, 2)
attn_bwd_dkv_tc(const __grid_constant__ CUtensorMap tm_qkv_row, const __grid_constant__ CUtensorMap tm_qkv_blk,
                const __grid_constant__ CUtensorMap tm_do_blk, const float* __restrict__ lse, const float* __restrict__ delta,
                bf16* __restrict__ dqkv, int tokens, int heads, const bf16* __restrict__ qkv, const float* __restrict__ sc, float eps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = sK + ROW_BYTES;
  uint8_t* sQdO = sV + ROW_BYTES;            // stage s: Q_j at sQdO + s*2*BLK, dO_j after it
  uint8_t* sPt = sQdO + 2 * 2 * BLK_BYTES;   // [128 keys x 64 queries] bf16 K-major
  uint8_t* sdSt = sPt + ROW_BYTES;
  float* sL = reinterpret_cast<float*>(sdSt + ROW_BYTES);  // [2][64]
  float* sD = sL + 2 * CB;                                  // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * CB);
  uint64_t* bar_kv = bars;
  uint64_t* qd_full = bars + 1;
  uint64_t* qd_empty = bars + 3;
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* p_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int D = heads * HD, k0 = kt * RT, nqb = tokens / CB, row_base = n * tokens;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv_row);
    prefetch_tmap(&tm_qkv_blk);
    prefetch_tmap(&tm_do_blk);
    mbar_init(bar_kv, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * ROW_BYTES);
    tma_load_2d(sK, &tm_qkv_row, bar_kv, D + h * HD, row_base + k0);
    tma_load_2d(sV, &tm_qkv_row, bar_kv, 2 * D + h * HD, row_base + k0);
    for (int j = 0; j < nqb; ++j) {
      const int s = j & 1;
      mbar_wait(&qd_empty[s], ((j >> 1) & 1) ^ 1);
      uint8_t* dst = sQdO + s * 2 * BLK_BYTES;
      mbar_arrive_expect_tx(&qd_full[s], 2 * BLK_BYTES);
      tma_load_2d(dst, &tm_qkv_blk, &qd_full[s], h * HD, row_base + j * CB);
      tma_load_2d(dst + BLK_BYTES, &tm_do_blk, &qd_full[s], h * HD, row_base + j * CB);
    }
  } else if (warp == 1) {
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    constexpr uint32_t idesc_s = make_idesc_bf16(RT, CB, 0, 0);  // S^T = K Q_j^T, dP^T = V dO_j^T
    constexpr uint32_t idesc_a = make_idesc_bf16(RT, HD, 0, 1);  // dV += P^T dO_j, dK += dS^T Q_j : B MN-major
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), pt_addr = smem_u32(sPt), dst_addr = smem_u32(sdSt);
    auto scores = [&](int j) {
      const int s = j & 1;
      mbar_wait(&qd_full[s], (j >> 1) & 1);
      mbar_wait(s_empty, (j & 1) ^ 1);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQdO + s * 2 * BLK_BYTES), do_addr = q_addr + BLK_BYTES;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {  // S^T and dP^T chains interleaved (independent accumulators)
        if (leader) umma_ss(tmem_base, make_smem_desc(k_addr + k * 32, 16, 1024), make_smem_desc(q_addr + k * 32, 16, 1024), idesc_s, k != 0);
        if (leader) umma_ss(tmem_base + 64, make_smem_desc(v_addr + k * 32, 16, 1024), make_smem_desc(do_addr + k * 32, 16, 1024), idesc_s, k != 0);
      }
      if (leader) umma_commit(s_full);
    };
    mbar_wait(bar_kv, 0);
    scores(0);
    for (int j = 0; j < nqb; ++j) {
      const int s = j & 1;
      if (j + 1 < nqb) scores(j + 1);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQdO + s * 2 * BLK_BYTES), do_addr = q_addr + BLK_BYTES;
#pragma unroll
      for (int k = 0; k < CB / 16; ++k) {  // dV and dK chains interleaved
        if (leader) umma_ss(tmem_base + 192, make_smem_desc(pt_addr + k * 32, 16, 1024), make_smem_desc(do_addr + k * 2048, 1024, 1024), idesc_a,
                (j | k) != 0);
        if (leader) umma_ss(tmem_base + 128, make_smem_desc(dst_addr + k * 32, 16, 1024), make_smem_desc(q_addr + k * 2048, 1024, 1024), idesc_a,
                (j | k) != 0);
      }
      if (leader) umma_commit(&qd_empty[s]);
      if (leader) umma_commit(p_empty);
    }
    if (leader) umma_commit(o_full);
  } else if (warp >= 2) {
    const int qq = warp & 3;
    const int r = qq * 32 + lane;  // key row of the tile
    const int tid = threadIdx.x - 64;  // 0..127 among the softmax warps
    const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
    const float c1 = 0.125f * LOG2E;
    for (int j = 0; j < nqb; ++j) {
      const int s = j & 1;
      // per-query L and delta of this block -> smem (double-buffered), visible to the 128 softmax threads
      {
        const size_t qrow = (size_t)row_base + j * CB + (tid & 63);
        if (tid < 64) sL[s * CB + tid] = lse[qrow * heads + h] * LOG2E;
        else sD[s * CB + (tid - 64)] = delta[qrow * heads + h];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t pp[32], pd[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dp[32];
        tmem_ld32(t_lane + half * 32, sv);
        tmem_ld32(t_lane + 64 + half * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = half * 32 + 2 * i;
          float p0 = ex2(fmaf(__uint_as_float(sv[2 * i]), c1, -sL[s * CB + c]));
          float p1 = ex2(fmaf(__uint_as_float(sv[2 * i + 1]), c1, -sL[s * CB + c + 1]));
          float d0 = p0 * (__uint_as_float(dp[2 * i]) - sD[s * CB + c]), d1 = p1 * (__uint_as_float(dp[2 * i + 1]) - sD[s * CB + c + 1]);
          pp[half * 16 + i] = pack_bf16(p0, p1);
          pd[half * 16 + i] = pack_bf16(d0, d1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      mbar_wait(p_empty, (j & 1) ^ 1);
      store_row_sw128(sPt, r, pp);
      store_row_sw128(sdSt, r, pd);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t a0[32], a1[32];
    const bool row_ok = k0 + r < tokens;
    const size_t grow = (size_t)row_base + k0 + r;
    tmem_ld32(t_lane + 128, a0);
    tmem_ld32(t_lane + 160, a1);
    tmem_ld_wait();
    if (row_ok) {
      if (sc) store_out_row_qknorm(dqkv + grow * 3 * D + D + h * HD, qkv + grow * 3 * D + D + h * HD, a0, a1, 0.125f,
                                   sc[grow * 2 * heads + heads + h], eps);
      else store_out_row(dqkv + grow * 3 * D + D + h * HD, a0, a1, 0.125f);
    }
    tmem_ld32(t_lane + 192, a0);
    tmem_ld32(t_lane + 224, a1);
    tmem_ld_wait();
    if (row_ok) store_out_row(dqkv + grow * 3 * D + 2 * D + h * HD, a0, a1, 1.0f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

int encode2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld_elems * 2};
  const uint32_t box[2] = {HD, box_rows};
  return (int)mapdit_encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
}  // namespace

extern "C" int mapdit_qk_norm_bwd(void* dqkv, const void* qkv, const float* sc, int m, int d, int head_dim, float eps, int dtype,
                                  void* stream);

// sc == nullptr: gradients w.r.t. the normalised q^, k^ (and v).  sc != nullptr: the q/k normalisation backward is applied
// as well (fused into the dq / dk epilogues on the tcgen05 path, a separate kernel behind the CUDA-core path).
static int attn_bwd_impl(const void* qkv, const void* o, const void* dout, const float* lse, const float* sc, float eps, void* dqkv,
                         float* delta, int n_samples, int tokens, int heads, int head_dim, int dtype, void* stream) {
  MAPDIT_REQUIRE(qkv && o && dout && lse && dqkv && delta && n_samples > 0 && tokens > 0, "cos_attn_bwd: bad args");
  if (!(dtype == MAPDIT_BF16 && head_dim == HD && tokens % CB == 0) || (mapdit_variant() & MAPDIT_VAR_DOT_ATTN)) {
    int rc = (dtype == MAPDIT_BF16 && mapdit_attn_mma_supported(tokens, head_dim))
                 ? mapdit_attn_mma_bwd(qkv, o, dout, lse, dqkv, delta, n_samples, tokens, heads, head_dim, stream)
                 : mapdit_attn_bwd_simt(qkv, o, dout, lse, dqkv, delta, n_samples, tokens, heads, head_dim, dtype, stream);
    if (rc != MAPDIT_OK || !sc) return rc;
    return mapdit_qk_norm_bwd(dqkv, qkv, sc, n_samples * tokens, heads * head_dim, head_dim, eps, dtype, stream);
  }
  const int D = heads * HD;
  const uint64_t rows = (uint64_t)n_samples * tokens;
  CUtensorMap t_qkv_row, t_qkv_blk, t_do_row, t_do_blk;
  int e = encode2d(&t_qkv_row, qkv, 3 * D, rows, 3 * D, RT) | encode2d(&t_qkv_blk, qkv, 3 * D, rows, 3 * D, CB) |
          encode2d(&t_do_row, dout, D, rows, D, RT) | encode2d(&t_do_blk, dout, D, rows, D, CB);
  if (e != 0) {
    mapdit_set_error("cos_attn_bwd: cuTensorMapEncodeTiled failed");
    return MAPDIT_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      mapdit_set_error("cos_attn_bwd: cudaFuncSetAttribute failed");
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((tokens + RT - 1) / RT, heads, n_samples);
  cudaStream_t s = (cudaStream_t)stream;
  attn_bwd_dq_tc<<<grid, NTHREADS, DQ_SMEM, s>>>(t_qkv_row, t_qkv_blk, t_do_row, (const bf16*)o, (const bf16*)dout, lse, delta,
                                                 (bf16*)dqkv, tokens, heads, (const bf16*)qkv, sc, eps);
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd(dq)");
  attn_bwd_dkv_tc<<<grid, NTHREADS, DKV_SMEM, s>>>(t_qkv_row, t_qkv_blk, t_do_blk, lse, delta, (bf16*)dqkv, tokens, heads,
                                                   (const bf16*)qkv, sc, eps);
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd(dkv)");
  return MAPDIT_OK;
}

extern "C" int mapdit_cos_attn_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta,
                                   int n_samples, int tokens, int heads, int head_dim, int dtype, void* stream) {
  return attn_bwd_impl(qkv, o, dout, lse, nullptr, 0.f, dqkv, delta, n_samples, tokens, heads, head_dim, dtype, stream);
}
extern "C" int mapdit_cos_attn_bwd_qknorm(const void* qkv, const void* o, const void* dout, const float* lse, const float* sc, float eps,
                                          void* dqkv, float* delta, int n_samples, int tokens, int heads, int head_dim, int dtype,
                                          void* stream) {
  MAPDIT_REQUIRE(sc != nullptr, "cos_attn_bwd_qknorm: sc is required");
  return attn_bwd_impl(qkv, o, dout, lse, sc, eps, dqkv, delta, n_samples, tokens, heads, head_dim, dtype, stream);
}

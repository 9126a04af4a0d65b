// Backward of cosine attention on tcgen05 (head_dim 64 or 72, tokens a multiple of 64, bf16 operands, fp32 accumulation).
//
//   logits = q·k/8,  P = exp(logits - L) (L = saved log-sum-exp),  delta_i = dO_i·O_i
//   dV = P^T dO,  dP = dO V^T,  dS = P (dP - delta),  dQ = dS K / 8,  dK = dS^T Q / 8
//
// General path: a coalesced delta kernel, then two kernels shaped like the forward kernel (TMA producer warp, MMA issuer warp,
// eight softmax warps: two per TMEM lane quarter, a thread owns 32 columns of one row):
//   * dq kernel  : CTA = (sample, head, 128 queries); rows = queries.  S = Q K_j^T and dP = dO V_j^T per 64-key block,
//                  dS (bf16) goes back into TMEM as the A operand of dQ += dS K_j, K_j re-read MN-major from its TMA tile.
//   * dkv kernel : CTA = (sample, head, 128 keys); rows = keys.  S^T = K Q_j^T and dP^T = V dO_j^T per 64-query block, so
//                  P^T and dS^T are produced directly in the row = key layout the dV += P^T dO_j and dK += dS^T Q_j MMAs
//                  need (written in place over S^T / dP^T in TMEM); dO_j and Q_j are re-read MN-major from their TMA tiles.
//                  No transposes, no atomics.
//   Every A operand lives in TMEM (the resident Q, dO / K, V tiles are copied there once per CTA): see attn_bwd_dq_tc.
// tokens == 256 at head_dim 64: one fused persistent kernel (attn_bwd_fused_tc and its successors) further down.
#include "tc_common.cuh"

int mapdit_attn_bwd_simt(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta, int n_samples,
                         int tokens, int heads, int head_dim, int dtype, void* stream);

int mapdit_attn_mma_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta, int n, int tokens,
                        int heads, int hd, void* stream);
bool mapdit_attn_mma_supported(int tokens, int hd);

int g_mapdit_attn_bwd_fused = 3;  // mapdit_set_option("attn_bwd_fused", 0..3): tokens == 256: 0 = dq + dkv kernel pair, 1 = fused kernel,
                                  // 2 = fused kernel with a dedicated read-out warpgroup, 3 = that with 64-query half-iterations and
                                  // P^T / dS^T in TMEM

extern long long* g_attn_dbg;  // developer timeline hook (mapdit_attn_debug_buffer)

namespace {
using namespace tc;

constexpr int HD = 64, RT = 128, CB = 64;     // row tile (TMEM lanes), column block
constexpr int ROW_BYTES = RT * HD * 2;        // 16 KB  [128 x 64] bf16
constexpr int BLK_BYTES = CB * HD * 2;        // 8 KB   [64 x 64] bf16
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
// Where a finished gradient row of the dq / dkv kernels goes: row r of a shared-memory tile laid out like the TMA operand tiles (64-channel
// SWIZZLE_128B panels; chunk = 8 channels = 16 bytes), from which each warp's 32-row slab leaves by TMA store.  (Row-per-thread 16-byte
// global stores are 32 line transactions per instruction: the read-out was bound by them.)
struct RowOut {
  uint8_t* p0;  // row r of panel 0
  uint8_t* p1;  // row r of panel 1 (head_dim 72: channels 64..71)
  int sw;       // r & 7
  __device__ __forceinline__ uint4* chunk(int c) const { return reinterpret_cast<uint4*>((c < 8 ? p0 : p1) + (((c & 7) ^ sw) << 4)); }
};
__device__ __forceinline__ void store_out_row(const RowOut& dst, const uint32_t (&a)[32], const uint32_t (&b)[32], float sc) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(a[8 * c]) * sc, __uint_as_float(a[8 * c + 1]) * sc);
    u.y = pack_bf16(__uint_as_float(a[8 * c + 2]) * sc, __uint_as_float(a[8 * c + 3]) * sc);
    u.z = pack_bf16(__uint_as_float(a[8 * c + 4]) * sc, __uint_as_float(a[8 * c + 5]) * sc);
    u.w = pack_bf16(__uint_as_float(a[8 * c + 6]) * sc, __uint_as_float(a[8 * c + 7]) * sc);
    *dst.chunk(c) = u;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(b[8 * c]) * sc, __uint_as_float(b[8 * c + 1]) * sc);
    u.y = pack_bf16(__uint_as_float(b[8 * c + 2]) * sc, __uint_as_float(b[8 * c + 3]) * sc);
    u.z = pack_bf16(__uint_as_float(b[8 * c + 4]) * sc, __uint_as_float(b[8 * c + 5]) * sc);
    u.w = pack_bf16(__uint_as_float(b[8 * c + 6]) * sc, __uint_as_float(b[8 * c + 7]) * sc);
    *dst.chunk(4 + c) = u;
  }
}

// dq / dk epilogue with the backward of the q/k L2 normalisation fused in (src/layers/attention.py:43-45; closed form in
// backward.cu qk_norm_bwd): the thread owns the whole 64-channel gradient row G = acc * att_scale of the NORMALISED head
// y = sqrt(hd) v / (r + eps); with s = sqrt(hd)/(r+eps) saved by the forward, dv = s (G - y (y.G)(r+eps)/(hd r)).
__device__ __forceinline__ void store_out_row_qknorm(const RowOut& dst, const bf16* __restrict__ yrow, const uint32_t (&a)[32],
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
    *dst.chunk(c) = u;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c]), -s * coef * y[32 + 8 * c]), fmaf(ga, __uint_as_float(b[8 * c + 1]), -s * coef * y[32 + 8 * c + 1]));
    u.y = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c + 2]), -s * coef * y[32 + 8 * c + 2]), fmaf(ga, __uint_as_float(b[8 * c + 3]), -s * coef * y[32 + 8 * c + 3]));
    u.z = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c + 4]), -s * coef * y[32 + 8 * c + 4]), fmaf(ga, __uint_as_float(b[8 * c + 5]), -s * coef * y[32 + 8 * c + 5]));
    u.w = pack_bf16(fmaf(ga, __uint_as_float(b[8 * c + 6]), -s * coef * y[32 + 8 * c + 6]), fmaf(ga, __uint_as_float(b[8 * c + 7]), -s * coef * y[32 + 8 * c + 7]));
    *dst.chunk(4 + c) = u;
  }
}

// ------------------------------------------------------------------------------------------------ dQ / dK, dV kernel pair
// Head dimension 64 or 72 (DiT-XL).  At 72 every operand tile is TWO 64-channel SWIZZLE_128B panels: the tensor maps are 3-D
// {72 channels, heads, rows}, the box that starts at channel 64 reads channels 64..71 and the TMA unit zero-fills the rest, so in
// shared memory a head is 128 channels wide with zeros behind channel 72 (no padded copy in HBM).  Contractions over channels
// (S, dP) take five K = 16 steps, accumulators over channels (dQ, dK, dV) are 80 columns wide (64 + the first 16 of the second
// panel through the descriptor's leading-dimension stride); columns 72..79 are exact zeros and are not stored.
template <int HDV>
struct BCfg {
  static constexpr int PANELS = HDV == 64 ? 1 : 2;
  static constexpr int KST = (HDV + 15) / 16;            // channel k-steps of S / dP: 4 | 5
  static constexpr int NACC = HDV == 64 ? 64 : 80;       // accumulator columns of dQ / dK / dV
  static constexpr int PROW = RT * 128, PBLK = CB * 128;  // one panel of a 128-row tile / a 64-row block
  static constexpr int ROWB = PROW * PANELS, BLKB = PBLK * PANELS;
  // One CTA per SM.  Ring of column-block stages (K_j | V_j, or Q_j | dO_j): with two stages a block's tiles were requested only one
  // block ahead and ncu showed 37 % of the samples in the softmax warps' wait for S
  // (the dq kernel holds K_j until dQ(j), which is issued behind S / dP of block j+2: three live blocks + two requested ahead)
  static constexpr int DQ_NST = 5, DKV_NST = 4;
  // softmax warps: two per TMEM lane quarter (= per SM sub-partition, so that one hides the tcgen05.ld / MUFU latency of the other), each
  // thread 32 of the 64 columns of its row of the logit block
  static constexpr int NSW = 8;
  static constexpr int NTHR = 64 + NSW * 32;
  static constexpr int DQ_SMEM = 2 * ROWB + DQ_NST * 2 * BLKB + 1024 + 256;
  static constexpr int DKV_SMEM = 2 * ROWB + DKV_NST * 2 * BLKB + 2 * 2 * CB * 4 + 1024 + 256;
  // TMEM: two (S [0, 64) | dP [64, 128)) buffers, then the accumulators (dQ, or dK and dV).  scores(j+1) runs on the tensor core while
  // the softmax warps still work on block j.
  static constexpr uint32_t T_ACC = 256;
  // behind the accumulators: the CTA's two resident [128 x head_dim] tiles (Q and dO, or K and V) as TMEM A operands, KST x 8 columns
  // of packed bf16 pairs each (see tile_to_tmem)
  static constexpr uint32_t A_COLS = KST * 8;
  static constexpr uint32_t TMEM = 512;
};
__host__ __device__ constexpr float att_scale_of(int hdv) { return hdv == 64 ? 0.125f : 0.11785113019775793f; }  // 1/sqrt(hd)
__host__ __device__ constexpr float sqrt_hd_of(int hdv) { return hdv == 64 ? 8.0f : 8.48528137423857f; }

// one operand tile (PANELS x [rows x 64]) by TMA.  HDV 64: 2-D map {3D | D columns, rows}, column = part * D + h * 64.  HDV 72: 3-D
// map {72, heads of the tensor, rows}, head index = part * heads + h, one load per panel.
template <int HDV>
__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int part, int h, int heads, int row, int panel_bytes) {
  if constexpr (HDV == 64) {
    tma_load_2d(dst, m, bar, part * heads * 64 + h * 64, row);
  } else {
#pragma unroll
    for (int p = 0; p < 2; ++p) tma_load_3d(dst + p * panel_bytes, m, bar, 64 * p, part * heads + h, row);
  }
}
// one warp's 32-row slab (rows 32 qq ..) of a staging tile -> global by TMA store, one bulk group.  The map's box is {64 channels, (1
// head,) 32 rows}; at head_dim 72 the second panel's box starts at channel 64 and the store clips it at the head's 72 channels.
template <int HDV>
__device__ __forceinline__ void store_slab(const CUtensorMap* m, const uint8_t* tile, int qq, int part, int h, int heads, int row, int panel_bytes) {
  if constexpr (HDV == 64) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m),
                 "r"(smem_u32(tile + qq * 4096)), "r"(part * heads * 64 + h * 64), "r"(row)
                 : "memory");
  } else {
#pragma unroll
    for (int p = 0; p < 2; ++p)
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)m),
                   "r"(smem_u32(tile + p * panel_bytes + qq * 4096)), "r"(64 * p), "r"(part * heads + h), "r"(row)
                   : "memory");
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// accumulator row (a | b | tail: NACC fp32) of thread = row -> HDV bf16, scaled
template <int HDV>
__device__ __forceinline__ void store_acc_row(const RowOut& dst, const uint32_t (&a)[32], const uint32_t (&b)[32], const uint32_t (&t)[8], float sc) {
  store_out_row(dst, a, b, sc);
  if constexpr (HDV != 64) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(t[0]) * sc, __uint_as_float(t[1]) * sc);
    u.y = pack_bf16(__uint_as_float(t[2]) * sc, __uint_as_float(t[3]) * sc);
    u.z = pack_bf16(__uint_as_float(t[4]) * sc, __uint_as_float(t[5]) * sc);
    u.w = pack_bf16(__uint_as_float(t[6]) * sc, __uint_as_float(t[7]) * sc);
    *dst.chunk(8) = u;
  }
}
// the same with the backward of the q/k L2 normalisation (see store_out_row_qknorm) for a 72-channel head
__device__ __forceinline__ void store_acc_row_qknorm72(const RowOut& dst, const bf16* __restrict__ yrow, const uint32_t (&a)[32], const uint32_t (&b)[32],
                                                       const uint32_t (&t)[8], float att_scale, float s, float eps) {
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const uint4 u = reinterpret_cast<const uint4*>(yrow)[c];
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 y2 = __bfloat1622float2(h2[e]);
      const int i = 8 * c + 2 * e;
      const float g0 = __uint_as_float(i < 32 ? a[i & 31] : (i < 64 ? b[i & 31] : t[i & 7]));
      const float g1 = __uint_as_float(i + 1 < 32 ? a[(i + 1) & 31] : (i + 1 < 64 ? b[(i + 1) & 31] : t[(i + 1) & 7]));
      dot = fmaf(y2.x, g0, fmaf(y2.y, g1, dot));
    }
  }
  dot *= att_scale;
  const float rpe = 8.48528137423857f / s;
  const float r = fmaxf(rpe - eps, 1e-30f);
  const float coef = dot * rpe / (72.0f * r);
  const float ga = s * att_scale, sc_y = -s * coef;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const uint4 u = reinterpret_cast<const uint4*>(yrow)[c];
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 y2 = __bfloat1622float2(h2[e]);
      const int i = 8 * c + 2 * e;
      const float g0 = __uint_as_float(i < 32 ? a[i & 31] : (i < 64 ? b[i & 31] : t[i & 7]));
      const float g1 = __uint_as_float(i + 1 < 32 ? a[(i + 1) & 31] : (i + 1 < 64 ? b[(i + 1) & 31] : t[(i + 1) & 7]));
      w[e] = pack_bf16(fmaf(ga, g0, sc_y * y2.x), fmaf(ga, g1, sc_y * y2.y));
    }
    *dst.chunk(c) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// Row r of a resident [128 x head_dim] tile (TMA-written SWIZZLE_128B panels) -> TMEM lane r, columns [0, KST*8): packed bf16 pairs in
// channel order, which is the A-operand layout of tcgen05.mma (k-step kk = columns kk*8 .. kk*8+7).  The tile is the A operand of every
// S / dP MMA of the CTA: read from TMEM it costs no shared-memory bandwidth, which is what bounds these kernels (an MMA over a
// [128 x 64] logit block is 32 tensor-core cycles but fetched 4 KB of A and 2 KB of B).  Channels 72..79 are the TMA unit's zero fill.
template <int HDV>
__device__ __forceinline__ void tile_to_tmem(uint32_t taddr, const uint8_t* tile, int r) {
  using B = BCfg<HDV>;
#pragma unroll
  for (int kk = 0; kk < B::KST; ++kk) {
    const uint8_t* prow = tile + (kk >> 2) * B::PROW + (r >> 3) * 1024 + (r & 7) * 128;
    const int c = (kk & 3) * 2;
    const uint4 u0 = *reinterpret_cast<const uint4*>(prow + ((c ^ (r & 7)) << 4));
    const uint4 u1 = *reinterpret_cast<const uint4*>(prow + (((c + 1) ^ (r & 7)) << 4));
    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
    tmem_st8(taddr + kk * 8, w);
  }
}

// ------------------------------------------------------------------------------------------------ dQ
// dS never touches shared memory: the softmax threads that own a query row read its S and dP values from TMEM and write the 64 bf16
// dS values (packed pairs, 32 columns) into one of two small TMEM tiles; the dQ += dS K_j MMAs take that tile as their A operand
// straight from TMEM.  With the [128 x 64] logit blocks of this kernel an MMA is
// only 32-40 tensor-core cycles long but fetches 6 KB of operands, so shared-memory bandwidth (MMA operand reads + the staging tile's
// stores and re-reads + TMA writes: 134 KB per block, 128 B/clk) was the bound of the first version, and every other shared-memory
// access (mbarrier polls, fences) queued behind it (tools/attn_bwd_xl_timeline.py, tools/probes/mma_rate_probe.cu: 48 cycles per
// [128 x 64 x 16] MMA with both operands in shared memory, 32 with A in TMEM).
// Issue order per block j, once the softmax warps have published dS(j) (which also says they are done reading S / dP(j)):
// S / dP(j+2) into the buffer of block j FIRST, then dQ(j) -- the softmax warps wait for S(j+2) while nothing waits for dQ(j).  (One
// commit behind both, so that seeing S(j+2) implies the dS tile may be overwritten, saves a barrier but shows S(j+2) 160 cycles later:
// measured slower.)
// Persistent: one CTA per SM walks (query tile, head, sample) items.  With a CTA per item a third of its 25 k cycles were prologue
// (TMA of Q / dO, cold K/V ring), epilogue and the gap to the next CTA, none of it overlapped (225 KB of shared memory and 480 TMEM
// columns: one CTA per SM).  Here the K/V ring keeps running across items (all block counters are global), the next item's Q / dO land
// in the staging tiles while the current item computes, and they are copied into the TMEM A tiles as soon as the current item's last
// S / dP MMAs have completed -- before its dQ is read out.
template <int HDV>
__global__ void __launch_bounds__(BCfg<HDV>::NTHR, 1)
attn_bwd_dq_tc(const __grid_constant__ CUtensorMap tm_qkv_row, const __grid_constant__ CUtensorMap tm_qkv_blk,
               const __grid_constant__ CUtensorMap tm_do_row, const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ lse,
               const float* __restrict__ delta, int tokens, int heads, int n_samples,
               const bf16* __restrict__ qkv, const float* __restrict__ sc, float eps, long long* __restrict__ dbg) {
  using B = BCfg<HDV>;
  // optional timeline of CTA 0 (tools/attn_bwd_xl_timeline.py): dbg[role*256 + 4*g + e] = clock64 at event e of the CTA's g-th key
  // block (counted across its items); role 3: item-level events of the first items
#define DQ_STAMP(role, g, e)                                                                                          \
  do {                                                                                                                \
    if (dbg && blockIdx.x == 0 && (g) < 64 && (threadIdx.x & 31) == 0) dbg[(role) * 256 + 4 * (g) + (e)] = clock64(); \
  } while (0)
  constexpr int ROWB = B::ROWB, BLKB = B::BLKB, PROW = B::PROW, PBLK = B::PBLK, NST = B::DQ_NST, NSW = B::NSW;
  constexpr uint32_t T_DQ = B::T_ACC, T_QA = T_DQ + B::NACC, T_DOA = T_QA + B::A_COLS, T_DS = T_DOA + B::A_COLS;  // dQ | Q, dO (A operands) | 2 x dS
  static_assert(T_DS + 2 * 32 <= B::TMEM, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;         // staging of the next item's Q / dO tiles (TMA -> here -> TMEM)
  uint8_t* sdO = sQ + ROWB;
  uint8_t* sKV = sdO + ROWB;  // stage s: K_j at sKV + s*2*BLKB, V_j after it
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NST * 2 * BLKB);
  uint64_t* bar_q = bars;                 // an item's Q / dO tiles have landed in the staging tiles
  uint64_t* kv_full = bars + 1;           // [NST]
  uint64_t* kv_empty = kv_full + NST;     // [NST]
  uint64_t* s_full = kv_empty + NST;      // [2]  S / dP of a block are in TMEM
  uint64_t* ds_full = s_full + 2;         // [2]  dS of a block is in TMEM (and its S / dP have been read)
  uint64_t* ds_empty = ds_full + 2;       // [2]  dQ(j) has read the dS tile
  uint64_t* o_full = ds_empty + 2;        // an item's dQ is complete
  uint64_t* a_full = o_full + 1;          // an item's Q and dO are in TMEM
  uint64_t* stage_free = a_full + 1;      // ... and the staging tiles may take the next item's
  uint64_t* acc_free = stage_free + 1;    // an item's dQ has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HDV, nkb = tokens / CB, nqt = (tokens + RT - 1) / RT;
  const int total_items = nqt * heads * n_samples;
  const int my_items = (int)blockIdx.x < total_items ? (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // item -> query tile (fastest: the CTAs that run together share K / V in L2), head, sample
  auto decode = [&](int it, int& q0, int& h, int& row_base) {
    const int item = blockIdx.x + it * gridDim.x;
    const int qt = item % nqt, rest = item / nqt;
    h = rest % heads;
    row_base = (rest / heads) * tokens;
    q0 = qt * RT;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv_row);
    prefetch_tmap(&tm_qkv_blk);
    prefetch_tmap(&tm_do_row);
    prefetch_tmap(&tm_out);
    mbar_init(bar_q, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&ds_full[i], NSW);
      mbar_init(&ds_empty[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(a_full, NSW);
    mbar_init(stage_free, NSW);
    mbar_init(acc_free, NSW / 2);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<B::TMEM>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      auto load_rows = [&](int it) {
        int q0, h, row_base;
        decode(it, q0, h, row_base);
        mbar_arrive_expect_tx(bar_q, 2 * ROWB);
        load_tile<HDV>(sQ, &tm_qkv_row, bar_q, 0, h, heads, row_base + q0, PROW);
        load_tile<HDV>(sdO, &tm_do_row, bar_q, 0, h, heads, row_base + q0, PROW);
      };
      if (my_items > 0) load_rows(0);
      uint32_t g = 0;
      for (int it = 0; it < my_items; ++it) {
        int q0, h, row_base;
        decode(it, q0, h, row_base);
        for (int j = 0; j < nkb; ++j, ++g) {
          const int s = g % NST;
          mbar_wait(&kv_empty[s], ((g / NST) & 1) ^ 1);
          DQ_STAMP(2, g, 0);
          uint8_t* dst = sKV + s * 2 * BLKB;
          mbar_arrive_expect_tx(&kv_full[s], 2 * BLKB);
          load_tile<HDV>(dst, &tm_qkv_blk, &kv_full[s], 1, h, heads, row_base + j * CB, PBLK);
          load_tile<HDV>(dst + BLKB, &tm_qkv_blk, &kv_full[s], 2, h, heads, row_base + j * CB, PBLK);
        }
        if (it + 1 < my_items) {  // the ring throttles this loop: we get here a few blocks before the item ends
          mbar_wait(stage_free, it & 1);
          load_rows(it + 1);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    constexpr uint32_t idesc_s = make_idesc_bf16(RT, CB, 0, 0);       // S / dP: A (Q, dO) from TMEM, B K-major
    constexpr uint32_t idesc_q = make_idesc_bf16(RT, B::NACC, 0, 1);  // dQ += dS K_j: dS from TMEM, K_j MN-major (d contiguous)
    // descriptors of ring slot 0; every MMA operand is one of these moved by a compile-time or per-block offset
    const uint64_t d_kv = make_smem_desc(smem_u32(sKV), 16, 1024);     // K-major view of a K / V block (S, dP)
    const uint64_t d_kmn = make_smem_desc(smem_u32(sKV), PBLK, 1024);  // MN-major view of a K block (dQ): LBO = panel stride
    auto scores = [&](uint32_t g) {  // kv_full of the block's ring slot has been waited for
      const uint32_t s = g % NST, b = g & 1;
      const uint64_t d_k = desc_advance(d_kv, s * 2 * BLKB), d_v = desc_advance(d_k, BLKB);
      const uint32_t t_s = tmem_base + b * 128;
      // the S and dP chains accumulate into different TMEM tiles: issued alternately so consecutive MMAs are independent
#pragma unroll
      for (int k = 0; k < B::KST; ++k) {
        const uint32_t bo = (k >> 2) * PBLK + (k & 3) * 32;  // panel + 16-channel step
        if (leader) umma_ts(t_s, tmem_base + T_QA + k * 8, desc_advance(d_k, bo), idesc_s, k != 0);
        if (leader) umma_ts(t_s + 64, tmem_base + T_DOA + k * 8, desc_advance(d_v, bo), idesc_s, k != 0);
      }
      if (leader) umma_commit(&s_full[b]);
    };
    uint32_t g0 = 0;
    for (int it = 0; it < my_items; ++it, g0 += nkb) {
      mbar_wait(a_full, it & 1);
      tc_fence_after();
      mbar_wait(&kv_full[g0 % NST], (g0 / NST) & 1);
      scores(g0);
      if (nkb > 1) {
        mbar_wait(&kv_full[(g0 + 1) % NST], ((g0 + 1) / NST) & 1);
        scores(g0 + 1);
      }
      for (int j = 0; j < nkb; ++j) {
        const uint32_t g = g0 + j, s = g % NST, b = g & 1;
        // the ring wait of block j+2 in the shadow of the softmax warps' work on block j, not behind it (every mbarrier wait of this warp,
        // even on a completed phase, is 100+ cycles)
        if (j + 2 < nkb) mbar_wait(&kv_full[(g + 2) % NST], ((g + 2) / NST) & 1);
        DQ_STAMP(0, g, 0);
        mbar_wait_spin(&ds_full[b], (g >> 1) & 1);
        tc_fence_after();
        DQ_STAMP(0, g, 1);
        if (j + 2 < nkb) scores(g + 2);
        if (j == 0 && it > 0) {  // the previous item's dQ has been read out of the accumulator
          mbar_wait(acc_free, (it - 1) & 1);
          tc_fence_after();
        }
        const uint64_t d_b = desc_advance(d_kmn, s * 2 * BLKB);
#pragma unroll
        for (int k = 0; k < CB / 16; ++k)  // 16 keys = 8 packed columns
          if (leader) umma_ts(tmem_base + T_DQ, tmem_base + T_DS + b * 32 + k * 8, desc_advance(d_b, k * 2048), idesc_q, (j | k) != 0);
        if (leader) umma_commit(&kv_empty[s]);
        if (leader) umma_commit(&ds_empty[b]);
        DQ_STAMP(0, g, 2);
      }
      if (leader) umma_commit(o_full);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps / epilogue
    const int qq = warp & 3;
    const int h0 = (warp - 2) >> 2;  // which 32 of the 64 columns of a logit block this warp owns
    const int r = qq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
    const float c1 = att_scale_of(HDV) * LOG2E;
    // the row of Q (first four warps) or of dO (the other four) of item `it`: staging tile -> TMEM A operand.  Only once the S / dP
    // MMAs of the previous item have completed (its last s_full has been seen by this thread).
    auto rows_to_tmem = [&](int it) {
      mbar_wait(bar_q, it & 1);
      if (h0 == 0) tile_to_tmem<HDV>(t_lane + T_QA, sQ, r);
      else tile_to_tmem<HDV>(t_lane + T_DOA, sdO, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_full);
        // the staging tiles are free for the producer again -- except that the first four warps put the current item's dQ into the
        // Q tile next and release it when their TMA store has read it (`release_out`)
        if (h0 == 1 || it == 0) mbar_arrive(stage_free);
      }
    };
    bool out_pending = false;  // lane 0 of the first four warps: a dQ slab store of this thread may still be reading the staging tile
    auto release_out = [&]() {
      if (out_pending) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(stage_free);
        out_pending = false;
      }
    };
    // delta_i = dO_i . O_i (attn_delta_kernel) and L_i of this thread's row, requested one item ahead
    auto row_stats = [&](int it, float& L, float& dl) {
      int q0, h, row_base;
      decode(it, q0, h, row_base);
      L = 0.f, dl = 0.f;
      if (q0 + r < tokens) {
        const size_t grow = (size_t)row_base + q0 + r;
        L = lse[grow * heads + h];
        dl = delta[grow * heads + h];
      }
    };
    float L_nxt = 0.f, dl_nxt = 0.f;
    if (my_items > 0) {
      row_stats(0, L_nxt, dl_nxt);
      rows_to_tmem(0);
    }
    uint32_t g0 = 0;
    for (int it = 0; it < my_items; ++it, g0 += nkb) {
      int q0, h, row_base;
      decode(it, q0, h, row_base);
      const bool row_ok = q0 + r < tokens;
      const size_t grow = (size_t)row_base + q0 + r;
      const float dl = dl_nxt, c2 = L_nxt * LOG2E;
      if (it + 1 < my_items) row_stats(it + 1, L_nxt, dl_nxt);
      for (int j = 0; j < nkb; ++j) {
        const uint32_t g = g0 + j, b = g & 1;
        if (warp == 2) DQ_STAMP(1, g, 0);
        mbar_wait(&s_full[b], (g >> 1) & 1);
        tc_fence_after();
        if (warp == 2) DQ_STAMP(1, g, 1);
        uint32_t sv[32], dp[32], pk[16];
        tmem_ld32(t_lane + b * 128 + h0 * 32, sv);
        tmem_ld32(t_lane + b * 128 + 64 + h0 * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2(fmaf(__uint_as_float(sv[2 * i]), c1, -c2)), p1 = ex2(fmaf(__uint_as_float(sv[2 * i + 1]), c1, -c2));
          float d0 = p0 * (__uint_as_float(dp[2 * i]) - dl), d1 = p1 * (__uint_as_float(dp[2 * i + 1]) - dl);
          pk[i] = pack_bf16(d0, d1);
        }
        if (warp == 2) DQ_STAMP(1, g, 2);
        mbar_wait(&ds_empty[b], ((g >> 1) & 1) ^ 1);  // dQ of two blocks ago has read the tile (long ago)
        tc_fence_after();
        tmem_st16(t_lane + T_DS + b * 32 + h0 * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[b]);
        if (warp == 2) DQ_STAMP(1, g, 3);
        if (j == 0) release_out();  // the previous item's dQ store has long read its slab by now
      }
      // the item's last S / dP MMAs have completed (we saw its last s_full): the A tiles may take the next item's rows.  Ahead of
      // the read-out below, so that the tensor core has S / dP of the next item's first blocks to do meanwhile.
      if (it + 1 < my_items) rows_to_tmem(it + 1);
      if (warp == 2 && it < 16) DQ_STAMP(3, it, 0);
      // read-out by the first four warps: a whole gradient row per thread (the q-norm backward needs its dot product with q^)
      if (h0 == 0) {
        uint32_t a0[32], a1[32], tl[8];
        mbar_wait(o_full, it & 1);
        tc_fence_after();
        tmem_ld32(t_lane + T_DQ, a0);
        tmem_ld32(t_lane + T_DQ + 32, a1);
        if constexpr (HDV != 64) tmem_ld8(t_lane + T_DQ + 64, tl);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);
        if (warp == 2 && it < 16) DQ_STAMP(3, it, 1);
        // the row goes into the (free again) Q staging tile, the warp's 32-row slab from there to dqkv by TMA.  row_ok is uniform over
        // a warp (tokens % 64 == 0): slabs past the sample's end are not stored.
        if (row_ok) {
          uint8_t* prow = sQ + (r >> 3) * 1024 + (r & 7) * 128;
          const RowOut dst{prow, prow + PROW, r & 7};
          if constexpr (HDV == 64) {
            if (sc) store_out_row_qknorm(dst, qkv + grow * 3 * D + h * HDV, a0, a1, 0.125f, sc[grow * 2 * heads + h], eps);
            else store_out_row(dst, a0, a1, 0.125f);
          } else {
            if (sc) store_acc_row_qknorm72(dst, qkv + grow * 3 * D + h * HDV, a0, a1, tl, att_scale_of(HDV), sc[grow * 2 * heads + h], eps);
            else store_acc_row<HDV>(dst, a0, a1, tl, att_scale_of(HDV));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (row_ok) {
            store_slab<HDV>(&tm_out, sQ, qq, 0, h, heads, row_base + q0 + qq * 32, PROW);
            out_pending = true;
          }
          if (!row_ok && it + 1 < my_items) mbar_arrive(stage_free);
        }
        if (warp == 2 && it < 16) DQ_STAMP(3, it, 2);
      }
    }
  }
  if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores complete before the CTA exits
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<B::TMEM>(tmem_base);
#undef DQ_STAMP
}

// ------------------------------------------------------------------------------------------------ dK, dV
// Rows (TMEM lanes) are keys.  P^T and dS^T go back into TMEM in place, over the S^T and dP^T columns they were computed from, and are
// the A operands of dV += P^T dO_j and dK += dS^T Q_j.  Persistent over (key tile, head, sample) items like attn_bwd_dq_tc.
template <int HDV>
__global__ void __launch_bounds__(BCfg<HDV>::NTHR, 1)
attn_bwd_dkv_tc(const __grid_constant__ CUtensorMap tm_qkv_row, const __grid_constant__ CUtensorMap tm_qkv_blk,
                const __grid_constant__ CUtensorMap tm_do_blk, const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ lse,
                const float* __restrict__ delta, int tokens, int heads, int n_samples, const bf16* __restrict__ qkv,
                const float* __restrict__ sc, float eps, long long* __restrict__ dbg) {
  using B = BCfg<HDV>;
  // optional timeline of CTA 0, roles 4..7 of the buffer (see attn_bwd_dq_tc)
#define DKV_STAMP(role, g, e)                                                                                             \
  do {                                                                                                                    \
    if (dbg && blockIdx.x == 0 && (g) < 64 && (threadIdx.x & 31) == 0) dbg[(4 + (role)) * 256 + 4 * (g) + (e)] = clock64(); \
  } while (0)
  constexpr int ROWB = B::ROWB, BLKB = B::BLKB, PROW = B::PROW, PBLK = B::PBLK, NST = B::DKV_NST, NSW = B::NSW;
  constexpr uint32_t T_DK = B::T_ACC, T_DV = T_DK + B::NACC, T_KA = T_DV + B::NACC, T_VA = T_KA + B::A_COLS;  // dK | dV | K, V as A operands
  static_assert(T_VA + B::A_COLS <= B::TMEM, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;         // staging of the next item's K / V tiles (TMA -> here -> TMEM)
  uint8_t* sV = sK + ROWB;
  uint8_t* sQdO = sV + ROWB;  // stage s: Q_j at sQdO + s*2*BLKB, dO_j after it
  float* sL = reinterpret_cast<float*>(sQdO + NST * 2 * BLKB);  // [2][64]
  float* sD = sL + 2 * CB;                                       // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * CB);
  uint64_t* bar_kv = bars;               // an item's K / V tiles have landed in the staging tiles
  uint64_t* qd_full = bars + 1;          // [NST]
  uint64_t* qd_empty = qd_full + NST;    // [NST]
  uint64_t* s_full = qd_empty + NST;     // [2]
  uint64_t* p_full = s_full + 2;         // [2]
  uint64_t* o_full = p_full + 2;         // an item's dK and dV are complete
  uint64_t* a_full = o_full + 1;         // an item's K and V are in TMEM
  uint64_t* stage_free = a_full + 1;     // ... and the staging tiles may take the next item's
  uint64_t* acc_free = stage_free + 1;   // an item's dK and dV have been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HDV, nqb = tokens / CB, nkt = (tokens + RT - 1) / RT;
  const int total_items = nkt * heads * n_samples;
  const int my_items = (int)blockIdx.x < total_items ? (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto decode = [&](int it, int& k0, int& h, int& row_base) {
    const int item = blockIdx.x + it * gridDim.x;
    const int kt = item % nkt, rest = item / nkt;
    h = rest % heads;
    row_base = (rest / heads) * tokens;
    k0 = kt * RT;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv_row);
    prefetch_tmap(&tm_qkv_blk);
    prefetch_tmap(&tm_do_blk);
    prefetch_tmap(&tm_out);
    mbar_init(bar_kv, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], NSW);
    }
    mbar_init(o_full, 1);
    mbar_init(a_full, NSW);
    mbar_init(stage_free, NSW);
    mbar_init(acc_free, NSW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<B::TMEM>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      auto load_rows = [&](int it) {
        int k0, h, row_base;
        decode(it, k0, h, row_base);
        mbar_arrive_expect_tx(bar_kv, 2 * ROWB);
        load_tile<HDV>(sK, &tm_qkv_row, bar_kv, 1, h, heads, row_base + k0, PROW);
        load_tile<HDV>(sV, &tm_qkv_row, bar_kv, 2, h, heads, row_base + k0, PROW);
      };
      if (my_items > 0) load_rows(0);
      uint32_t g = 0;
      for (int it = 0; it < my_items; ++it) {
        int k0, h, row_base;
        decode(it, k0, h, row_base);
        for (int j = 0; j < nqb; ++j, ++g) {
          const int s = g % NST;
          mbar_wait(&qd_empty[s], ((g / NST) & 1) ^ 1);
          DKV_STAMP(2, g, 0);
          uint8_t* dst = sQdO + s * 2 * BLKB;
          mbar_arrive_expect_tx(&qd_full[s], 2 * BLKB);
          load_tile<HDV>(dst, &tm_qkv_blk, &qd_full[s], 0, h, heads, row_base + j * CB, PBLK);
          load_tile<HDV>(dst + BLKB, &tm_do_blk, &qd_full[s], 0, h, heads, row_base + j * CB, PBLK);
        }
        if (it + 1 < my_items) {
          mbar_wait(stage_free, it & 1);
          load_rows(it + 1);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    constexpr uint32_t idesc_s = make_idesc_bf16(RT, CB, 0, 0);       // S^T = K Q_j^T, dP^T = V dO_j^T
    constexpr uint32_t idesc_a = make_idesc_bf16(RT, B::NACC, 0, 1);  // dV += P^T dO_j, dK += dS^T Q_j : A from TMEM, B MN-major
    const uint64_t d_qd = make_smem_desc(smem_u32(sQdO), 16, 1024);      // K-major view of a Q / dO block (S^T, dP^T)
    const uint64_t d_qdmn = make_smem_desc(smem_u32(sQdO), PBLK, 1024);  // MN-major view (dK, dV): LBO = panel stride
    auto scores = [&](uint32_t g) {  // qd_full of the block's ring slot has been waited for
      const uint32_t s = g % NST, b = g & 1;
      const uint64_t d_q = desc_advance(d_qd, s * 2 * BLKB), d_do = desc_advance(d_q, BLKB);
      const uint32_t t_s = tmem_base + b * 128;
#pragma unroll
      for (int k = 0; k < B::KST; ++k) {  // S^T and dP^T chains interleaved (independent accumulators)
        const uint32_t bo = (k >> 2) * PBLK + (k & 3) * 32;
        if (leader) umma_ts(t_s, tmem_base + T_KA + k * 8, desc_advance(d_q, bo), idesc_s, k != 0);
        if (leader) umma_ts(t_s + 64, tmem_base + T_VA + k * 8, desc_advance(d_do, bo), idesc_s, k != 0);
      }
      if (leader) umma_commit(&s_full[b]);
    };
    uint32_t g0 = 0;
    for (int it = 0; it < my_items; ++it, g0 += nqb) {
      mbar_wait(a_full, it & 1);
      tc_fence_after();
      mbar_wait(&qd_full[g0 % NST], (g0 / NST) & 1);
      scores(g0);
      if (nqb > 1) mbar_wait(&qd_full[(g0 + 1) % NST], ((g0 + 1) / NST) & 1);
      for (int j = 0; j < nqb; ++j) {
        const uint32_t g = g0 + j, s = g % NST, b = g & 1;
        DKV_STAMP(0, g, 3);
        if (j + 1 < nqb) scores(g + 1);  // into the other buffer: its P^T / dS^T were read by the MMAs of block j-1, issued before this
        // the ring wait of block j+2 in the shadow of the softmax warps' work on block j (see attn_bwd_dq_tc)
        if (j + 2 < nqb) mbar_wait(&qd_full[(g + 2) % NST], ((g + 2) / NST) & 1);
        DKV_STAMP(0, g, 0);
        mbar_wait_spin(&p_full[b], (g >> 1) & 1);
        tc_fence_after();
        DKV_STAMP(0, g, 1);
        if (j == 0 && it > 0) {  // the previous item's dK / dV have been read out of the accumulators
          mbar_wait(acc_free, (it - 1) & 1);
          tc_fence_after();
        }
        const uint64_t d_q = desc_advance(d_qdmn, s * 2 * BLKB), d_do = desc_advance(d_q, BLKB);
        const uint32_t t_p = tmem_base + b * 128;  // P^T over the S^T columns, dS^T over the dP^T columns
#pragma unroll
        for (int k = 0; k < CB / 16; ++k) {  // dV and dK chains interleaved
          const uint32_t ko = (k >> 1) * 32 + (k & 1) * 8;
          if (leader) umma_ts(tmem_base + T_DV, t_p + ko, desc_advance(d_do, k * 2048), idesc_a, (j | k) != 0);
          if (leader) umma_ts(tmem_base + T_DK, t_p + 64 + ko, desc_advance(d_q, k * 2048), idesc_a, (j | k) != 0);
        }
        if (leader) umma_commit(&qd_empty[s]);
        DKV_STAMP(0, g, 2);
      }
      if (leader) umma_commit(o_full);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps / epilogue
    const int qq = warp & 3;
    const int h0 = (warp - 2) >> 2;    // which 32 of the 64 query columns of a block this warp owns
    const int r = qq * 32 + lane;      // key row of the tile
    const int tid = threadIdx.x - 64;  // 0..255 among the softmax warps
    const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
    const float c1 = att_scale_of(HDV) * LOG2E;
    // the row of K (first four warps) or of V (the other four) of item `it`: staging tile -> TMEM A operand, once the S^T / dP^T MMAs
    // of the previous item have completed (its last s_full has been seen by this thread)
    auto rows_to_tmem = [&](int it) {
      mbar_wait(bar_kv, it & 1);
      if (h0 == 0) tile_to_tmem<HDV>(t_lane + T_KA, sK, r);
      else tile_to_tmem<HDV>(t_lane + T_VA, sV, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_full);
        // the staging tiles take the current item's dK (K tile) and dV (V tile) next: released to the producer when the TMA stores
        // have read them (`release_out`)
        if (it == 0) mbar_arrive(stage_free);
      }
    };
    bool out_pending = false;  // lane 0: a slab store of this thread may still be reading the staging tile
    auto release_out = [&]() {
      if (out_pending) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(stage_free);
        out_pending = false;
      }
    };
    // per-query L (threads 0-63) and delta (64-127) of a block: 64 different lines of global memory each.  The values of the CTA's next
    // block (of the next item behind an item's last) are requested while the current one is processed.  (Nothing may depend on the
    // loaded value before the next block's store: the first version multiplied by log2 e right behind the load and the warp sat out the
    // load latency, 1 100 cycles per block, on that multiply.)
    auto fetch = [&](int it, int j) {
      int k0, h, row_base;
      decode(it, k0, h, row_base);
      const size_t qrow = (size_t)row_base + j * CB + (tid & 63);
      return tid < 64 ? lse[qrow * heads + h] : delta[qrow * heads + h];
    };
    float ld_val = 0.f;
    if (my_items > 0) {
      if (tid < 128) ld_val = fetch(0, 0);
      rows_to_tmem(0);
    }
    uint32_t g0 = 0;
    for (int it = 0; it < my_items; ++it, g0 += nqb) {
      int k0, h, row_base;
      decode(it, k0, h, row_base);
      for (int j = 0; j < nqb; ++j) {
        const uint32_t g = g0 + j;
        const int s = g & 1, b = g & 1;
        // -> smem (double-buffered), visible to the softmax threads
        if (tid < 64) sL[s * CB + tid] = ld_val * LOG2E;
        else if (tid < 128) sD[s * CB + (tid - 64)] = ld_val;
        if (tid < 128) {
          if (j + 1 < nqb) ld_val = fetch(it, j + 1);
          else if (it + 1 < my_items) ld_val = fetch(it + 1, 0);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NSW * 32) : "memory");
        if (warp == 2) DKV_STAMP(1, g, 0);
        mbar_wait(&s_full[b], (g >> 1) & 1);
        tc_fence_after();
        if (warp == 2) DKV_STAMP(1, g, 1);
        uint32_t sv[32], dp[32], pp[16], pd[16];
        tmem_ld32(t_lane + b * 128 + h0 * 32, sv);
        tmem_ld32(t_lane + b * 128 + 64 + h0 * 32, dp);
        tmem_ld_wait();
        const float4* pL = reinterpret_cast<const float4*>(sL + s * CB + h0 * 32);
        const float4* pD = reinterpret_cast<const float4*>(sD + s * CB + h0 * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // four query columns per trip: L and delta as one 16-byte broadcast read each
          const float4 l4 = pL[i], d4 = pD[i];
          float p0 = ex2(fmaf(__uint_as_float(sv[4 * i]), c1, -l4.x)), p1 = ex2(fmaf(__uint_as_float(sv[4 * i + 1]), c1, -l4.y));
          float p2 = ex2(fmaf(__uint_as_float(sv[4 * i + 2]), c1, -l4.z)), p3 = ex2(fmaf(__uint_as_float(sv[4 * i + 3]), c1, -l4.w));
          pp[2 * i] = pack_bf16(p0, p1);
          pp[2 * i + 1] = pack_bf16(p2, p3);
          pd[2 * i] = pack_bf16(p0 * (__uint_as_float(dp[4 * i]) - d4.x), p1 * (__uint_as_float(dp[4 * i + 1]) - d4.y));
          pd[2 * i + 1] = pack_bf16(p2 * (__uint_as_float(dp[4 * i + 2]) - d4.z), p3 * (__uint_as_float(dp[4 * i + 3]) - d4.w));
        }
        // in place, over columns only this thread reads: 32 queries = 16 packed columns at the start of the warp's 32-column group
        if (warp == 2) DKV_STAMP(1, g, 2);
        tmem_st16(t_lane + b * 128 + h0 * 32, pp);
        tmem_st16(t_lane + b * 128 + 64 + h0 * 32, pd);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);
        if (warp == 2) DKV_STAMP(1, g, 3);
        if (j == 0) release_out();  // the previous item's dK / dV store has long read its slab by now
      }
      // the item's last S^T / dP^T MMAs have completed: the A tiles may take the next item's rows, ahead of the read-out below
      if (it + 1 < my_items) rows_to_tmem(it + 1);
      if (warp == 2 && it < 16) DKV_STAMP(3, it, 0);
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      uint32_t a0[32], a1[32], tl[8];
      const bool row_ok = k0 + r < tokens;
      const size_t grow = (size_t)row_base + k0 + r;
      // the first four warps write dK, the other four dV
      tmem_ld32(t_lane + (h0 == 0 ? T_DK : T_DV), a0);
      tmem_ld32(t_lane + (h0 == 0 ? T_DK : T_DV) + 32, a1);
      if constexpr (HDV != 64) tmem_ld8(t_lane + (h0 == 0 ? T_DK : T_DV) + 64, tl);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
      if (warp == 2 && it < 16) DKV_STAMP(3, it, 1);
      // rows go into the (free again) staging tiles -- dK over the K tile, dV over the V tile -- and each warp's 32-row slab from there to
      // dqkv by TMA.  row_ok is uniform over a warp (tokens % 64 == 0): slabs past the sample's end are not stored.
      uint8_t* stile = h0 == 0 ? sK : sV;
      if (row_ok) {
        uint8_t* prow = stile + (r >> 3) * 1024 + (r & 7) * 128;
        const RowOut dst{prow, prow + PROW, r & 7};
        if (h0 == 0) {
          if constexpr (HDV == 64) {
            if (sc) store_out_row_qknorm(dst, qkv + grow * 3 * D + D + h * HDV, a0, a1, 0.125f, sc[grow * 2 * heads + heads + h], eps);
            else store_out_row(dst, a0, a1, 0.125f);
          } else {
            if (sc) store_acc_row_qknorm72(dst, qkv + grow * 3 * D + D + h * HDV, a0, a1, tl, att_scale_of(HDV), sc[grow * 2 * heads + heads + h], eps);
            else store_acc_row<HDV>(dst, a0, a1, tl, att_scale_of(HDV));
          }
        } else {
          store_acc_row<HDV>(dst, a0, a1, tl, 1.0f);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (row_ok) {
          store_slab<HDV>(&tm_out, stile, qq, h0 == 0 ? 1 : 2, h, heads, row_base + k0 + qq * 32, PROW);
          out_pending = true;
        } else if (it + 1 < my_items) {
          mbar_arrive(stage_free);
        }
      }
      if (warp == 2 && it < 16) DKV_STAMP(3, it, 2);
    }
  }
  if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores complete before the CTA exits
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<B::TMEM>(tmem_base);
#undef DKV_STAMP
}

// ------------------------------------------------------------------------------------------------ fused dQ, dK, dV (tokens == 256)
// One persistent CTA per SM walks (sample, head) items; everything an item needs stays on chip: Q, K, V, dO as eight
// [128 x 64] TMA tiles (128 KB), S^T / dP^T / dV / dK / dQ_0 / dQ_1 in the 512 TMEM columns.  Rows (TMEM lanes) are KEYS:
//   S^T = K_kt Q_qb^T, dP^T = V_kt dO_qb^T                      (M = 128 keys, N = 128 queries, K = 64)
//   P^T = exp2(S^T c - L_q), dS^T = P^T (dP^T - delta_q)         8 softmax warps, bf16 -> two [128 x 128] staging tiles
//   dV_kt += P^T dO_qb, dK_kt += dS^T Q_qb                       (A K-major from the staging tiles, B = dO / Q tile MN-major)
//   dQ_qb += dS K_kt                                             (A = the SAME dS^T tile read MN-major, B = K tile MN-major)
// 5 GEMMs per logit block instead of the 7 of the dq + dkv kernel pair (S and dP are not recomputed) and Q/K/V/dO are read
// from HBM once.  Iteration order (kt, qb): even items (0,0) (0,1) (1,1) (1,0), odd items (0,1) (0,0) (1,0) (1,1): the query
// block an item finishes with is the one the next item needs last, so every input tile has about one iteration between
// its release and the first MMA that needs its refill.  Each tile has its own full/empty barrier pair.
// Outputs never touch the LSU's global path (row-per-thread 16-byte stores cost 32 line transactions per instruction and
// were 60 % of the first version's time): a finished accumulator row is converted in registers, written into a dead
// input tile (dK in place over the K tile whose rows are also the k^ the q/k-norm backward needs, dV over the V tile, dQ
// into a 16 KB staging tile) and leaves through per-warp TMA stores; the tile is released to the producer when the store
// has read it.
constexpr int F_T = 256;
constexpr int F_TILE = RT * HD * 2;              // 16 KB
constexpr int F_STG = 2 * F_TILE;                // [128 keys x 128 queries] bf16 = two 64-query panels
constexpr int F_NTHREADS = 320;                  // warp 0 TMA, warp 1 MMA, warps 2-9 softmax / epilogue
constexpr int F_SMEM = 9 * F_TILE + 2 * F_STG + 4 * F_T * 4 + 256 + 1024;
enum { TK0 = 0, TK1 = 1, TV0 = 2, TV1 = 3, TQ0 = 4, TQ1 = 5, TD0 = 6, TD1 = 7, TOUT = 8 };

__device__ __forceinline__ void load_row_sw128(const uint8_t* tile, int r, float (&y)[64]) {
  const uint8_t* prow = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(prow + ((c ^ (r & 7)) << 4));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = __bfloat1622float2(h2[e]);
      y[8 * c + 2 * e] = t.x;
      y[8 * c + 2 * e + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void load_row_global(const bf16* __restrict__ yrow, float (&y)[64]) {
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
}
// accumulator row (a | b = 64 fp32) -> 64 bf16 packed; plain scale, or the q/k-normalisation backward (see store_out_row_qknorm)
__device__ __forceinline__ void pack_row(const uint32_t (&a)[32], const uint32_t (&b)[32], float scale, uint32_t (&out)[32]) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    out[c] = pack_bf16(__uint_as_float(a[2 * c]) * scale, __uint_as_float(a[2 * c + 1]) * scale);
    out[16 + c] = pack_bf16(__uint_as_float(b[2 * c]) * scale, __uint_as_float(b[2 * c + 1]) * scale);
  }
}
__device__ __forceinline__ void pack_row_qknorm(const uint32_t (&a)[32], const uint32_t (&b)[32], const float (&y)[64], float att_scale,
                                                float s, float eps, uint32_t (&out)[32]) {
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) dot = fmaf(y[c], __uint_as_float(a[c]), fmaf(y[32 + c], __uint_as_float(b[c]), dot));
  dot *= att_scale;
  const float rpe = 8.0f / s;
  const float r = fmaxf(rpe - eps, 1e-30f);
  const float sco = -s * (dot * rpe / (64.0f * r));
  const float ga = s * att_scale;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    out[c] = pack_bf16(fmaf(ga, __uint_as_float(a[2 * c]), sco * y[2 * c]), fmaf(ga, __uint_as_float(a[2 * c + 1]), sco * y[2 * c + 1]));
    out[16 + c] = pack_bf16(fmaf(ga, __uint_as_float(b[2 * c]), sco * y[32 + 2 * c]), fmaf(ga, __uint_as_float(b[2 * c + 1]), sco * y[32 + 2 * c + 1]));
  }
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"((uint64_t)m), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(F_NTHREADS, 1)
attn_bwd_fused_tc(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                  const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ lse, const float* __restrict__ delta, int heads,
                  int total_items, const bf16* __restrict__ qkv, const float* __restrict__ sc, float eps, long long* __restrict__ dbg) {
  // optional timeline of CTA 0 (tools/attn_bwd_timeline.py): dbg[role*256 + 4*g + e] = clock64 at event e of iteration g
#define FSTAMP(role, g, e) \
  do { \
    if (dbg && blockIdx.x == 0 && (g) < 64 && lane == 0) dbg[(role) * 256 + 4 * (g) + (e)] = clock64(); \
  } while (0)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sTile = smem;                      // 8 input tiles + the dQ output staging tile, index = T* enum
  uint8_t* sPt = sTile + 9 * F_TILE;
  uint8_t* sdSt = sPt + F_STG;
  float* sL = reinterpret_cast<float*>(sdSt + F_STG);  // [2][256] log2-domain log-sum-exp of the item's queries
  float* sDl = sL + 2 * F_T;                            // [2][256] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDl + 2 * F_T);
  uint64_t* full = bars;         // [8]
  uint64_t* empty = bars + 8;    // [8]
  uint64_t* s_full = bars + 16;
  uint64_t* s_empty = bars + 17;
  uint64_t* p_full = bars + 18;
  uint64_t* p_empty = bars + 19;
  uint64_t* acc_free = bars + 20;  // dV / dK accumulators read out (8 warps)
  uint64_t* dq_free = bars + 21;   // the previous item's last dQ accumulator read out (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int my_items = blockIdx.x < total_items ? (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int G = my_items * 4;  // (kt, qb) iterations of this CTA
  // iteration i of item `it`: key tile i >> 1, query block {0,1,1,0}[i] (even items) / {1,0,0,1}[i] (odd items)
  auto qb_of = [](int g) { return (((g & 3) == 1 || (g & 3) == 2) ? 1 : 0) ^ ((g >> 2) & 1); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_out);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full[i], 1);
      // K / V tiles are also written and TMA-stored by four epilogue warps before the producer may refill them; the Q tiles
      // are copied (q^ rows for the dQ epilogue) by four warps
      mbar_init(&empty[i], i < TD0 ? 5 : 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(acc_free, 8);
    mbar_init(dq_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  constexpr uint32_t C_S = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

  if (warp == 0 && lane == 0) {
    for (int it = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int n = item / heads, h = item - n * heads;
      const uint32_t ph = it & 1;
      const int f = it & 1;  // first query block of this item = the one the previous item released first
      // The smem tiles are single-buffered, so a tile's refill is issued only ~one iteration before its first MMA: far less
      // than an HBM round trip under load (2-4 us measured).  Pull the NEXT item's 128 KB into L2 now, one whole item ahead.
      if (it + 1 < my_items) {
        const int nitem = item + gridDim.x;
        const int nn = nitem / heads, nh = nitem - nn * heads;
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          const int row = nn * F_T + rb * RT;
          tma_prefetch_l2_2d(&tm_qkv, nh * HD, row);
          tma_prefetch_l2_2d(&tm_qkv, D + nh * HD, row);
          tma_prefetch_l2_2d(&tm_qkv, 2 * D + nh * HD, row);
          tma_prefetch_l2_2d(&tm_do, nh * HD, row);
        }
      }
      // tiles in the order the previous item releases them
      const int order[8] = {TK0, TV0, TQ0 + f, TD0 + f, TQ0 + (f ^ 1), TD0 + (f ^ 1), TK1, TV1};
      for (int k = 0; k < 8; ++k) {
        const int t = order[k];
        mbar_wait(&empty[t], ph ^ 1);
        if (dbg && blockIdx.x == 0 && it < 8) dbg[3 * 256 + it * 8 + k] = clock64();
        mbar_arrive_expect_tx(&full[t], F_TILE);
        const int row = n * F_T + (t & 1) * RT;
        if (t >= TD0) tma_load_2d(sTile + t * F_TILE, &tm_do, &full[t], h * HD, row);
        else tma_load_2d(sTile + t * F_TILE, &tm_qkv, &full[t], (t >= TQ0 ? 0 : (t >= TV0 ? 2 * D : D)) + h * HD, row);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    constexpr uint32_t idesc_s = make_idesc_bf16(RT, 128, 0, 0);  // S^T / dP^T: both operands K-major
    constexpr uint32_t idesc_a = make_idesc_bf16(RT, HD, 0, 1);   // dV / dK: A K-major (staging), B MN-major
    constexpr uint32_t idesc_q = make_idesc_bf16(RT, HD, 1, 1);   // dQ: A = dS^T read MN-major, B = K MN-major
    const uint32_t tile0 = smem_u32(sTile), pt_addr = smem_u32(sPt), dst_addr = smem_u32(sdSt);
    auto scores = [&](int g) {
      const int i = g & 3, kt = i >> 1, qb = qb_of(g);
      const uint32_t ph = (g >> 2) & 1;
      if ((i & 1) == 0) {  // first use of this key tile
        mbar_wait(&full[TK0 + kt], ph);
        mbar_wait(&full[TV0 + kt], ph);
      }
      if (kt == 0) {       // first use of this query block
        mbar_wait(&full[TQ0 + qb], ph);
        mbar_wait(&full[TD0 + qb], ph);
      }
      mbar_wait(s_empty, (g & 1) ^ 1);
      tc_fence_after();
      const uint32_t k_addr = tile0 + (TK0 + kt) * F_TILE, v_addr = tile0 + (TV0 + kt) * F_TILE;
      const uint32_t q_addr = tile0 + (TQ0 + qb) * F_TILE, do_addr = tile0 + (TD0 + qb) * F_TILE;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {  // the two chains write different accumulators: issued alternately
        if (leader) umma_ss(tmem_base + C_S, make_smem_desc(k_addr + k * 32, 16, 1024), make_smem_desc(q_addr + k * 32, 16, 1024), idesc_s, k != 0);
        if (leader) umma_ss(tmem_base + C_DP, make_smem_desc(v_addr + k * 32, 16, 1024), make_smem_desc(do_addr + k * 32, 16, 1024), idesc_s, k != 0);
      }
      if (leader) umma_commit(s_full);
    };
    // are the input tiles iteration g touches for the first time loaded?  (non-blocking, warp-uniform)
    auto tiles_ready = [&](int g) {
      const int i = g & 3, kt = i >> 1, qb = qb_of(g);
      const uint32_t ph = (g >> 2) & 1;
      bool ok = true;
      if ((i & 1) == 0) ok = mbar_test_wait(smem_u32(&full[TK0 + kt]), ph) && mbar_test_wait(smem_u32(&full[TV0 + kt]), ph);
      if (kt == 0) ok = ok && mbar_test_wait(smem_u32(&full[TQ0 + qb]), ph) && mbar_test_wait(smem_u32(&full[TD0 + qb]), ph);
      return __all_sync(0xffffffffu, ok) != 0;
    };
    if (G > 0) scores(0);
    for (int g = 0; g < G; ++g) {
      // S / dP of the next iteration are issued as soon as their input tiles have landed -- ahead of this iteration's
      // dV / dK / dQ MMAs if possible (the softmax warps then never wait for S), otherwise while polling for P / dS, and
      // at the latest after them (they must not queue behind a tile that is still in flight)
      const int i = g & 3, kt = i >> 1, qb = qb_of(g);
      bool issued = g + 1 >= G;
      auto phase2_ready = [&]() {
        bool ok = mbar_test_wait(smem_u32(p_full), g & 1);
        // dV / dK restart at i == 2 and at i == 0: the epilogue warps must have read the previous accumulators; the dQ
        // columns of iteration 1 held the previous item's first block
        if (i == 2) ok = ok && mbar_test_wait(smem_u32(acc_free), 0);
        if (i == 0 && g > 0) ok = ok && mbar_test_wait(smem_u32(acc_free), 1);
        if (i == 1 && g > 4) ok = ok && mbar_test_wait(smem_u32(dq_free), ((g >> 2) - 1) & 1);
        return __all_sync(0xffffffffu, ok) != 0;
      };
      {
        const long long t0 = clock64();
        for (;;) {
          if (!issued && tiles_ready(g + 1)) {
            scores(g + 1);
            issued = true;
          }
          if (phase2_ready()) break;
          if (clock64() - t0 > 4000000000LL) {
            if (lane == 0) printf("mapdit: attn_bwd_fused MMA warp timed out (block %d iteration %d)\n", blockIdx.x, g);
            __trap();
          }
        }
      }
      FSTAMP(0, g, 0);
      FSTAMP(0, g, 1);
      tc_fence_after();
      const uint32_t k_addr = tile0 + (TK0 + kt) * F_TILE;
      const uint32_t q_addr = tile0 + (TQ0 + qb) * F_TILE, do_addr = tile0 + (TD0 + qb) * F_TILE;
#pragma unroll
      for (int k = 0; k < 128 / 16; ++k) {
        const uint32_t a_off = (k >> 2) * F_TILE + (k & 3) * 32;  // K-major: 64-query panel, 16 queries = 32 bytes
        if (leader) umma_ss(tmem_base + C_DV, make_smem_desc(pt_addr + a_off, 16, 1024), make_smem_desc(do_addr + k * 2048, 1024, 1024), idesc_a,
                ((i & 1) | k) != 0);
        if (leader) umma_ss(tmem_base + C_DK, make_smem_desc(dst_addr + a_off, 16, 1024), make_smem_desc(q_addr + k * 2048, 1024, 1024), idesc_a,
                ((i & 1) | k) != 0);
        if (leader) umma_ss(tmem_base + C_DQ + qb * HD, make_smem_desc(dst_addr + k * 2048, F_TILE, 1024), make_smem_desc(k_addr + k * 2048, 1024, 1024),
                idesc_q, (kt | k) != 0);
      }
      if (leader) umma_commit(p_empty);
      FSTAMP(0, g, 2);
      if (i == 1) {
        if (leader) umma_commit(&empty[TK0]);
        if (leader) umma_commit(&empty[TV0]);
      } else if (i == 2) {  // the second query block of the item is finished
        if (leader) umma_commit(&empty[TQ0 + qb]);
        if (leader) umma_commit(&empty[TD0 + qb]);
      } else if (i == 3) {
        if (leader) umma_commit(&empty[TQ0 + qb]);
        if (leader) umma_commit(&empty[TD0 + qb]);
        if (leader) umma_commit(&empty[TK1]);
        if (leader) umma_commit(&empty[TV1]);
      }
      if (!issued) scores(g + 1);
    }
  } else if (warp >= 2) {
    const int wq = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;    // which 64-query half of the 128-query block
    const int r = wq * 32 + lane;        // row of the tile
    const int tid = threadIdx.x - 64;    // 0..255
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float c1 = 0.125f * LOG2E;
    float nxt_l = 0.f, nxt_d = 0.f;  // the next item's L / delta of query `tid`, in flight across one iteration
    auto fetch_ld = [&](int item) {  // per-query L (log2 domain) and delta of an item: issue the loads ...
      const int n = item / heads, h = item - n * heads;
      const size_t qrow = (size_t)n * F_T + tid;
      // volatile asm pins the loads here and their first use in commit_ld: left to the compiler, the multiply below is
      // scheduled right behind the load and the warp sits out a whole HBM round trip (2-4 k cycles under load)
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nxt_l) : "l"(lse + qrow * heads + h));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nxt_d) : "l"(delta + qrow * heads + h));
    };
    auto commit_ld = [&](int buf) {  // ... and park them in shared memory once they have arrived
      asm volatile("" : "+f"(nxt_l), "+f"(nxt_d));
      sL[buf + tid] = nxt_l * LOG2E;
      sDl[buf + tid] = nxt_d;
    };
    int rel_a = -1, rel_b = -1;  // input tiles this warp TMA-stored from and still has to hand back to the producer
    auto release_pending = [&]() {  // called where the warp is about to wait anyway: the stores have long read their smem
      if (rel_a >= 0) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&empty[rel_a]);
          if (rel_b >= 0) mbar_arrive(&empty[rel_b]);
        }
        rel_a = rel_b = -1;
      }
    };
    auto signal_free = [&](uint64_t* bar) {  // the accumulator(s) just read are in registers: the MMA warp may overwrite them
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // packed bf16 row -> row r of smem tile `stile`, then this warp's 32-row slab -> dqkv[grow0 + wq*32 .., gcol .. gcol+64) by TMA
    auto store_slab = [&](uint8_t* stile, const uint32_t (&out)[32], size_t grow0, int gcol) {
      store_row_sw128(stile, r, out);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_store_2d(&tm_out, stile + wq * 32 * 128, gcol, (int)(grow0 + wq * 32));
    };
    // finish a dq / dk accumulator row already in registers: q/k-normalisation backward against the normalised row that sits
    // in `stile` (k^ in its input tile, q^ stashed in the staging tile), result written over it
    auto finish_qk = [&](const uint32_t (&a0)[32], const uint32_t (&a1)[32], uint8_t* stile, size_t grow0, int gcol, int sc_col) {
      uint32_t out[32];
      if (sc) {
        float y[64];
        load_row_sw128(stile, r, y);
        pack_row_qknorm(a0, a1, y, 0.125f, sc[(grow0 + r) * 2 * heads + sc_col], eps, out);
      } else {
        pack_row(a0, a1, 0.125f, out);
      }
      store_slab(stile, out, grow0, gcol);
    };
    // this thread's q^ row of query block qb -> the dQ staging tile (the dQ epilogue normalises against it and overwrites it)
    auto stash_q = [&](int qb) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous dQ store has left the tile
      __syncwarp();
      const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128;
      const uint4* src = reinterpret_cast<const uint4*>(sTile + (TQ0 + qb) * F_TILE + off);
      uint4* dst = reinterpret_cast<uint4*>(sTile + TOUT * F_TILE + off);
      uint4 v[8];  // chunk order rotated by the row so the 32 lanes of an access hit all banks (a plain copy: any order works)
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = src[c ^ (r & 7)];
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c ^ (r & 7)] = v[c];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[TQ0 + qb]);
    };
    if (G > 0) {
      fetch_ld(blockIdx.x);
      commit_ld(0);
    }
    for (int g = 0; g < G; ++g) {
      const int i = g & 3, qb = qb_of(g), it = g >> 2;
      const int item = blockIdx.x + it * gridDim.x;
      const int buf = (it & 1) * F_T;
      if (warp == 2 && i == 2) FSTAMP(2, g, 1);
      release_pending();
      if (warp == 2 && i == 2) FSTAMP(2, g, 2);
      if (i == 0) {
        if (warp == 2) FSTAMP(2, g, 1);
        if (g > 0 && half == 0) stash_q(qb_of(g - 1));  // the previous item's first block, for its dQ read-out below
        if (warp == 2) FSTAMP(2, g, 2);
        asm volatile("bar.sync 1, 256;" ::: "memory");  // L / delta of this item (written an iteration or more ago) visible
      }
      if (i == 2 && it + 1 < my_items) fetch_ld(item + gridDim.x);
      if (warp == 2) FSTAMP(1, g, 0);
      mbar_wait(s_full, g & 1);
      if (warp == 2) FSTAMP(1, g, 1);
      tc_fence_after();
      uint32_t pp[32], pd[32];
      const float* Lq = sL + buf + qb * RT + half * 64;
      const float* Dq = sDl + buf + qb * RT + half * 64;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t sv[32], dp[32];
        tmem_ld32(t_lane + C_S + half * 64 + c2 * 32, sv);
        tmem_ld32(t_lane + C_DP + half * 64 + c2 * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // four queries per step: their L and delta are one 16-byte broadcast load each
          const int c = c2 * 32 + 4 * j;
          const float4 l4 = *reinterpret_cast<const float4*>(Lq + c), d4 = *reinterpret_cast<const float4*>(Dq + c);
          const float p0 = ex2(fmaf(__uint_as_float(sv[4 * j]), c1, -l4.x));
          const float p1 = ex2(fmaf(__uint_as_float(sv[4 * j + 1]), c1, -l4.y));
          const float p2 = ex2(fmaf(__uint_as_float(sv[4 * j + 2]), c1, -l4.z));
          const float p3 = ex2(fmaf(__uint_as_float(sv[4 * j + 3]), c1, -l4.w));
          pp[c2 * 16 + 2 * j] = pack_bf16(p0, p1);
          pp[c2 * 16 + 2 * j + 1] = pack_bf16(p2, p3);
          pd[c2 * 16 + 2 * j] = pack_bf16(p0 * (__uint_as_float(dp[4 * j]) - d4.x), p1 * (__uint_as_float(dp[4 * j + 1]) - d4.y));
          pd[c2 * 16 + 2 * j + 1] = pack_bf16(p2 * (__uint_as_float(dp[4 * j + 2]) - d4.z), p3 * (__uint_as_float(dp[4 * j + 3]) - d4.w));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      if (warp == 2) FSTAMP(1, g, 2);
      if (g > 0) mbar_wait(p_empty, (g - 1) & 1);  // the MMAs that read the staging tiles (and completed dV/dK/dQ) have retired
      store_row_sw128(sPt + half * F_TILE, r, pp);
      store_row_sw128(sdSt + half * F_TILE, r, pd);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 2) FSTAMP(1, g, 3);
      if (i == 2 && it + 1 < my_items) commit_ld(buf ^ F_T);
      // ---- accumulator read-out.  half 0 warps: dV_0 | dQ (second block) | dQ (first block); half 1 warps: dK_0 | - | dV_1, dK_1
      if (i == 1) {          // the second query block is resident (S of this iteration used it): stash its q^ rows
        if (half == 0) stash_q(qb);
        if (warp == 2) FSTAMP(2, g, 0);
      } else if (i == 2) {   // key tile 0 finished with iteration 1
        const int n = item / heads, h = item - n * heads;
        const size_t grow0 = (size_t)n * F_T;
        uint32_t a0[32], a1[32];
        tc_fence_after();
        tmem_ld32(t_lane + (half ? C_DK : C_DV), a0);
        tmem_ld32(t_lane + (half ? C_DK : C_DV) + 32, a1);
        tmem_ld_wait();
        signal_free(acc_free);
        if (half == 0) {
          uint32_t out[32];
          pack_row(a0, a1, 1.0f, out);
          store_slab(sTile + TV0 * F_TILE, out, grow0, 2 * D + h * HD);
          rel_a = TV0;
        } else {
          finish_qk(a0, a1, sTile + TK0 * F_TILE, grow0, D + h * HD, heads + h);
          rel_a = TK0;
        }
        if (warp == 2) FSTAMP(2, g, 0);
      } else if (i == 3) {   // the item's second query block finished with iteration 2 (no MMA overwrites dQ before the next item)
        if (half == 0) {
          const int n = item / heads, h = item - n * heads;
          const int qs = qb_of(g - 1);
          uint32_t a0[32], a1[32];
          tc_fence_after();
          tmem_ld32(t_lane + C_DQ + qs * HD, a0);
          tmem_ld32(t_lane + C_DQ + qs * HD + 32, a1);
          tmem_ld_wait();
          finish_qk(a0, a1, sTile + TOUT * F_TILE, (size_t)n * F_T + qs * RT, h * HD, h);
        }
        if (warp == 2) FSTAMP(2, g, 0);
      } else if (g > 0) {    // i == 0: previous item: key tile 1 and its first query block finished with its iteration 3
        const int pit = item - gridDim.x;
        const int n = pit / heads, h = pit - n * heads;
        tc_fence_after();
        if (half == 0) {
          const int qf = qb_of(g - 1);
          uint32_t a0[32], a1[32];
          signal_free(acc_free);  // (8 arrivals per phase; these warps do not read dV / dK here)
          tmem_ld32(t_lane + C_DQ + qf * HD, a0);
          tmem_ld32(t_lane + C_DQ + qf * HD + 32, a1);
          tmem_ld_wait();
          signal_free(dq_free);
          finish_qk(a0, a1, sTile + TOUT * F_TILE, (size_t)n * F_T + qf * RT, h * HD, h);
        } else {
          const size_t grow0 = (size_t)n * F_T + RT;
          uint32_t a0[32], a1[32], outv[32];
          tmem_ld32(t_lane + C_DV, a0);
          tmem_ld32(t_lane + C_DV + 32, a1);
          tmem_ld_wait();
          pack_row(a0, a1, 1.0f, outv);
          tmem_ld32(t_lane + C_DK, a0);
          tmem_ld32(t_lane + C_DK + 32, a1);
          tmem_ld_wait();
          signal_free(acc_free);
          store_slab(sTile + TV1 * F_TILE, outv, grow0, 2 * D + h * HD);
          finish_qk(a0, a1, sTile + TK1 * F_TILE, grow0, D + h * HD, heads + h);
          rel_a = TK1;
          rel_b = TV1;
        }
        if (warp == 2) FSTAMP(2, g, 0);
      }
    }
    if (G > 0) {  // last item: key tile 1 and the first query block
      mbar_wait(p_empty, (G - 1) & 1);
      tc_fence_after();
      const int item = blockIdx.x + (my_items - 1) * gridDim.x;
      const int n = item / heads, h = item - n * heads;
      uint32_t a0[32], a1[32];
      if (half == 0) {
        const int qf = qb_of(G - 1);
        stash_q(qf);
        tmem_ld32(t_lane + C_DQ + qf * HD, a0);
        tmem_ld32(t_lane + C_DQ + qf * HD + 32, a1);
        tmem_ld_wait();
        finish_qk(a0, a1, sTile + TOUT * F_TILE, (size_t)n * F_T + qf * RT, h * HD, h);
      } else {
        const size_t grow0 = (size_t)n * F_T + RT;
        uint32_t outv[32];
        tmem_ld32(t_lane + C_DV, a0);
        tmem_ld32(t_lane + C_DV + 32, a1);
        tmem_ld_wait();
        pack_row(a0, a1, 1.0f, outv);
        store_slab(sTile + TV1 * F_TILE, outv, grow0, 2 * D + h * HD);
        tmem_ld32(t_lane + C_DK, a0);
        tmem_ld32(t_lane + C_DK + 32, a1);
        tmem_ld_wait();
        finish_qk(a0, a1, sTile + TK1 * F_TILE, grow0, D + h * HD, heads + h);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ fused backward, second cut
// Same data flow, tiles, barriers and MMA schedule as attn_bwd_fused_tc, but the accumulator read-outs have their own four
// warps (10-13) so the eight softmax warps only do  wait S -> exp -> staging -> publish.  14 warps put four warps on two of
// the SM sub-partitions, which caps every thread at 128 registers: the exp pass reads TMEM 16 columns at a time, and a q/k-norm
// read-out makes two passes over its accumulator (dot product, then the output) in 32-column halves instead of holding the whole row.
constexpr int F2_NTHREADS = 448;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 bf16 (columns 32*hf .. 32*hf+31) of row r of a [128 x 64] SWIZZLE_128B tile <-> registers
__device__ __forceinline__ void load_half_row_sw128(const uint8_t* tile, int r, int hf, float (&y)[32]) {
  const uint8_t* prow = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(prow + (((hf * 4 + c) ^ (r & 7)) << 4));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = __bfloat1622float2(h2[e]);
      y[8 * c + 2 * e] = t.x;
      y[8 * c + 2 * e + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void store_half_row_sw128(uint8_t* tile, int r, int hf, const uint32_t (&pk)[16]) {
  uint8_t* prow = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(prow + (((hf * 4 + c) ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

__global__ void __launch_bounds__(F2_NTHREADS, 1)
attn_bwd_fused2_tc(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ lse, const float* __restrict__ delta, int heads,
                   int total_items, const float* __restrict__ sc, float eps, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sTile = smem;                      // 8 input tiles + the dQ staging tile, index = T* enum
  uint8_t* sPt = sTile + 9 * F_TILE;
  uint8_t* sdSt = sPt + F_STG;
  float* sL = reinterpret_cast<float*>(sdSt + F_STG);  // [2][256] log2-domain log-sum-exp of the item's queries
  float* sDl = sL + 2 * F_T;                            // [2][256] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDl + 2 * F_T);
  uint64_t* full = bars;            // [8] input tile landed
  uint64_t* empty = bars + 8;       // [8] input tile may be refilled
  uint64_t* s_full = bars + 16;     // S^T / dP^T of iteration g complete
  uint64_t* s_empty = bars + 17;    // ... read by the 8 softmax warps
  uint64_t* p_full = bars + 18;     // [2] 64-query panel `half` of the P^T / dS^T staging tiles of iteration g written (4 warps each)
  uint64_t* p_empty = bars + 20;    // [2] ... consumed by the dV / dK / dQ MMAs (dQ reads both panels in every k-step: both complete together)
  uint64_t* acc_free = bars + 22;   // dV / dK read out (4 epilogue warps), twice per item
  uint64_t* dqs_free = bars + 23;   // dQ of the item's second query block read out
  uint64_t* dqf_free = bars + 24;   // dQ of the item's first query block read out
  uint64_t* kv0_ready = bars + 25;  // MMA -> epilogue: dV_0 / dK_0 complete (after iteration 1)
  uint64_t* dqs_ready = bars + 26;  // dQ (second block) complete (after iteration 2)
  uint64_t* fin_ready = bars + 27;  // dV_1 / dK_1 / dQ (first block) complete (after iteration 3)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int my_items = blockIdx.x < total_items ? (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int G = my_items * 4;
  auto qb_of = [](int g) { return (((g & 3) == 1 || (g & 3) == 2) ? 1 : 0) ^ ((g >> 2) & 1); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_out);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], i < TD0 ? 5 : 1);  // K / V / Q tiles: the MMA warp + four epilogue warps
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(acc_free, 4);
    mbar_init(dqs_free, 4);
    mbar_init(dqf_free, 4);
    mbar_init(kv0_ready, 1);
    mbar_init(dqs_ready, 1);
    mbar_init(fin_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  constexpr uint32_t C_S = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < my_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int n = item / heads, h = item - n * heads;
        const uint32_t ph = it & 1;
        const int f = it & 1;
        if (it + 1 < my_items) {  // next item's 128 KB into L2, one whole item ahead (see attn_bwd_fused_tc)
          const int nitem = item + gridDim.x;
          const int nn = nitem / heads, nh = nitem - nn * heads;
#pragma unroll
          for (int rb = 0; rb < 2; ++rb) {
            const int row = nn * F_T + rb * RT;
            tma_prefetch_l2_2d(&tm_qkv, nh * HD, row);
            tma_prefetch_l2_2d(&tm_qkv, D + nh * HD, row);
            tma_prefetch_l2_2d(&tm_qkv, 2 * D + nh * HD, row);
            tma_prefetch_l2_2d(&tm_do, nh * HD, row);
          }
        }
        const int order[8] = {TK0, TV0, TQ0 + f, TD0 + f, TQ0 + (f ^ 1), TD0 + (f ^ 1), TK1, TV1};
        for (int k = 0; k < 8; ++k) {
          const int t = order[k];
          mbar_wait(&empty[t], ph ^ 1);
          mbar_arrive_expect_tx(&full[t], F_TILE);
          const int row = n * F_T + (t & 1) * RT;
          if (t >= TD0) tma_load_2d(sTile + t * F_TILE, &tm_do, &full[t], h * HD, row);
          else tma_load_2d(sTile + t * F_TILE, &tm_qkv, &full[t], (t >= TQ0 ? 0 : (t >= TV0 ? 2 * D : D)) + h * HD, row);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(RT, 128, 0, 0);
    constexpr uint32_t idesc_a = make_idesc_bf16(RT, HD, 0, 1);
    constexpr uint32_t idesc_q = make_idesc_bf16(RT, HD, 1, 1);
    const uint32_t tile0 = smem_u32(sTile), pt_addr = smem_u32(sPt), dst_addr = smem_u32(sdSt);
    auto scores = [&](int g) {
      const int i = g & 3, kt = i >> 1, qb = qb_of(g);
      const uint32_t ph = (g >> 2) & 1;
      if ((i & 1) == 0) {
        mbar_wait(&full[TK0 + kt], ph);
        mbar_wait(&full[TV0 + kt], ph);
      }
      if (kt == 0) {
        mbar_wait(&full[TQ0 + qb], ph);
        mbar_wait(&full[TD0 + qb], ph);
      }
      mbar_wait(s_empty, (g & 1) ^ 1);
      tc_fence_after();
      const uint32_t k_addr = tile0 + (TK0 + kt) * F_TILE, v_addr = tile0 + (TV0 + kt) * F_TILE;
      const uint32_t q_addr = tile0 + (TQ0 + qb) * F_TILE, do_addr = tile0 + (TD0 + qb) * F_TILE;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        if (leader) umma_ss(tmem_base + C_S, make_smem_desc(k_addr + k * 32, 16, 1024), make_smem_desc(q_addr + k * 32, 16, 1024), idesc_s, k != 0);
        if (leader) umma_ss(tmem_base + C_DP, make_smem_desc(v_addr + k * 32, 16, 1024), make_smem_desc(do_addr + k * 32, 16, 1024), idesc_s, k != 0);
      }
      if (leader) umma_commit(s_full);
    };
    auto tiles_ready = [&](int g) {
      const int i = g & 3, kt = i >> 1, qb = qb_of(g);
      const uint32_t ph = (g >> 2) & 1;
      bool ok = true;
      if ((i & 1) == 0) ok = mbar_test_wait(smem_u32(&full[TK0 + kt]), ph) && mbar_test_wait(smem_u32(&full[TV0 + kt]), ph);
      if (kt == 0) ok = ok && mbar_test_wait(smem_u32(&full[TQ0 + qb]), ph) && mbar_test_wait(smem_u32(&full[TD0 + qb]), ph);
      return __all_sync(0xffffffffu, ok) != 0;
    };
    if (G > 0) scores(0);
    for (int g = 0; g < G; ++g) {
      const int i = g & 3, kt = i >> 1, qb = qb_of(g), it = g >> 2;
      bool issued = g + 1 >= G;
      auto phase2_ready = [&]() {
        bool ok = mbar_test_wait(smem_u32(&p_full[0]), g & 1) && mbar_test_wait(smem_u32(&p_full[1]), g & 1);
        if (i == 2) ok = ok && mbar_test_wait(smem_u32(acc_free), 0);                         // dV_0 / dK_0 read out
        if (i == 0 && it > 0) ok = ok && mbar_test_wait(smem_u32(acc_free), 1)                // previous item: dV_1 / dK_1 ...
                                   && mbar_test_wait(smem_u32(dqs_free), (it - 1) & 1);        // ... and the dQ this block's columns held
        if (i == 1 && it > 0) ok = ok && mbar_test_wait(smem_u32(dqf_free), (it - 1) & 1);
        return __all_sync(0xffffffffu, ok) != 0;
      };
      {
        const long long t0 = clock64();
        for (;;) {
          if (!issued && tiles_ready(g + 1)) {
            scores(g + 1);
            issued = true;
          }
          if (phase2_ready()) break;
          if (clock64() - t0 > 4000000000LL) {
            if (lane == 0) printf("mapdit: attn_bwd_fused2 MMA warp timed out (block %d iteration %d)\n", blockIdx.x, g);
            __trap();
          }
        }
      }
      FSTAMP(0, g, 1);
      tc_fence_after();
      const uint32_t k_addr = tile0 + (TK0 + kt) * F_TILE;
      const uint32_t q_addr = tile0 + (TQ0 + qb) * F_TILE, do_addr = tile0 + (TD0 + qb) * F_TILE;
      // three independent accumulation chains, issued round-robin (back-to-back MMAs into one accumulator serialise)
#pragma unroll
      for (int k = 0; k < 128 / 16; ++k) {
        const uint32_t a_off = (k >> 2) * F_TILE + (k & 3) * 32;
        if (leader) umma_ss(tmem_base + C_DV, make_smem_desc(pt_addr + a_off, 16, 1024), make_smem_desc(do_addr + k * 2048, 1024, 1024), idesc_a,
                ((i & 1) | k) != 0);
        if (leader) umma_ss(tmem_base + C_DK, make_smem_desc(dst_addr + a_off, 16, 1024), make_smem_desc(q_addr + k * 2048, 1024, 1024), idesc_a,
                ((i & 1) | k) != 0);
        if (leader) umma_ss(tmem_base + C_DQ + qb * HD, make_smem_desc(dst_addr + k * 2048, F_TILE, 1024), make_smem_desc(k_addr + k * 2048, 1024, 1024),
                idesc_q, (kt | k) != 0);
      }
      if (leader) umma_commit(&p_empty[0]);
      if (leader) umma_commit(&p_empty[1]);
      FSTAMP(0, g, 2);
      if (i == 1) {
        if (leader) umma_commit(kv0_ready);
        if (leader) umma_commit(&empty[TK0]);
        if (leader) umma_commit(&empty[TV0]);
      } else if (i == 2) {
        if (leader) umma_commit(dqs_ready);
        if (leader) umma_commit(&empty[TQ0 + qb]);
        if (leader) umma_commit(&empty[TD0 + qb]);
      } else if (i == 3) {
        if (leader) umma_commit(fin_ready);
        if (leader) umma_commit(&empty[TQ0 + qb]);
        if (leader) umma_commit(&empty[TD0 + qb]);
        if (leader) umma_commit(&empty[TK1]);
        if (leader) umma_commit(&empty[TV1]);
      }
      if (!issued) scores(g + 1);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ softmax warps: S^T, dP^T -> P^T, dS^T staging tiles
    const int wq = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = wq * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float c1 = 0.125f * LOG2E;
    // per-query L and delta of an item go global -> shared with cp.async: no register holds them across the exp pass (under the
    // 128-register cap the compiler spilled such a register right behind the load, stalling the warp for an HBM round trip)
    auto fetch_ld = [&](int item, int buf) {
      const int n = item / heads, h = item - n * heads;
      const size_t qrow = (size_t)n * F_T + tid;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sL + buf + tid)), "l"(lse + qrow * heads + h) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sDl + buf + tid)), "l"(delta + qrow * heads + h) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto land_ld = [&](int buf) {  // own element arrived: natural log -> log2 domain, then the 256-thread barrier publishes it
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      sL[buf + tid] *= LOG2E;
    };
    if (G > 0) fetch_ld(blockIdx.x, 0);
    for (int g = 0; g < G; ++g) {
      const int i = g & 3, qb = qb_of(g), it = g >> 2;
      const int item = blockIdx.x + it * gridDim.x;
      const int buf = (it & 1) * F_T;
      if (i == 0) {
        land_ld(buf);
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (i == 2 && it + 1 < my_items) fetch_ld(item + gridDim.x, buf ^ F_T);
      if (warp == 2) FSTAMP(1, g, 0);
      mbar_wait(s_full, g & 1);
      if (warp == 2) FSTAMP(1, g, 1);
      tc_fence_after();
      const float* Lq = sL + buf + qb * RT + half * 64;
      const float* Dq = sDl + buf + qb * RT + half * 64;
      // 32 queries at a time, straight into the staging tiles (96 live registers under the 128 cap; holding the whole 64-query
      // row needed 16-column TMEM loads and four wait round trips: 2.0 k instead of 1.3 k cycles per pass)
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t sv[32], dp[32], pp[16], pd[16];
        tmem_ld32(t_lane + C_S + half * 64 + c2 * 32, sv);
        tmem_ld32(t_lane + C_DP + half * 64 + c2 * 32, dp);
        tmem_ld_wait();
        if (c2 == 1) {  // last TMEM read of this iteration: S / dP may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c2 * 32 + 4 * j;
          const float4 l4 = *reinterpret_cast<const float4*>(Lq + c), d4 = *reinterpret_cast<const float4*>(Dq + c);
          const float p0 = ex2(fmaf(__uint_as_float(sv[4 * j]), c1, -l4.x));
          const float p1 = ex2(fmaf(__uint_as_float(sv[4 * j + 1]), c1, -l4.y));
          const float p2 = ex2(fmaf(__uint_as_float(sv[4 * j + 2]), c1, -l4.z));
          const float p3 = ex2(fmaf(__uint_as_float(sv[4 * j + 3]), c1, -l4.w));
          pp[2 * j] = pack_bf16(p0, p1);
          pp[2 * j + 1] = pack_bf16(p2, p3);
          pd[2 * j] = pack_bf16(p0 * (__uint_as_float(dp[4 * j]) - d4.x), p1 * (__uint_as_float(dp[4 * j + 1]) - d4.y));
          pd[2 * j + 1] = pack_bf16(p2 * (__uint_as_float(dp[4 * j + 2]) - d4.z), p3 * (__uint_as_float(dp[4 * j + 3]) - d4.w));
        }
        if (c2 == 0) {
          if (warp == 2) FSTAMP(1, g, 2);
          if (g > 0) mbar_wait(&p_empty[half], (g - 1) & 1);  // the MMAs that read this panel of the staging tiles have retired
        }
        store_half_row_sw128(sPt + half * F_TILE, r, c2, pp);
        store_half_row_sw128(sdSt + half * F_TILE, r, c2, pd);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[half]);
      if (warp == 2) FSTAMP(1, g, 3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 10-13: accumulator read-outs, q^ stashes
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    auto signal_free = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto slab_store = [&](uint8_t* stile, size_t grow0, int gcol) {  // rows of this warp written: 32-row slab -> dqkv by TMA
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_store_2d(&tm_out, stile + wq * 32 * 128, gcol, (int)(grow0 + wq * 32));
    };
    auto release = [&](int t0, int t1) {  // the stores issued so far have read their smem: hand input tiles back
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&empty[t0]);
        if (t1 >= 0) mbar_arrive(&empty[t1]);
      }
      __syncwarp();
    };
    // dV: plain copy of a finished accumulator into (dead) tile `stile`
    auto readout_plain = [&](uint32_t col, uint8_t* stile, float scale) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t a[32], out[16];
        tmem_ld32(t_lane + col + 32 * hf, a);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) out[c] = pack_bf16(__uint_as_float(a[2 * c]) * scale, __uint_as_float(a[2 * c + 1]) * scale);
        store_half_row_sw128(stile, r, hf, out);
      }
    };
    // dQ / dK: q/k-normalisation backward against the normalised row in `stile` (see store_out_row_qknorm), written over it.
    // Two passes over the accumulator in 32-column halves (register cap 128); `done` is signalled after the last TMEM read.
    auto readout_qk = [&](uint32_t col, uint8_t* stile, float s, uint64_t* done) {
      float dot = 0.f;
      if (sc) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t a[32];
          float y[32];
          tmem_ld32(t_lane + col + 32 * hf, a);
          load_half_row_sw128(stile, r, hf, y);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) dot = fmaf(y[c], __uint_as_float(a[c]), dot);
        }
      }
      dot *= 0.125f;
      const float rpe = 8.0f / s;
      const float rr = fmaxf(rpe - eps, 1e-30f);
      const float sco = sc ? -s * (dot * rpe / (64.0f * rr)) : 0.f;
      const float ga = (sc ? s : 1.0f) * 0.125f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t a[32], out[16];
        float y[32];
        tmem_ld32(t_lane + col + 32 * hf, a);
        load_half_row_sw128(stile, r, hf, y);
        tmem_ld_wait();
        if (hf == 1 && done) signal_free(done);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          out[c] = pack_bf16(fmaf(ga, __uint_as_float(a[2 * c]), sco * y[2 * c]), fmaf(ga, __uint_as_float(a[2 * c + 1]), sco * y[2 * c + 1]));
        store_half_row_sw128(stile, r, hf, out);
      }
    };
    auto stash_q = [&](int qb) {  // this thread's q^ row of query block qb -> the dQ staging tile; then the Q tile may be refilled
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128;
      const uint4* src = reinterpret_cast<const uint4*>(sTile + (TQ0 + qb) * F_TILE + off);
      uint4* dst = reinterpret_cast<uint4*>(sTile + TOUT * F_TILE + off);
      uint4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = src[c ^ (r & 7)];
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c ^ (r & 7)] = v[c];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[TQ0 + qb]);
    };
    for (int it = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int n = item / heads, h = item - n * heads;
      const uint32_t ph = it & 1;
      const int qf = it & 1, qs = qf ^ 1;  // first / second query block of this item
      const size_t row0 = (size_t)n * F_T;
      float s_q = 1.f, s_k = 1.f;
      // ---- second query block resident from iteration 1 on: stash its q^ rows (dQ staging tile is free: its last store was read)
      mbar_wait(&full[TQ0 + qs], ph);
      stash_q(qs);
      // ---- key tile 0 complete after iteration 1
      if (sc) s_k = sc[(row0 + r) * 2 * heads + heads + h];
      mbar_wait(kv0_ready, ph);
      tc_fence_after();
      readout_plain(C_DV, sTile + TV0 * F_TILE, 1.0f);
      readout_qk(C_DK, sTile + TK0 * F_TILE, s_k, acc_free);
      slab_store(sTile + TV0 * F_TILE, row0, 2 * D + h * HD);
      slab_store(sTile + TK0 * F_TILE, row0, D + h * HD);
      release(TK0, TV0);  // the next item needs these two first: hand them back before anything else
      if (warp == 10) FSTAMP(2, 4 * it, 0);
      // ---- dQ of the second block complete after iteration 2
      if (sc) s_q = sc[(row0 + qs * RT + r) * 2 * heads + h];
      mbar_wait(dqs_ready, ph);
      tc_fence_after();
      readout_qk(C_DQ + qs * HD, sTile + TOUT * F_TILE, s_q, dqs_free);
      slab_store(sTile + TOUT * F_TILE, row0 + qs * RT, h * HD);
      if (warp == 10) FSTAMP(2, 4 * it, 1);
      // ---- first query block: stash (its tile is still resident: this warp's arrival is part of its release)
      stash_q(qf);
      // ---- key tile 1 and the first block's dQ complete after iteration 3
      if (sc) {
        s_k = sc[(row0 + RT + r) * 2 * heads + heads + h];
        s_q = sc[(row0 + qf * RT + r) * 2 * heads + h];
      }
      mbar_wait(fin_ready, ph);
      tc_fence_after();
      readout_plain(C_DV, sTile + TV1 * F_TILE, 1.0f);
      readout_qk(C_DK, sTile + TK1 * F_TILE, s_k, acc_free);
      slab_store(sTile + TV1 * F_TILE, row0 + RT, 2 * D + h * HD);
      slab_store(sTile + TK1 * F_TILE, row0 + RT, D + h * HD);
      readout_qk(C_DQ + qf * HD, sTile + TOUT * F_TILE, s_q, dqf_free);
      slab_store(sTile + TOUT * F_TILE, row0 + qf * RT, h * HD);
      release(TK1, TV1);
      if (warp == 10) FSTAMP(2, 4 * it, 2);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ fused backward, third cut
// The producer and the read-out warps of attn_bwd_fused2_tc, a different inner loop.  fused2 handed the whole [128 keys x 128 queries]
// P^T and dS^T tiles through shared memory: 64 KB of stores, a proxy fence, 24 MMAs with both operands in shared memory (48 cycles
// each, shared-memory bound) and no second S^T / dP^T buffer, so softmax and tensor core strictly alternated (4.6 k cycles per
// iteration, tools/attn_bwd_timeline.py).  Here an iteration is split into two HALF-iterations of 64 queries: S^T / dP^T are 64 TMEM
// columns each, which leaves room for two buffers; P^T and dS^T go back into TMEM in place and are the A operands of the dV / dK MMAs
// (as in attn_bwd_dkv_tc); only dS^T is also staged in shared memory, for the transposed read of the dQ MMA (issued every second
// half-iteration over both panels), in one of two staging tiles.
__global__ void __launch_bounds__(F2_NTHREADS, 1)
attn_bwd_fused3_tc(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ lse, const float* __restrict__ delta, int heads,
                   int total_items, const float* __restrict__ sc, float eps, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sTile = smem;                      // 8 input tiles + the dQ staging tile, index = T* enum
  uint8_t* sStg = sTile + 9 * F_TILE;  // 2 x dS^T staging tile ([128 keys x 128 queries] as two 64-query panels), for the dQ MMAs only
  float* sL = reinterpret_cast<float*>(sStg + 2 * F_STG);  // [2][256] log2-domain log-sum-exp of the item's queries
  float* sDl = sL + 2 * F_T;                            // [2][256] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDl + 2 * F_T);
  uint64_t* full = bars;            // [8] input tile landed
  uint64_t* empty = bars + 8;       // [8] input tile may be refilled
  uint64_t* s_full = bars + 16;     // [2] S^T / dP^T of half-iteration u complete in TMEM buffer u & 1
  uint64_t* p_full = bars + 18;     // [2] P^T / dS^T of half-iteration u written (TMEM in place; dS^T also into its staging panel), 8 warps
  uint64_t* stg_empty = bars + 20;  // [2] the dQ MMAs of iteration g have read staging tile g & 1
  uint64_t* acc_free = bars + 22;   // dV / dK read out (4 epilogue warps), twice per item
  uint64_t* dqs_free = bars + 23;   // dQ of the item's second query block read out
  uint64_t* dqf_free = bars + 24;   // dQ of the item's first query block read out
  uint64_t* kv0_ready = bars + 25;  // MMA -> epilogue: dV_0 / dK_0 complete (after iteration 1)
  uint64_t* dqs_ready = bars + 26;  // dQ (second block) complete (after iteration 2)
  uint64_t* fin_ready = bars + 27;  // dV_1 / dK_1 / dQ (first block) complete (after iteration 3)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int my_items = blockIdx.x < total_items ? (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int G = my_items * 8;  // half-iterations u = 8 * it + w: w >> 1 = the (key tile, query block) step of the item, w & 1 = which 64 queries of the block
  auto qb_of = [](int g) { return (((g & 3) == 1 || (g & 3) == 2) ? 1 : 0) ^ ((g >> 2) & 1); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_out);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], i < TD0 ? 5 : 1);  // K / V / Q tiles: the MMA warp + four epilogue warps
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&stg_empty[i], 1);
    }
    mbar_init(acc_free, 4);
    mbar_init(dqs_free, 4);
    mbar_init(dqf_free, 4);
    mbar_init(kv0_ready, 1);
    mbar_init(dqs_ready, 1);
    mbar_init(fin_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // two (S^T [0, 64) | dP^T [64, 128)) buffers for alternate half-iterations, then the accumulators
  constexpr uint32_t C_DV = 256, C_DK = 320, C_DQ = 384;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < my_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int n = item / heads, h = item - n * heads;
        const uint32_t ph = it & 1;
        const int f = it & 1;
        if (it + 1 < my_items) {  // next item's 128 KB into L2, one whole item ahead (see attn_bwd_fused_tc)
          const int nitem = item + gridDim.x;
          const int nn = nitem / heads, nh = nitem - nn * heads;
#pragma unroll
          for (int rb = 0; rb < 2; ++rb) {
            const int row = nn * F_T + rb * RT;
            tma_prefetch_l2_2d(&tm_qkv, nh * HD, row);
            tma_prefetch_l2_2d(&tm_qkv, D + nh * HD, row);
            tma_prefetch_l2_2d(&tm_qkv, 2 * D + nh * HD, row);
            tma_prefetch_l2_2d(&tm_do, nh * HD, row);
          }
        }
        const int order[8] = {TK0, TV0, TQ0 + f, TD0 + f, TQ0 + (f ^ 1), TD0 + (f ^ 1), TK1, TV1};
        for (int k = 0; k < 8; ++k) {
          const int t = order[k];
          mbar_wait(&empty[t], ph ^ 1);
          mbar_arrive_expect_tx(&full[t], F_TILE);
          const int row = n * F_T + (t & 1) * RT;
          if (t >= TD0) tma_load_2d(sTile + t * F_TILE, &tm_do, &full[t], h * HD, row);
          else tma_load_2d(sTile + t * F_TILE, &tm_qkv, &full[t], (t >= TQ0 ? 0 : (t >= TV0 ? 2 * D : D)) + h * HD, row);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(RT, 64, 0, 0);   // S^T = K Q^T, dP^T = V dO^T over 64 queries
    constexpr uint32_t idesc_a = make_idesc_bf16(RT, HD, 0, 1);   // dV += P^T dO, dK += dS^T Q: A from TMEM, B MN-major
    constexpr uint32_t idesc_q = make_idesc_bf16(RT, HD, 1, 1);   // dQ += dS K: A = the dS^T staging tile read MN-major, B MN-major
    const uint64_t d_tile = make_smem_desc(smem_u32(sTile), 16, 1024);      // K-major view of input tile 0
    const uint64_t d_tmn = make_smem_desc(smem_u32(sTile), 1024, 1024);     // MN-major view of input tile 0 (B of dV / dK / dQ)
    const uint64_t d_stg = make_smem_desc(smem_u32(sStg), F_TILE, 1024);    // MN-major view of staging tile 0 (A of dQ): LBO = panel stride
    auto scores = [&](int u) {
      const int w = u & 7, pos = w >> 1, sub = w & 1, kt = pos >> 1, qb = qb_of(u >> 1), b = u & 1;
      const uint32_t ph = (u >> 3) & 1;
      if ((w & 3) == 0) {
        mbar_wait(&full[TK0 + kt], ph);
        mbar_wait(&full[TV0 + kt], ph);
      }
      if (kt == 0 && sub == 0) {
        mbar_wait(&full[TQ0 + qb], ph);
        mbar_wait(&full[TD0 + qb], ph);
      }
      const uint64_t d_k = desc_advance(d_tile, (TK0 + kt) * F_TILE), d_v = desc_advance(d_tile, (TV0 + kt) * F_TILE);
      const uint64_t d_q = desc_advance(d_tile, (TQ0 + qb) * F_TILE + sub * (F_TILE / 2));
      const uint64_t d_do = desc_advance(d_tile, (TD0 + qb) * F_TILE + sub * (F_TILE / 2));
      const uint32_t t_s = tmem_base + b * 128;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        if (leader) umma_ss(t_s, desc_advance(d_k, k * 32), desc_advance(d_q, k * 32), idesc_s, k != 0);
        if (leader) umma_ss(t_s + 64, desc_advance(d_v, k * 32), desc_advance(d_do, k * 32), idesc_s, k != 0);
      }
      if (leader) umma_commit(&s_full[b]);
    };
    if (G > 0) scores(0);
    if (G > 1) scores(1);
    for (int u = 0; u < G; ++u) {
      const int w = u & 7, pos = w >> 1, sub = w & 1, kt = pos >> 1, g = u >> 1, qb = qb_of(g), it = u >> 3, b = u & 1;
      // accumulators this half-iteration starts from zero must have been read out
      if (w == 4) mbar_wait(acc_free, 0);                                   // dV_0 / dK_0
      if (w == 0 && it > 0) mbar_wait(acc_free, 1);                          // previous item: dV_1 / dK_1
      if (w == 1 && it > 0) mbar_wait(dqs_free, (it - 1) & 1);               // ... the dQ these columns held (its second block)
      if (w == 3 && it > 0) mbar_wait(dqf_free, (it - 1) & 1);               // ... and its first block
      mbar_wait_spin(&p_full[b], (u >> 1) & 1);
      FSTAMP(0, u, 1);
      tc_fence_after();
      const uint64_t d_do = desc_advance(d_tmn, (TD0 + qb) * F_TILE + sub * (F_TILE / 2));
      const uint64_t d_q = desc_advance(d_tmn, (TQ0 + qb) * F_TILE + sub * (F_TILE / 2));
      const uint32_t t_p = tmem_base + b * 128;  // P^T over the S^T columns, dS^T over the dP^T columns (see the softmax warps)
#pragma unroll
      for (int k = 0; k < 64 / 16; ++k) {  // 16 queries per k-step
        const uint32_t ko = (k >> 1) * 32 + (k & 1) * 8;
        if (leader) umma_ts(tmem_base + C_DV, t_p + ko, desc_advance(d_do, k * 2048), idesc_a, ((w & 3) | k) != 0);
        if (leader) umma_ts(tmem_base + C_DK, t_p + 64 + ko, desc_advance(d_q, k * 2048), idesc_a, ((w & 3) | k) != 0);
      }
      if (sub == 1) {  // both 64-query panels of the block's dS^T are staged: dQ_qb += dS K_kt over the 128 keys
        const uint64_t d_a = desc_advance(d_stg, (g & 1) * F_STG), d_kk = desc_advance(d_tmn, (TK0 + kt) * F_TILE);
#pragma unroll
        for (int k = 0; k < 128 / 16; ++k)
          if (leader) umma_ss(tmem_base + C_DQ + qb * HD, desc_advance(d_a, k * 2048), desc_advance(d_kk, k * 2048), idesc_q, (kt | k) != 0);
        if (leader) umma_commit(&stg_empty[g & 1]);
      }
      // (kv0_ready only behind the dQ MMAs of w == 3, the last readers of the K_0 tile: the read-out warps overwrite that tile with dK_0.
      // Committing it ahead of them passed the tests three times and then corrupted 4 % of the gradient when the timing shifted.)
      if (w == 3 && leader) umma_commit(kv0_ready);
      FSTAMP(0, u, 2);
      if (w == 3) {
        if (leader) umma_commit(&empty[TK0]);
        if (leader) umma_commit(&empty[TV0]);
      } else if (w == 5) {
        if (leader) umma_commit(dqs_ready);
        if (leader) umma_commit(&empty[TQ0 + qb]);
        if (leader) umma_commit(&empty[TD0 + qb]);
      } else if (w == 7) {
        if (leader) umma_commit(fin_ready);
        if (leader) umma_commit(&empty[TQ0 + qb]);
        if (leader) umma_commit(&empty[TD0 + qb]);
        if (leader) umma_commit(&empty[TK1]);
        if (leader) umma_commit(&empty[TV1]);
      }
      // S^T / dP^T two half-iterations ahead, into the buffer whose P^T / dS^T the TS MMAs above read: the tensor core runs one thread's
      // MMAs in issue order.  (Ahead of the dQ MMAs instead of behind them: measured slower.)
      if (u + 2 < G) scores(u + 2);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ softmax warps: S^T, dP^T -> P^T, dS^T staging tiles
    const int wq = warp & 3;
    const int h0 = (warp - 2) >> 2;  // which 32 of the 64 queries of a half-iteration
    const int r = wq * 32 + lane;    // key row
    const int tid = threadIdx.x - 64;
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float c1 = 0.125f * LOG2E;
    // per-query L and delta of an item go global -> shared with cp.async: no register holds them across the exp pass
    auto fetch_ld = [&](int item, int buf) {
      const int n = item / heads, h = item - n * heads;
      const size_t qrow = (size_t)n * F_T + tid;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sL + buf + tid)), "l"(lse + qrow * heads + h) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sDl + buf + tid)), "l"(delta + qrow * heads + h) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto land_ld = [&](int buf) {  // own element arrived: natural log -> log2 domain, then the 256-thread barrier publishes it
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      sL[buf + tid] *= LOG2E;
    };
    if (G > 0) fetch_ld(blockIdx.x, 0);
    for (int u = 0; u < G; ++u) {
      const int w = u & 7, sub = w & 1, g = u >> 1, qb = qb_of(g), it = u >> 3, b = u & 1;
      const int item = blockIdx.x + it * gridDim.x;
      const int buf = (it & 1) * F_T;
      if (w == 0) {
        land_ld(buf);
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (w == 4 && it + 1 < my_items) fetch_ld(item + gridDim.x, buf ^ F_T);
      if (warp == 2) FSTAMP(1, u, 0);
      mbar_wait(&s_full[b], (u >> 1) & 1);
      if (warp == 2) FSTAMP(1, u, 1);
      tc_fence_after();
      const float4* pL = reinterpret_cast<const float4*>(sL + buf + qb * RT + sub * 64 + h0 * 32);
      const float4* pD = reinterpret_cast<const float4*>(sDl + buf + qb * RT + sub * 64 + h0 * 32);
      uint32_t sv[32], dp[32], pp[16], pd[16];
      tmem_ld32(t_lane + b * 128 + h0 * 32, sv);
      tmem_ld32(t_lane + b * 128 + 64 + h0 * 32, dp);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 l4 = pL[j], d4 = pD[j];
        const float p0 = ex2(fmaf(__uint_as_float(sv[4 * j]), c1, -l4.x));
        const float p1 = ex2(fmaf(__uint_as_float(sv[4 * j + 1]), c1, -l4.y));
        const float p2 = ex2(fmaf(__uint_as_float(sv[4 * j + 2]), c1, -l4.z));
        const float p3 = ex2(fmaf(__uint_as_float(sv[4 * j + 3]), c1, -l4.w));
        pp[2 * j] = pack_bf16(p0, p1);
        pp[2 * j + 1] = pack_bf16(p2, p3);
        pd[2 * j] = pack_bf16(p0 * (__uint_as_float(dp[4 * j]) - d4.x), p1 * (__uint_as_float(dp[4 * j + 1]) - d4.y));
        pd[2 * j + 1] = pack_bf16(p2 * (__uint_as_float(dp[4 * j + 2]) - d4.z), p3 * (__uint_as_float(dp[4 * j + 3]) - d4.w));
      }
      if (warp == 2) FSTAMP(1, u, 2);
      // in place, over columns only this thread reads: 32 queries = 16 packed columns at the start of the warp's 32-column group; these
      // are the A operands of the dV / dK MMAs
      tmem_st16(t_lane + b * 128 + h0 * 32, pp);
      tmem_st16(t_lane + b * 128 + 64 + h0 * 32, pd);
      // dS^T once more into panel `sub` of the block's staging tile: the dQ MMAs read it transposed (MN-major), which TMEM cannot serve
      if (sub == 0 && g >= 2) mbar_wait(&stg_empty[g & 1], ((g >> 1) & 1) ^ 1);  // the dQ MMAs of iteration g - 2 have read the tile
      store_half_row_sw128(sStg + (g & 1) * F_STG + sub * F_TILE, r, h0, pd);
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      if (warp == 2) FSTAMP(1, u, 3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 10-13: accumulator read-outs, q^ stashes
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    auto signal_free = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto slab_store = [&](uint8_t* stile, size_t grow0, int gcol) {  // rows of this warp written: 32-row slab -> dqkv by TMA
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_store_2d(&tm_out, stile + wq * 32 * 128, gcol, (int)(grow0 + wq * 32));
    };
    auto release = [&](int t0, int t1) {  // the stores issued so far have read their smem: hand input tiles back
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&empty[t0]);
        if (t1 >= 0) mbar_arrive(&empty[t1]);
      }
      __syncwarp();
    };
    // dV: plain copy of a finished accumulator into (dead) tile `stile`
    auto readout_plain = [&](uint32_t col, uint8_t* stile, float scale) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t a[32], out[16];
        tmem_ld32(t_lane + col + 32 * hf, a);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) out[c] = pack_bf16(__uint_as_float(a[2 * c]) * scale, __uint_as_float(a[2 * c + 1]) * scale);
        store_half_row_sw128(stile, r, hf, out);
      }
    };
    // dQ / dK: q/k-normalisation backward against the normalised row in `stile` (see store_out_row_qknorm), written over it.  The whole
    // accumulator row goes into registers first and `done` is signalled right behind that one TMEM round trip: the MMA warp waits for
    // it before the next key tile / item may accumulate into these columns (the two-pass read-out of fused2 signalled ~1.5 k cycles later).
    auto readout_qk = [&](uint32_t col, uint8_t* stile, float s, uint64_t* done) {
      uint32_t a0[32], a1[32];
      tmem_ld32(t_lane + col, a0);
      tmem_ld32(t_lane + col + 32, a1);
      tmem_ld_wait();
      if (done) signal_free(done);
      float dot = 0.f;
      if (sc) {
        float y[32];
        load_half_row_sw128(stile, r, 0, y);
#pragma unroll
        for (int c = 0; c < 32; ++c) dot = fmaf(y[c], __uint_as_float(a0[c]), dot);
        load_half_row_sw128(stile, r, 1, y);
#pragma unroll
        for (int c = 0; c < 32; ++c) dot = fmaf(y[c], __uint_as_float(a1[c]), dot);
      }
      dot *= 0.125f;
      const float rpe = 8.0f / s;
      const float rr = fmaxf(rpe - eps, 1e-30f);
      const float sco = sc ? -s * (dot * rpe / (64.0f * rr)) : 0.f;
      const float ga = (sc ? s : 1.0f) * 0.125f;
      {
        uint32_t out[16];
        float y[32];
        load_half_row_sw128(stile, r, 0, y);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          out[c] = pack_bf16(fmaf(ga, __uint_as_float(a0[2 * c]), sco * y[2 * c]), fmaf(ga, __uint_as_float(a0[2 * c + 1]), sco * y[2 * c + 1]));
        store_half_row_sw128(stile, r, 0, out);
        load_half_row_sw128(stile, r, 1, y);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          out[c] = pack_bf16(fmaf(ga, __uint_as_float(a1[2 * c]), sco * y[2 * c]), fmaf(ga, __uint_as_float(a1[2 * c + 1]), sco * y[2 * c + 1]));
        store_half_row_sw128(stile, r, 1, out);
      }
    };
    auto stash_q = [&](int qb) {  // this thread's q^ row of query block qb -> the dQ staging tile; then the Q tile may be refilled
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128;
      const uint4* src = reinterpret_cast<const uint4*>(sTile + (TQ0 + qb) * F_TILE + off);
      uint4* dst = reinterpret_cast<uint4*>(sTile + TOUT * F_TILE + off);
      uint4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = src[c ^ (r & 7)];
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c ^ (r & 7)] = v[c];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[TQ0 + qb]);
    };
    for (int it = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int n = item / heads, h = item - n * heads;
      const uint32_t ph = it & 1;
      const int qf = it & 1, qs = qf ^ 1;  // first / second query block of this item
      const size_t row0 = (size_t)n * F_T;
      float s_q = 1.f, s_k = 1.f;
      // ---- second query block resident from iteration 1 on: stash its q^ rows (dQ staging tile is free: its last store was read)
      mbar_wait(&full[TQ0 + qs], ph);
      stash_q(qs);
      // ---- key tile 0 complete after iteration 1
      if (sc) s_k = sc[(row0 + r) * 2 * heads + heads + h];
      mbar_wait(kv0_ready, ph);
      tc_fence_after();
      readout_plain(C_DV, sTile + TV0 * F_TILE, 1.0f);
      readout_qk(C_DK, sTile + TK0 * F_TILE, s_k, acc_free);
      slab_store(sTile + TV0 * F_TILE, row0, 2 * D + h * HD);
      slab_store(sTile + TK0 * F_TILE, row0, D + h * HD);
      release(TK0, TV0);  // the next item needs these two first: hand them back before anything else
      if (warp == 10) FSTAMP(2, 8 * it, 0);
      // ---- dQ of the second block complete after iteration 2
      if (sc) s_q = sc[(row0 + qs * RT + r) * 2 * heads + h];
      mbar_wait(dqs_ready, ph);
      tc_fence_after();
      readout_qk(C_DQ + qs * HD, sTile + TOUT * F_TILE, s_q, dqs_free);
      slab_store(sTile + TOUT * F_TILE, row0 + qs * RT, h * HD);
      if (warp == 10) FSTAMP(2, 8 * it, 1);
      // ---- first query block: stash (its tile is still resident: this warp's arrival is part of its release)
      stash_q(qf);
      // ---- key tile 1 and the first block's dQ complete after iteration 3
      if (sc) {
        s_k = sc[(row0 + RT + r) * 2 * heads + heads + h];
        s_q = sc[(row0 + qf * RT + r) * 2 * heads + h];
      }
      mbar_wait(fin_ready, ph);
      tc_fence_after();
      readout_plain(C_DV, sTile + TV1 * F_TILE, 1.0f);
      readout_qk(C_DK, sTile + TK1 * F_TILE, s_k, acc_free);
      slab_store(sTile + TV1 * F_TILE, row0 + RT, 2 * D + h * HD);
      slab_store(sTile + TK1 * F_TILE, row0 + RT, D + h * HD);
      readout_qk(C_DQ + qf * HD, sTile + TOUT * F_TILE, s_q, dqf_free);
      slab_store(sTile + TOUT * F_TILE, row0 + qf * RT, h * HD);
      release(TK1, TV1);
      if (warp == 10) FSTAMP(2, 8 * it, 2);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <int HDV>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                                         long long m_heads) {
  // delta[row, head] = dO_row,head . O_row,head over HDV channels (128 or 144 contiguous bytes of each tensor; the [rows, heads * HDV]
  // tensors are one contiguous run of (row, head) segments).  LPP lanes share one (row, head): lane l reads the 16-byte chunk l % LPP
  // (nine of sixteen lanes at HDV 72), so a load instruction of a warp covers whole contiguous segments (a thread-per-row layout touches
  // 32 different lines per instruction and ran at 60 % of the HBM peak; inside the dq kernel it cost 12 000 cycles per CTA); all 16 loads
  // of a thread are issued before the first use.
  constexpr int LPP = HDV == 64 ? 8 : 16, NCH = HDV / 8, PPI = 32 / LPP;  // lanes per pair, chunks per pair, pairs per load instruction
  const int lane = threadIdx.x & 31, sub = lane / LPP, ch = lane % LPP;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long base = warp_id * (8 * PPI);
  if (base >= m_heads) return;
  uint4 a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long i = base + k * PPI + sub;
    const bool ok = i < m_heads && ch < NCH;
    a[k] = ok ? reinterpret_cast<const uint4*>(o + i * HDV)[ch] : make_uint4(0, 0, 0, 0);
    b[k] = ok ? reinterpret_cast<const uint4*>(dout + i * HDV)[ch] : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a[k]);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b[k]);
    float dl = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(ha[e]), y = __bfloat1622float2(hb[e]);
      dl = fmaf(x.x, y.x, fmaf(x.y, y.y, dl));
    }
#pragma unroll
    for (int m = 1; m < LPP; m <<= 1) dl += __shfl_xor_sync(0xffffffffu, dl, m);
    const long long i = base + k * PPI + sub;
    if (ch == 0 && i < m_heads) delta[i] = dl;
  }
}
template <int HDV>
static void launch_delta(const void* o, const void* dout, float* delta, long long m_heads, cudaStream_t s) {
  constexpr int PAIRS_PER_BLOCK = 8 * (HDV == 64 ? 4 : 2) * 8;  // 8 warps x 8 loads x pairs per load instruction
  attn_delta_kernel<HDV><<<(unsigned)((m_heads + PAIRS_PER_BLOCK - 1) / PAIRS_PER_BLOCK), 256, 0, s>>>((const bf16*)o, (const bf16*)dout, delta, m_heads);
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

// head_dim 72 (DiT-XL): the dq + dkv kernel pair on two-panel operand tiles, 3-D tensor maps {72 channels, heads, rows}
// grid of the persistent dq / dkv kernels: one CTA per SM (or per item if there are fewer)
static int pair_grid(int tokens, int heads, int n_samples) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)((tokens + RT - 1) / RT) * heads * n_samples;
  return (int)(items < sms ? items : sms);
}

static int attn_bwd_pair72(const void* qkv, const void* o, const void* dout, const float* lse, const float* sc, float eps, void* dqkv,
                           float* delta, int n_samples, int tokens, int heads, void* stream) {
  MAPDIT_REQUIRE(o != nullptr, "cos_attn_bwd: o is required for head_dim 72");
  const int hd = 72, D = heads * hd;
  const uint64_t rows = (uint64_t)n_samples * tokens;
  auto enc = [&](CUtensorMap* m, const void* base, int n_heads, uint32_t box_rows) {
    const uint64_t dims[3] = {(uint64_t)hd, (uint64_t)n_heads, rows};
    const uint64_t strides[2] = {(uint64_t)hd * 2, (uint64_t)n_heads * hd * 2};
    const uint32_t box[3] = {64, 1, box_rows};
    return (int)mapdit_encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  CUtensorMap t_qkv_row, t_qkv_blk, t_do_row, t_do_blk, t_out;  // t_out: dqkv in 32-row slabs (per-warp TMA stores of the read-out)
  if (enc(&t_qkv_row, qkv, 3 * heads, RT) | enc(&t_qkv_blk, qkv, 3 * heads, CB) | enc(&t_do_row, dout, heads, RT) | enc(&t_do_blk, dout, heads, CB) |
      enc(&t_out, dqkv, 3 * heads, 32)) {
    mapdit_set_error("cos_attn_bwd(hd 72): cuTensorMapEncodeTiled failed");
    return MAPDIT_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_tc<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<72>::DQ_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_tc<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<72>::DKV_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      mapdit_set_error("cos_attn_bwd(hd 72): cudaFuncSetAttribute failed");
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = pair_grid(tokens, heads, n_samples);  // persistent: one CTA per SM over (row tile, head, sample) items
  launch_delta<72>(o, dout, delta, (long long)rows * heads, s);
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd(delta, hd 72)");
  attn_bwd_dq_tc<72><<<grid, BCfg<72>::NTHR, BCfg<72>::DQ_SMEM, s>>>(t_qkv_row, t_qkv_blk, t_do_row, t_out, lse, delta, tokens, heads,
                                                                   n_samples, (const bf16*)qkv, sc, eps, g_attn_dbg);
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd(dq, hd 72)");
  attn_bwd_dkv_tc<72><<<grid, BCfg<72>::NTHR, BCfg<72>::DKV_SMEM, s>>>(t_qkv_row, t_qkv_blk, t_do_blk, t_out, lse, delta, tokens, heads,
                                                                     n_samples, (const bf16*)qkv, sc, eps, g_attn_dbg);
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd(dkv, hd 72)");
  return MAPDIT_OK;
}

// sc == nullptr: gradients w.r.t. the normalised q^, k^ (and v).  sc != nullptr: the q/k normalisation backward is applied
// as well (fused into the dq / dk epilogues on the tcgen05 path, a separate kernel behind the CUDA-core path).
static int attn_bwd_impl(const void* qkv, const void* o, const void* dout, const float* lse, const float* sc, float eps, void* dqkv,
                         float* delta, int n_samples, int tokens, int heads, int head_dim, int dtype, void* stream) {
  MAPDIT_REQUIRE(qkv && dout && lse && dqkv && delta && n_samples > 0 && tokens > 0, "cos_attn_bwd: bad args");
  // o == NULL: `delta` already holds dO.O (the out-proj dgrad GEMM produced it, MAPDIT_EPI_STORE_DELTA): only the fused kernel
  // takes that shortcut, every other path computes delta itself from o
  const bool fused_path = dtype == MAPDIT_BF16 && head_dim == HD && tokens == F_T && g_mapdit_attn_bwd_fused && !(mapdit_variant() & MAPDIT_VAR_DOT_ATTN);
  MAPDIT_REQUIRE(o || fused_path, "cos_attn_bwd: o may only be omitted (delta precomputed) on the fused tokens == 256 path");
  if (dtype == MAPDIT_BF16 && head_dim == 72 && tokens % CB == 0 && !(mapdit_variant() & MAPDIT_VAR_DOT_ATTN))
    return attn_bwd_pair72(qkv, o, dout, lse, sc, eps, dqkv, delta, n_samples, tokens, heads, stream);  // DiT-XL
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
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<64>::DQ_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<64>::DKV_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      mapdit_set_error("cos_attn_bwd: cudaFuncSetAttribute failed");
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (tokens == F_T && g_mapdit_attn_bwd_fused) {
    static bool fattr = false;
    if (!fattr) {
      if (cudaFuncSetAttribute(attn_bwd_fused_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess) {
        mapdit_set_error("cos_attn_bwd(fused): cudaFuncSetAttribute failed");
        return MAPDIT_ERR_CUDA;
      }
      fattr = true;
    }
    if (o) {
      const long long mh = (long long)rows * heads;
      launch_delta<64>(o, dout, delta, mh, s);
      MAPDIT_LAUNCH_CHECK("cos_attn_bwd(delta)");
    }
    const int items = n_samples * heads;
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    CUtensorMap t_out;  // per-warp stores of 32-row slabs of a [128 x 64] smem tile
    if (encode2d(&t_out, dqkv, 3 * D, rows, 3 * D, 32) != 0) {
      mapdit_set_error("cos_attn_bwd(fused): cuTensorMapEncodeTiled failed");
      return MAPDIT_ERR_CUDA;
    }
    if (g_mapdit_attn_bwd_fused >= 3) {
      static bool f3attr = false;
      if (!f3attr) {
        if (cudaFuncSetAttribute(attn_bwd_fused3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess) {
          mapdit_set_error("cos_attn_bwd(fused3): cudaFuncSetAttribute failed");
          return MAPDIT_ERR_CUDA;
        }
        f3attr = true;
      }
      attn_bwd_fused3_tc<<<items < sms ? items : sms, F2_NTHREADS, F_SMEM, s>>>(t_qkv_row, t_do_row, t_out, lse, delta, heads, items, sc, eps,
                                                                              g_attn_dbg);
      MAPDIT_LAUNCH_CHECK("cos_attn_bwd(fused3)");
      return MAPDIT_OK;
    }
    if (g_mapdit_attn_bwd_fused >= 2) {
      static bool f2attr = false;
      if (!f2attr) {
        if (cudaFuncSetAttribute(attn_bwd_fused2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess) {
          mapdit_set_error("cos_attn_bwd(fused2): cudaFuncSetAttribute failed");
          return MAPDIT_ERR_CUDA;
        }
        f2attr = true;
      }
      attn_bwd_fused2_tc<<<items < sms ? items : sms, F2_NTHREADS, F_SMEM, s>>>(t_qkv_row, t_do_row, t_out, lse, delta, heads, items, sc, eps,
                                                                              g_attn_dbg);
      MAPDIT_LAUNCH_CHECK("cos_attn_bwd(fused2)");
      return MAPDIT_OK;
    }
    attn_bwd_fused_tc<<<items < sms ? items : sms, F_NTHREADS, F_SMEM, s>>>(t_qkv_row, t_do_row, t_out, lse, delta, heads, items,
                                                                           (const bf16*)qkv, sc, eps, g_attn_dbg);
    MAPDIT_LAUNCH_CHECK("cos_attn_bwd(fused)");
    return MAPDIT_OK;
  }
  const int grid = pair_grid(tokens, heads, n_samples);
  launch_delta<64>(o, dout, delta, (long long)rows * heads, s);
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd(delta)");
  CUtensorMap t_slab;  // dqkv in 32-row slabs (per-warp TMA stores of the read-out)
  if (encode2d(&t_slab, dqkv, 3 * D, rows, 3 * D, 32) != 0) {
    mapdit_set_error("cos_attn_bwd: cuTensorMapEncodeTiled failed");
    return MAPDIT_ERR_CUDA;
  }
  attn_bwd_dq_tc<64><<<grid, BCfg<64>::NTHR, BCfg<64>::DQ_SMEM, s>>>(t_qkv_row, t_qkv_blk, t_do_row, t_slab, lse, delta, tokens, heads,
                                                                   n_samples, (const bf16*)qkv, sc, eps, g_attn_dbg);
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd(dq)");
  attn_bwd_dkv_tc<64><<<grid, BCfg<64>::NTHR, BCfg<64>::DKV_SMEM, s>>>(t_qkv_row, t_qkv_blk, t_do_blk, t_slab, lse, delta, tokens, heads,
                                                                     n_samples, (const bf16*)qkv, sc, eps, g_attn_dbg);
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

// K4, second generation: cosine attention forward with one CTA per SM and two query tiles in flight
// (head_dim 64, tokens a multiple of 256).  Replaces src/layers/attention.py:43-49 like attention_tc.cu, which stays for
// token counts that are only a multiple of 64.
//
// What ncu said about the first kernel (profiles/r1_attn_ncu_summary.md): XU 44 %, tensor 20 %, no saturated pipe — the
// softmax warps spent their time on the S-ready wait, the P store to shared memory + fence.proxy.async, the row-sum
// exchange and the per-tile epilogue, with the two CTAs of an SM drifting into lock-step.  This kernel removes those:
//   * one work item = (sample, head, 256 queries): two 128-row query tiles A and B share every 64-key K/V block (half
//     the K/V traffic); each tile has a double-buffered S in TMEM, so S(g+1) is computed while softmax works on S(g)
//     (with single buffers ncu showed the softmax warps 38 % of their samples waiting for the next S);
//   * P never touches shared memory: it is written back over its own S columns in TMEM (tcgen05.st, bf16 pairs) and is
//     the A operand of the PV MMA straight from TMEM;
//   * the row sum comes out of the tensor core: V is extended by a constant panel of ones (N = 80), so column 64 of the
//     O accumulator is sum_k P[r,k] of exactly the bf16 P the numerator used — no FADDs, no cross-warp exchange;
//   * a third warpgroup normalises and stores O, so the softmax warpgroups go straight on to the next item.
// Warps: 0-3 softmax A, 4-7 softmax B, 8-11 epilogue (TMEM lane quarter = warp % 4), 12 TMA producer, 13 / 14 MMA issuers of
// tile A / B.  One issuer per tile because a clock64 timeline of the kernel showed the issue of these small MMAs
// (N = 64 / 80: 32-40 tensor-pipe clocks each) costing ~80-100 clocks of the issuing thread apiece: a single issuer serving
// both tiles was the bottleneck (softmax warps 38 % of their samples waiting for S, tensor pipe 22 % active).
// (Sixteen softmax warps, two per lane quarter splitting the key columns behind a pair barrier, measured 11 % slower.)
// TMEM (512 columns): S_A/P_A buffers [0,64) [64,128)  S_B/P_B [128,192) [192,256)  O_A [256,336)  O_B [352,432).
// Logits are bounded (|q.k|/8 <= 8): p = exp(logit - 8), no running max, no rescaling of O.
#include "tc_common.cuh"

namespace {
using namespace tc;

#ifndef ATTN2_KB
#define ATTN2_KB 128
#endif
constexpr int QT = 128, KB = ATTN2_KB;  // keys per step: 128 with one S buffer per tile, or 64 with two
constexpr int SBUF = 128 / KB;          // S buffers per tile (each tile owns 128 TMEM columns)
// Head dimension 64: one 64-channel SWIZZLE_128B panel per operand tile.  Head dimension 72 (DiT-XL): TWO panels per tile — the
// tensor maps are 3-D {72 channels, 3H heads, rows}, so the box that starts at channel 64 reads channels 64..71 and the TMA
// unit zero-fills the 56 channels past the head's end: in shared memory every head is 128 channels wide with zeros behind
// channel 72, without a padded copy in HBM.  S = Q K^T then contracts over 80 channels (five K = 16 steps, the last one over the
// second panel's first 16 channels), PV produces 80 output columns (64 + 16 of the second V panel through the descriptor's
// leading-dimension stride, which the 64-channel kernel points at its constant ones panel instead).
template <int HDV>
struct ACfg {
  static constexpr int PANELS = HDV == 64 ? 1 : 2;
  static constexpr int KSTEPS = (HDV + 15) / 16;                // 4 | 5
  static constexpr int TILE_Q = QT * 64 * 2 * PANELS;           // one 128-query tile: 16 | 32 KB
  static constexpr int Q_BYTES = 2 * TILE_Q;                    // both query tiles of an item
  static constexpr int K_BYTES = KB * 64 * 2 * PANELS;          // 16 | 32 KB at 128 keys
  static constexpr int KV_BYTES = 2 * K_BYTES;                  // K block then V block
  static constexpr int ONES_BYTES = HDV == 64 ? KB * 128 : 0;   // [keys x 128 B] of bf16 1.0: the row-sum panel of the 64-channel kernel
  static constexpr int QBUF = HDV == 64 ? 2 : 1;                // query buffers (the two-panel tiles leave room for one)
  static constexpr int NS = HDV == 64 ? (KB == 64 ? 4 : 3) : 2;  // K/V stages
  static constexpr int RS_BYTES = HDV == 64 ? 0 : 2 * 2 * QT * 4;  // row sums [item parity][tile][row] (no ones panel at 72)
  // output staging: a finished O row is normalised in registers, written into a SWIZZLE_128B tile and leaves through per-warp TMA
  // stores of 32-row slabs (row-per-thread 16-byte global stores are 32 line transactions per instruction; they cost the kernel 7 %,
  // mostly by slowing the softmax warps' shared-memory traffic while an item's O was written).  64 channels: one tile per query
  // tile; 72: one tile + a dense [128 x 8] tail (channels 64..71) shared by both query tiles.
  static constexpr int OUT_TILE = QT * 128;
  static constexpr int OUT_BYTES = HDV == 64 ? 2 * OUT_TILE : OUT_TILE + QT * 16;
  static constexpr int SMEM_BYTES = QBUF * Q_BYTES + NS * KV_BYTES + OUT_BYTES + ONES_BYTES + RS_BYTES + 1024 + 512;
};
constexpr int NTHREADS = 15 * 32;
constexpr int W_EPI = 8, W_TMA = 12, W_MMA = 13;  // warps 13 and 14 issue the MMAs of tile A and tile B
constexpr uint32_t TMEM_COLS = 512;
__host__ __device__ constexpr uint32_t col_s(int x, int b) { return (uint32_t)(x * 128 + b * KB); }
__host__ __device__ constexpr uint32_t col_o(int x) { return x ? 352u : 256u; }
constexpr int ON = 80;  // PV accumulator width: 64 channels + 16 row-sum columns

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_o_tail, float* __restrict__ lse, int tokens,
                int heads, int n_samples, long long* __restrict__ dbg) {
  using A = ACfg<HD>;
  constexpr int Q_BYTES = A::Q_BYTES, K_BYTES = A::K_BYTES, KV_BYTES = A::KV_BYTES, ONES_BYTES = A::ONES_BYTES, NS = A::NS, QBUF = A::QBUF;
  constexpr int PANEL_Q = QT * 128, PANEL_K = KB * 128;  // bytes of one 64-channel panel of a query tile / a key block
  // optional timeline of CTA 0 (tools/attn_timeline.py): dbg[role*256 + 4*g + e] = clock64 at event e of step g
#define DBG(role, g, e)                                                                             \
  do {                                                                                               \
    if (dbg && blockIdx.x == 0 && (g) < 64 && lane == 0) dbg[(role) * 256 + 4 * (g) + (e)] = clock64(); \
  } while (0)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // [QBUF buffers][tile A | tile B], a tile = PANELS x [128 x 64]
  uint8_t* sKV = sQ + QBUF * Q_BYTES;       // stage s: K at sKV + s*KV_BYTES, V right after
  uint8_t* sOut = sKV + NS * KV_BYTES;      // output staging tile(s), then (head_dim 72) the tail
  uint8_t* sOnes = sOut + A::OUT_BYTES;
  float* sRS = reinterpret_cast<float*>(sOnes + ONES_BYTES);  // head_dim 72 only
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ONES_BYTES + A::RS_BYTES);
  uint64_t* q_full = bars;              // [2]
  uint64_t* q_empty = q_full + 2;       // [2]
  uint64_t* kv_full = q_empty + 2;      // [NS]
  uint64_t* kv_empty = kv_full + NS;    // [NS]
  uint64_t* s_full = kv_empty + NS;     // [tile][buffer]
  uint64_t* p_full = s_full + 4;        // [tile][buffer]
  uint64_t* o_full = p_full + 4;        // [2]
  uint64_t* o_empty = o_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const float sqrt_hd = HD == 64 ? 8.0f : 8.48528137423857f;  // |logit| <= sqrt(head_dim) for L2-normalised q, k
  const int nkb = tokens / KB;
  const int npair = tokens / (2 * QT);
  const int total_items = npair * heads * n_samples;

  // constant ones panel (generic-proxy writes, made visible to the tensor core's async proxy below)
  for (int i = threadIdx.x; i < ONES_BYTES / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async();
  if (warp == W_TMA && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_kv);
    prefetch_tmap(&tm_o);
    if constexpr (HD != 64) prefetch_tmap(&tm_o_tail);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);
      mbar_init(&s_full[2 * i], 1);
      mbar_init(&s_full[2 * i + 1], 1);
      mbar_init(&p_full[2 * i], 4);
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);  // one commit per MMA issuer
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == W_TMA) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int pr = item % npair, rest = item / npair, h = rest % heads, n = rest / heads;
        const int row_base = n * tokens;
        const int qb = it % QBUF;
        mbar_wait(&q_empty[qb], ((it / QBUF) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], Q_BYTES);
        if constexpr (HD == 64) {
          tma_load_2d(sQ + qb * Q_BYTES, &tm_q, &q_full[qb], h * HD, row_base + pr * 2 * QT);
        } else {  // per query tile: panel 0 (channels 0..63), panel 1 (channels 64..71 + zero fill); box = 128 rows
#pragma unroll
          for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int p = 0; p < 2; ++p)
              tma_load_3d(sQ + qb * Q_BYTES + x * A::TILE_Q + p * PANEL_Q, &tm_q, &q_full[qb], 64 * p, h, row_base + pr * 2 * QT + x * QT);
        }
        for (int j = 0; j < nkb; ++j, ++g) {
          const int s = g % NS;
          mbar_wait(&kv_empty[s], ((g / NS) & 1) ^ 1);
          uint8_t* dst = sKV + s * KV_BYTES;
          mbar_arrive_expect_tx(&kv_full[s], KV_BYTES);
          if constexpr (HD == 64) {
            tma_load_2d(dst, &tm_kv, &kv_full[s], D + h * HD, row_base + j * KB);
            tma_load_2d(dst + K_BYTES, &tm_kv, &kv_full[s], 2 * D + h * HD, row_base + j * KB);
          } else {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              tma_load_3d(dst + p * PANEL_K, &tm_kv, &kv_full[s], 64 * p, heads + h, row_base + j * KB);
              tma_load_3d(dst + K_BYTES + p * PANEL_K, &tm_kv, &kv_full[s], 64 * p, 2 * heads + h, row_base + j * KB);
            }
          }
        }
      }
    }
  } else if (warp >= W_MMA) {
    {
      const int x = warp - W_MMA;  // this issuer's query tile
      // ------------------------------------------------ MMA issuer: a flat software pipeline over (item, key block) steps.
      // The WHOLE warp runs this loop (waits included) and one elected lane issues the tcgen05 instructions: inside an
      // `if (lane == 0)` region the compiler cannot prove the descriptors warp-uniform and wraps every UTCHMMA operand in an
      // ELECT / R2UR.BROADCAST waterfall loop, which for these small MMAs (32-40 tensor clocks each) cost more than the MMA.
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = make_idesc_bf16(QT, KB, 0, 0);  // S = Q K^T, both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, ON, 0, 1);  // O = P V, P from TMEM, V MN-major (+ ones panel)
      const int my_items = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const uint32_t total_steps = (uint32_t)my_items * nkb;
      // second N panel of the PV B operand: the constant ones panel (64 channels: column 64 = row sum) or V's own second panel (72)
      const uint32_t v_lbo = HD == 64 ? 0u : (uint32_t)PANEL_K;
      const uint32_t ones_addr = smem_u32(sOnes);
      // descriptors of Q buffer 0 (this issuer's tile) and of ring stage 0; every MMA operand is one of these moved by a byte offset
      // (desc_advance: one add).  Rebuilding the descriptors per MMA made the issue loop twice as long as the MMAs it issues
      // (tools/attn_timeline.py: 1.3-1.5 k cycles for 12 MMAs worth 580 tensor-core cycles).
      const uint64_t d_q0 = make_smem_desc(smem_u32(sQ + x * A::TILE_Q), 16, 1024);
      const uint64_t d_k0 = make_smem_desc(smem_u32(sKV), 16, 1024);
      const uint64_t d_v0 = make_smem_desc(smem_u32(sKV + K_BYTES), HD == 64 ? ones_addr - smem_u32(sKV + K_BYTES) : v_lbo, 1024);
      auto issue_s = [&](uint32_t g) {  // S_x(g) = Q_x K_g^T into buffer g & 1, then signal tile x's softmax warps
        const uint32_t it = g / nkb, b = g % SBUF;
        if (g - it * nkb == 0) mbar_wait(&q_full[it % QBUF], (it / QBUF) & 1);
        mbar_wait(&kv_full[g % NS], (g / NS) & 1);
        tc_fence_after();
        const uint64_t d_q = desc_advance(d_q0, (it % QBUF) * Q_BYTES), d_k = desc_advance(d_k0, (g % NS) * KV_BYTES);
#pragma unroll
        for (int k = 0; k < A::KSTEPS; ++k)  // 16 channels per MMA; step 4 (head_dim 72) is the first 16 channels of the second panel
          if (leader)
            umma_ss(tmem_base + col_s(x, b), desc_advance(d_q, (k >> 2) * PANEL_Q + (k & 3) * 32),
                    desc_advance(d_k, (k >> 2) * PANEL_K + (k & 3) * 32), idesc_s, k != 0);
        if (leader) umma_commit(&s_full[2 * x + b]);
        // the query tiles of an item are free once S_B of its last key block has been issued
        if (leader && g - it * nkb + 1 == (uint32_t)nkb) umma_commit(&q_empty[it % QBUF]);
      };
      for (uint32_t g = 0; g < (uint32_t)SBUF && g < total_steps; ++g) issue_s(g);
      for (uint32_t g = 0; g < total_steps; ++g) {
        const uint32_t it = g / nkb, j = g - it * nkb, b = g % SBUF;
        const bool last_j = (j + 1 == (uint32_t)nkb);
        // head_dim 64: the second N panel is the fixed ones panel, so the leading-dimension offset shrinks as the stage address grows
        const uint32_t v_off = (g % NS) * KV_BYTES;
        const uint64_t d_v = desc_advance(d_v0, v_off) - (HD == 64 ? (uint64_t)(v_off >> 4) << 16 : 0ull);
        {
          if (j == 0) mbar_wait(&o_empty[x], (it & 1) ^ 1);  // the epilogue has read the previous item's O_x
          mbar_wait(&p_full[2 * x + b], (g / SBUF) & 1);
          tc_fence_after();
          DBG(x, g, 0);  // MMA warp: P_x(g) seen
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)  // 16 keys per MMA: P advances 8 TMEM columns, V two 8-row groups
            if (leader)
              umma_ts(tmem_base + col_o(x), tmem_base + col_s(x, b) + k * 8, desc_advance(d_v, k * 2048), idesc_o, (j | k) != 0);
          if (leader && last_j) umma_commit(&o_full[x]);
          if (leader) umma_commit(&kv_empty[g % NS]);  // V_g is done with (K_g since S_x(g), two steps ago)
          if (g + SBUF < total_steps) issue_s(g + SBUF);  // reuses buffer b right behind the PV that read P from it
          DBG(x, g, 1);  // MMA warp: PV_x(g) and S_x(g+2) issued
        }
      }
    }
  } else if (warp < W_EPI) {
    // ------------------------------------------------ softmax warpgroups: thread = query row, 64 logits per step
    const int x = warp >> 2, qq = warp & 3;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
    const float c1 = (1.0f / sqrt_hd) * 1.4426950408889634f, c2 = sqrt_hd * 1.4426950408889634f;
    float rsum = 0.f;  // head_dim 72: this row's sum of exponentials over the item's key blocks
    const int my_items = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total_steps = (uint32_t)my_items * nkb;
    // explicit ping-pong: the two warpgroups take turns in the exponential phase (named barriers 1 + x: "tile x may go"),
    // so they never share the MUFU and one tile's exp phase covers the other's MMA phase; left alone they fall into
    // lock-step (timeline: both 2x slower in the exp phase, then both idle while the tensor core works)
    if (x == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");  // tile A goes first
    for (uint32_t g = 0; g < total_steps; ++g) {
      const uint32_t b = g % SBUF, t_s = t_lane + col_s(x, b);
      if (qq == 0) DBG(2 + x, g, 0);  // softmax: starts waiting for S_x(g)
      mbar_wait(&s_full[2 * x + b], (g / SBUF) & 1);
      tc_fence_after();
      if (x == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
      if (qq == 0) DBG(2 + x, g, 1);  // softmax: S_x(g) seen and our turn
      // 32-column chunks, software pipelined: the tcgen05.ld of chunk c+1 is in flight while chunk c is exponentiated.
      // P (bf16 pairs) overwrites the first half of its own S buffer: chunk c lands on columns [16c, 16c+16), all already in
      // registers, while the load in flight reads columns >= 32(c+1).
      uint32_t sa[32], sb[32];
      tmem_ld32(t_s, sa);
#pragma unroll
      for (int c = 0; c < KB / 32; ++c) {
        uint32_t(&cur)[32] = (c & 1) ? sb : sa;
        uint32_t(&nxt)[32] = (c & 1) ? sa : sb;
        tmem_ld_wait();
        if (c + 1 < KB / 32) tmem_ld32(t_s + (c + 1) * 32, nxt);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pa = ex2f(fmaf(__uint_as_float(cur[2 * i]), c1, -c2)), pb = ex2f(fmaf(__uint_as_float(cur[2 * i + 1]), c1, -c2));
          if constexpr (HD != 64) rsum += pa + pb;
          pk[i] = pack_bf16(pa, pb);
        }
        tmem_st16(t_s + c * 16, pk);
      }
      if constexpr (HD != 64) {  // last key block of the item: hand the row sum to the epilogue warp of this lane quarter
        const uint32_t it = g / nkb;
        if (g - it * nkb + 1 == (uint32_t)nkb) {
          sRS[((it & 1) * 2 + x) * QT + qq * 32 + lane] = rsum;
          rsum = 0.f;
        }
      }
      if (x == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");  // the exponentials are issued: the other tile's turn
      else asm volatile("bar.arrive 1, 256;" ::: "memory");
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * x + b]);
      if (qq == 0) DBG(2 + x, g, 2);  // softmax: P_x(g) published
    }
  } else {
    // ------------------------------------------------ epilogue warpgroup: O / rowsum -> global, log-sum-exp
    const int qq = warp & 3;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
    uint32_t it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int pr = item % npair, rest = item / npair, h = rest % heads, n = rest / heads;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const size_t grow = (size_t)n * tokens + pr * 2 * QT + x * QT + qq * 32 + lane;
        mbar_wait(&o_full[x], it & 1);
        tc_fence_after();
        uint32_t ov[32], rs;
        tmem_ld32(t_lane + col_o(x), ov);
        if constexpr (HD == 64) tmem_ld1(t_lane + col_o(x) + 64, rs);
        tmem_ld_wait();
        float total;
        if constexpr (HD == 64) total = __uint_as_float(rs);
        else total = sRS[((it & 1) * 2 + x) * QT + qq * 32 + lane];  // written before the last P was published (ordered by the barriers)
        const float inv = 1.0f / total;
        if (lse) lse[grow * heads + h] = sqrt_hd + logf(total);
        // this warp's 32-row slab of the staging tile: the store that last read it has to be through with it (64 channels: the
        // previous item's tile x, one store group back; 72: the other query tile, the last group)
        uint8_t* stile = sOut + (HD == 64 ? x * A::OUT_TILE : 0);
        if (lane == 0) {
          if constexpr (HD == 64) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        const int r = qq * 32 + lane;
        uint8_t* prow = stile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t tail[8];
          if (half == 1) {
            tmem_ld32(t_lane + col_o(x) + 32, ov);
            if constexpr (HD != 64)
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(tail[0]), "=r"(tail[1]), "=r"(tail[2]), "=r"(tail[3]), "=r"(tail[4]), "=r"(tail[5]), "=r"(tail[6]), "=r"(tail[7])
                           : "r"(t_lane + col_o(x) + 64)
                           : "memory");
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[x]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(ov[8 * c]) * inv, __uint_as_float(ov[8 * c + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(ov[8 * c + 2]) * inv, __uint_as_float(ov[8 * c + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(ov[8 * c + 4]) * inv, __uint_as_float(ov[8 * c + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(ov[8 * c + 6]) * inv, __uint_as_float(ov[8 * c + 7]) * inv);
            *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (r & 7)) << 4)) = u;
          }
          if constexpr (HD != 64) {
            if (half == 1) {  // channels 64..71
              uint4 u;
              u.x = pack_bf16(__uint_as_float(tail[0]) * inv, __uint_as_float(tail[1]) * inv);
              u.y = pack_bf16(__uint_as_float(tail[2]) * inv, __uint_as_float(tail[3]) * inv);
              u.z = pack_bf16(__uint_as_float(tail[4]) * inv, __uint_as_float(tail[5]) * inv);
              u.w = pack_bf16(__uint_as_float(tail[6]) * inv, __uint_as_float(tail[7]) * inv);
              *reinterpret_cast<uint4*>(sOut + A::OUT_TILE + r * 16) = u;
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const int grow0 = n * tokens + pr * 2 * QT + x * QT + qq * 32;
          if constexpr (HD == 64) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm_o),
                         "r"(smem_u32(stile + qq * 4096)), "r"(h * HD), "r"(grow0)
                         : "memory");
          } else {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&tm_o),
                         "r"(smem_u32(stile + qq * 4096)), "r"(0), "r"(h), "r"(grow0)
                         : "memory");
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&tm_o_tail),
                         "r"(smem_u32(sOut + A::OUT_TILE + qq * 512)), "r"(64), "r"(h), "r"(grow0)
                         : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores complete before the CTA exits
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<TMEM_COLS>(tmem_base);
}
}  // namespace

bool mapdit_attn_tc2_supported(int tokens, int hd) { return (hd == 64 || hd == 72) && tokens % (2 * QT) == 0 && tokens >= 2 * QT; }

long long* g_attn_dbg = nullptr;  // also stamped by attn_bwd_fused_tc (attention_bwd_tc.cu)
extern "C" int mapdit_attn_debug_buffer(void* p) {  // developer hook: timeline buffer of >= 1024 int64 (or null)
  g_attn_dbg = (long long*)p;
  return MAPDIT_OK;
}

namespace {
template <int HDV>
int launch_attn_tc2(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& to, const CUtensorMap& to_tail, float* lse, int n,
                    int tokens, int heads, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc2_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<HDV>::SMEM_BYTES);
    if (e != cudaSuccess) {
      mapdit_set_error("attn_tc2_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int items = (tokens / (2 * QT)) * heads * n;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = items < sms ? items : sms;  // persistent, one CTA per SM (512 TMEM columns)
  attn_tc2_kernel<HDV><<<grid, NTHREADS, ACfg<HDV>::SMEM_BYTES, (cudaStream_t)stream>>>(tq, tkv, to, to_tail, lse, tokens, heads, n, g_attn_dbg);
  return MAPDIT_OK;
}
}  // namespace

int mapdit_attn_tc2_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, void* stream) {
  const int D = heads * hd;
  CUtensorMap tq, tkv, to, to_tail;  // to / to_tail: the output through per-warp TMA stores of 32-row slabs
  CUresult r1, r2, r3, r4;
  if (hd == 64) {
    const uint64_t dims[2] = {(uint64_t)3 * D, (uint64_t)n * tokens};
    const uint64_t strides[1] = {(uint64_t)3 * D * 2};
    const uint32_t box_q[2] = {64, 2 * QT}, box_kv[2] = {64, KB};
    r1 = mapdit_encode_tmap(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
    r2 = mapdit_encode_tmap(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
    const uint64_t odims[2] = {(uint64_t)D, (uint64_t)n * tokens};
    const uint64_t ostrides[1] = {(uint64_t)D * 2};
    const uint32_t box_o[2] = {64, 32};
    r3 = mapdit_encode_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, o, odims, ostrides, box_o, CU_TENSOR_MAP_SWIZZLE_128B);
    r4 = r3;
    to_tail = to;
  } else {
    // {channel within head, head of q|k|v, row}: a 64-channel box at channel 64 runs past the head's 72 channels and is zero-filled there
    const uint64_t dims[3] = {(uint64_t)hd, (uint64_t)3 * heads, (uint64_t)n * tokens};
    const uint64_t strides[2] = {(uint64_t)hd * 2, (uint64_t)3 * D * 2};
    const uint32_t box_q[3] = {64, 1, QT}, box_kv[3] = {64, 1, KB};
    r1 = mapdit_encode_tmap(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
    r2 = mapdit_encode_tmap(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
    // output {channel within head, head, row}: the 64-channel box of a 72-channel head is clipped by the store, the tail box adds 64..71
    const uint64_t odims[3] = {(uint64_t)hd, (uint64_t)heads, (uint64_t)n * tokens};
    const uint64_t ostrides[2] = {(uint64_t)hd * 2, (uint64_t)D * 2};
    const uint32_t box_o[3] = {64, 1, 32}, box_t[3] = {8, 1, 32};
    r3 = mapdit_encode_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, o, odims, ostrides, box_o, CU_TENSOR_MAP_SWIZZLE_128B);
    r4 = mapdit_encode_tmap(&to_tail, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, o, odims, ostrides, box_t, CU_TENSOR_MAP_SWIZZLE_NONE);
  }
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS || r3 != CUDA_SUCCESS || r4 != CUDA_SUCCESS) {
    mapdit_set_error("attn_tc2_fwd: cuTensorMapEncodeTiled failed (%d, %d, %d, %d)", (int)r1, (int)r2, (int)r3, (int)r4);
    return MAPDIT_ERR_CUDA;
  }
  const int rc = hd == 64 ? launch_attn_tc2<64>(tq, tkv, to, to_tail, lse, n, tokens, heads, stream)
                           : launch_attn_tc2<72>(tq, tkv, to, to_tail, lse, n, tokens, heads, stream);
  if (rc != MAPDIT_OK) return rc;
  MAPDIT_LAUNCH_CHECK("attn_tc2_fwd");
  return MAPDIT_OK;
}

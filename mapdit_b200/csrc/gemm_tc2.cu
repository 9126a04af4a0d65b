// 2-CTA variant of the bf16 tcgen05 GEMM: a CTA pair (cluster of 2, one TPC) computes a 256 x BN tile with
// tcgen05.mma.cta_group::2.  Each CTA loads its own 128 rows of A and only HALF of the B tile; the tensor core
// reads both halves across the pair, so every SM ingests (128 + BN/2) rows per k-block instead of (128 + BN):
// the 1-CTA kernel is operand-delivery bound (profiles/r1_gemm_tc_ncu_summary.md), this cuts that traffic by 1/3
// and frees shared memory for a 7-stage ring.  Same warp roles, barriers and fused epilogues as gemm_tc.cu; the
// leader CTA (cluster rank 0) issues the MMAs, commits are multicast to both CTAs' barriers, and both CTAs'
// epilogue warps release the accumulator stage on the leader's barrier.
#include "gemm_epilogue.cuh"

extern long long* g_attn_dbg;  // developer timeline hook (mapdit_attn_debug_buffer)
extern int g_mapdit_gemm_fused_resid;  // developer switch (mapdit_set_option "gemm_fused_resid"): 0 = first-generation epilogue

namespace {
using namespace tc;
using namespace gemm_epi;

constexpr int BM = 128, BK = 64, UK = 16;  // BM = rows per CTA (256 per pair)
constexpr int NUM_THREADS = 384;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-parity bit of a shared::cluster address -> leader CTA

// FUSED = the second-generation residual epilogues (run_tile_fused_resid): per-warp staging of 5 tile buffers + the per-sample
// vectors instead of the 2 + 2 buffers of the generic path; one operand stage less for BN = 256 (4 instead of 5)
template <int BN, int KIND>
struct Cfg2 {
  static constexpr bool FUSED = KIND == 1;  // KIND: 0 = generic epilogues, 1 = second-generation residual epilogues, 2 = generic + STORE_DELTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int BH_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;
  static constexpr int EPI_BYTES = FUSED ? fr_bytes(BN) : STG_BYTES + RSTG_BYTES;
  static constexpr int MAX_STAGES = (227 * 1024 - 2048 - EPI_BYTES) / STAGE_BYTES;  // generic: 5 for BN = 256 (6 measured no faster)
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 512;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs load into their own smem; the transaction bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

template <int BN, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ EpiTmaps etm, const EpiParams ep,
                int num_m_blocks, int num_n_blocks, int num_k_blocks, long long* __restrict__ dbg) {
  using C = Cfg2<BN, KIND>;
  constexpr bool FUSED = C::FUSED;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;
  // generic: output staging (8 x 4 KB) then residual tiles (ResidLoader, 8 x 4 KB); FUSED: 8 x 10 KB tile buffers then the vectors
  uint8_t* rstaging = staging + (FUSED ? 8 * FR_BUF_BYTES_PER_WARP : STG_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + C::EPI_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* rbar = tempty + 2;  // [8 warps][3] (generic path uses [8][2])
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_tiles = ((num_m_blocks + 1) >> 1) * num_n_blocks;  // 256-row tiles

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    prefetch_tmap(&etm.resid);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 16);  // 8 epilogue warps of each CTA of the pair
    }
    for (int i = 0; i < 24; ++i) mbar_init(&rbar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int m_pair = tile / num_n_blocks, n_blk = tile - m_pair * num_n_blocks;
      const int m_blk = 2 * m_pair + (int)rank;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::STAGE_BYTES;
        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
        tma_load_2d_2sm(sa, &tma_a, &full[stage], kb * BK, m_blk * BM);
        tma_load_2d_2sm(sa + C::A_BYTES, &tma_b, &full[stage], kb * BK, n_blk * BN + (int)rank * (BN / 2));
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    // ------------------------------------------------ MMA issuer (leader CTA only)
    constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0, acc = 0, acc_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const int tcount = (tile - cluster_id) / num_clusters;
      if (dbg && blockIdx.x == 0 && lane == 0 && tcount < 16) dbg[512 + tcount * 4 + 0] = clock64();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
        const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UK; ++k)
          if (leader) umma_ss_2cta(d_tmem, make_smem_desc(a_addr + k * UK * 2, 16, 1024), make_smem_desc(b_addr + k * UK * 2, 16, 1024), idesc,
                       (kb | k) != 0 ? 1u : 0u);
        if (leader) umma_commit_2cta(&empty[stage]);
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) umma_commit_2cta(&tfull[acc]);
      if (dbg && blockIdx.x == 0 && lane == 0 && tcount < 16) dbg[512 + tcount * 4 + 1] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int q = warp & 3, half = (warp - 4) >> 2;
    uint32_t acc = 0, acc_phase = 0;
    float gsc = 0.f, inv_den = 1.f;
    if (ep.epilogue == MAPDIT_EPI_RESID_MOD) {
      gsc = *ep.gain;
      inv_den = 1.0f / mod_den(gsc);
    }
    Stager st{staging + (FUSED ? 0 : (warp - 4) * STG_BYTES_PER_WARP), 0u, lane, 0};
    ResidLoader rl{rstaging + (FUSED ? 0 : (warp - 4) * RSTG_BYTES_PER_WARP), rbar + 2 * (warp - 4), &etm.resid, 0u, 0u, lane, 0};
    FusedResid fr;
    if constexpr (FUSED) {
      fr.bufs = staging + (warp - 4) * FR_BUF_BYTES_PER_WARP;
      fr.vec = reinterpret_cast<float*>(rstaging + (warp - 4) * fr_vec_bytes_per_warp(BN));
      fr.bars = rbar + FR_XBUFS * (warp - 4);
      fr.consumed = 0u;
      fr.lane = lane;
      fr.first_tile = cluster_id;
      fr.tile_stride = num_clusters;
      fr.total_tiles = total_tiles;
      fr.num_n_blocks = num_n_blocks;
      fr.row_in_pair = (int)rank * BM + q * 32;
      fr.half = half;
      fr.nv_tile = -1;
      fr.fine = nullptr;
      fr_issue<BN>(ep, etm, fr, 0u);  // the residual stream runs one chunk ahead of its consumer
    }
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int m_pair = tile / num_n_blocks, n_blk = tile - m_pair * num_n_blocks;
      const int row = (2 * m_pair + (int)rank) * BM + q * 32 + lane;
      st.row0 = (2 * m_pair + (int)rank) * BM + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const int tcount = (tile - cluster_id) / num_clusters;
      if (dbg && blockIdx.x == 0 && lane == 0 && tcount < 16) dbg[(warp - 4) * 64 + tcount * 4 + 0] = clock64();
      auto wait_acc = [&]() {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (dbg && blockIdx.x == 0 && lane == 0 && tcount < 16) dbg[(warp - 4) * 64 + tcount * 4 + 1] = clock64();
      };
      if constexpr (FUSED) {
        // fine-grained stamps of warp 4 (first column half) and warp 8 (second) for tiles 2 and 3 of CTA 0
        fr.fine = (dbg && blockIdx.x == 0 && (warp == 4 || warp == 8) && (tcount == 2 || tcount == 3))
                      ? dbg + 576 + ((warp == 8 ? 2 : 0) + (tcount - 2)) * 32 : nullptr;
        run_tile_fused_resid<BN>(ep, etm, fr, t_row, tile, gsc, inv_den, wait_acc, [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty[acc]);  // released as soon as the tile's last accumulator chunk is in registers
        });
        if (dbg && blockIdx.x == 0 && lane == 0 && tcount < 16) dbg[(warp - 4) * 64 + tcount * 4 + 2] = clock64();
      } else {
        // the STORE_DELTA branch lives in its own instantiation: compiled into the generic one it pushed the kernel past its 168
        // registers (104 bytes of spills)
        run_tile<BN, KIND == 2>(ep, etm, st, t_row, row, n_blk, half, gsc, inv_den, wait_acc, &rl);
        tc_fence_before();
        __syncwarp();
        if (dbg && blockIdx.x == 0 && lane == 0 && tcount < 16) dbg[(warp - 4) * 64 + tcount * 4 + 2] = clock64();
        if (lane == 0) mbar_arrive_leader(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (FUSED) fr.drain();
    else st.drain();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the peer may still read its smem / signal its barriers
  if (warp == 2) tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base);
}

template <int BN, int KIND>
int launch2(const mapdit_gemm_args* g, const EpiParams& ep, cudaStream_t stream, int num_sms) {
  using C = Cfg2<BN, KIND>;
  CUtensorMap ta, tb;
  const uint64_t dims_a[2] = {(uint64_t)g->k, (uint64_t)g->m}, dims_b[2] = {(uint64_t)g->k, (uint64_t)g->n};
  const uint64_t str_a[1] = {(uint64_t)g->lda * 2}, str_b[1] = {(uint64_t)g->ldb * 2};
  const uint32_t box_a[2] = {BK, BM}, box_b[2] = {BK, BN / 2};
  CUresult r1 = mapdit_encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->a, dims_a, str_a, box_a, CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r2 = mapdit_encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->b, dims_b, str_b, box_b, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    mapdit_set_error("gemm_bf16(2cta): cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
    return MAPDIT_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BN, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      mapdit_set_error("gemm_bf16(2cta): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int mb = (g->m + BM - 1) / BM, nb = (g->n + BN - 1) / BN, kb = (g->k + BK - 1) / BK;
  const int tiles = ((mb + 1) / 2) * nb;
  const int max_clusters = num_sms / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  EpiTmaps etm;
  if (make_store_maps(&etm, ep) != 0) {
    mapdit_set_error("gemm_bf16(2cta): cuTensorMapEncodeTiled (store maps) failed");
    return MAPDIT_ERR_CUDA;
  }
  gemm_tc2_kernel<BN, KIND><<<2 * clusters, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, etm, ep, mb, nb, kb, g_attn_dbg);
  return MAPDIT_OK;
}

// the second-generation residual epilogue needs a warp's 32 rows inside one sample and 16-byte aligned per-sample vectors
bool fused_resid_ok(const mapdit_gemm_args* g, const EpiParams& ep) {
  if (!g_mapdit_gemm_fused_resid) return false;
  // 1 (default) = only where the epilogue is the bottleneck: a short main loop per tile (K x tile width <= 1024 x 256: the out-proj
  // GEMMs, and fc2 of the D = 384 models on 128-wide tiles); with K = 4 D >= 3072 (fc2 of DiT-B and up) the main loop hides the
  // first-generation epilogue and keeps its fifth operand stage.  2 = always (A/B)
  const long long bn_tile = (g->n % 256 == 0 || g->n > 1024) ? 256 : (g->n % 192 == 0 ? 192 : 128);  // an upper bound is enough here
  if (g_mapdit_gemm_fused_resid == 1 && (long long)g->k * bn_tile > 1024 * 256) return false;
  if (ep.epilogue != MAPDIT_EPI_RESID && ep.epilogue != MAPDIT_EPI_RESID_MOD && ep.epilogue != MAPDIT_EPI_RESID_ROT) return false;
  if (ep.tokens % 32 != 0 || ep.ldmod % 4 != 0 || ep.ldshift % 4 != 0 || ep.N % 4 != 0) return false;
  auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
  return ep.resid != nullptr && al16(ep.gate) && al16(ep.shift) && al16(ep.scale);
}
template <int BN>
int launch2_any(const mapdit_gemm_args* g, const EpiParams& ep, cudaStream_t stream, int num_sms) {
  if (ep.epilogue == MAPDIT_EPI_STORE_DELTA) return launch2<BN, 2>(g, ep, stream, num_sms);
  return fused_resid_ok(g, ep) ? launch2<BN, 1>(g, ep, stream, num_sms) : launch2<BN, 0>(g, ep, stream, num_sms);
}
}  // namespace

// returns MAPDIT_ERR_UNSUPPORTED (without setting an error) when the shape should go to the 1-CTA kernel
int mapdit_gemm_bf16_2cta(const mapdit_gemm_args* g, const gemm_epi::EpiParams& ep, cudaStream_t stream, int num_sms) {
  const int mb = (g->m + BM - 1) / BM;
  // 256-wide tiles also when N is only a multiple of 128 and the clipped last tile wastes <= 12.5 % of the MMA work
  // (N = 1152, 3456: DiT-S qkv, DiT-XL): half the B-operand traffic per FLOP of the 128-wide kernel
  const int nb256 = (g->n + 255) / 256;
  const bool wide_ok = g->n % 256 == 0 || (g->n % 128 == 0 && (long long)nb256 * 256 * 8 <= (long long)g->n * 9);
  extern int g_mapdit_gemm_2cta_bn;  // developer switch (mapdit_set_option "gemm_2cta_bn"): 0 = auto, 128 / 256 = force that tile width
  if (g_mapdit_gemm_2cta_bn == 128 && g->n % 128 == 0) return launch2_any<128>(g, ep, stream, num_sms);
  // 192-wide tiles where they divide N and save whole rounds of the persistent grid (cost model: rounds x tile width) — but only
  // for the narrow models' shapes (N not a multiple of 256, or K <= 512).  Measured: DiT-S/2 training 16.80 -> 16.04 ms; on the
  // N = 768 GEMMs of DiT-B/2 (K >= 768) the extra re-reads of the A tile cost more than the 46 idle CTA pairs of the eleventh round
  // of a 256 x 256 tiling (out-proj 0.0952 -> 0.0973 ms, fc2 0.2439 -> 0.2550 ms), so those keep 256.
  const long long pairs = (mb + 1) / 2, cl = num_sms / 2;
  auto cost = [&](int bn) { return ((pairs * ((g->n + bn - 1) / bn) + cl - 1) / cl) * bn; };
  const bool ok192 = g->n % 192 == 0 && pairs * (g->n / 192) >= cl;
  if (g_mapdit_gemm_2cta_bn == 192 && ok192) return launch2_any<192>(g, ep, stream, num_sms);
  if (g_mapdit_gemm_2cta_bn == 0 && ok192 && (g->n % 256 != 0 || g->k <= 512) && (!wide_ok || cost(192) * 100 <= cost(256) * 97)) {
    const bool narrow_ok = g->n % 128 == 0 && pairs * (g->n / 128) >= num_sms;
    if (wide_ok || !narrow_ok || cost(192) * 100 <= cost(128) * 97) return launch2_any<192>(g, ep, stream, num_sms);
  }
  if (wide_ok && (long long)((mb + 1) / 2) * nb256 >= num_sms / 2) return launch2_any<256>(g, ep, stream, num_sms);
  if (g->n % 128 == 0 && (long long)((mb + 1) / 2) * (g->n / 128) >= num_sms) return launch2_any<128>(g, ep, stream, num_sms);
  return MAPDIT_ERR_UNSUPPORTED;
}

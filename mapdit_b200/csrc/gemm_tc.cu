// K2/K3: persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores with fused epilogues.
//
//   D[M,N] = A[M,K] · B[N,K]^T      A,B bf16 K-major in HBM, fp32 accumulation in TMEM
//
// One CTA per SM walks a static tile schedule (n fastest so the CTAs of one wave share the A panel in
// L2).  Warp 0 lane 0: TMA producer (cp.async.bulk.tensor, SWIZZLE_128B, STAGES-deep mbarrier ring);
// warp 1 lane 0: tcgen05.mma issuer (UMMA 128 x BN x 16, accumulators double-buffered in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1); warp 2: TMEM allocator; warps 4-7:
// epilogue (tcgen05.ld 32x32b -> registers -> fused elementwise math -> 16-byte global stores).
//
// Fused epilogues (reference code they replace):
//   QKNORM     q,k heads L2-normalised in registers        src/layers/attention.py:43-45
//   MPSILU     silu(acc)/0.596                              src/basic/mp_silu.py:7
//   RESID_MOD  x' = mp_sum(x, gate*acc, 0.3); h = modulate(x', shift, scale, g)
//                                                           src/blocks/dit_block.py:35-36, src/utils.py:11-16
#include <string.h>

#include "gemm_epilogue.cuh"

extern "C" int mapdit_qk_normalize(void* qkv, int m, int d, int head_dim, float eps, int dtype, void* stream);

namespace {
using namespace tc;
using namespace gemm_epi;

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int NUM_THREADS = 384;  // warps 0-3: TMA / MMA / TMEM alloc / spare, warps 4-11: epilogue
constexpr int SMEM_LIMIT = 227 * 1024;

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_STAGES = (SMEM_LIMIT - 2048 - STG_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES /*epilogue staging*/ + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ EpiTmaps etm, const EpiParams ep,
               int num_m_blocks, int num_n_blocks, int num_k_blocks) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;  // 8 x 4 KB, 1024-aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + STG_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = num_m_blocks * num_n_blocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
        tma_load_2d(sa, &tma_a, &full[stage], kb * BK, m_blk * BM);
        tma_load_2d(sa + C::A_BYTES, &tma_b, &full[stage], kb * BK, n_blk * BN);
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
        const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          const uint64_t adesc = make_smem_desc(a_addr + k * UK * 2, 16, 1024);
          const uint64_t bdesc = make_smem_desc(b_addr + k * UK * 2, 16, 1024);
          if (leader) umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (leader) umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (TMEM lane quarter = warp % 4)
    // 8 epilogue warps: TMEM lane quarter = warp % 4, the two warps of a quarter take alternate column chunks
    const int q = warp & 3, half = (warp - 4) >> 2;
    uint32_t acc = 0, acc_phase = 0;
    float gsc = 0.f, inv_den = 1.f;
    if (ep.epilogue == MAPDIT_EPI_RESID_MOD) {
      gsc = *ep.gain;
      inv_den = 1.0f / mod_den(gsc);
    }
    Stager st{staging + (warp - 4) * STG_BYTES_PER_WARP, 0u, lane, 0};
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
      const int row = m_blk * BM + q * 32 + lane;
      st.row0 = m_blk * BM + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      run_tile<BN>(ep, etm, st, t_row, row, n_blk, half, gsc, inv_den, [&]() {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
      });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    st.drain();  // all TMA stores of this warp have completed before the CTA may exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN>
int launch(const mapdit_gemm_args* g, const EpiParams& ep, cudaStream_t stream, int num_sms) {
  using C = Cfg<BN>;
  CUtensorMap ta, tb;
  const uint64_t dims_a[2] = {(uint64_t)g->k, (uint64_t)g->m}, dims_b[2] = {(uint64_t)g->k, (uint64_t)g->n};
  const uint64_t str_a[1] = {(uint64_t)g->lda * 2}, str_b[1] = {(uint64_t)g->ldb * 2};
  const uint32_t box_a[2] = {BK, BM}, box_b[2] = {BK, BN};
  CUresult r1 = mapdit_encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->a, dims_a, str_a, box_a, CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r2 = mapdit_encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->b, dims_b, str_b, box_b, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    mapdit_set_error("gemm_bf16: cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
    return MAPDIT_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      mapdit_set_error("gemm_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int mb = (g->m + BM - 1) / BM, nb = (g->n + BN - 1) / BN, kb = (g->k + BK - 1) / BK;
  const int tiles = mb * nb;
  const int grid = tiles < num_sms ? tiles : num_sms;
  EpiTmaps etm;
  if (make_store_maps(&etm, ep) != 0) {
    mapdit_set_error("gemm_bf16: cuTensorMapEncodeTiled (store maps) failed");
    return MAPDIT_ERR_CUDA;
  }
  gemm_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, etm, ep, mb, nb, kb);
  return MAPDIT_OK;
}

int num_sms_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
}  // namespace

CUresult mapdit_encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* base, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p)
      return CUDA_ERROR_NOT_FOUND;
    fn = (EncodeFn)p;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return fn(map, dt, rank, const_cast<void*>(base), (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes,
            (const cuuint32_t*)box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int mapdit_gemm_bf16_2cta(const mapdit_gemm_args* g, const gemm_epi::EpiParams& ep, cudaStream_t stream, int num_sms);
extern int g_mapdit_attn_v2;
extern int g_mapdit_attn_bwd_fused;
int g_mapdit_gemm_2cta_bn = 0;
int g_mapdit_gemm_fused_resid = 1;  // second-generation residual epilogues of the CTA-pair kernel (gemm_epilogue.cuh)
static int g_use_2cta = 1;  // CTA-pair kernel by default where the shape qualifies (A/B: bench.py --gemm-2cta 0)

// runtime switches (benchmark A/B): "gemm_2cta" = 0/1
extern "C" int mapdit_set_option(const char* name, int value) {
  if (name && !strcmp(name, "gemm_2cta")) {
    g_use_2cta = value;
    return MAPDIT_OK;
  }
  if (name && !strcmp(name, "attn_v2")) {
    g_mapdit_attn_v2 = value;
    return MAPDIT_OK;
  }
  if (name && !strcmp(name, "gemm_2cta_bn")) {
    g_mapdit_gemm_2cta_bn = value;
    return MAPDIT_OK;
  }
  if (name && !strcmp(name, "gemm_fused_resid")) {
    g_mapdit_gemm_fused_resid = value;
    return MAPDIT_OK;
  }
  if (name && !strcmp(name, "attn_bwd_fused")) {
    g_mapdit_attn_bwd_fused = value;
    return MAPDIT_OK;
  }
  mapdit_set_error("set_option: unknown option");
  return MAPDIT_ERR_ARG;
}

extern "C" int mapdit_gemm_bf16(const mapdit_gemm_args* g, void* stream) {
  MAPDIT_REQUIRE(g && g->a && g->b && g->out && g->m > 0 && g->n > 0 && g->k > 0, "gemm_bf16: bad args");
  MAPDIT_REQUIRE(g->k % 8 == 0 && g->lda % 8 == 0 && g->ldb % 8 == 0 && g->n % 8 == 0 && g->ldo % 8 == 0,
                 "gemm_bf16: k, n and leading dimensions must be multiples of 8 (16-byte TMA / vector alignment)");
  MAPDIT_REQUIRE(((uintptr_t)g->a & 15) == 0 && ((uintptr_t)g->b & 15) == 0 && ((uintptr_t)g->out & 15) == 0,
                 "gemm_bf16: pointers must be 16-byte aligned");
  int epi = g->epilogue;
  bool post_qknorm = false;
  if (epi == MAPDIT_EPI_QKNORM && (g->head_dim != 64 || g->n % 64 != 0)) {
    epi = MAPDIT_EPI_STORE;  // head_dim 72 (DiT-XL): plain store, then the row-wise normalise kernel
    post_qknorm = true;
  }
  if (epi == MAPDIT_EPI_SILU_BWD) MAPDIT_REQUIRE(g->resid != nullptr, "gemm_bf16: SILU_BWD epilogue needs the pre-activation in `resid`");
  MAPDIT_REQUIRE(epi >= MAPDIT_EPI_STORE && epi <= MAPDIT_EPI_STORE_DELTA, "gemm_bf16: unknown epilogue");
  if (epi == MAPDIT_EPI_STORE_DELTA)
    MAPDIT_REQUIRE(g->resid && g->aux && g->n % 64 == 0, "gemm_bf16: STORE_DELTA epilogue needs o in `resid`, delta in `aux` and N % 64 == 0");
  if (epi == MAPDIT_EPI_RESID || epi == MAPDIT_EPI_RESID_MOD || epi == MAPDIT_EPI_RESID_ROT) {
    MAPDIT_REQUIRE(g->resid && g->gate && g->tokens > 0 && g->ldmod % 4 == 0, "gemm_bf16: residual epilogue needs resid/gate/tokens");
    if (epi == MAPDIT_EPI_RESID_MOD) MAPDIT_REQUIRE(g->out2 && g->shift && g->scale && g->gain, "gemm_bf16: modulate epilogue needs out2/shift/scale/gain");
    if (epi == MAPDIT_EPI_RESID_ROT)
      MAPDIT_REQUIRE(g->out2 && g->shift && g->ldrot % 4 == 0 && ((uintptr_t)g->shift & 15) == 0,
                     "gemm_bf16: rotation epilogue needs out2 and the 16-byte aligned (cos, sin) table in `shift`");
  }
  MAPDIT_REQUIRE(epi == MAPDIT_EPI_STORE || g->out_dtype == MAPDIT_BF16, "gemm_bf16: fused epilogues write bf16");
  EpiParams ep;
  ep.out = g->out; ep.out2 = g->out2; ep.resid = g->resid; ep.gate = g->gate; ep.shift = g->shift; ep.scale = g->scale;
  ep.gain = g->gain; ep.aux = g->aux; ep.ldo = g->ldo; ep.ldmod = g->ldmod; ep.ldshift = (epi == MAPDIT_EPI_RESID_ROT && g->ldrot > 0) ? g->ldrot : g->ldmod; ep.M = g->m; ep.N = g->n; ep.tokens = g->tokens > 0 ? g->tokens : 1;
  ep.qk_cols = g->qk_cols; ep.epilogue = epi; ep.out_f32 = (g->out_dtype == MAPDIT_F32); ep.eps = g->eps; ep.variant = mapdit_variant();

  const int sms = num_sms_cached();
  cudaStream_t s = (cudaStream_t)stream;
  if (g_use_2cta) {  // CTA-pair kernel for the large-M GEMMs; small / odd shapes fall through to the 1-CTA kernel
    int rc2 = mapdit_gemm_bf16_2cta(g, ep, s, sms);
    if (rc2 == MAPDIT_OK) {
      MAPDIT_LAUNCH_CHECK("gemm_bf16(2cta)");
      if (post_qknorm) return mapdit_qk_normalize(g->out, g->m, g->qk_cols / 2, g->head_dim, g->eps, MAPDIT_BF16, stream);
      return MAPDIT_OK;
    }
    if (rc2 != MAPDIT_ERR_UNSUPPORTED) return rc2;
  }
  const int mb = (g->m + BM - 1) / BM;
  const bool need64 = (epi == MAPDIT_EPI_QKNORM || epi == MAPDIT_EPI_STORE_DELTA);
  // largest BN that divides N and still yields >= 2 waves of tiles; otherwise the smallest legal one
  const int cand[5] = {256, 192, 128, 64, 32};
  int bn = 0;
  for (int i = 0; i < 5; ++i) {
    int c = cand[i];
    if (g->n % c != 0 || (need64 && c % 64 != 0)) continue;
    bn = c;
    if ((long long)mb * (g->n / c) >= 2LL * sms) break;
  }
  if (bn == 0) bn = need64 ? 64 : 32;  // N tail handled by column masking
  int rc;
  switch (bn) {
    case 256: rc = launch<256>(g, ep, s, sms); break;
    case 192: rc = launch<192>(g, ep, s, sms); break;
    case 128: rc = launch<128>(g, ep, s, sms); break;
    case 64: rc = launch<64>(g, ep, s, sms); break;
    default: rc = launch<32>(g, ep, s, sms); break;
  }
  if (rc != MAPDIT_OK) return rc;
  MAPDIT_LAUNCH_CHECK("gemm_bf16");
  if (post_qknorm) return mapdit_qk_normalize(g->out, g->m, g->qk_cols / 2, g->head_dim, g->eps, MAPDIT_BF16, stream);
  return MAPDIT_OK;
}

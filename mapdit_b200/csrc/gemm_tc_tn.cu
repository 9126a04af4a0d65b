// Weight-gradient GEMM on tcgen05:  C[N_out, K_in] (fp32) = dY[M, N_out]^T · X[M, K_in]   (contraction over the M tokens)
//
// Both operands are read as they lie in HBM (token-major, row-major): for the MMA they are "MN-major"
// (the non-contracted dimension is the contiguous one), which tcgen05 supports for 16-bit types, so no
// transposed copies of the activations are ever written.  TMA boxes of {64 channels x 64 tokens} land in
// the 128B-swizzled MN-major canonical layout (8-token groups of 1 KB, 64-channel panels of 8 KB:
// SBO = 1024, LBO = 8192).  Split-K over the token dimension fills the 148 SMs when the weight has few
// output tiles; partial sums are combined with 16-byte fp32 reductions into a zeroed C.
// Replaces autograd's x^T·dy of F.linear (src/basic/mp_linear.py:46,75).
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int NUM_THREADS = 256;
constexpr int PANEL_BYTES = 64 * BK * 2;  // one {64 channels x 64 tokens} TMA box = 8 KB

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_STAGES = (227 * 1024 - 2048) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, float* __restrict__ C, long long ldc,
               int n_out, int k_in, int num_m_blocks, int num_n_blocks, int num_k_blocks, int splits) {
  using Cf = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cf::STAGES * Cf::STAGE_BYTES);
  uint64_t* empty = full + Cf::STAGES;
  uint64_t* tfull = empty + Cf::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = num_m_blocks * num_n_blocks * splits;
  const int kb_per_split = (num_k_blocks + splits - 1) / splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cf::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cf::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // item -> (tile, split); tile -> (m_blk over N_out, n_blk over K_in)
  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / splits, sp = item - tile * splits;
      const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
      const int kb0 = sp * kb_per_split, kb1 = min(num_k_blocks, kb0 + kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cf::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[stage], Cf::STAGE_BYTES);
#pragma unroll
        for (int p = 0; p < BM / 64; ++p) tma_load_2d(sa + p * PANEL_BYTES, &tma_a, &full[stage], m_blk * BM + p * 64, kb * BK);
#pragma unroll
        for (int p = 0; p < BN / 64; ++p)
          tma_load_2d(sa + Cf::A_BYTES + p * PANEL_BYTES, &tma_b, &full[stage], n_blk * BN + p * 64, kb * BK);
        if (++stage == Cf::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);  // both operands MN-major
    int stage = 0;
    uint32_t phase = 0, acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / splits, sp = item - tile * splits;
      const int kb0 = sp * kb_per_split, kb1 = min(num_k_blocks, kb0 + kb_per_split);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cf::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cf::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {  // 16 tokens = two 8-row groups = 2 KB inside every panel
          const uint64_t adesc = make_smem_desc(a_addr + k * 2048, PANEL_BYTES, 1024);
          const uint64_t bdesc = make_smem_desc(b_addr + k * 2048, PANEL_BYTES, 1024);
          if (leader) umma_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        if (leader) umma_commit(&empty[stage]);
        if (++stage == Cf::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) umma_commit(&tfull[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / splits, sp = item - tile * splits;
      const int m_blk = tile / num_n_blocks, n_blk = tile - m_blk * num_n_blocks;
      const int kb0 = sp * kb_per_split;
      const bool has_work = kb0 < num_k_blocks;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c, r);
        tmem_ld_wait();
        const int col = n_blk * BN + c;
        if (row >= n_out || col >= k_in || !has_work) continue;
        float* dst = C + (long long)row * ldc + col;
        const int nvalid = min(32, k_in - col);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g * 4 >= nvalid) break;
          float4 v = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]),
                                 __uint_as_float(r[g * 4 + 3]));
          if (splits == 1) *reinterpret_cast<float4*>(dst + g * 4) = v;
          else atomicAdd(reinterpret_cast<float4*>(dst + g * 4), v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cf::TMEM_COLS>(tmem_base);
}

template <int BN>
int launch(const void* dy, long long ldy, const void* x, long long ldx, float* c, long long ldc, int m_tokens, int n_out, int k_in,
           cudaStream_t stream, int sms) {
  using Cf = Cfg<BN>;
  CUtensorMap ta, tb;
  const uint64_t da[2] = {(uint64_t)n_out, (uint64_t)m_tokens}, db[2] = {(uint64_t)k_in, (uint64_t)m_tokens};
  const uint64_t sa[1] = {(uint64_t)ldy * 2}, sb[1] = {(uint64_t)ldx * 2};
  const uint32_t box[2] = {64, BK};
  CUresult r1 = mapdit_encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, da, sa, box, CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r2 = mapdit_encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, db, sb, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    mapdit_set_error("gemm_bf16_tn: cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
    return MAPDIT_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM_BYTES);
    if (e != cudaSuccess) {
      mapdit_set_error("gemm_bf16_tn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int mb = (n_out + BM - 1) / BM, nb = (k_in + BN - 1) / BN, kb = (m_tokens + BK - 1) / BK;
  const int tiles = mb * nb;
  // Split-K so that the persistent grid runs whole waves: cost ~ ceil(items / SMs) rounds x (k-blocks per item).  The former
  // "2 x SMs / tiles" rule gave 2.2-2.4 waves (3 rounds) for the qkv / fc1 / fc2 weight gradients of DiT-B/2.
  int splits = 1;
  if (tiles < 2 * sms) {
    long long best_cost = -1;
    const int max_s = kb < 32 ? kb : 32;
    for (int sp = 1; sp <= max_s; ++sp) {
      const int per = (kb + sp - 1) / sp;
      const int real = (kb + per - 1) / per;  // every split non-empty
      if (real != sp) continue;
      const long long rounds = ((long long)tiles * sp + sms - 1) / sms;
      const long long cost = rounds * (per + 6);  // + a few k-blocks' worth of per-item prologue / epilogue / reduction
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        splits = sp;
      }
    }
  }
  if (splits > 1) {
    cudaError_t e = cudaMemset2DAsync(c, (size_t)ldc * 4, 0, (size_t)k_in * 4, (size_t)n_out, stream);
    if (e != cudaSuccess) {
      mapdit_set_error("gemm_bf16_tn: memset: %s", cudaGetErrorString(e));
      return MAPDIT_ERR_CUDA;
    }
  }
  const int items = tiles * splits;
  const int grid = items < sms ? items : sms;
  gemm_tn_kernel<BN><<<grid, NUM_THREADS, Cf::SMEM_BYTES, stream>>>(ta, tb, c, ldc, n_out, k_in, mb, nb, kb, splits);
  return MAPDIT_OK;
}
}  // namespace

extern "C" int mapdit_gemm_bf16_tn(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* c, int64_t ldc, int m_tokens,
                                   int n_out, int k_in, void* stream) {
  MAPDIT_REQUIRE(dy && x && c && m_tokens > 0 && n_out > 0 && k_in > 0, "gemm_bf16_tn: bad args");
  MAPDIT_REQUIRE(ldy % 8 == 0 && ldx % 8 == 0 && n_out % 8 == 0 && k_in % 8 == 0 && ldc % 4 == 0,
                 "gemm_bf16_tn: dimensions must be multiples of 8 (16-byte TMA / vector alignment)");
  MAPDIT_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)c & 15) == 0, "gemm_bf16_tn: 16-byte alignment");
  int sms = 148;
  {
    static int cached = 0;
    if (!cached) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
      if (cached <= 0) cached = 148;
    }
    sms = cached;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  // widest tile that divides K_in: 192 covers the D = 384 / 1152 models (DiT-S, DiT-XL), whose weights are not multiples of 256 columns
  if (k_in % 256 == 0) rc = launch<256>(dy, ldy, x, ldx, c, ldc, m_tokens, n_out, k_in, s, sms);
  else if (k_in % 192 == 0) rc = launch<192>(dy, ldy, x, ldx, c, ldc, m_tokens, n_out, k_in, s, sms);
  else if (k_in % 128 == 0) rc = launch<128>(dy, ldy, x, ldx, c, ldc, m_tokens, n_out, k_in, s, sms);
  else rc = launch<64>(dy, ldy, x, ldx, c, ldc, m_tokens, n_out, k_in, s, sms);
  if (rc != MAPDIT_OK) return rc;
  MAPDIT_LAUNCH_CHECK("gemm_bf16_tn");
  return MAPDIT_OK;
}

// K4: cosine attention on the 5th-gen tensor cores (head_dim 64, tokens a multiple of 64).
//
// qkv[N*T, 3D] bf16 holds L2-normalised q,k heads (done in the registers of the producing GEMM's
// epilogue, gemm_tc.cu EPI_QKNORM), so logits = q·k/8 are bounded by +-8 and softmax needs no running
// max: p = exp(logit - 8) never overflows and the sum over <= 1024 keys stays in fp32 range
// (SURVEY.md §A.4).  Replaces src/layers/attention.py:43-49.
//
// One CTA = one (sample, head, 128-query tile).  Key/value blocks of 64 stream through a 2-stage TMA
// ring.  warp 0: TMA producer; warp 1: tcgen05.mma issuer (S = Q K^T into a double-buffered TMEM
// tile, O += P V with V as an MN-major B operand); warps 2-5: softmax (tcgen05.ld S row -> exp2 ->
// bf16 P written to shared memory in the 128B-swizzled K-major layout the MMA reads) and the final
// O / rowsum epilogue.  Two CTAs fit per SM (80 KB smem, 256 TMEM columns each) so one CTA's
// exponentials overlap the other's MMAs and loads.
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int HD = 64, QT = 128, KB = 64;
constexpr int Q_BYTES = QT * HD * 2;   // 16 KB
constexpr int KV_BYTES = KB * HD * 2;  // 8 KB each
constexpr int P_BYTES = QT * KB * 2;   // 16 KB
constexpr int NS = 3;  // K/V ring depth (TMA latency ~ 2 us >> one block's compute)
constexpr int SMEM_BYTES = Q_BYTES + NS * 2 * KV_BYTES + 2 * P_BYTES + 2 * 2 * QT * 4 /*row-sum exchange*/ + 1024 + 256;
constexpr int NTHREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two warps per TMEM lane quarter, 32 key columns each)
constexpr uint32_t TMEM_COLS = 256;  // S0 [0,64) S1 [64,128) O [128,192)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(NTHREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, bf16* __restrict__ o,
               float* __restrict__ lse, int tokens, int heads, int n_samples) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Q_BYTES;            // stage s: K at sKV + s*2*KV_BYTES, V right after
  uint8_t* sP = sKV + NS * 2 * KV_BYTES;  // 2 buffers
  float* sRS = reinterpret_cast<float*>(sP + 2 * P_BYTES);  // [2 (item parity)][2 (column half)][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRS + 2 * 2 * QT);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;          // [NS]
  uint64_t* kv_empty = kv_full + NS;     // [NS]
  uint64_t* s_full = kv_empty + NS;      // [2]
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;
  uint64_t* q_empty = o_full + 1;
  uint64_t* o_empty = q_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int nkb = tokens / KB;
  const int nqt = (tokens + QT - 1) / QT;
  const int total_items = nqt * heads * n_samples;  // persistent: each CTA walks items blockIdx.x, +gridDim.x, ...

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_kv);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
      mbar_init(&p_full[i], 8);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------ TMA producer (runs ahead into the next item)
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int qt = item % nqt, rest = item / nqt, h = rest % heads, n = rest / heads;
      const int row_base = n * tokens;
      mbar_wait(q_empty, (it & 1) ^ 1);
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &tm_q, q_full, h * HD, row_base + qt * QT);
      for (int j = 0; j < nkb; ++j, ++g) {
        const int s = g % NS;
        mbar_wait(&kv_empty[s], ((g / NS) & 1) ^ 1);
        uint8_t* k_dst = sKV + s * 2 * KV_BYTES;
        mbar_arrive_expect_tx(&kv_full[s], 2 * KV_BYTES);
        tma_load_2d(k_dst, &tm_kv, &kv_full[s], D + h * HD, row_base + j * KB);
        tma_load_2d(k_dst + KV_BYTES, &tm_kv, &kv_full[s], 2 * D + h * HD, row_base + j * KB);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();  // whole warp runs the loop, one lane issues (see tc_common.cuh)
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(QT, KB, 0, 0);   // S = Q K^T : both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(QT, HD, 0, 1);   // O = P V   : V is MN-major (d contiguous)
    const uint32_t q_addr = smem_u32(sQ);
    auto issue_s = [&](uint32_t gg) {
      const int s = gg & 1, kvs = gg % NS;
      mbar_wait(&kv_full[kvs], (gg / NS) & 1);
      mbar_wait(&s_empty[s], ((gg >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sKV + kvs * 2 * KV_BYTES);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        if (leader) umma_ss(tmem_base + s * KB, make_smem_desc(q_addr + k * 32, 16, 1024), make_smem_desc(k_addr + k * 32, 16, 1024), idesc_s,
                k != 0);
      if (leader) umma_commit(&s_full[s]);
    };
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      mbar_wait(q_full, it & 1);
      issue_s(g);
      if (nkb == 1) if (leader) umma_commit(q_empty);  // Q tile free once the item's last S MMA retires
      for (int j = 0; j < nkb; ++j) {
        const uint32_t gg = g + j;
        const int s = gg & 1;
        if (j + 1 < nkb) {
          issue_s(gg + 1);
          if (j + 2 == nkb) if (leader) umma_commit(q_empty);
        }
        if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);  // previous item's O has been read out of TMEM
        mbar_wait(&p_full[s], (gg >> 1) & 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + s * P_BYTES);
        const int kvs = gg % NS;
        const uint32_t v_addr = smem_u32(sKV + kvs * 2 * KV_BYTES + KV_BYTES);
#pragma unroll
        for (int k = 0; k < KB / 16; ++k)  // 16 keys per MMA: P advances 32 B along K, V advances two 8-row groups
          if (leader) umma_ss(tmem_base + 128, make_smem_desc(p_addr + k * 32, 16, 1024), make_smem_desc(v_addr + k * 2048, 1024, 1024), idesc_o,
                  (j | k) != 0);
        if (leader) umma_commit(&kv_empty[kvs]);
        if (leader) umma_commit(&p_empty[s]);
      }
      if (leader) umma_commit(o_full);
      g += nkb;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------ softmax + epilogue
    // TMEM lane quarter = warp % 4; the two warps of a quarter split the 64 key columns (and the 64 output channels)
    const int qq = warp & 3, ch = (warp - 2) >> 2;  // ch = column half 0/1
    const int r = qq * 32 + lane;                    // query row inside the tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
    const float c1 = 0.125f * 1.4426950408889634f, c2 = 8.0f * 1.4426950408889634f;
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int qt = item % nqt, rest = item / nqt, h = rest % heads, n = rest / heads;
      const int row_base = n * tokens, q0 = qt * QT;
      float rowsum = 0.f;
      for (int j = 0; j < nkb; ++j) {
        const uint32_t gg = g + j;
        const int s = gg & 1;
        const uint32_t ph = (gg >> 1) & 1;
        mbar_wait(&s_full[s], ph);
        tc_fence_after();
        uint32_t a0[32];
        tmem_ld32(t_lane + s * KB + ch * 32, a0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2(fmaf(__uint_as_float(a0[2 * i]), c1, -c2)), p1 = ex2(fmaf(__uint_as_float(a0[2 * i + 1]), c1, -c2));
          rowsum += p0 + p1;  // fp32 sum of the unrounded probabilities (the bf16 rounding of P is unbiased noise on top)
          pk[i] = pack_bf16(p0, p1);
        }
        mbar_wait(&p_empty[s], ph ^ 1);
        // row r of the [128 x 64] bf16 K-major SWIZZLE_128B tile: 16-byte chunk c lives at c ^ (r % 8); this warp owns chunks 4ch..4ch+3
        uint8_t* prow = sP + s * P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(prow + (((ch * 4 + c) ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
      }
      g += nkb;
      // combine the two column halves' row sums (double-buffered by item parity; 256 softmax threads)
      float* rs = sRS + (it & 1) * 2 * QT;
      rs[ch * QT + r] = rowsum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float total = rs[r] + rs[QT + r];
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      uint32_t o0[32];
      tmem_ld32(t_lane + 128 + ch * 32, o0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (q0 + r < tokens) {
        const float inv = 1.0f / total;
        if (lse && ch == 0) lse[(size_t)(row_base + q0 + r) * heads + h] = 8.0f + logf(total);
        bf16* dst = o + (size_t)(row_base + q0 + r) * D + h * HD + ch * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}
}  // namespace

bool mapdit_attn_tc_supported(int tokens, int hd) { return hd == HD && tokens % KB == 0 && tokens >= KB; }

int mapdit_attn_tc_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, void* stream) {
  const int D = heads * hd;
  CUtensorMap tq, tkv;
  const uint64_t dims[2] = {(uint64_t)3 * D, (uint64_t)n * tokens};
  const uint64_t strides[1] = {(uint64_t)3 * D * 2};
  const uint32_t box_q[2] = {HD, QT}, box_kv[2] = {HD, KB};
  CUresult r1 = mapdit_encode_tmap(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r2 = mapdit_encode_tmap(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    mapdit_set_error("attn_tc_fwd: cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
    return MAPDIT_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      mapdit_set_error("attn_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MAPDIT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int items = ((tokens + QT - 1) / QT) * heads * n;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = items < 2 * sms ? items : 2 * sms;  // persistent, two CTAs per SM
  attn_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tkv, (bf16*)o, lse, tokens, heads, n);
  MAPDIT_LAUNCH_CHECK("attn_tc_fwd");
  return MAPDIT_OK;
}

// Fused epilogues shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels (see gemm_tc.cu for what each one replaces).
//
// Data path per epilogue warp: tcgen05.ld (32 lanes x 32 fp32 columns) -> registers -> elementwise math -> bf16 ->
// a private, double-buffered 32x32 staging tile in shared memory (64-byte rows, SWIZZLE_64B so the row-per-lane writes
// are bank-conflict free) -> one TMA store per tile chunk (cp.async.bulk.tensor, bulk-group tracked).  The TMA engine
// writes full lines and clips rows >= M / columns >= N, so the LSU never sees the scattered 16-byte row stores that
// bounded the K = 768 GEMMs.  fp32 outputs (the modulation GEMM) are 128 contiguous bytes per lane and go out directly.
#pragma once
#include "tc_common.cuh"

namespace gemm_epi {
using namespace tc;

constexpr int STG_BYTES_PER_WARP = 2 * 32 * 64;  // two 32x32 bf16 buffers
constexpr int STG_BYTES = 8 * STG_BYTES_PER_WARP;  // 8 epilogue warps

struct EpiParams {
  void* out;
  void* out2;
  const void* resid;
  const float* gate;
  const float* shift;
  const float* scale;
  const float* gain;
  void* aux;
  long long ldo, ldmod;
  long long ldshift;  // leading dimension of `shift` (= ldmod except for the RESID_ROT (cos, sin) table)
  int M, N, tokens, qk_cols, epilogue, out_f32;
  float eps;
  int variant;  // MAPDIT_VAR_* word of the launching thread
};

// TMA store descriptors of the (up to) three bf16 [M, N] outputs: out, out2, aux; load descriptor of the bf16 [M, N] residual
struct alignas(64) EpiTmaps {
  CUtensorMap out, out2, aux, resid;
};

// 16-byte streaming global load that does not allocate in L1
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// sigmoid(x) = 0.5 + 0.5 tanh(x/2): ONE MUFU op (tanh.approx, rel. error 2^-11, far below the bf16 output rounding) instead of
// the ex2 + rcp pair of x/(1+exp(-x)).  The N = 3072 epilogues were MUFU-bound: 2 x 128 x 256 MUFU ops per tile at 16/clk
// = 4096 clk against a 6144-clk main loop, next to the epilogue's own issue slots.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float silu_fast(float x, float mul) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_fast(h), h) * mul;
}

struct Stager {
  uint8_t* base;  // this warp's 4 KB
  uint32_t parity;
  int lane;
  int row0;  // first global row of this warp's 32 rows

  // store 32 consecutive columns [col, col+32) of this warp's 32 rows (lane = row) as bf16 through TMA
  __device__ __forceinline__ void store(const CUtensorMap* map, const float (&f)[32], int col) {
    uint8_t* buf = base + parity * 2048;
    parity ^= 1;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store that last used `buf` has read it
    __syncwarp();
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 u;
      u.x = pack_bf16(f[c * 8 + 0], f[c * 8 + 1]);
      u.y = pack_bf16(f[c * 8 + 2], f[c * 8 + 3]);
      u.z = pack_bf16(f[c * 8 + 4], f[c * 8 + 5]);
      u.w = pack_bf16(f[c * 8 + 6], f[c * 8 + 7]);
      *reinterpret_cast<uint4*>(buf + lane * 64 + ((c ^ sw) << 4)) = u;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map),
                   "r"(smem_u32(buf)), "r"(col), "r"(row0)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  __device__ __forceinline__ void drain() {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
};

// Residual tiles by TMA (2-CTA kernel): the row-per-lane 16-byte loads of `resid` cost 32 line transactions per instruction in
// the LSU -- 4096 LSU cycles per 128 x 256 tile, the size of a whole K = 768 main loop, which is why the fused out-proj GEMM
// ran at half the speed of the same GEMM with a plain store.  Each epilogue warp instead has two private 32 x 32 bf16 buffers
// (64-byte rows, SWIZZLE_64B like the output staging) filled by cp.async.bulk.tensor one chunk ahead; a lane then reads its
// row with four conflict-free 16-byte shared-memory loads.
constexpr int RSTG_BYTES_PER_WARP = 2 * 32 * 64;
constexpr int RSTG_BYTES = 8 * RSTG_BYTES_PER_WARP;
struct ResidLoader {
  uint8_t* base;   // this warp's 4 KB
  uint64_t* bars;  // this warp's two mbarriers (count 1)
  const CUtensorMap* map;
  uint32_t issued, consumed;
  int lane, row0;

  __device__ __forceinline__ void issue(int col) {
    const uint32_t b = issued & 1;
    __syncwarp();  // every lane has finished reading what this buffer held two chunks ago
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[b], 32 * 64);
      tma_load_2d(base + b * 2048, map, &bars[b], col, row0);
    }
    ++issued;
  }
  __device__ __forceinline__ void consume(uint4 (&pre)[4]) {
    const uint32_t b = consumed & 1;
    mbar_wait(&bars[b], (consumed >> 1) & 1);
    const uint8_t* row = base + b * 2048 + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c) pre[c] = *reinterpret_cast<const uint4*>(row + ((c ^ sw) << 4));
    ++consumed;
  }
};

__device__ __forceinline__ void store_row32_f32(void* base, long long off, const float (&f)[32], int nvalid) {
  float* p = reinterpret_cast<float*>(base) + off;
#pragma unroll
  for (int g = 0; g < 8; ++g)
    if (g * 4 < nvalid) *reinterpret_cast<float4*>(p + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
}

// nvalid must be warp-uniform.  The full-chunk case takes eight unpredicated 16-byte loads: with a per-load predicate the
// compiler re-materialises the global-memory descriptor (two R2UR) in front of every LDG, which serialised the loads and left
// their latency fully exposed three times per chunk (ncu: the FFMA/FMUL behind these loads were the top stall of the epilogue).
__device__ __forceinline__ void load_row32_f32(const float* p, float (&f)[32], int nvalid) {
  if (nvalid >= 32) {
    float4 u[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) u[g] = reinterpret_cast<const float4*>(p)[g];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      f[g * 4] = u[g].x;
      f[g * 4 + 1] = u[g].y;
      f[g * 4 + 2] = u[g].z;
      f[g * 4 + 3] = u[g].w;
    }
    return;
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g * 4 < nvalid) u = *reinterpret_cast<const float4*>(p + g * 4);
    f[g * 4] = u.x;
    f[g * 4 + 1] = u.y;
    f[g * 4 + 2] = u.z;
    f[g * 4 + 3] = u.w;
  }
}

// One epilogue warp's share of one accumulator tile.  `half` = which of the two warps of a TMEM lane quarter (alternate
// column chunks).  `wait_acc()` blocks until the accumulator is complete; it is called after the first residual prefetch
// has been issued so the global-load latency overlaps the tail of the main loop.
template <int BN, bool WITH_DELTA = true, typename WaitFn>
__device__ __forceinline__ void run_tile(const EpiParams& ep, const EpiTmaps& tm, Stager& st, uint32_t t_row, int row, int n_blk, int half,
                                         float gsc, float inv_den, WaitFn wait_acc, ResidLoader* rl = nullptr) {
  const bool mod2 = ep.epilogue == MAPDIT_EPI_RESID_MOD || ep.epilogue == MAPDIT_EPI_RESID_ROT;  // writes h to out2
  const bool reads_resid = ep.epilogue == MAPDIT_EPI_RESID || mod2 || ep.epilogue == MAPDIT_EPI_SILU_BWD;
  const bool row_ok = row < ep.M;
  const long long sample = row_ok ? row / ep.tokens : 0;
  uint4 pre[4];
  auto prefetch = [&](int c) {
    const int col = n_blk * BN + c;
    if (rl) {  // TMA path: the chunk lands in this warp's staging buffer (columns / rows past the edge are zero-filled)
      if (col < ep.N) {
        rl->row0 = st.row0;
        rl->issue(col);
      }
      return;
    }
    const int nvalid = min(32, ep.N - col);
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.resid) + (long long)row * ep.ldo + col);
#pragma unroll
    for (int g = 0; g < 4; ++g) pre[g] = (row_ok && g * 8 < nvalid) ? src[g] : make_uint4(0, 0, 0, 0);
  };
  // The per-sample vectors (gate, shift, scale: one 128-byte line per 32-column chunk each) are read by every lane at the
  // same address right before use; ncu showed their L2 round trips exposed twice per chunk (long-scoreboard stalls on the
  // first FFMA/FMUL).  Lanes 0-2 pull the NEXT chunk's lines into L1 one chunk ahead; the streaming residual loads bypass
  // L1 allocation so the ~24 KB of L1 left beside the operand ring keeps them.
  const bool has_vecs = ep.epilogue == MAPDIT_EPI_RESID || mod2;
  auto prefetch_vecs = [&](int c) {
    const int col = n_blk * BN + c;
    const int lane = st.lane;
    if (!has_vecs || c >= BN || col >= ep.N || lane > 2) return;
    const float* base = lane == 0 ? ep.gate : (lane == 1 ? ep.shift : ep.scale);
    if (lane > 0 && (!mod2 || base == nullptr)) return;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(base + sample * (lane == 1 ? ep.ldshift : ep.ldmod) + col));
  };
  // each of the two warps of a lane quarter owns one contiguous half of the tile's columns: consecutive chunks (and their
  // TMA stores / residual loads) then touch the two halves of the same 128-byte lines back to back
  constexpr int CSPAN = BN >= 64 ? BN / 2 : BN;  // BN = 32: only `half` 0 has a chunk
  const int c_begin = half * CSPAN, c_end = BN >= 64 ? c_begin + CSPAN : (half == 0 ? BN : 0);
  prefetch_vecs(c_begin);
  if (reads_resid && c_begin < c_end) prefetch(c_begin);
  if (WITH_DELTA && ep.epilogue == MAPDIT_EPI_STORE_DELTA && rl && half * 64 < BN && n_blk * BN + half * 64 < ep.N) {
    rl->row0 = st.row0;
    rl->issue(n_blk * BN + half * 64);
    rl->issue(n_blk * BN + half * 64 + 32);
  }
  const float silu_mul = (ep.variant & MAPDIT_VAR_PLAIN_SILU) ? 1.0f : 1.0f / MP_SILU_DIV;
  const bool plain_res = ep.variant & MAPDIT_VAR_PLAIN_RESID;
  const float res_a = plain_res ? 1.0f : (1.0f - MP_RES_T) / MP_RES_DEN, res_b = plain_res ? 1.0f : MP_RES_T / MP_RES_DEN;
  const float mod_a = (1.0f - gsc) * inv_den, mod_b = gsc * inv_den;
  wait_acc();

  if (ep.epilogue == MAPDIT_EPI_QKNORM) {
    // 64 columns (= one head) at a time
    if constexpr (BN % 64 == 0) {
      for (int c = half * 64; c < BN; c += 128) {
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + c, r0);
        tmem_ld32(t_row + c + 32, r1);
        tmem_ld_wait();
        const int col = n_blk * BN + c;
        float f0[32], f1[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f0[j] = __uint_as_float(r0[j]);
          f1[j] = __uint_as_float(r1[j]);
        }
        if (col < ep.qk_cols) {
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) ss = fmaf(f0[j], f0[j], fmaf(f1[j], f1[j], ss));
          const float sc = 8.0f / (sqrtf(ss) + ep.eps);  // sqrt(head_dim = 64)
          if (ep.aux && row_ok) reinterpret_cast<float*>(ep.aux)[(long long)row * (ep.qk_cols >> 6) + (col >> 6)] = sc;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f0[j] *= sc;
            f1[j] *= sc;
          }
        }
        st.store(&tm.out, f0, col);
        st.store(&tm.out, f1, col + 32);
      }
    }
    return;
  }
  if (WITH_DELTA && ep.epilogue == MAPDIT_EPI_STORE_DELTA) {
    // out = acc; delta[row, head] = sum over the head's 64 columns of bf16(acc) * o (`resid`): the out-proj dgrad emits the
    // attention backward's delta = dO.O, so no separate pass re-reads dO and O from HBM
    if constexpr (BN % 64 == 0) {
      for (int c = half * 64; c < BN; c += 128) {
        const int col = n_blk * BN + c;
        if (col >= ep.N) break;  // warp-uniform (N % 64 == 0)
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + c, r0);
        tmem_ld32(t_row + c + 32, r1);
        uint4 oa[4], ob[4];
        if (rl) {  // both 32-column chunks were requested before the accumulator wait (first group) or one group ago
          rl->consume(oa);
          rl->consume(ob);
          if (c + 128 < BN && col + 128 < ep.N) {
            rl->issue(col + 128);
            rl->issue(col + 160);
          }
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.resid) + (long long)row * ep.ldo + col);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            oa[g] = row_ok ? src[g] : make_uint4(0, 0, 0, 0);
            ob[g] = row_ok ? src[4 + g] : make_uint4(0, 0, 0, 0);
          }
        }
        tmem_ld_wait();
        float dl = 0.f;
        // one 32-column chunk at a time (keeps the live registers below the 168 the kernel has): the attention backward reads dO
        // as bf16, so the dot product uses the rounded values
        auto chunk = [&](const uint32_t (&r)[32], const uint4 (&ov)[4], int ccol) {
          float f[32];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&ov[g]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = g * 8 + 2 * j;
              const float2 d = __bfloat1622float2(__floats2bfloat162_rn(__uint_as_float(r[e]), __uint_as_float(r[e + 1])));
              const float2 a = __bfloat1622float2(ho[j]);
              f[e] = d.x;
              f[e + 1] = d.y;
              dl = fmaf(d.x, a.x, fmaf(d.y, a.y, dl));
            }
          }
          st.store(&tm.out, f, ccol);
        };
        chunk(r0, oa, col);
        chunk(r1, ob, col + 32);
        if (row_ok) reinterpret_cast<float*>(ep.aux)[(long long)row * (ep.N >> 6) + (col >> 6)] = dl;
      }
    }
    return;
  }
  for (int c = c_begin; c < c_end; c += 32) {
    uint32_t r[32];
    tmem_ld32(t_row + c, r);
    tmem_ld_wait();
    const int col = n_blk * BN + c;
    const int nvalid = min(32, ep.N - col);
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
    float xo[32];
    if (reads_resid) {
      if (rl) {
        // next chunk first (its buffer was consumed one chunk ago), then this chunk out of shared memory
        if (c + 32 < c_end) prefetch(c + 32);
        if (col < ep.N) rl->consume(pre);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&pre[g]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t2 = __bfloat1622float2(hp[j]);
          xo[g * 8 + 2 * j] = t2.x;
          xo[g * 8 + 2 * j + 1] = t2.y;
        }
      }
      if (!rl && c + 32 < c_end) prefetch(c + 32);  // next chunk's residual while this one is processed
    }
    if (c + 32 < c_end) prefetch_vecs(c + 32);
    if (nvalid <= 0) continue;  // warp-uniform
    if (ep.epilogue == MAPDIT_EPI_STORE) {
      if (ep.out_f32) {
        if (row_ok) store_row32_f32(ep.out, (long long)row * ep.ldo + col, f, nvalid);
      } else {
        st.store(&tm.out, f, col);
      }
    } else if (ep.epilogue == MAPDIT_EPI_MPSILU) {
      if (ep.out2) st.store(&tm.out2, f, col);  // pre-activation for the backward
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = silu_fast(f[j], silu_mul);
      st.store(&tm.out, f, col);
    } else if (ep.epilogue == MAPDIT_EPI_SILU_BWD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float sg = fmaf(0.5f, tanh_fast(0.5f * xo[j]), 0.5f);
        f[j] *= sg * fmaf(xo[j], 1.0f - sg, 1.0f) * silu_mul;
      }
      st.store(&tm.out, f, col);
    } else {  // RESID / RESID_MOD / RESID_ROT
      float gt[32];
      if (ep.aux) st.store(&tm.aux, f, col);  // raw branch output, needed for d(gate)
      load_row32_f32(ep.gate + sample * ep.ldmod + col, gt, nvalid);
      // x' = (0.7 x + 0.3 gate acc)/den as  (b gate) acc + a x  with the constants folded into warp-uniform scalars: two
      // instructions per element.  (The literal form, with its `plain_res` select and lerp's two-sided formula, compiled to ~13
      // mostly predicated instructions per element: the fused epilogue issued 28 k warp instructions per 128 x 256 tile, more
      // issue cycles than the K = 768 main loop has.)
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaf(gt[j] * res_b, f[j], res_a * xo[j]);
      st.store(&tm.out, f, col);
      if (ep.epilogue == MAPDIT_EPI_RESID_MOD) {
        // h = lerp(x' scale, shift, g)/den(g) = x' (scale (1-g)/den) + shift g/den
        load_row32_f32(ep.shift + sample * ep.ldmod + col, xo, nvalid);
        load_row32_f32(ep.scale + sample * ep.ldmod + col, gt, nvalid);
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], gt[j] * mod_a, xo[j] * mod_b);
        st.store(&tm.out2, f, col);
      } else if (ep.epilogue == MAPDIT_EPI_RESID_ROT) {
        // rotation modulation (UNPINNED): channel pair (2i, 2i+1) of x' rotated by theta_i; xo = 16 (cos, sin) pairs
        load_row32_f32(ep.shift + sample * ep.ldshift + col, xo, nvalid);
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const float a = f[2 * p], b = f[2 * p + 1];
          f[2 * p] = a * xo[2 * p] - b * xo[2 * p + 1];
          f[2 * p + 1] = fmaf(a, xo[2 * p + 1], b * xo[2 * p]);
        }
        if (ep.scale) {
          load_row32_f32(ep.scale + sample * ep.ldmod + col, gt, nvalid);
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= gt[j];
        }
        st.store(&tm.out2, f, col);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Fused residual epilogues, second generation (2-CTA kernel, tokens % 32 == 0 so that a warp's 32 rows lie in one sample).
//
// What bounded the first generation (15-17 k cycles per 256 x 256 tile against a 6-9 k main loop, the MMA warp idle
// on its accumulator stages) and what replaces it:
//   * per-sample vectors (gate | shift | scale, or gate | (cos, sin) | scale): every lane used to read them from global memory
//     right before use, three exposed L2 round trips per 32-column chunk.  Each warp now stages the vectors of its whole column
//     half once per tile in shared memory, already multiplied by the warp-uniform constants; the global loads for the NEXT tile
//     are issued during the last chunk of the current one and only consumed at the head of the next;
//   * the residual arrives by TMA as one flat stream over all chunks of all tiles of this warp, one chunk ahead in three
//     rotating 32 x 32 buffers, so the first chunk of a tile is in flight during the last chunk of the previous tile (a clock64
//     timeline showed 1.5-2 k cycles of exposed TMA latency at the head of every tile before);
//   * up to three TMA stores per chunk rotated through two staging buffers, so every chunk waited for a store issued moments
//     earlier.  Now no buffer is reused within a chunk: x' is written back over the residual rows it was computed from (the
//     lane that read a row writes it) and stored from there, aux and h have their own buffer each, and every
//     cp.async.bulk.wait_group is for a store issued a whole chunk earlier; one fence.proxy.async covers aux and x';
//   * accumulator chunks are read one ahead and the accumulator stage is released as soon as the last chunk is in registers.
// (Tried and dropped: coalesced st.global through a transpose buffer instead of TMA stores: 580-860 cycles per 32 x 32 tile in
// the LSU against ~200 for fence + TMA store issue; 0.113 ms against 0.095 ms for the out-proj GEMM.)
// The arithmetic (and therefore every output bit) is the same as run_tile's.
constexpr int FR_XBUFS = 3;
constexpr int FR_BUF_BYTES_PER_WARP = (FR_XBUFS + 2) * 2048;
__host__ __device__ constexpr int fr_vec_bytes_per_warp(int bn) { return 3 * (bn / 2) * 4; }
__host__ __device__ constexpr int fr_bytes(int bn) { return 8 * (FR_BUF_BYTES_PER_WARP + fr_vec_bytes_per_warp(bn)); }

__device__ __forceinline__ void bulk_wait_read(int k) {  // k warp-uniform, 0..3
  if (k <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (k == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else if (k == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

struct FusedResid {
  uint8_t* bufs;   // this warp's (FR_XBUFS + 2) x 2 KB: X[0..2] residual landing + x' staging, then aux, then h
  float* vec;      // this warp's 3 x (BN/2) floats
  uint64_t* bars;  // this warp's FR_XBUFS mbarriers (count 1)
  uint32_t consumed;
  int lane;
  // the warp's chunk stream: tile k of this CTA pair is first_tile + k * tile_stride; chunk s belongs to tile s / NCHUNK
  int first_tile, tile_stride, total_tiles, num_n_blocks, row_in_pair;  // row_in_pair = rank * 128 + quarter * 32
  int half;
  float4 nv[3];     // raw vectors of the next tile, prefetched (valid when nv_tile == that tile)
  int nv_tile;
  long long* fine;  // developer timeline (tools/gemm_timeline.py): 8 clock64 stamps per chunk, or null

  __device__ __forceinline__ void drain() {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  __device__ __forceinline__ void stamp(int ci, int e) {
    if (fine && lane == 0) fine[ci * 8 + e] = clock64();
  }
};

// raw per-sample vectors of `tile` for this lane's four columns: v[0] = gate, MOD: v[1] = scale, v[2] = shift; ROT: v[1] = (cos, sin)
// pairs, v[2] = scale or 1.  Only loads: the values are first used (fr_stage_vectors) at the head of the tile they belong to.
template <int BN>
__device__ __forceinline__ void fr_load_vectors(const EpiParams& ep, const FusedResid& fr, int tile, float4 (&v)[3]) {
  constexpr int CSPAN = BN / 2;
  static_assert(CSPAN / 4 <= 32, "one float4 per lane covers the column half");
  const int m_pair = tile / fr.num_n_blocks, n_blk = tile - m_pair * fr.num_n_blocks;
  const int row0 = m_pair * 256 + fr.row_in_pair;
  const int col = n_blk * BN + fr.half * CSPAN + 4 * fr.lane;
  const long long sample = row0 < ep.M ? row0 / ep.tokens : 0;
  const bool mod = ep.epilogue == MAPDIT_EPI_RESID_MOD, rot = ep.epilogue == MAPDIT_EPI_RESID_ROT;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  v[0] = z; v[1] = z; v[2] = z;
  if (4 * fr.lane < CSPAN && col < ep.N) {
    v[0] = *reinterpret_cast<const float4*>(ep.gate + sample * ep.ldmod + col);
    if (mod) {
      v[1] = *reinterpret_cast<const float4*>(ep.scale + sample * ep.ldmod + col);
      v[2] = *reinterpret_cast<const float4*>(ep.shift + sample * ep.ldmod + col);
    } else if (rot) {
      v[1] = *reinterpret_cast<const float4*>(ep.shift + sample * ep.ldshift + col);
      v[2] = ep.scale ? *reinterpret_cast<const float4*>(ep.scale + sample * ep.ldmod + col) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
  }
}
// constants folded in (v0 = gate b; MOD: v1 = scale (1-g)/den, v2 = shift g/den) -> this warp's shared-memory vectors
template <int BN>
__device__ __forceinline__ void fr_stage_vectors(const EpiParams& ep, FusedResid& fr, float gsc, float inv_den) {
  constexpr int CSPAN = BN / 2;
  const float res_b = (ep.variant & MAPDIT_VAR_PLAIN_RESID) ? 1.0f : MP_RES_T / MP_RES_DEN;
  float4 g = fr.nv[0], v1 = fr.nv[1], v2 = fr.nv[2];
  g.x *= res_b; g.y *= res_b; g.z *= res_b; g.w *= res_b;
  if (ep.epilogue == MAPDIT_EPI_RESID_MOD) {
    const float mod_a = (1.0f - gsc) * inv_den, mod_b = gsc * inv_den;
    v1.x *= mod_a; v1.y *= mod_a; v1.z *= mod_a; v1.w *= mod_a;
    v2.x *= mod_b; v2.y *= mod_b; v2.z *= mod_b; v2.w *= mod_b;
  }
  if (4 * fr.lane < CSPAN) {
    *reinterpret_cast<float4*>(fr.vec + 4 * fr.lane) = g;
    *reinterpret_cast<float4*>(fr.vec + CSPAN + 4 * fr.lane) = v1;
    *reinterpret_cast<float4*>(fr.vec + 2 * CSPAN + 4 * fr.lane) = v2;
  }
}

// request chunk `s` of this warp's residual stream (no-op past the last tile); clipped column chunks load column 0 so that
// every chunk of every tile completes exactly one phase of its barrier
template <int BN>
__device__ __forceinline__ void fr_issue(const EpiParams& ep, const EpiTmaps& tm, FusedResid& fr, uint32_t s) {
  constexpr int CSPAN = BN / 2, NCHUNK = CSPAN / 32;
  const int tile = fr.first_tile + (int)(s / NCHUNK) * fr.tile_stride;
  if (tile >= fr.total_tiles) return;
  const int ci = (int)(s % NCHUNK);
  const int m_pair = tile / fr.num_n_blocks, n_blk = tile - m_pair * fr.num_n_blocks;
  int col = n_blk * BN + fr.half * CSPAN + 32 * ci;
  if (col >= ep.N) col = 0;
  const uint32_t b = s % FR_XBUFS;
  if (fr.lane == 0) {
    mbar_arrive_expect_tx(&fr.bars[b], 32 * 64);
    tma_load_2d(fr.bufs + b * 2048, &tm.resid, &fr.bars[b], col, m_pair * 256 + fr.row_in_pair);
  }
}

template <int BN, typename WaitFn, typename ReleaseFn>
__device__ __forceinline__ void run_tile_fused_resid(const EpiParams& ep, const EpiTmaps& tm, FusedResid& fr, uint32_t t_row, int tile,
                                                     float gsc, float inv_den, WaitFn wait_acc, ReleaseFn release_acc) {
  constexpr int CSPAN = BN / 2, NCHUNK = CSPAN / 32;
  const int lane = fr.lane, half = fr.half;
  const int m_pair = tile / fr.num_n_blocks, n_blk = tile - m_pair * fr.num_n_blocks;
  const int row0 = m_pair * 256 + fr.row_in_pair;
  const int col0 = n_blk * BN + half * CSPAN;
  const bool mod = ep.epilogue == MAPDIT_EPI_RESID_MOD, rot = ep.epilogue == MAPDIT_EPI_RESID_ROT;
  const bool has_h = mod || rot, has_aux = ep.aux != nullptr;
  const int groups = 1 + (has_h ? 1 : 0) + (has_aux ? 1 : 0);  // bulk groups committed per chunk (empty ones for clipped chunks)
  const bool rows_ok = row0 < ep.M;  // M % 32 == 0: a warp's rows are all inside or all outside
  const float res_a = (ep.variant & MAPDIT_VAR_PLAIN_RESID) ? 1.0f : (1.0f - MP_RES_T) / MP_RES_DEN;
  uint8_t* abuf = fr.bufs + FR_XBUFS * 2048;
  uint8_t* hbuf = abuf + 2048;
  const int sw = (lane >> 1) & 3;

  // ---- this tile's per-sample vectors -> shared memory (loaded during the previous tile's last chunk when there was one)
  if (fr.nv_tile != tile) fr_load_vectors<BN>(ep, fr, tile, fr.nv);
  __syncwarp();  // every lane is done reading the previous tile's vectors
  fr_stage_vectors<BN>(ep, fr, gsc, inv_den);
  __syncwarp();
  wait_acc();

  // accumulator chunks are read one ahead: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
  uint32_t ra[32], rb[32];
  tmem_ld32(t_row + half * CSPAN, ra);
#pragma unroll
  for (int ci = 0; ci < NCHUNK; ++ci) {
    const int col = col0 + 32 * ci;
    uint32_t(&r)[32] = (ci & 1) ? rb : ra;
    uint32_t(&rn)[32] = (ci & 1) ? ra : rb;
    const bool last = ci + 1 == NCHUNK;
    // this chunk's residual rows -> registers
    const uint32_t b = fr.consumed % FR_XBUFS;
    fr.stamp(ci, 0);
    mbar_wait(&fr.bars[b], (fr.consumed / FR_XBUFS) & 1);
    fr.stamp(ci, 1);
    uint8_t* xrow = fr.bufs + b * 2048 + lane * 64;
    uint4 pre[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pre[k] = *reinterpret_cast<const uint4*>(xrow + ((k ^ sw) << 4));
    // stores issued a whole chunk ago have read their buffers: aux(s-1) [-> abuf is free], x'(s-2) [-> the next landing buffer]
    if (lane == 0) bulk_wait_read(has_aux ? groups - 1 : groups);
    __syncwarp();
    fr_issue<BN>(ep, tm, fr, fr.consumed + 1);
    ++fr.consumed;
    fr.stamp(ci, 2);
    if (last) {  // the next tile's vectors: issued here, first used at the head of that tile
      const int nt = tile + fr.tile_stride;
      if (nt < fr.total_tiles) {
        fr_load_vectors<BN>(ep, fr, nt, fr.nv);
        fr.nv_tile = nt;
      }
    }
    tmem_ld_wait();
    if (!last) tmem_ld32(t_row + half * CSPAN + 32 * (ci + 1), rn);
    else release_acc();  // every tcgen05.ld of this tile has completed: the MMA warp may reuse the accumulator stage
    fr.stamp(ci, 3);
    if (col >= ep.N || !rows_ok) {  // clipped chunk (warp-uniform): nothing to write, but the bulk-group count per chunk stays uniform
      if (lane == 0)
        for (int gI = 0; gI < groups; ++gI) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      continue;
    }
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
    if (has_aux) {  // raw branch output, needed for d(gate)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 u;
        u.x = pack_bf16(f[8 * k + 0], f[8 * k + 1]);
        u.y = pack_bf16(f[8 * k + 2], f[8 * k + 3]);
        u.z = pack_bf16(f[8 * k + 4], f[8 * k + 5]);
        u.w = pack_bf16(f[8 * k + 6], f[8 * k + 7]);
        *reinterpret_cast<uint4*>(abuf + lane * 64 + ((k ^ sw) << 4)) = u;
      }
    }
    // x' = (b gate) acc + a x, written back over the residual rows and stored from there
    const float4* vg = reinterpret_cast<const float4*>(fr.vec + 32 * ci);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&pre[k]);
      const float4 g0 = vg[2 * k], g1 = vg[2 * k + 1];
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x2 = __bfloat1622float2(hp[j]);
        f[8 * k + 2 * j] = fmaf(gg[2 * j], f[8 * k + 2 * j], res_a * x2.x);
        f[8 * k + 2 * j + 1] = fmaf(gg[2 * j + 1], f[8 * k + 2 * j + 1], res_a * x2.y);
      }
      uint4 u;
      u.x = pack_bf16(f[8 * k + 0], f[8 * k + 1]);
      u.y = pack_bf16(f[8 * k + 2], f[8 * k + 3]);
      u.z = pack_bf16(f[8 * k + 4], f[8 * k + 5]);
      u.w = pack_bf16(f[8 * k + 6], f[8 * k + 7]);
      *reinterpret_cast<uint4*>(xrow + ((k ^ sw) << 4)) = u;
    }
    fr.stamp(ci, 4);
    fence_proxy_async();  // one fence for the aux and the x' rows
    __syncwarp();
    if (lane == 0) {
      if (has_aux) tma_store_2d(&tm.aux, abuf, col, row0);
      tma_store_2d(&tm.out, fr.bufs + b * 2048, col, row0);
    }
    fr.stamp(ci, 5);
    if (has_h) {
      if (lane == 0) bulk_wait_read(groups - 1);  // h(s-1) has been read out of hbuf
      __syncwarp();
      fr.stamp(ci, 6);
      const float4* v1 = reinterpret_cast<const float4*>(fr.vec + CSPAN + 32 * ci);
      const float4* v2 = reinterpret_cast<const float4*>(fr.vec + 2 * CSPAN + 32 * ci);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 a0 = v1[2 * k], a1 = v1[2 * k + 1], b0 = v2[2 * k], b1 = v2[2 * k + 1];
        const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float h[8];
        if (mod) {  // h = x' (scale (1-g)/den) + shift g/den
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = fmaf(f[8 * k + j], aa[j], bb[j]);
        } else {  // rotation of channel pairs by (cos, sin) = (aa[2p], aa[2p+1]), then the optional scale
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float xa = f[8 * k + 2 * p], xb = f[8 * k + 2 * p + 1];
            h[2 * p] = (xa * aa[2 * p] - xb * aa[2 * p + 1]) * bb[2 * p];
            h[2 * p + 1] = fmaf(xa, aa[2 * p + 1], xb * aa[2 * p]) * bb[2 * p + 1];
          }
        }
        uint4 u;
        u.x = pack_bf16(h[0], h[1]);
        u.y = pack_bf16(h[2], h[3]);
        u.z = pack_bf16(h[4], h[5]);
        u.w = pack_bf16(h[6], h[7]);
        *reinterpret_cast<uint4*>(hbuf + lane * 64 + ((k ^ sw) << 4)) = u;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_store_2d(&tm.out2, hbuf, col, row0);
      fr.stamp(ci, 7);
    }
  }
}

// host: descriptors for the bf16 outputs (unused slots alias `out` so every map is valid)
inline int make_store_maps(EpiTmaps* tm, const EpiParams& ep) {
  const uint64_t dims[2] = {(uint64_t)ep.N, (uint64_t)ep.M};
  const uint64_t strides[1] = {(uint64_t)ep.ldo * 2};
  const uint32_t box[2] = {32, 32};
  void* outs[4] = {ep.out, ep.out2 ? ep.out2 : ep.out,
                   (ep.aux && ep.epilogue != MAPDIT_EPI_QKNORM && ep.epilogue != MAPDIT_EPI_STORE_DELTA) ? ep.aux : ep.out,
                   ep.resid ? const_cast<void*>(ep.resid) : ep.out};
  CUtensorMap* maps[4] = {&tm->out, &tm->out2, &tm->aux, &tm->resid};
  for (int i = 0; i < 4; ++i) {
    CUresult r = mapdit_encode_tmap(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, outs[i], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (r != CUDA_SUCCESS) return (int)r;
  }
  return 0;
}
}  // namespace gemm_epi

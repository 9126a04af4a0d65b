// Fused epilogues shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels (see gemm_tc.cu for what each one replaces).
#pragma once
#include "tc_common.cuh"

namespace gemm_epi {
using namespace tc;

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
  int M, N, tokens, qk_cols, epilogue, out_f32;
  float eps;
};

__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)) * (1.0f / MP_SILU_DIV); }

// store up to 32 consecutive columns of one row (nvalid multiple of 8)
__device__ __forceinline__ void store_row32(void* base, long long off, const float (&f)[32], int nvalid, bool as_f32) {
  if (as_f32) {
    float* p = reinterpret_cast<float*>(base) + off;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      if (g * 4 < nvalid) *reinterpret_cast<float4*>(p + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
  } else {
    bf16* p = reinterpret_cast<bf16*>(base) + off;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g * 8 < nvalid) {
        uint4 u;
        u.x = pack_bf16(f[g * 8 + 0], f[g * 8 + 1]);
        u.y = pack_bf16(f[g * 8 + 2], f[g * 8 + 3]);
        u.z = pack_bf16(f[g * 8 + 4], f[g * 8 + 5]);
        u.w = pack_bf16(f[g * 8 + 6], f[g * 8 + 7]);
        *reinterpret_cast<uint4*>(p + g * 8) = u;
      }
  }
}

__device__ __forceinline__ void load_row32_bf16(const void* base, long long off, float (&f)[32], int nvalid) {
  const bf16* p = reinterpret_cast<const bf16*>(base) + off;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (g * 8 < nvalid) u = *reinterpret_cast<const uint4*>(p + g * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t = __bfloat1622float2(h[j]);
      f[g * 8 + 2 * j] = t.x;
      f[g * 8 + 2 * j + 1] = t.y;
    }
  }
}

__device__ __forceinline__ void load_row32_f32(const float* p, float (&f)[32], int nvalid) {
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


// One epilogue warp's share of one accumulator tile.  `q` = TMEM lane quarter (warp % 4), `half` = which of the two warps
// of that quarter (alternate column chunks).  `wait_acc()` blocks until the accumulator is complete; it is called after
// the first residual prefetch has been issued so the global-load latency overlaps the tail of the main loop.
template <int BN, typename WaitFn>
__device__ __forceinline__ void run_tile(const EpiParams& ep, uint32_t t_row, int row, int n_blk, int half, float gsc, float inv_den,
                                         WaitFn wait_acc) {
  const bool reads_resid = ep.epilogue == MAPDIT_EPI_RESID || ep.epilogue == MAPDIT_EPI_RESID_MOD || ep.epilogue == MAPDIT_EPI_SILU_BWD;
  const bool row_ok = row < ep.M;
  const long long sample = row_ok ? row / ep.tokens : 0;
  {
      // the residual / pre-activation row segment of the first chunk is fetched before waiting for the accumulator
      uint4 pre[4];
      auto prefetch = [&](int c) {
        const int col = n_blk * BN + c;
        const int nvalid = min(32, ep.N - col);
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.resid) + (long long)row * ep.ldo + col);
#pragma unroll
        for (int g = 0; g < 4; ++g) pre[g] = (row_ok && g * 8 < nvalid) ? src[g] : make_uint4(0, 0, 0, 0);
      };
      if (reads_resid && half * 32 < BN) prefetch(half * 32);
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
            if (row_ok && col < ep.N) {
              store_row32(ep.out, (long long)row * ep.ldo + col, f0, min(32, ep.N - col), false);
              if (col + 32 < ep.N) store_row32(ep.out, (long long)row * ep.ldo + col + 32, f1, min(32, ep.N - col - 32), false);
            }
          }
        }
      } else {
        for (int c = half * 32; c < BN; c += 64) {
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
            if (c + 64 < BN) prefetch(c + 64);  // next chunk's residual while this one is processed
          }
          if (!row_ok || nvalid <= 0) continue;
          const long long off = (long long)row * ep.ldo + col;
          if (ep.epilogue == MAPDIT_EPI_STORE) {
            store_row32(ep.out, off, f, nvalid, ep.out_f32 != 0);
          } else if (ep.epilogue == MAPDIT_EPI_MPSILU) {
            if (ep.out2) store_row32(ep.out2, off, f, nvalid, false);  // pre-activation for the backward
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = silu_fast(f[j]);
            store_row32(ep.out, off, f, nvalid, false);
          } else if (ep.epilogue == MAPDIT_EPI_SILU_BWD) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float sg = __fdividef(1.0f, 1.0f + __expf(-xo[j]));
              f[j] *= sg * fmaf(xo[j], 1.0f - sg, 1.0f) * (1.0f / MP_SILU_DIV);
            }
            store_row32(ep.out, off, f, nvalid, false);
          } else {  // RESID / RESID_MOD
            float gt[32];
            if (ep.aux) store_row32(ep.aux, off, f, nvalid, false);  // raw branch output, needed for d(gate)
            load_row32_f32(ep.gate + sample * ep.ldmod + col, gt, nvalid);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaf(MP_RES_T, gt[j] * f[j] - xo[j], xo[j]) * (1.0f / MP_RES_DEN);
            store_row32(ep.out, off, f, nvalid, false);
            if (ep.epilogue == MAPDIT_EPI_RESID_MOD) {
              load_row32_f32(ep.shift + sample * ep.ldmod + col, xo, nvalid);
              load_row32_f32(ep.scale + sample * ep.ldmod + col, gt, nvalid);
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = lerp_t(f[j] * gt[j], xo[j], gsc) * inv_den;
              store_row32(ep.out2, off, f, nvalid, false);
            }
          }
        }
      }
  }
}
}  // namespace gemm_epi

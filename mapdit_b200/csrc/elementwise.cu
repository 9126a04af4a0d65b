// Elementwise / row-wise kernels of the MaP-DiT hot path (HBM-bound; vectorised where it matters).
// Each entry point cites the reference code it replaces in include/mapdit.h.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// error / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void mapdit_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void mapdit_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static thread_local int g_variant = 0;
int mapdit_variant() { return g_variant; }
extern "C" int mapdit_set_variant(int flags) {
  int old = g_variant;
  g_variant = flags;
  return old;
}

extern "C" const char* mapdit_last_error(void) { return g_err; }
extern "C" int mapdit_abi_version(void) { return 1; }
extern "C" int mapdit_sizeof_gemm_args(void) { return (int)sizeof(mapdit_gemm_args); }
extern "C" int64_t mapdit_launch_count(void) { return (int64_t)g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// K1 weight normalisation.  One CTA per row.  The row is re-read from L1/L2 in each pass
// (rows are <= 18 KB), so DRAM traffic stays at 4 B read + outputs per element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(float* __restrict__ w, int cols, float eps, int force,
                                                              float* __restrict__ eff_f32, bf16* __restrict__ eff_bf16,
                                                              bf16* __restrict__ eff_bf16_t, float* __restrict__ inv_norm,
                                                              int rows, int64_t ld_t) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  float* row = w + (size_t)r * cols;
  float ss = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    float v = row[i];
    ss = __fmaf_rn(v, v, ss);
  }
  float nrm = sqrtf(block_sum(ss, red));
  float inv;
  if (force < 0) {  // use_weight_normalization=False: the effective weight is the raw weight
    inv = 1.0f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float e = row[i];
      if (eff_f32) eff_f32[(size_t)r * cols + i] = e;
      if (eff_bf16) eff_bf16[(size_t)r * cols + i] = __float2bfloat16_rn(e);
      if (eff_bf16_t) eff_bf16_t[(size_t)i * ld_t + r] = __float2bfloat16_rn(e);
    }
  } else if (force) {
    // w <- w*sqrt(n)/(||w||+eps)  (src/utils.py:19-23), then eff is computed from the forced row
    const float sq = sqrtf((float)cols);
    const float den = nrm + eps;
    float ss2 = 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float v = (row[i] * sq) / den;
      ss2 = __fmaf_rn(v, v, ss2);
    }
    float nrm2 = sqrtf(block_sum(ss2, red));
    inv = 1.0f / (nrm2 + eps);
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float v = (row[i] * sq) / den;
      row[i] = v;
      float e = v * inv;
      if (eff_f32) eff_f32[(size_t)r * cols + i] = e;
      if (eff_bf16) eff_bf16[(size_t)r * cols + i] = __float2bfloat16_rn(e);
      if (eff_bf16_t) eff_bf16_t[(size_t)i * ld_t + r] = __float2bfloat16_rn(e);
    }
  } else {
    inv = 1.0f / (nrm + eps);
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float e = row[i] * inv;
      if (eff_f32) eff_f32[(size_t)r * cols + i] = e;
      if (eff_bf16) eff_bf16[(size_t)r * cols + i] = __float2bfloat16_rn(e);
      if (eff_bf16_t) eff_bf16_t[(size_t)i * ld_t + r] = __float2bfloat16_rn(e);
    }
  }
  if (inv_norm && threadIdx.x == 0) inv_norm[r] = inv;
}

// Tiled flavour for the big block weights (cols % 4 == 0), two launches:
//  (1) wn_norms_kernel: one warp per row, float4 loads, 4 loads in flight per lane -> {den, inv} per row into a scratch;
//  (2) wn_apply_kernel: one CTA per 64x64 tile: float4 write-back / fp32 copy, 8-byte bf16 stores, and the transposed
//      bf16 copy (the dgrad operand) through a shared-memory transpose so that it leaves as 32-byte row segments instead
//      of scattered 2-byte stores.
constexpr int WN_ROWS = 64, WN_TCOLS = 64, WN_PITCH = WN_TCOLS + 2;
__device__ __forceinline__ void wn_norms_body(const float* __restrict__ w, int rows, int cols, float eps, int force,
                                              float2* __restrict__ scratch, float* __restrict__ inv_norm, int group) {
  const int r = group * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float4* row4 = reinterpret_cast<const float4*>(w + (size_t)r * cols);
  const int n4 = cols / 4;
  const float sq = sqrtf((float)cols);
  float ss = 0.f;
#pragma unroll 4
  for (int i = lane; i < n4; i += 32) {
    const float4 v = row4[i];
    ss = __fmaf_rn(v.x, v.x, ss);
    ss = __fmaf_rn(v.y, v.y, ss);
    ss = __fmaf_rn(v.z, v.z, ss);
    ss = __fmaf_rn(v.w, v.w, ss);
  }
  const float nrm = sqrtf(warp_sum(ss));
  float den = 1.f, inv;
  if (force < 0) {
    inv = 1.0f;
  } else if (force) {
    den = nrm + eps;
    float ss2 = 0.f;
#pragma unroll 4
    for (int i = lane; i < n4; i += 32) {
      const float4 v = row4[i];
      const float a = (v.x * sq) / den, b = (v.y * sq) / den, c = (v.z * sq) / den, d = (v.w * sq) / den;
      ss2 = __fmaf_rn(a, a, ss2);
      ss2 = __fmaf_rn(b, b, ss2);
      ss2 = __fmaf_rn(c, c, ss2);
      ss2 = __fmaf_rn(d, d, ss2);
    }
    inv = 1.0f / (sqrtf(warp_sum(ss2)) + eps);
  } else {
    inv = 1.0f / (nrm + eps);
  }
  if (lane == 0) {
    scratch[r] = make_float2(den, inv);
    if (inv_norm) inv_norm[r] = inv;
  }
}
__global__ void __launch_bounds__(256) wn_norms_kernel(const float* __restrict__ w, int rows, int cols, float eps, int force,
                                                       float2* __restrict__ scratch, float* __restrict__ inv_norm) {
  wn_norms_body(w, rows, cols, eps, force, scratch, inv_norm, blockIdx.x);
}

__device__ __forceinline__ void wn_apply_body(float* __restrict__ w, int rows, int cols, int force, const float2* __restrict__ scratch,
                                              float* __restrict__ eff_f32, bf16* __restrict__ eff_bf16, bf16* __restrict__ eff_bf16_t,
                                              int64_t ld_t, int tile_r, int tile_c, bf16* tile) {
  const int r0 = tile_r * WN_ROWS, c0 = tile_c * WN_TCOLS;
  const float sq = sqrtf((float)cols);
  const int tr = threadIdx.x >> 4, tc = (threadIdx.x & 15) * 4;  // 16 rows x 16 float4 per pass, 4 passes
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = tr + 16 * k, r = r0 + rr, c = c0 + tc;
    if (r < rows && c < cols) {
      const size_t off = (size_t)r * cols + c;
      const float2 di = scratch[r];
      float4 v = *reinterpret_cast<const float4*>(w + off);
      if (force > 0) {
        v.x = (v.x * sq) / di.x;
        v.y = (v.y * sq) / di.x;
        v.z = (v.z * sq) / di.x;
        v.w = (v.w * sq) / di.x;
        *reinterpret_cast<float4*>(w + off) = v;
      }
      const float4 e = make_float4(v.x * di.y, v.y * di.y, v.z * di.y, v.w * di.y);
      if (eff_f32) *reinterpret_cast<float4*>(eff_f32 + off) = e;
      const __nv_bfloat162 lo = __floats2bfloat162_rn(e.x, e.y), hi = __floats2bfloat162_rn(e.z, e.w);
      if (eff_bf16) {
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&lo);
        u.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(eff_bf16 + off) = u;
      }
      if (eff_bf16_t) {
        __nv_bfloat162* t2 = reinterpret_cast<__nv_bfloat162*>(tile + rr * WN_PITCH + tc);
        t2[0] = lo;
        t2[1] = hi;
      }
    }
  }
  if (eff_bf16_t) {
    __syncthreads();
    const bool t_vec = r0 + WN_ROWS <= rows && (ld_t % 8) == 0 && ((reinterpret_cast<uintptr_t>(eff_bf16_t) & 15) == 0);
    const int c = threadIdx.x >> 2, seg = (threadIdx.x & 3) * 16;  // output row c0+c, source rows seg..seg+15
    if (c0 + c < cols) {
      bf16* dst = eff_bf16_t + (size_t)(c0 + c) * ld_t + r0 + seg;
      if (t_vec) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 b2;
          b2.x = tile[(seg + 2 * j) * WN_PITCH + c];
          b2.y = tile[(seg + 2 * j + 1) * WN_PITCH + c];
          pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else {
        for (int j = 0; j < 16; ++j)
          if (r0 + seg + j < rows) dst[j] = tile[(seg + j) * WN_PITCH + c];
      }
    }
  }
}
__global__ void __launch_bounds__(256) wn_apply_kernel(float* __restrict__ w, int rows, int cols, int force,
                                                       const float2* __restrict__ scratch, float* __restrict__ eff_f32,
                                                       bf16* __restrict__ eff_bf16, bf16* __restrict__ eff_bf16_t, int64_t ld_t) {
  __shared__ __align__(16) bf16 tile[WN_ROWS * WN_PITCH];
  wn_apply_body(w, rows, cols, force, scratch, eff_f32, eff_bf16, eff_bf16_t, ld_t, blockIdx.x, blockIdx.y, tile);
}

// Multi-tensor flavour: every tiled-eligible weight of the model in TWO launches (the training step re-normalises 60+
// matrices; one launch pair per matrix left the kernels latency bound at ~1.1 TB/s).  `descs` is a device table of
// 10 x int64 rows {w, eff_f32, eff_bf16, eff_bf16_t, ld_t, rows, cols, group0, tile0, row0}: group0 / tile0 are this
// tensor's first CTA index in the norms / apply grids, row0 its first row in the {den, inv} scratch.
struct WnDesc {
  float* w;
  float* eff_f32;
  bf16* eff_bf16;
  bf16* eff_bf16_t;
  long long ld_t, rows, cols, group0, tile0, row0;
};
__device__ __forceinline__ int wn_find(const WnDesc* __restrict__ d, int n, long long idx, bool tiles) {
  int lo = 0, hi = n - 1;  // last descriptor whose first CTA index is <= idx
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((tiles ? d[mid].tile0 : d[mid].group0) <= idx) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}
__global__ void __launch_bounds__(256) wn_norms_multi_kernel(const WnDesc* __restrict__ descs, int n, float eps, int force,
                                                             float2* __restrict__ scratch) {
  const WnDesc d = descs[wn_find(descs, n, blockIdx.x, false)];
  wn_norms_body(d.w, (int)d.rows, (int)d.cols, eps, force, scratch + d.row0, nullptr, (int)(blockIdx.x - d.group0));
}
__global__ void __launch_bounds__(256) wn_apply_multi_kernel(const WnDesc* __restrict__ descs, int n, int force,
                                                             const float2* __restrict__ scratch) {
  __shared__ __align__(16) bf16 tile[WN_ROWS * WN_PITCH];
  const WnDesc d = descs[wn_find(descs, n, blockIdx.x, true)];
  const int t = (int)(blockIdx.x - d.tile0), tiles_c = ((int)d.cols + WN_TCOLS - 1) / WN_TCOLS;
  wn_apply_body(d.w, (int)d.rows, (int)d.cols, force, scratch + d.row0, d.eff_f32, d.eff_bf16, d.eff_bf16_t, d.ld_t, t / tiles_c,
                t % tiles_c, tile);
}
extern "C" int mapdit_weight_norm_fwd_multi(const void* descs, int n_desc, int total_groups, int total_tiles, float eps, int force,
                                            void* scratch, void* stream) {
  MAPDIT_REQUIRE(descs && scratch && n_desc > 0 && total_groups > 0 && total_tiles > 0, "weight_norm_fwd_multi: bad args");
  wn_norms_multi_kernel<<<total_groups, 256, 0, (cudaStream_t)stream>>>((const WnDesc*)descs, n_desc, eps, force, (float2*)scratch);
  MAPDIT_LAUNCH_CHECK("weight_norm_fwd_multi(norms)");
  wn_apply_multi_kernel<<<total_tiles, 256, 0, (cudaStream_t)stream>>>((const WnDesc*)descs, n_desc, force, (const float2*)scratch);
  MAPDIT_LAUNCH_CHECK("weight_norm_fwd_multi(apply)");
  return MAPDIT_OK;
}

// per-row {den, inv} scratch of the tiled path: grown on demand outside stream capture, one per device
static float2* wn_scratch(int rows, cudaStream_t stream) {
  static float2* buf[16] = {};
  static int cap[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return nullptr;
  if (cap[dev] >= rows) return buf[dev];
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
  const int want = rows < 65536 ? 65536 : rows;
  float2* p = nullptr;
  if (cudaMalloc(&p, (size_t)want * sizeof(float2)) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  // the old buffer (if any) may still be in use by queued kernels: leak it (happens at most a few times per process)
  buf[dev] = p;
  cap[dev] = want;
  return p;
}

extern "C" int mapdit_weight_norm_fwd(float* w, int rows, int cols, float eps, int force, float* eff_f32, void* eff_bf16,
                                      void* eff_bf16_t, int64_t ld_t, float* inv_norm, void* stream) {
  MAPDIT_REQUIRE(w && rows > 0 && cols > 0, "weight_norm_fwd: bad args");
  const bool aligned = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(eff_f32)) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(eff_bf16) & 7) == 0;
  float2* scratch = (cols % 4 == 0 && rows >= 16 && aligned) ? wn_scratch(rows, (cudaStream_t)stream) : nullptr;
  if (scratch) {
    wn_norms_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, eps, force, scratch, inv_norm);
    MAPDIT_LAUNCH_CHECK("weight_norm_fwd(norms)");
    if (force > 0 || eff_f32 || eff_bf16 || eff_bf16_t) {
      dim3 grid((rows + WN_ROWS - 1) / WN_ROWS, (cols + WN_TCOLS - 1) / WN_TCOLS);
      wn_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, force, scratch, eff_f32, (bf16*)eff_bf16, (bf16*)eff_bf16_t,
                                                             ld_t > 0 ? ld_t : rows);
      MAPDIT_LAUNCH_CHECK("weight_norm_fwd(apply)");
    }
    return MAPDIT_OK;
  }
  weight_norm_fwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(w, cols, eps, force, eff_f32, (bf16*)eff_bf16,
                                                                  (bf16*)eff_bf16_t, inv_norm, rows, ld_t > 0 ? ld_t : rows);
  MAPDIT_LAUNCH_CHECK("weight_norm_fwd");
  return MAPDIT_OK;
}

// grad_v = (G - v (v.G)/(r (r+eps)))/(r+eps)     (autograd of src/utils.py:19-23 / sqrt(n))
__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              float* __restrict__ gv, int cols, float eps, int accumulate) {
  __shared__ float red[32];
  const size_t off = (size_t)blockIdx.x * cols;
  float ss = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    float a = v[off + i];
    ss = __fmaf_rn(a, a, ss);
    dot = __fmaf_rn(a, g[off + i], dot);
  }
  float r = sqrtf(block_sum(ss, red));
  dot = block_sum(dot, red);
  float inv = 1.0f / (r + eps);
  float coef = dot / (fmaxf(r, 1e-30f) * (r + eps));
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    float val = (g[off + i] - v[off + i] * coef) * inv;
    gv[off + i] = accumulate ? gv[off + i] + val : val;
  }
}

// float4 flavour, one warp per row (cols % 4 == 0)
__global__ void __launch_bounds__(256) weight_norm_bwd_vec_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                                  float* __restrict__ gv, int rows, int cols, float eps,
                                                                  int accumulate) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const size_t off = (size_t)r * cols;
  const float4* v4 = reinterpret_cast<const float4*>(v + off);
  const float4* g4 = reinterpret_cast<const float4*>(g + off);
  float4* o4 = reinterpret_cast<float4*>(gv + off);
  float ss = 0.f, dot = 0.f;
#pragma unroll 4
  for (int i = lane; i < cols / 4; i += 32) {
    const float4 a = v4[i], b = g4[i];
    ss = __fmaf_rn(a.x, a.x, ss);
    ss = __fmaf_rn(a.y, a.y, ss);
    ss = __fmaf_rn(a.z, a.z, ss);
    ss = __fmaf_rn(a.w, a.w, ss);
    dot = __fmaf_rn(a.x, b.x, dot);
    dot = __fmaf_rn(a.y, b.y, dot);
    dot = __fmaf_rn(a.z, b.z, dot);
    dot = __fmaf_rn(a.w, b.w, dot);
  }
  const float rn = sqrtf(warp_sum(ss));
  dot = warp_sum(dot);
  const float inv = 1.0f / (rn + eps);
  const float coef = dot / (fmaxf(rn, 1e-30f) * (rn + eps));
#pragma unroll 4
  for (int i = lane; i < cols / 4; i += 32) {
    const float4 a = v4[i], b = g4[i];
    float4 o = make_float4((b.x - a.x * coef) * inv, (b.y - a.y * coef) * inv, (b.z - a.z * coef) * inv, (b.w - a.w * coef) * inv);
    if (accumulate) {
      const float4 p = o4[i];
      o.x += p.x;
      o.y += p.y;
      o.z += p.z;
      o.w += p.w;
    }
    o4[i] = o;
  }
}

extern "C" int mapdit_weight_norm_bwd(const float* v, const float* g_eff, float* grad_v, int rows, int cols, float eps,
                                      int accumulate, void* stream) {
  MAPDIT_REQUIRE(v && g_eff && grad_v && rows > 0 && cols > 0, "weight_norm_bwd: bad args");
  const bool al16 = ((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(g_eff) | reinterpret_cast<uintptr_t>(grad_v)) & 15) == 0;
  if (cols % 4 == 0 && cols >= 128 && al16) {
    weight_norm_bwd_vec_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(v, g_eff, grad_v, rows, cols, eps, accumulate);
    MAPDIT_LAUNCH_CHECK("weight_norm_bwd");
    return MAPDIT_OK;
  }
  weight_norm_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(v, g_eff, grad_v, cols, eps, accumulate);
  MAPDIT_LAUNCH_CHECK("weight_norm_bwd");
  return MAPDIT_OK;
}

// Multi-tensor, IN-PLACE flavour: the weight-gradient GEMMs write d(effective weight) straight into the gradient buffers and one
// launch per block turns all of them into d(raw weight) (was one launch per weight: 66 per DiT-B/2 step).  One warp per row; a
// row is read twice (dot products, then the projection) and each element is overwritten by the thread that read it.
struct WnBwdItem {
  const float* v;
  float* g;
  long long rows, cols, group0;  // group0 = first 8-row group of this tensor in the launch
};
__global__ void __launch_bounds__(256) weight_norm_bwd_multi_kernel(const WnBwdItem* __restrict__ table, int n_items, float eps) {
  int it = 0;
  while (it + 1 < n_items && (long long)blockIdx.x >= table[it + 1].group0) ++it;
  const WnBwdItem t = table[it];
  const long long r = ((long long)blockIdx.x - t.group0) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= t.rows) return;
  const int cols = (int)t.cols;
  const float4* v4 = reinterpret_cast<const float4*>(t.v + r * cols);
  float4* g4 = reinterpret_cast<float4*>(t.g + r * cols);
  float ss = 0.f, dot = 0.f;
#pragma unroll 4
  for (int i = lane; i < cols / 4; i += 32) {
    const float4 a = v4[i], b = g4[i];
    ss = __fmaf_rn(a.x, a.x, ss);
    ss = __fmaf_rn(a.y, a.y, ss);
    ss = __fmaf_rn(a.z, a.z, ss);
    ss = __fmaf_rn(a.w, a.w, ss);
    dot = __fmaf_rn(a.x, b.x, dot);
    dot = __fmaf_rn(a.y, b.y, dot);
    dot = __fmaf_rn(a.z, b.z, dot);
    dot = __fmaf_rn(a.w, b.w, dot);
  }
  const float rn = sqrtf(warp_sum(ss));
  dot = warp_sum(dot);
  const float inv = 1.0f / (rn + eps);
  const float coef = dot / (fmaxf(rn, 1e-30f) * (rn + eps));
#pragma unroll 4
  for (int i = lane; i < cols / 4; i += 32) {
    const float4 a = v4[i], b = g4[i];
    g4[i] = make_float4((b.x - a.x * coef) * inv, (b.y - a.y * coef) * inv, (b.z - a.z * coef) * inv, (b.w - a.w * coef) * inv);
  }
}
extern "C" int mapdit_weight_norm_bwd_multi(const void* table, int n_items, int n_groups, float eps, void* stream) {
  MAPDIT_REQUIRE(table && n_items > 0 && n_groups > 0, "weight_norm_bwd_multi: bad args");
  weight_norm_bwd_multi_kernel<<<n_groups, 256, 0, (cudaStream_t)stream>>>((const WnBwdItem*)table, n_items, eps);
  MAPDIT_LAUNCH_CHECK("weight_norm_bwd_multi");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// K3 standalone: modulate, residual, mp_silu, qk-normalise, casts
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void modulate_kernel(const T* __restrict__ x, T* __restrict__ h, const float* __restrict__ shift,
                                const float* __restrict__ scale, const float* __restrict__ gain, int64_t ldmod, int64_t total,
                                int d, int tokens) {
  const float g = *gain;
  const float den = mod_den(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / d;
    int col = (int)(i - row * d);
    int64_t n = row / tokens;
    st_act(h + i, modulate_f(ld_act(x + i), shift[n * ldmod + col], scale[n * ldmod + col], g, den));
  }
}

extern "C" int mapdit_modulate_fwd(const void* x, void* h, const float* shift, const float* scale, const float* gain,
                                   int64_t ldmod, int m, int d, int tokens, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && h && shift && scale && gain && m > 0 && d > 0 && tokens > 0, "modulate_fwd: bad args");
  int64_t total = (int64_t)m * d;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == MAPDIT_F32)
    modulate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)h, shift, scale, gain, ldmod, total, d, tokens);
  else
    modulate_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)h, shift, scale, gain, ldmod, total, d, tokens);
  MAPDIT_LAUNCH_CHECK("modulate_fwd");
  return MAPDIT_OK;
}

template <typename T>
__global__ void resid_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ xo, const float* __restrict__ gate,
                             int64_t ldmod, int64_t total, int d, int tokens, int var) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / d;
    int col = (int)(i - row * d);
    int64_t n = row / tokens;
    st_act(xo + i, resid_v(ld_act(x + i), gate[n * ldmod + col], ld_act(y + i), var));
  }
}

extern "C" int mapdit_resid_fwd(const void* x, const void* y, void* xout, const float* gate, int64_t ldmod, int m, int d,
                                int tokens, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && y && xout && gate && m > 0 && d > 0 && tokens > 0, "resid_fwd: bad args");
  int64_t total = (int64_t)m * d;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == MAPDIT_F32)
    resid_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)y, (float*)xout, gate, ldmod, total, d, tokens, mapdit_variant());
  else
    resid_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)y, (bf16*)xout, gate, ldmod, total, d, tokens, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("resid_fwd");
  return MAPDIT_OK;
}

template <typename TI, typename TO>
__global__ void mp_silu_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n, int var) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_act(y + i, mp_silu_v(ld_act(x + i), var));
}
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_act(y + i, ld_act(x + i));
}

static inline int ew_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

extern "C" int mapdit_mp_silu_fwd(const void* x, void* y, int64_t n, int in_dtype, int out_dtype, void* stream) {
  MAPDIT_REQUIRE(x && y && n > 0, "mp_silu_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int grid = ew_grid(n);
  if (in_dtype == MAPDIT_F32 && out_dtype == MAPDIT_F32) mp_silu_kernel<float, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, n, mapdit_variant());
  else if (in_dtype == MAPDIT_F32) mp_silu_kernel<float, bf16><<<grid, 256, 0, s>>>((const float*)x, (bf16*)y, n, mapdit_variant());
  else if (out_dtype == MAPDIT_F32) mp_silu_kernel<bf16, float><<<grid, 256, 0, s>>>((const bf16*)x, (float*)y, n, mapdit_variant());
  else mp_silu_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, n, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("mp_silu_fwd");
  return MAPDIT_OK;
}

extern "C" int mapdit_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  MAPDIT_REQUIRE(src && dst && n > 0, "cast: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int grid = ew_grid(n);
  if (src_dtype == MAPDIT_F32 && dst_dtype == MAPDIT_BF16) cast_kernel<float, bf16><<<grid, 256, 0, s>>>((const float*)src, (bf16*)dst, n);
  else if (src_dtype == MAPDIT_BF16 && dst_dtype == MAPDIT_F32) cast_kernel<bf16, float><<<grid, 256, 0, s>>>((const bf16*)src, (float*)dst, n);
  else if (src_dtype == MAPDIT_F32) cast_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, (float*)dst, n);
  else cast_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)src, (bf16*)dst, n);
  MAPDIT_LAUNCH_CHECK("cast");
  return MAPDIT_OK;
}

// strided 2-D cast (a column slice of a [rows, ld] matrix): the per-block slices of the modulation-vector gradients
template <typename TI, typename TO>
__global__ void cast2d_kernel(const TI* __restrict__ x, int64_t ldx, TO* __restrict__ y, int64_t ldy, int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    st_act(y + r * ldy + c, ld_act(x + r * ldx + c));
  }
}
extern "C" int mapdit_cast_2d(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows, int cols, int src_dtype,
                              int dst_dtype, void* stream) {
  MAPDIT_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, "cast_2d: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = ew_grid((int64_t)rows * cols);
  if (src_dtype == MAPDIT_F32 && dst_dtype == MAPDIT_BF16)
    cast2d_kernel<float, bf16><<<grid, 256, 0, s>>>((const float*)src, ld_src, (bf16*)dst, ld_dst, rows, cols);
  else if (src_dtype == MAPDIT_BF16 && dst_dtype == MAPDIT_F32)
    cast2d_kernel<bf16, float><<<grid, 256, 0, s>>>((const bf16*)src, ld_src, (float*)dst, ld_dst, rows, cols);
  else if (src_dtype == MAPDIT_F32)
    cast2d_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, ld_src, (float*)dst, ld_dst, rows, cols);
  else
    cast2d_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)src, ld_src, (bf16*)dst, ld_dst, rows, cols);
  MAPDIT_LAUNCH_CHECK("cast_2d");
  return MAPDIT_OK;
}

// one warp per (row, head) of the q and k thirds of qkv[M, 3D]
template <typename T>
__global__ void qk_normalize_kernel(T* __restrict__ qkv, int64_t n_heads_total, int heads2, int d, int hd, float eps) {
  int lane = threadIdx.x & 31;
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= n_heads_total) return;
  int64_t row = wid / heads2;
  int hh = (int)(wid - row * heads2);  // 0..2H-1 : q heads then k heads (contiguous columns [0, 2D))
  T* p = qkv + row * (int64_t)(3 * d) + (int64_t)hh * hd;
  float v[4];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    v[j] = (c < hd) ? ld_act(p + c) : 0.f;
    ss = __fmaf_rn(v[j], v[j], ss);
  }
  float nrm = sqrtf(warp_sum(ss));
  float sq = sqrtf((float)hd), den = nrm + eps;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    if (c < hd) st_act(p + c, (v[j] * sq) / den);
  }
}

int mapdit_qk_normalize_vec(void* qkv, float* sc, int m, int d, int head_dim, float eps, void* stream);  // backward.cu
extern "C" int mapdit_qk_normalize(void* qkv, int m, int d, int head_dim, float eps, int dtype, void* stream) {
  MAPDIT_REQUIRE(qkv && m > 0 && d > 0 && head_dim > 0 && head_dim <= 128 && d % head_dim == 0, "qk_normalize: bad args");
  if (dtype == MAPDIT_BF16 && head_dim % 8 == 0 && d % 8 == 0 && ((uintptr_t)qkv & 15) == 0) {  // 16-byte chunks, 16 lanes per head
    mapdit_qk_normalize_vec(qkv, nullptr, m, d, head_dim, eps, stream);
    MAPDIT_LAUNCH_CHECK("qk_normalize(vec)");
    return MAPDIT_OK;
  }
  int heads2 = 2 * (d / head_dim);
  int64_t total = (int64_t)m * heads2;
  int64_t blocks = (total * 32 + 255) / 256;
  if (dtype == MAPDIT_F32)
    qk_normalize_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)qkv, total, heads2, d, head_dim, eps);
  else
    qk_normalize_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)qkv, total, heads2, d, head_dim, eps);
  MAPDIT_LAUNCH_CHECK("qk_normalize");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// patch embed: x0 = mp_sum(patchify(x)|1 · Wx_eff^T, pos, 0.5)  (+ optional modulate for block 0)
// CTA = 16 tokens x 256 output channels; patches staged in smem.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float round_act(float v, float*) { return v; }
__device__ __forceinline__ float round_act(float v, bf16*) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void st_act4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st_act4(bf16* p, const float (&v)[4]) {
  uint2 u;
  *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v[0], v[1]);
  *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

// Thread = 4 tokens x 4 consecutive channels (64 channel groups x 4 token groups per CTA): a warp stores 256 contiguous bytes per
// instruction (the former thread-per-channel layout stored 2 bytes per thread and re-read x0 to see its rounding: 0.21 ms per
// forward for 200 MB of output).  d % 4 == 0 is required for the vector path (all registry models); otherwise VEC = false.
constexpr int PE_GROUPS = 8;  // token groups (of 16) per CTA on the vector path

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ wx,
                                                          const float* __restrict__ pos, T* __restrict__ x0, T* __restrict__ h,
                                                          const float* __restrict__ shift, const float* __restrict__ scale,
                                                          const float* __restrict__ gain, int64_t ldmod, int64_t m_total, int C,
                                                          int S, int p, int d, int var) {
  extern __shared__ float sp[];  // [16][K1] patches of one token group, then (VEC) [K1][256] weights of this CTA's channel slab
  const int g = S / p, T_ = g * g, K = p * p * C, K1 = K + 1;
  float gn = 0.f, den = 1.f;
  if (h) { gn = *gain; den = mod_den(gn); }
  // patches of 16 consecutive tokens -> sp.  All index arithmetic in 32 bits (a 64-bit division costs ~100 instructions and the
  // first version did one per element and per output token: more instructions than the FMAs)
  auto gather = [&](int64_t tok0) {
    for (int i = threadIdx.x; i < 16 * K1; i += blockDim.x) {
      const int tl = i / K1, f = i - tl * K1;
      const int64_t tok = tok0 + tl;
      float v = 0.f;
      if (tok < m_total) {
        if (f == K) v = 1.0f;  // bias column (src/dit.py:82)
        else {
          const unsigned tk = (unsigned)tok;
          const unsigned n = tk / (unsigned)T_, tt = tk - n * (unsigned)T_;
          const unsigned hh = tt / (unsigned)g, ww = tt - hh * (unsigned)g;
          const unsigned pp = (unsigned)f / (unsigned)C, c = (unsigned)f - pp * (unsigned)C, p1 = pp / (unsigned)p, p2 = pp - p1 * (unsigned)p;
          v = x[(((size_t)n * C + c) * S + (hh * p + p1)) * (size_t)S + (ww * p + p2)];
        }
      }
      sp[i] = v;
    }
  };
  if (VEC) {
    // CTA = PE_GROUPS x 16 tokens x 256 channels; the slab's weights are staged once as [k][channel] (one 16-byte shared-memory
    // load per k per thread; per-thread strided global weight loads left the kernel latency bound at 0.21 ms per forward)
    float* sw = sp + 16 * K1;
    const int c0 = blockIdx.y * 256;
    for (int i = threadIdx.x; i < 256 * K1; i += blockDim.x) {
      const int cl = i / K1, k = i - cl * K1;
      sw[k * 256 + cl] = (c0 + cl < d) ? wx[(size_t)(c0 + cl) * K1 + k] : 0.f;
    }
    const int cl = (threadIdx.x & 63) * 4, col = c0 + cl;
    const int j0 = (threadIdx.x >> 6) * 4;  // first of this thread's 4 tokens of a group
    for (int grp = 0; grp < PE_GROUPS; ++grp) {
      const int64_t tok0 = ((int64_t)blockIdx.x * PE_GROUPS + grp) * 16;
      if (tok0 >= m_total) break;
      __syncthreads();  // previous group's patches consumed (and, first trip, nothing)
      gather(tok0);
      __syncthreads();
      if (col >= d) continue;
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
      for (int k = 0; k < K1; ++k) {  // same k order per output as the scalar path
        const float4 w4 = *reinterpret_cast<const float4*>(sw + k * 256 + cl);
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = sp[(j0 + j) * K1 + k];
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[j][c] = __fmaf_rn(a, w[c], acc[j][c]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t tok = tok0 + j0 + j;
        if (tok >= m_total) break;
        const unsigned n = (unsigned)tok / (unsigned)T_;
        const unsigned tt = (unsigned)tok - n * (unsigned)T_;
        const float4 pv = *reinterpret_cast<const float4*>(pos + (size_t)tt * d + col);
        const float pvv[4] = {pv.x, pv.y, pv.z, pv.w};
        float v[4], hv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (var & MAPDIT_VAR_PLAIN_POS) v[c] = acc[j][c] + pvv[c];
          else if (sizeof(T) == 2) v[c] = lerp_t(acc[j][c], pvv[c], 0.5f) * (1.0f / MP_HALF_DEN);  // bf16 output: rounding dominates
          else v[c] = lerp_t(acc[j][c], pvv[c], 0.5f) / MP_HALF_DEN;
        }
        st_act4(x0 + tok * d + col, v);
        if (h) {  // modulate sees the value as stored in the residual stream
#pragma unroll
          for (int c = 0; c < 4; ++c)
            hv[c] = modulate_f(round_act(v[c], (T*)nullptr), shift[n * ldmod + col + c], scale[n * ldmod + col + c], gn, den);
          st_act4(h + tok * d + col, hv);
        }
      }
    }
    return;
  }
  const int64_t tok0 = (int64_t)blockIdx.x * 16;
  gather(tok0);
  __syncthreads();
  const int col = blockIdx.y * 256 + threadIdx.x;
  if (col >= d) return;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const float* wr = wx + (size_t)col * K1;
  for (int k = 0; k < K1; ++k) {
    float w = wr[k];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = __fmaf_rn(sp[j * K1 + k], w, acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    int64_t tok = tok0 + j;
    if (tok >= m_total) break;
    int64_t n = tok / T_;
    int tt = (int)(tok - n * T_);
    const float pv = pos[(size_t)tt * d + col];
    float v = (var & MAPDIT_VAR_PLAIN_POS) ? acc[j] + pv : lerp_t(acc[j], pv, 0.5f) / MP_HALF_DEN;
    st_act(x0 + tok * d + col, v);
    if (h) st_act(h + tok * d + col, modulate_f(round_act(v, (T*)nullptr), shift[n * ldmod + col], scale[n * ldmod + col], gn, den));
  }
}

extern "C" int mapdit_patch_embed(const float* x, const float* wx_eff, const float* pos, void* x0, void* h, const float* shift,
                                  const float* scale, const float* gain, int64_t ldmod, int n_samples, int channels,
                                  int input_size, int patch, int d, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && wx_eff && pos && x0 && n_samples > 0 && input_size % patch == 0, "patch_embed: bad args");
  int g = input_size / patch;
  int64_t m_total = (int64_t)n_samples * g * g;
  MAPDIT_REQUIRE(m_total < (1LL << 31), "patch_embed: more than 2^31 tokens");
  int K1 = patch * patch * channels + 1;
  const bool vec = d % 4 == 0 && ((uintptr_t)x0 & 15) == 0 && (!h || ((uintptr_t)h & 15) == 0) && ((uintptr_t)pos & 15) == 0 &&
                   (size_t)(16 + 256) * K1 * sizeof(float) <= 48 * 1024;
  size_t smem = (size_t)(16 + (vec ? 256 : 0)) * K1 * sizeof(float);
  const int64_t groups = (m_total + 15) / 16;
  dim3 grid((unsigned)(vec ? (groups + PE_GROUPS - 1) / PE_GROUPS : groups), (unsigned)((d + 255) / 256));
  const int var = mapdit_variant();
  cudaStream_t s = (cudaStream_t)stream;
#define PE_LAUNCH(TT, V) patch_embed_kernel<TT, V><<<grid, 256, smem, s>>>(x, wx_eff, pos, (TT*)x0, (TT*)h, shift, scale, gain, ldmod, m_total, channels, input_size, patch, d, var)
  if (dtype == MAPDIT_F32) {
    if (vec) PE_LAUNCH(float, true); else PE_LAUNCH(float, false);
  } else {
    if (vec) PE_LAUNCH(bf16, true); else PE_LAUNCH(bf16, false);
  }
#undef PE_LAUNCH
  MAPDIT_LAUNCH_CHECK("patch_embed");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// conditioning path
// ------------------------------------------------------------------------------------------------
__global__ void fourier_kernel(const int64_t* __restrict__ t, const float* __restrict__ scale, const float* __restrict__ shift,
                               float* __restrict__ e, int n, int ch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ch) return;
  int r = i / ch, j = i - r * ch;
  // no FMA contraction: fl(fl(t*scale)+shift), accurate cosf
  float arg = __fadd_rn(__fmul_rn((float)t[r], scale[j]), shift[j]);
  e[i] = 1.4142135623730951f * cosf(arg);
}
extern "C" int mapdit_fourier(const int64_t* t, const float* scale, const float* shift, float* e, int n, int channels, void* stream) {
  MAPDIT_REQUIRE(t && scale && shift && e && n > 0 && channels > 0, "fourier: bad args");
  fourier_kernel<<<(n * channels + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, scale, shift, e, n, channels);
  MAPDIT_LAUNCH_CHECK("fourier");
  return MAPDIT_OK;
}

// Sinusoidal timestep features of the vanilla DiT (use_mp_embedding=False; UNPINNED, Peebles & Xie timestep_embedding):
// e[n, j] = cos(t f_j), e[n, half + j] = sin(t f_j), f_j = exp(-ln(max_period) j / half)
__global__ void timestep_sincos_kernel(const int64_t* __restrict__ t, float* __restrict__ e, int n, int dim, float max_period) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= n * half) return;
  int r = i / half, j = i - r * half;
  const float f = expf(-logf(max_period) * (float)j / (float)half);
  const float arg = __fmul_rn((float)t[r], f);
  e[(size_t)r * dim + j] = cosf(arg);
  e[(size_t)r * dim + half + j] = sinf(arg);
}
extern "C" int mapdit_timestep_sincos(const int64_t* t, float* e, int n, int dim, float max_period, void* stream) {
  MAPDIT_REQUIRE(t && e && n > 0 && dim > 0 && dim % 2 == 0, "timestep_sincos: bad args");
  timestep_sincos_kernel<<<(n * (dim / 2) + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, e, n, dim, max_period);
  MAPDIT_LAUNCH_CHECK("timestep_sincos");
  return MAPDIT_OK;
}

// adaLN with LayerNorm (use_no_layernorm=False; UNPINNED, vanilla DiT block): h = LN(x) (1 + scale) + shift, LN without
// affine, eps 1e-6, biased variance.  One warp per token row; stats[row] = {mean, rstd} (nullable) for the backward.
template <typename T>
__global__ void __launch_bounds__(256) ln_modulate_kernel(const T* __restrict__ x, T* __restrict__ h, const float* __restrict__ shift,
                                                          const float* __restrict__ scale, float2* __restrict__ stats, int64_t ldmod,
                                                          int m, int d, int tokens) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  const T* xr = x + (size_t)row * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += ld_act(xr + i);
  const float mean = warp_sum(s) / (float)d;
  float v = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float c = ld_act(xr + i) - mean;
    v = __fmaf_rn(c, c, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)d + 1e-6f);
  if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
  const int64_t n = row / tokens;
  for (int i = lane; i < d; i += 32) {
    const float xh = (ld_act(xr + i) - mean) * rstd;
    st_act(h + (size_t)row * d + i, __fmaf_rn(xh, 1.0f + scale[n * ldmod + i], shift[n * ldmod + i]));
  }
}
extern "C" int mapdit_ln_modulate_fwd(const void* x, void* h, const float* shift, const float* scale, float* stats, int64_t ldmod,
                                      int m, int d, int tokens, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && h && shift && scale && m > 0 && d > 0 && tokens > 0, "ln_modulate_fwd: bad args");
  const int grid = (m + 7) / 8;
  if (dtype == MAPDIT_F32)
    ln_modulate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)h, shift, scale, (float2*)stats, ldmod, m, d, tokens);
  else
    ln_modulate_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)h, shift, scale, (float2*)stats, ldmod, m, d, tokens);
  MAPDIT_LAUNCH_CHECK("ln_modulate_fwd");
  return MAPDIT_OK;
}

__global__ void __launch_bounds__(256) embed_rows_kernel(const int64_t* __restrict__ idx, const uint8_t* __restrict__ drop,
                                                         int64_t null_idx, const float* __restrict__ table, float* __restrict__ out,
                                                         int d, float eps, int var) {
  __shared__ float red[32];
  int n = blockIdx.x;
  int64_t id = idx[n];
  if (drop && drop[n]) id = null_idx;
  const float* row = table + id * d;
  if (var & MAPDIT_VAR_PLAIN_EMBED) {  // nn.Embedding: plain gather
    for (int i = threadIdx.x; i < d; i += blockDim.x) out[(size_t)n * d + i] = row[i];
    return;
  }
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) ss = __fmaf_rn(row[i], row[i], ss);
  float nrm = sqrtf(block_sum(ss, red));
  float sq = sqrtf((float)d), den = nrm + eps;
  for (int i = threadIdx.x; i < d; i += blockDim.x) out[(size_t)n * d + i] = (row[i] * sq) / den;
}
extern "C" int mapdit_embed_rows(const int64_t* idx, const uint8_t* drop_mask, int64_t null_idx, const float* table, float* out,
                                 int n, int d, float eps, void* stream) {
  MAPDIT_REQUIRE(idx && table && out && n > 0 && d > 0, "embed_rows: bad args");
  embed_rows_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(idx, drop_mask, null_idx, table, out, d, eps, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("embed_rows");
  return MAPDIT_OK;
}

__global__ void cond_combine_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c,
                                    float* __restrict__ cs32, bf16* __restrict__ cs16, int64_t n, int var) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = (var & MAPDIT_VAR_PLAIN_EMBED) ? a[i] + b[i] : lerp_t(a[i], b[i], 0.5f) / MP_HALF_DEN;
  if (c) c[i] = v;
  float s = mp_silu_v(v, var);
  if (cs32) cs32[i] = s;
  if (cs16) cs16[i] = __float2bfloat16_rn(s);
}
extern "C" int mapdit_cond_combine(const float* a, const float* b, float* c, float* cs_f32, void* cs_bf16, int64_t n, void* stream) {
  MAPDIT_REQUIRE(a && b && n > 0, "cond_combine: bad args");
  cond_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, c, cs_f32, (bf16*)cs_bf16, n, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("cond_combine");
  return MAPDIT_OK;
}

// one CTA per sample: adim (<=32) dot products of length d, then sigmoid(sum_j dot_j*ref_j/sqrt(adim))
__global__ void __launch_bounds__(256) mp_scale_kernel(const float* __restrict__ c, const float* __restrict__ w,
                                                       const float* __restrict__ ref, float* __restrict__ s, int d, int adim) {
  __shared__ float red[32];
  int n = blockIdx.x;
  float angle = 0.f;
  for (int j = 0; j < adim; ++j) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) acc = __fmaf_rn(c[(size_t)n * d + i], w[(size_t)j * d + i], acc);
    acc = block_sum(acc, red);
    angle = __fmaf_rn(acc, ref[j], angle);
  }
  if (threadIdx.x == 0) {
    float a = angle / sqrtf((float)adim);
    s[n] = 1.0f / (1.0f + expf(-a));
  }
}
extern "C" int mapdit_mp_scale(const float* c, const float* w_eff, const float* ref, float* s, int n, int d, int adim, void* stream) {
  MAPDIT_REQUIRE(c && w_eff && ref && s && n > 0 && d > 0 && adim > 0, "mp_scale: bad args");
  mp_scale_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(c, w_eff, ref, s, d, adim);
  MAPDIT_LAUNCH_CHECK("mp_scale");
  return MAPDIT_OK;
}

// out[n, part*C + c, hh*p+p1, ww*p+p2] = lin[tok, part*p*p*C + (p1*p+p2)*C + c] * s_part[n]
template <typename T>
__global__ void final_unpatchify_kernel(const T* __restrict__ lin, const float* __restrict__ s_mu, const float* __restrict__ s_sg,
                                        float* __restrict__ out, int64_t total, int C, int S, int p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes the OUTPUT (coalesced writes): [n, 2C, S, S]
  int xx = (int)(i % S);
  int64_t r = i / S;
  int yy = (int)(r % S);
  r /= S;
  int ch = (int)(r % (2 * C));
  int64_t n = r / (2 * C);
  int part = ch / C, c = ch - part * C;
  int g = S / p, hh = yy / p, p1 = yy - hh * p, ww = xx / p, p2 = xx - ww * p;
  int64_t tok = n * (g * g) + hh * g + ww;
  int ppc = p * p * C;
  float v = ld_act(lin + tok * (2 * ppc) + part * ppc + (p1 * p + p2) * C + c);
  out[i] = v * (part ? s_sg[n] : s_mu[n]);
}
extern "C" int mapdit_final_unpatchify(const void* lin, const float* s_mu, const float* s_sigma, float* out, int n_samples,
                                       int channels, int input_size, int patch, int dtype, void* stream) {
  MAPDIT_REQUIRE(lin && s_mu && s_sigma && out && n_samples > 0, "final_unpatchify: bad args");
  int64_t total = (int64_t)n_samples * 2 * channels * input_size * input_size;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == MAPDIT_F32)
    final_unpatchify_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)lin, s_mu, s_sigma, out, total, channels, input_size, patch);
  else
    final_unpatchify_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)lin, s_mu, s_sigma, out, total, channels, input_size, patch);
  MAPDIT_LAUNCH_CHECK("final_unpatchify");
  return MAPDIT_OK;
}

// eps channels of both halves <- uncond + s*(cond - uncond)   (src/dit.py:113-118)
__global__ void cfg_combine_kernel(float* __restrict__ out, int n_half, int C, int hw, float s) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t per = (int64_t)C * hw;
  if (i >= (int64_t)n_half * per) return;
  int64_t n = i / per, r = i - n * per;
  int64_t a = n * (2 * per) + r, b = (n + n_half) * (2 * per) + r;
  float cond = out[a], unc = out[b];
  float v = unc + s * (cond - unc);
  out[a] = v;
  out[b] = v;
}
extern "C" int mapdit_cfg_combine(float* out, int n_half, int channels, int hw, float cfg_scale, void* stream) {
  MAPDIT_REQUIRE(out && n_half > 0, "cfg_combine: bad args");
  int64_t total = (int64_t)n_half * channels * hw;
  cfg_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n_half, channels, hw, cfg_scale);
  MAPDIT_LAUNCH_CHECK("cfg_combine");
  return MAPDIT_OK;
}

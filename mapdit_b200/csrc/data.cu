// "Next" row N3: the data step in front of the training hot path.  The reference's CustomDataset
// (train.py:144-176) samples one latent per item on CPU workers: feature = mean + eps * std, then
// transforms.Normalize(stats.mean, stats.std) per channel.  Here the posterior tables live in HBM (ImageNet-128 latents
// at 32x32x4 fp32 are 2 x 21 GB: they fit next to the model in 180 GB) and a batch is ONE gather + sample + normalise
// kernel: 16 B/element in, 4 B/element out, no host involvement.
#include "common.cuh"

__global__ void __launch_bounds__(256) latent_sample_kernel(const float* __restrict__ means, const float* __restrict__ stds,
                                                            const int64_t* __restrict__ idx, const float* __restrict__ eps,
                                                            const float* __restrict__ ch_mean, const float* __restrict__ ch_std,
                                                            float* __restrict__ out, int64_t total4, int chw4, int hw4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / chw4;
    const int r = (int)(i - n * chw4);
    const int c = r / hw4;
    const int64_t src = idx[n] * chw4 + r;
    const float4 m = reinterpret_cast<const float4*>(means)[src], s = reinterpret_cast<const float4*>(stds)[src];
    const float4 e = reinterpret_cast<const float4*>(eps)[i];
    const float cm = ch_mean[c], cs = ch_std[c];
    float4 o;
    // mean + eps*std with torch's separate mul and add (no FMA contraction), then (x - mean_c) / std_c
    o.x = __fdiv_rn(__fsub_rn(__fadd_rn(m.x, __fmul_rn(e.x, s.x)), cm), cs);
    o.y = __fdiv_rn(__fsub_rn(__fadd_rn(m.y, __fmul_rn(e.y, s.y)), cm), cs);
    o.z = __fdiv_rn(__fsub_rn(__fadd_rn(m.z, __fmul_rn(e.z, s.z)), cm), cs);
    o.w = __fdiv_rn(__fsub_rn(__fadd_rn(m.w, __fmul_rn(e.w, s.w)), cm), cs);
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

extern "C" int mapdit_latent_sample(const float* means, const float* stds, const int64_t* idx, const float* eps, const float* ch_mean,
                                    const float* ch_std, float* out, int n, int channels, int hw, void* stream) {
  MAPDIT_REQUIRE(means && stds && idx && eps && ch_mean && ch_std && out && n > 0 && channels > 0 && hw > 0 && hw % 4 == 0,
                 "latent_sample: bad args (hw must be a multiple of 4)");
  const int64_t total4 = (int64_t)n * channels * hw / 4;
  const int64_t blocks = (total4 + 255) / 256;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  latent_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(means, stds, idx, eps, ch_mean, ch_std, out, total4, channels * hw / 4, hw / 4);
  MAPDIT_LAUNCH_CHECK("latent_sample");
  return MAPDIT_OK;
}

// Small kernels around the VAE legs of the pipeline (stable_diffusion.py:369-394, :793-834;
// ddim_inversion.py:29-31): pixel <-> activation conversions with the reference's rounding, the KL posterior sample,
// and a row softmax for the single-head mid-block attention (head dim 512 exceeds the fused attention kernel's TMEM budget,
// so that one layer runs as QK^T GEMM -> softmax -> PV GEMM).
#include "host_util.h"
#include "ptx.cuh"

namespace uv {

// In-place softmax(scale * x) over the columns of every row.  One CTA per row; cols <= 256 * kMaxPerThread.
static constexpr int kSmMaxPerThread = 64;
__global__ void softmax_rows_kernel(__half* __restrict__ x, int ld, int cols, float scale) {
  __shared__ float red[32];
  __half* row = x + (size_t)blockIdx.x * ld;
  float v[kSmMaxPerThread];
  float m = -INFINITY;
  int n = 0;
  for (int c = threadIdx.x; c < cols; c += blockDim.x, ++n) {
    v[n] = __half2float(row[c]) * scale;
    m = fmaxf(m, v[n]);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.0f;
  for (int i = 0; i < n; ++i) {
    v[i] = __expf(v[i] - m);
    s += v[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.0f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  const float inv = 1.0f / s;
  n = 0;
  for (int c = threadIdx.x; c < cols; c += blockDim.x, ++n) row[c] = __float2half_rn(v[n] * inv);
}

// decoder output rows [pixels, ld] (channels 0..2) -> uint8 [pixels, 3]: (x / 2 + 0.5).clamp(0, 1) in fp16, then
// round(255 * float(.)) half-to-even -- stable_diffusion.py:812-814 (get_images_from_latents)
__global__ void frames_to_u8_kernel(const __half* __restrict__ x, int ld, long long n, uint8_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * 3; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / 3;
    const int c = (int)(i - pix * 3);
    const __half h = __hadd(__hmul(x[pix * ld + c], __float2half(0.5f)), __float2half(0.5f));
    const float f = fminf(fmaxf(__half2float(h), 0.0f), 1.0f);
    out[i] = (uint8_t)rintf(f * 255.0f);
  }
}

// uint8 [pixels, 3] -> encoder input rows [pixels, cpad] fp16, channels >= 3 zero: image / 127.5 - 1.0 evaluated in
// double and rounded once (stable_diffusion.py:826-827: numpy float64, then .to(vae.dtype))
__global__ void u8_to_frames_kernel(const uint8_t* __restrict__ in, long long n, int cpad, __half* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * cpad; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cpad;
    const int c = (int)(i - pix * cpad);
    out[i] = c < 3 ? __double2half((double)in[pix * 3 + c] / 127.5 - 1.0) : __float2half(0.0f);
  }
}

// KL posterior sample of the encoder moments (diffusers DiagonalGaussianDistribution, third-party): rows [(f) hw, ld] =
// [mean C | logvar C]; z = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * scaling, written as (C, F, hw) -- the
// reference's "(b f) c h w -> b c f h w" (ddim_inversion.py:29-31).  noise: (F, C, hw) fp16 or null (the mode).
__global__ void vae_sample_kernel(const __half* __restrict__ mom, int ld, const __half* __restrict__ noise, int C, int F,
                                  int HW, float scaling, __half* __restrict__ out) {
  const long long total = (long long)C * F * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int f = (int)((i / HW) % F);
    const int c = (int)(i / ((long long)HW * F));
    const __half* r = mom + ((long long)f * HW + p) * ld;
    float z = __half2float(r[c]);
    if (noise) {
      const float lv = fminf(fmaxf(__half2float(r[C + c]), -30.0f), 20.0f);
      const __half sd = __float2half_rn(__expf(0.5f * lv));
      z = __half2float(__float2half_rn(z + __half2float(sd) * __half2float(noise[((long long)f * C + c) * HW + p])));
    }
    out[i] = __float2half_rn(z * scaling);
  }
}

static inline int grid_1d(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace uv

using namespace uv;

extern "C" int univst_softmax_rows_f16(void* X, int32_t ld, int32_t rows, int32_t cols, float scale, void* stream) {
  UV_REQUIRE(X && rows > 0 && cols > 0 && ld >= cols, "softmax_rows: bad arguments");
  UV_REQUIRE(cols <= 256 * kSmMaxPerThread, "softmax_rows: at most %d columns", 256 * kSmMaxPerThread);
  softmax_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>((__half*)X, ld, cols, scale);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_frames_to_u8(const void* X, int32_t ld, int64_t pixels, void* out, void* stream) {
  UV_REQUIRE(X && out && pixels > 0 && ld >= 3, "frames_to_u8: bad arguments");
  frames_to_u8_kernel<<<grid_1d(pixels * 3, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)X, ld, pixels, (uint8_t*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_u8_to_frames_f16(const void* in, int64_t pixels, int32_t cpad, void* out, void* stream) {
  UV_REQUIRE(in && out && pixels > 0 && cpad >= 3, "u8_to_frames: bad arguments");
  u8_to_frames_kernel<<<grid_1d(pixels * cpad, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in, pixels, cpad,
                                                                                   (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_vae_sample_f16(const void* moments, int32_t ld, const void* noise, int32_t C, int32_t F, int32_t HW,
                                     float scaling, void* out, void* stream) {
  UV_REQUIRE(moments && out && C > 0 && F > 0 && HW > 0 && ld >= 2 * C, "vae_sample: bad arguments");
  vae_sample_kernel<<<grid_1d((long long)C * F * HW, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)moments, ld, (const __half*)noise, C, F, HW, scaling, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

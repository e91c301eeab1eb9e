// Small HBM-bound kernels around the UNet: resampling of channels-last activations, latent pack / unpack
// between the reference's (B, 4, F, h, w) layout and the kernels' [(b f), h, w, C], the sinusoidal timestep
// embedding, and the per-step latent arithmetic of the denoising loop (mask blend, late latent AdaIN, DDIM step).
#include "xrank.cuh"
#include "ptx.cuh"

namespace uv {

// ---------------------------------------------------------------------------------------------- resampling
// nearest x2 upsample, NHWC (resnet.py:145, F.interpolate(scale_factor=2, mode="nearest"))
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int NB, int H, int W, int nvec) {
  const size_t total = (size_t)NB * 4 * H * W * nvec;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    size_t t = i / nvec;
    const int ox = (int)(t % (2 * W));
    t /= 2 * W;
    const int oy = (int)(t % (2 * H));
    const int n = (int)(t / (2 * H));
    y[i] = __ldg(x + (((size_t)n * H + (oy >> 1)) * W + (ox >> 1)) * nvec + v);
  }
}

// [NB, 2H, 2W, C] -> four parity planes [(row parity, col parity)][NB, H, W, C] (input of the stride-2 conv)
__global__ void space_to_depth2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int NB, int H, int W,
                                       int nvec) {
  const size_t plane = (size_t)NB * H * W * nvec;
  const size_t total = plane * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    size_t t = i / nvec;
    const int ix = (int)(t % (2 * W));
    t /= 2 * W;
    const int iy = (int)(t % (2 * H));
    const int n = (int)(t / (2 * H));
    const int p = (iy & 1) * 2 + (ix & 1);
    y[p * plane + (((size_t)n * H + (iy >> 1)) * W + (ix >> 1)) * nvec + v] = __ldg(x + i);
  }
}

// ---------------------------------------------------------------------------------------------- latents
struct PackSrc {
  const __half* z[4];
};
// out[(b f), hw, Cpad] = z_b[c, f, hw] (c < C), zero elsewhere
__global__ void pack_latents_kernel(PackSrc src, int B, int C, int F, int HW, int Cpad, __half* __restrict__ out) {
  const size_t total = (size_t)B * F * HW * Cpad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    size_t t = i / Cpad;
    const int hw = (int)(t % HW);
    t /= HW;
    const int f = (int)(t % F);
    const int b = (int)(t / F);
    out[i] = (c < C) ? src.z[b][((size_t)c * F + f) * HW + hw] : __float2half(0.0f);
  }
}

// y[b, c, f, hw] = x[(b f) hw, c]  (x row stride ld)
__global__ void unpack_latents_kernel(const __half* __restrict__ x, int ld, int B, int C, int F, int HW,
                                      __half* __restrict__ y) {
  const size_t total = (size_t)B * C * F * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int hw = (int)(i % HW);
    size_t t = i / HW;
    const int f = (int)(t % F);
    t /= F;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    y[i] = x[(((size_t)b * F + f) * HW + hw) * ld + c];
  }
}

// Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): out[b] = [cos(t f_k) | sin(t f_k)], f_k = exp(-ln(1e4) k / half)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B, int dim, __half* __restrict__ out) {
  const int half_dim = dim / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * half_dim; i += gridDim.x * blockDim.x) {
    const int b = i / half_dim, k = i % half_dim;
    const float freq = expf(-logf(10000.0f) * (float)k / (float)half_dim);
    const float e = t[b] * freq;
    out[(size_t)b * dim + k] = __float2half_rn(cosf(e));
    out[(size_t)b * dim + half_dim + k] = __float2half_rn(sinf(e));
  }
}

// bilinear (align_corners=False, no antialias) resize of a {0,1} mask, u8 [F, Hin, Win] -> fp16 [F, Hout, Wout]
// (stable_diffusion.py:689 F.interpolate(mask, size=latent, mode='bilinear'); load_mask src/util.py:133-144 maps
// any non-zero pixel to 1 through its uint8 * 255 wrap-around)
__global__ void mask_resize_kernel(const uint8_t* __restrict__ m, int F, int Hin, int Win, int Hout, int Wout,
                                   __half* __restrict__ out) {
  const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
  const int total = F * Hout * Wout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ox = i % Wout, oy = (i / Wout) % Hout, f = i / (Wout * Hout);
    float fy = ((float)oy + 0.5f) * sy - 0.5f, fx = ((float)ox + 0.5f) * sx - 0.5f;
    fy = fmaxf(fy, 0.0f);
    fx = fmaxf(fx, 0.0f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const uint8_t* p = m + (size_t)f * Hin * Win;
    const float v00 = p[y0 * Win + x0] ? 1.0f : 0.0f, v01 = p[y0 * Win + x1] ? 1.0f : 0.0f;
    const float v10 = p[y1 * Win + x0] ? 1.0f : 0.0f, v11 = p[y1 * Win + x1] ? 1.0f : 0.0f;
    const float v = (1.0f - ly) * ((1.0f - lx) * v00 + lx * v01) + ly * ((1.0f - lx) * v10 + lx * v11);
    out[i] = __float2half_rn(v);
  }
}

// out = (1 - m) * a + m * b, m[f, hw] broadcast over the C channels of (C, F, HW) latents (stable_diffusion.py:692)
__global__ void latent_blend_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                    const __half* __restrict__ m, int C, int FHW, __half* __restrict__ out) {
  const int total = C * FHW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const float w = __half2float(m[i % FHW]);
    out[i] = __float2half_rn((1.0f - w) * __half2float(a[i]) + w * __half2float(b[i]));
  }
}

// the same blend on frame-major (F, C, HW) latents -- the SD3 layout (custom_pipeline.py:300-303): m[f, hw] over channels
__global__ void latent_blend_fc_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                       const __half* __restrict__ m, int C, int HW, int total, __half* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const float w = __half2float(m[(i / (C * HW)) * HW + i % HW]);
    out[i] = __float2half_rn((1.0f - w) * __half2float(a[i]) + w * __half2float(b[i]));
  }
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.0f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
  return r;
}

// latent_adain (pnp_utils.py:128-139): out = IN_{F,H,W}(cnt) * std_s[c, f] + mean_s[c, f]; IN biased / eps 1e-5,
// style statistics unbiased over (H, W) per (channel, frame).  One block per channel.
__global__ void latent_adain_kernel(const __half* __restrict__ cnt, const __half* __restrict__ sty, int F, int HW,
                                    __half* __restrict__ out) {
  __shared__ float sh[32];
  __shared__ float s_mean[64], s_std[64];
  const int c = blockIdx.x;
  const __half* x = cnt + (size_t)c * F * HW;
  const __half* s = sty + (size_t)c * F * HW;
  const int n = F * HW;
  float a = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += __half2float(x[i]);
  const float mean = block_sum(a, sh) / (float)n;
  float q = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = __half2float(x[i]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum(q, sh) / (float)n + 1e-5f);
  for (int f = 0; f < F; ++f) {
    float sa = 0.0f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) sa += __half2float(s[(size_t)f * HW + i]);
    const float sm = block_sum(sa, sh) / (float)HW;
    float sq = 0.0f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float d = __half2float(s[(size_t)f * HW + i]) - sm;
      sq += d * d;
    }
    const float sv = block_sum(sq, sh) / (float)(HW - 1);
    if (threadIdx.x == 0) {
      // the reference keeps the statistics as fp16 tensors
      s_mean[f] = __half2float(__float2half_rn(sm));
      s_std[f] = __half2float(__float2half_rn(sqrtf(sv)));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int f = i / HW;
    const float v = __half2float(__float2half_rn((__half2float(x[i]) - mean) * rstd));
    out[(size_t)c * n + i] = __float2half_rn(v * s_std[f] + s_mean[f]);
  }
}

// DDIM step (eta = 0): x0 = (z - sqrt(1 - a_t) eps) / sqrt(a_t); z' = sqrt(a_p) x0 + sqrt(1 - a_p) eps.
// The same kernel is the inversion step with (a_t, a_p) swapped in meaning (ddim_inversion.py:190-204).
// eps is read straight from the channels-last conv_out buffer of the chosen branch.
__global__ void ddim_step_kernel(const __half* __restrict__ z, const __half* __restrict__ eps_nhwc, int ld,
                                 int branch_row0, int C, int F, int HW, float c_x0_z, float c_x0_e, float c_p_x0,
                                 float c_p_e, __half* __restrict__ z_out, __half* __restrict__ x0_out) {
  const int total = C * F * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int hw = i % HW, f = (i / HW) % F, c = i / (HW * F);
    const float e = __half2float(eps_nhwc[((size_t)branch_row0 + (size_t)f * HW + hw) * ld + c]);
    const float x0 = c_x0_z * __half2float(z[i]) - c_x0_e * e;
    if (x0_out) x0_out[i] = __float2half_rn(x0);
    z_out[i] = __float2half_rn(c_p_x0 * x0 + c_p_e * e);
  }
}

__global__ void axpby_kernel(const __half* __restrict__ a, const __half* __restrict__ b, float wa, float wb, int n,
                             __half* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = __float2half_rn(wa * __half2float(a[i]) + wb * __half2float(b[i]));
}

static inline int grid_for(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace uv

using namespace uv;

extern "C" int univst_upsample2x_f16(const void* X, int32_t NB, int32_t H, int32_t W, int32_t C, void* Y, void* stream) {
  UV_REQUIRE(X && Y && NB > 0 && H > 0 && W > 0 && C % 8 == 0, "upsample2x: bad arguments");
  const size_t total = (size_t)NB * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)X, (uint4*)Y, NB, H, W, C / 8);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_space_to_depth2_f16(const void* X, int32_t NB, int32_t Ho, int32_t Wo, int32_t C, void* Y,
                                          void* stream) {
  UV_REQUIRE(X && Y && NB > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, "space_to_depth2: bad arguments");
  const size_t total = (size_t)NB * 4 * Ho * Wo * (C / 8);
  space_to_depth2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)X, (uint4*)Y, NB, Ho, Wo,
                                                                               C / 8);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_pack_latents_f16(const void* const* Z, int32_t B, int32_t C, int32_t F, int32_t HW, int32_t Cpad,
                                       void* out, void* stream) {
  UV_REQUIRE(Z && out && B > 0 && B <= 4 && C > 0 && C <= Cpad, "pack_latents: bad arguments");
  PackSrc src{};
  for (int b = 0; b < B; ++b) {
    UV_REQUIRE(Z[b], "pack_latents: null latent");
    src.z[b] = (const __half*)Z[b];
  }
  const size_t total = (size_t)B * F * HW * Cpad;
  pack_latents_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, B, C, F, HW, Cpad, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_unpack_latents_f16(const void* X, int32_t ld, int32_t B, int32_t C, int32_t F, int32_t HW, void* Y,
                                         void* stream) {
  UV_REQUIRE(X && Y && B > 0 && C > 0 && C <= ld, "unpack_latents: bad arguments");
  const size_t total = (size_t)B * C * F * HW;
  unpack_latents_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)X, ld, B, C, F, HW,
                                                                             (__half*)Y);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_timestep_embedding_f16(const float* t, int32_t B, int32_t dim, void* out, void* stream) {
  UV_REQUIRE(t && out && B > 0 && dim % 2 == 0, "timestep_embedding: bad arguments");
  timestep_embedding_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(t, B, dim, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_mask_resize_u8(const uint8_t* mask, int32_t F, int32_t Hin, int32_t Win, int32_t Hout,
                                     int32_t Wout, void* out, void* stream) {
  UV_REQUIRE(mask && out && F > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "mask_resize: bad arguments");
  mask_resize_kernel<<<grid_for((size_t)F * Hout * Wout, 256), 256, 0, (cudaStream_t)stream>>>(mask, F, Hin, Win, Hout,
                                                                                              Wout, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_latent_blend_f16(const void* a, const void* b, const void* mask, int32_t C, int32_t F, int32_t HW,
                                       void* out, void* stream) {
  UV_REQUIRE(a && b && mask && out && C > 0 && F > 0 && HW > 0, "latent_blend: bad arguments");
  latent_blend_kernel<<<grid_for((size_t)C * F * HW, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)a, (const __half*)b, (const __half*)mask, C, F * HW, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_latent_blend_fc_f16(const void* a, const void* b, const void* mask, int32_t F, int32_t C, int32_t HW,
                                          void* out, void* stream) {
  UV_REQUIRE(a && b && mask && out && C > 0 && F > 0 && HW > 0 && (int64_t)F * C * HW < (1ll << 31), "latent_blend_fc: bad arguments");
  latent_blend_fc_kernel<<<grid_for((size_t)C * F * HW, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)a, (const __half*)b, (const __half*)mask, C, HW, F * C * HW, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_latent_adain_f16(const void* cnt, const void* sty, int32_t C, int32_t F, int32_t HW, void* out,
                                       void* stream) {
  UV_REQUIRE(cnt && sty && out && C > 0 && F > 0 && F <= 64 && HW > 1, "latent_adain: bad arguments (F <= 64)");
  latent_adain_kernel<<<C, 1024, 0, (cudaStream_t)stream>>>((const __half*)cnt, (const __half*)sty, F, HW, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_ddim_step_f16(const void* z, const void* eps_nhwc, int32_t ld, int32_t branch, int32_t C, int32_t F,
                                    int32_t HW, float alpha_t, float alpha_prev, void* z_out, void* x0_out,
                                    void* stream) {
  UV_REQUIRE(z && eps_nhwc && z_out && C > 0 && C <= ld && alpha_t > 0.0f && alpha_prev > 0.0f, "ddim_step: bad arguments");
  const float sa = sqrtf(alpha_t), sb = sqrtf(1.0f - alpha_t);
  ddim_step_kernel<<<grid_for((size_t)C * F * HW, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)z, (const __half*)eps_nhwc, ld, branch * F * HW, C, F, HW, 1.0f / sa, sb / sa, sqrtf(alpha_prev),
      sqrtf(1.0f - alpha_prev), (__half*)z_out, (__half*)x0_out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Frames <-> pixels exchange of the frame-sharded AnimateDiff motion modules, written straight into the
// peers' symmetric-memory buffers over NVLink: every 16-byte piece of a local row is stored at its place in the owner's
// final layout, so the permute-copy, the NCCL all-to-all and the permute-copy back of the library path are ONE pass
// (one read of the local rows, one write -- remote for (P - 1) / P of them).  The caller orders it with a cross-rank
// barrier on the stream (torch symmetric memory) before anybody reads.
//   dir 0: local rows (b, fl, pix)         -> rank pix / n, row (b, rank * Fl + fl, pix % n)      "all frames, my pixels"
//   dir 1: local rows (b, f_global, pix_l) -> rank f_global / Fl, row (b, f_global % Fl, rank * n + pix_l)
struct PeerPtrs {
  __half* p[16];
};

template <int DIR, bool SYNC>
__global__ void exchange_push_kernel(const __half* __restrict__ src, int ld, PeerPtrs dst, int rank, int P, int B, int Fl,
                                     int N, int C, XrankPeers xp) {
  const int vpr = C >> 3;                       // 16-byte vectors per row
  const int n = N / P;
  const long long rows = (long long)B * Fl * N;   // = B * (P Fl) * n
  const long long total = rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vpr;
    const int v = (int)(i - row * vpr);
    int owner;
    long long drow;
    if (DIR == 0) {
      const int pix = (int)(row % N);
      const long long bf = row / N;                // b * Fl + fl
      const int fl = (int)(bf % Fl), b = (int)(bf / Fl);
      owner = pix / n;
      drow = ((long long)b * (P * Fl) + rank * Fl + fl) * n + (pix - owner * n);
    } else {
      const int pl = (int)(row % n);
      const long long bf = row / n;                // b * (P Fl) + f_global
      const int fg = (int)(bf % (P * Fl)), b = (int)(bf / (P * Fl));
      owner = fg / Fl;
      drow = ((long long)b * Fl + (fg - owner * Fl)) * N + rank * n + pl;
    }
    const uint4 val = *reinterpret_cast<const uint4*>(src + row * ld + v * 8);
    *reinterpret_cast<uint4*>(dst.p[owner] + drow * C + v * 8) = val;
  }
  if (SYNC) xrank_kernel_tail(xp, rank, P);   // every rank's rows have landed when the kernel retires
}

// K/V halo of the frame-sharded sparse-causal attention: `nblk` blocks of [rows, cols] (one per branch: a
// boundary frame's K|V columns of the fused projection) are read once and stored into every non-null destination -- the
// next rank's "previous frame" bank, or, from rank 0, every rank's "first frame" bank -- over NVLink peer mappings.
__global__ void halo_push_kernel(const __half* __restrict__ src, int ld_src, long long src_blk, PeerPtrs dst, int P,
                                 int ld_dst, long long dst_blk, int nblk, int rows, int cols) {
  const int vpr = cols >> 3;
  const long long total = (long long)nblk * rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const long long rr = i / vpr;
    const int row = (int)(rr % rows), blk = (int)(rr / rows);
    const uint4 val = *reinterpret_cast<const uint4*>(src + (blk * src_blk + row) * ld_src + v * 8);
    const long long off = (blk * dst_blk + row) * ld_dst + v * 8;
    for (int r = 0; r < P; ++r)
      if (dst.p[r]) *reinterpret_cast<uint4*>(dst.p[r] + off) = val;
  }
}

extern "C" int univst_halo_push_f16(const void* src, int32_t ld_src, int64_t src_blk_rows, void* const* dst, int32_t P,
                                    int32_t ld_dst, int64_t dst_blk_rows, int32_t nblk, int32_t rows, int32_t cols,
                                    void* stream) {
  UV_REQUIRE(src && dst && P >= 1 && P <= 16 && nblk > 0 && rows > 0 && cols > 0, "halo_push: bad arguments (up to 16 ranks)");
  UV_REQUIRE(cols % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0, "halo_push: columns and row strides must be multiples of 8");
  UV_REQUIRE(((uintptr_t)src & 15) == 0, "halo_push: source must be 16-byte aligned");
  PeerPtrs pp{};
  int any = 0;
  for (int r = 0; r < P; ++r) {
    pp.p[r] = (__half*)dst[r];
    UV_REQUIRE(((uintptr_t)dst[r] & 15) == 0, "halo_push: destinations must be 16-byte aligned");
    any |= dst[r] != nullptr;
  }
  if (!any) return UNIVST_OK;
  const size_t total = (size_t)nblk * rows * (cols / 8);
  halo_push_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)src, ld_src, src_blk_rows, pp, P,
                                                                         ld_dst, dst_blk_rows, nblk, rows, cols);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

static int exchange_push(int32_t dir, const void* src, int32_t ld, void* const* dst, int32_t rank, int32_t P, int32_t B,
                         int32_t Fl, int32_t N, int32_t C, void* const* ctl, void* stream) {
  UV_REQUIRE(src && dst && (dir == 0 || dir == 1), "exchange_push: null pointer or bad direction");
  UV_REQUIRE(P >= 1 && P <= 16 && rank >= 0 && rank < P && B > 0 && Fl > 0 && N > 0 && N % P == 0,
             "exchange_push: up to 16 ranks, pixels divisible by the rank count");
  UV_REQUIRE(C > 0 && C % 8 == 0 && ld % 8 == 0 && ld >= C, "exchange_push: channels and row stride must be multiples of 8");
  PeerPtrs pp{};
  for (int r = 0; r < P; ++r) {
    UV_REQUIRE(dst[r], "exchange_push: null peer buffer");
    pp.p[r] = (__half*)dst[r];
  }
  const size_t total = (size_t)B * Fl * N * (C / 8);
  const int grid = grid_for(total, 256);
  XrankPeers xp{};
  if (ctl) {
    int r = fill_peers(xp, ctl, rank, P, "exchange_push");
    if (r) return r;
  }
  const __half* s = (const __half*)src;
  cudaStream_t st = (cudaStream_t)stream;
  if (dir == 0 && !ctl) exchange_push_kernel<0, false><<<grid, 256, 0, st>>>(s, ld, pp, rank, P, B, Fl, N, C, xp);
  if (dir == 1 && !ctl) exchange_push_kernel<1, false><<<grid, 256, 0, st>>>(s, ld, pp, rank, P, B, Fl, N, C, xp);
  if (dir == 0 && ctl) exchange_push_kernel<0, true><<<grid, 256, 0, st>>>(s, ld, pp, rank, P, B, Fl, N, C, xp);
  if (dir == 1 && ctl) exchange_push_kernel<1, true><<<grid, 256, 0, st>>>(s, ld, pp, rank, P, B, Fl, N, C, xp);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_exchange_push_f16(int32_t dir, const void* src, int32_t ld, void* const* dst, int32_t rank, int32_t P,
                                        int32_t B, int32_t Fl, int32_t N, int32_t C, void* stream) {
  return exchange_push(dir, src, ld, dst, rank, P, B, Fl, N, C, nullptr, stream);
}

extern "C" int univst_exchange_push_xrank_f16(int32_t dir, const void* src, int32_t ld, void* const* dst, void* const* ctl,
                                              int32_t rank, int32_t P, int32_t B, int32_t Fl, int32_t N, int32_t C,
                                              void* stream) {
  UV_REQUIRE(ctl, "exchange_push_xrank: null control blocks");
  return exchange_push(dir, src, ld, dst, rank, P, B, Fl, N, C, ctl, stream);
}

namespace uv {
// out = x + gate[row / rows_per_sample] * y   (the gated residuals of the SD3 MMDiT blocks: diffusers JointTransformerBlock)
__global__ void gated_add_kernel(const __half* __restrict__ x, const __half* __restrict__ y, const __half* __restrict__ gate,
                                 int ld_gate, int rows_per_sample, long long rows, int C, __half* __restrict__ out) {
  const int vpr = C >> 3;
  const long long total = rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vpr;
    const int v = (int)(i - row * vpr);
    const uint4 a = *reinterpret_cast<const uint4*>(x + row * C + v * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(y + row * C + v * 8);
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(gate + (row / rows_per_sample) * ld_gate + v * 8));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_half2(aw[j]), fb = unpack_half2(bw[j]), fg = unpack_half2(gw[j]);
      // the reference rounds gate * y to fp16 before the add (two tensor ops)
      const float2 p = unpack_half2(pack_half2(fg.x * fb.x, fg.y * fb.y));
      o[j] = pack_half2(fa.x + p.x, fa.y + p.y);
    }
    *reinterpret_cast<uint4*>(out + row * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

struct FloatVals {
  float v[64];
};
__global__ void set_floats_kernel(float* __restrict__ dst, FloatVals vals, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}
}  // namespace uv

extern "C" int univst_gated_add_f16(const void* x, const void* y, const void* gate, int32_t ld_gate, int32_t rows_per_sample,
                                    int64_t rows, int32_t C, void* out, void* stream) {
  UV_REQUIRE(x && y && gate && out && rows > 0 && C > 0 && C % 8 == 0 && ld_gate % 8 == 0 && rows_per_sample > 0 &&
                 ((uintptr_t)gate & 15) == 0,
             "gated_add: bad arguments (channels and gate row stride multiples of 8, 16-byte aligned gate)");
  gated_add_kernel<<<grid_for((size_t)rows * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)x, (const __half*)y, (const __half*)gate, ld_gate, rows_per_sample, rows, C, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_set_floats(float* dst, const float* vals, int32_t n, void* stream) {
  UV_REQUIRE(dst && vals && n > 0 && n <= 64, "set_floats: 1..64 values");
  FloatVals fv{};
  for (int i = 0; i < n; ++i) fv.v[i] = vals[i];
  set_floats_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(dst, fv, n);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_axpby_f16(const void* a, const void* b, float wa, float wb, int64_t n, void* out, void* stream) {
  UV_REQUIRE(a && b && out && n > 0 && n < (1ll << 31), "axpby: bad arguments");
  axpby_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)a, (const __half*)b, wa, wb,
                                                                         (int)n, (__half*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

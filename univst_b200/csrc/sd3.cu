// Elementwise pieces of the SD3 / SD3.5 joint-attention processors (reference: backbones/video_diffusion_sd3/
// pnp_utils.py) that have no counterpart in the SD / AnimateDiff path:
//   * per-head RMS norm of q and k (attn.norm_q / norm_k / norm_added_q / norm_added_k, called at :46-49, :92-95)
//   * the AdaIN-guided shift in the (frame, head, token, channel) layout (:180-193, attention_adain :287-300): the
//     style statistics are per (frame, head, channel) over the tokens, the content instance norm is over
//     (tokens, head_dim) jointly per (frame, head).
// Both run in place on the fused [tokens, ld] = [Q | K | V] projection buffer; HBM-bound.
#include "host_util.h"
#include "ptx.cuh"

namespace uv {

// one thread per (row, head) and tensor (blockIdx.y: 0 = Q block, 1 = K block): x <- x * rsqrt(mean(x^2) + eps) * w
__global__ void rmsnorm_heads_kernel(__half* __restrict__ x, int ld, int rows, int H, int d, int C,
                                     const __half* __restrict__ wq, const __half* __restrict__ wk, float eps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * H) return;
  const int row = idx / H, h = idx - row * H;
  const __half* w = blockIdx.y == 0 ? wq : wk;
  if (!w) return;
  __half* ptr = x + (size_t)row * ld + blockIdx.y * C + h * d;
  float ss = 0.0f;
  for (int c = 0; c < d; c += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(ptr + c);
    const uint32_t v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_half2(v[j]);
      ss += t.x * t.x + t.y * t.y;
    }
  }
  const float r = rsqrtf(ss / (float)d + eps);
  for (int c = 0; c < d; c += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(ptr + c);
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(w + c));
    const uint32_t v[4] = {u.x, u.y, u.z, u.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_half2(v[j]), s = unpack_half2(gw[j]);
      o[j] = pack_half2(t.x * r * s.x, t.y * r * s.y);
    }
    *reinterpret_cast<uint4*>(ptr + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// column sums of a [rows, ncols] block per frame: partial[f][chunk][ncols][2] = (sum, sum of squares)
__global__ void colsum_partial_kernel(const __half* __restrict__ x, int ld, size_t frame_stride, int rows, int ncols,
                                      int rows_per_chunk, float* __restrict__ partial) {
  const int f = blockIdx.z, chunk = blockIdx.y;
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (col >= ncols) return;
  const int r0 = chunk * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  const __half* base = x + (size_t)f * frame_stride + col;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(base + (size_t)r * ld));
    s0 += t.x;
    s1 += t.y;
    q0 += t.x * t.x;
    q1 += t.y * t.y;
  }
  float* out = partial + (((size_t)f * gridDim.y + chunk) * ncols + col) * 2;
  out[0] = s0;
  out[1] = q0;
  out[2] = s1;
  out[3] = q1;
}

// style: colstat[f][col][2] = (mean, unbiased std) over the tokens; edit: headstat[f][head'][2] = (mean, rstd) over
// (tokens, head_dim) with biased variance and eps 1e-5 (F.instance_norm of a 4-D tensor); head' runs over the K heads
// then the V heads.  grid (F), block 256.
__global__ void sd3_stats_final_kernel(const float* __restrict__ part_sty, const float* __restrict__ part_edit, int nchunks,
                                       int ncols, int rows, int d, float* __restrict__ colstat, float* __restrict__ headstat) {
  const int f = blockIdx.x;
  for (int col = threadIdx.x; col < ncols; col += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const float* pp = part_sty + (((size_t)f * nchunks + c) * ncols + col) * 2;
      s += pp[0];
      q += pp[1];
    }
    const float n = (float)rows, mean = s / n;
    colstat[((size_t)f * ncols + col) * 2] = mean;
    colstat[((size_t)f * ncols + col) * 2 + 1] = sqrtf(fmaxf((q - s * mean) / (n - 1.0f), 0.0f));
  }
  const int nheads = ncols / d;
  for (int h = threadIdx.x; h < nheads; h += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int c = 0; c < nchunks; ++c)
      for (int j = 0; j < d; ++j) {
        const float* pp = part_edit + (((size_t)f * nchunks + c) * ncols + h * d + j) * 2;
        s += pp[0];
        q += pp[1];
      }
    const float n = (float)rows * (float)d, mean = s / n;
    headstat[((size_t)f * nheads + h) * 2] = mean;
    headstat[((size_t)f * nheads + h) * 2 + 1] = rsqrtf(fmaxf(q / n - mean * mean, 0.0f) + 1e-5f);
  }
}

// one thread per (frame, token, 8 channels of Q|K|V): Q2 <- gamma (alpha Q0 + (1 - alpha) Q2);
// K2 <- beta ((K2 - mean_h) rstd_h sigma_c + mu_c) + (1 - beta) K1, V likewise.
__global__ void sd3_shift_kernel(__half* __restrict__ qkv, int ld, int F, int N, int C, int d,
                                 const float* __restrict__ colstat, const float* __restrict__ headstat, float alpha,
                                 float beta, float gamma) {
  const int nvec = 3 * C / 8;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)F * N * nvec) return;
  const int v = (int)(idx % nvec);
  const size_t tok = idx / nvec;
  const int f = (int)(tok / N);
  const int col = v * 8;
  const size_t branch = (size_t)F * N * ld;
  __half* p0 = qkv + tok * ld + col;   // content
  __half* p1 = p0 + branch;            // style
  __half* p2 = p1 + branch;            // edit
  const uint4 e = *reinterpret_cast<const uint4*>(p2);
  const uint32_t ew[4] = {e.x, e.y, e.z, e.w};
  uint32_t o[4];
  if (col < C) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p0));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_half2(aw[j]), y = unpack_half2(ew[j]);
      // two in-place fp16 assignments in the reference: the blend is rounded before the gamma scaling
      const float2 m = unpack_half2(pack_half2(alpha * x.x + (1.0f - alpha) * y.x, alpha * x.y + (1.0f - alpha) * y.y));
      o[j] = pack_half2(gamma * m.x, gamma * m.y);
    }
  } else {
    const int kc = col - C;               // column inside the K|V block
    const int hh = kc / d;                // K heads, then V heads
    const float mean = headstat[((size_t)f * (2 * C / d) + hh) * 2], rstd = headstat[((size_t)f * (2 * C / d) + hh) * 2 + 1];
    const uint4 su = __ldg(reinterpret_cast<const uint4*>(p1));
    const uint32_t sw[4] = {su.x, su.y, su.z, su.w};
    const float* cs = colstat + ((size_t)f * 2 * C + kc) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_half2(ew[j]), sy = unpack_half2(sw[j]);
      const float a = beta * ((t.x - mean) * rstd * cs[4 * j + 1] + cs[4 * j]) + (1.0f - beta) * sy.x;
      const float b = beta * ((t.y - mean) * rstd * cs[4 * j + 3] + cs[4 * j + 2]) + (1.0f - beta) * sy.y;
      o[j] = pack_half2(a, b);
    }
  }
  *reinterpret_cast<uint4*>(p2) = make_uint4(o[0], o[1], o[2], o[3]);
}

static constexpr int kSd3Chunks = 32;

}  // namespace uv

using namespace uv;

extern "C" int univst_rmsnorm_heads_f16(void* QKV, int32_t ld, int32_t rows, int32_t H, int32_t d, const void* wq,
                                        const void* wk, float eps, void* stream) {
  UV_REQUIRE(QKV && rows > 0 && H > 0 && d % 8 == 0 && ld % 8 == 0 && ld >= 2 * H * d, "rmsnorm_heads: bad shape");
  const int threads = 256;
  dim3 grid((unsigned)(((int64_t)rows * H + threads - 1) / threads), 2);
  rmsnorm_heads_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>((__half*)QKV, ld, rows, H, d, H * d, (const __half*)wq,
                                                                   (const __half*)wk, eps);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int64_t univst_sd3_shift_workspace_bytes(int32_t F, int32_t C, int32_t d) {
  // 2 x partial [F][chunks][2C][2] + colstat [F][2C][2] + headstat [F][2C/d][2]
  return ((int64_t)2 * F * kSd3Chunks * 2 * C * 2 + (int64_t)F * 2 * C * 2 + (int64_t)F * (2 * C / d) * 2) * sizeof(float);
}

extern "C" int univst_sd3_attn_shift_f16(void* QKV, int32_t ld, int32_t F, int32_t N, int32_t H, int32_t d, float alpha,
                                         float beta, float gamma, void* workspace, void* stream) {
  UV_REQUIRE(QKV && workspace, "sd3_attn_shift: null pointer");
  const int C = H * d;
  UV_REQUIRE(F > 0 && N > 1 && d % 8 == 0 && ld % 8 == 0 && ld >= 3 * C, "sd3_attn_shift: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  __half* base = (__half*)QKV;
  const int ncols = 2 * C;
  int nchunks = (N + 63) / 64;
  if (nchunks > kSd3Chunks) nchunks = kSd3Chunks;
  const int rows_per_chunk = (N + nchunks - 1) / nchunks;
  float* part_sty = (float*)workspace;
  float* part_edit = part_sty + (size_t)F * kSd3Chunks * ncols * 2;
  float* colstat = part_edit + (size_t)F * kSd3Chunks * ncols * 2;
  float* headstat = colstat + (size_t)F * ncols * 2;
  const size_t branch = (size_t)F * N * ld;
  dim3 grid((ncols / 2 + 127) / 128, nchunks, F);
  colsum_partial_kernel<<<grid, 128, 0, st>>>(base + branch + C, ld, (size_t)N * ld, N, ncols, rows_per_chunk, part_sty);
  UV_CHECK_CUDA(cudaGetLastError());
  colsum_partial_kernel<<<grid, 128, 0, st>>>(base + 2 * branch + C, ld, (size_t)N * ld, N, ncols, rows_per_chunk, part_edit);
  UV_CHECK_CUDA(cudaGetLastError());
  sd3_stats_final_kernel<<<F, 256, 0, st>>>(part_sty, part_edit, nchunks, ncols, N, d, colstat, headstat);
  UV_CHECK_CUDA(cudaGetLastError());
  const int64_t total = (int64_t)F * N * (3 * C / 8);
  sd3_shift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(base, ld, F, N, C, d, colstat, headstat, alpha, beta, gamma);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

// AdaIN-guided attention shift of the edit branch (reference: backbones/video_diffusion_sd/pnp_utils.py:47-57,
// attention_adain :114-125), applied in place on the fused [tokens, 3C] = [Q | K | V] projection buffer whose
// images are ordered (branch, frame) with branch 0 = content, 1 = style, 2 = edit:
//
//   Q2 <- gamma * (alpha * Q0 + (1 - alpha) * Q2)
//   K2 <- beta * (IN(K2) * sigma(K1) + mu(K1)) + (1 - beta) * K1          (V likewise)
//
// IN = per-token normalisation over the channel axis (biased variance, eps 1e-5: what F.instance_norm does to a
// (frames, tokens, channels) tensor); mu / sigma = per-(frame, channel) mean / unbiased std over the tokens of
// the style branch.  Two small HBM-bound kernels: column statistics of the style K|V, then one warp per token.
#include "host_util.h"
#include "ptx.cuh"

namespace uv {

static constexpr int kStatChunks = 32;

// grid (nchunks, F); x -> first column of the style K block of frame 0; partial[f][chunk][ncols][2]
__global__ void colstats_partial_kernel(const __half* __restrict__ x, int ld, size_t frame_stride, int rows, int nvec,
                                        int rows_par, int rows_per_chunk, float* __restrict__ partial) {
  const int f = blockIdx.y, chunk = blockIdx.x;
  const int v = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const int r_begin = chunk * rows_per_chunk, r_end = min(rows, r_begin + rows_per_chunk);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.0f;
  const __half* base = x + (size_t)f * frame_stride + v * 8;
  for (int r = r_begin + rl; r < r_end; r += rows_par) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * ld));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_half2(w[j]);
      s[2 * j] += t.x;
      s[2 * j + 1] += t.y;
      q[2 * j] += t.x * t.x;
      q[2 * j + 1] += t.y * t.y;
    }
  }
  // fold the row lanes of this block through shared memory
  extern __shared__ float sh[];  // [rows_par][nvec][16]
  float* mine = sh + ((size_t)rl * nvec + v) * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j] = s[j];
    mine[8 + j] = q[j];
  }
  __syncthreads();
  if (rl == 0) {
    for (int o = 1; o < rows_par; ++o) {
      const float* other = sh + ((size_t)o * nvec + v) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += other[j];
        q[j] += other[8 + j];
      }
    }
    float* out = partial + (((size_t)f * gridDim.x + chunk) * nvec + v) * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      out[j] = s[j];
      out[8 + j] = q[j];
    }
  }
}

// stats[f][col][2] = (mean, unbiased std)
__global__ void colstats_final_kernel(const float* __restrict__ partial, int nchunks, int nvec, int rows,
                                      float* __restrict__ stats) {
  const int f = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nvec * 8) return;
  const int v = col >> 3, j = col & 7;
  float s = 0.0f, q = 0.0f;
  for (int c = 0; c < nchunks; ++c) {
    const float* pp = partial + (((size_t)f * nchunks + c) * nvec + v) * 16;
    s += pp[j];
    q += pp[8 + j];
  }
  const float n = (float)rows;
  const float mean = s / n;
  const float var = fmaxf((q - s * mean) / (n - 1.0f), 0.0f);
  stats[((size_t)f * nvec * 8 + col) * 2] = mean;
  stats[((size_t)f * nvec * 8 + col) * 2 + 1] = sqrtf(var);
}

template <int MAXV>
__device__ __forceinline__ void shift_kv_row(const __half* __restrict__ sty, __half* __restrict__ edit,
                                             const float* __restrict__ st, int C, int lane, float beta) {
  // edit <- beta * (IN(edit) * sigma + mu) + (1 - beta) * sty
  const int nvec = C >> 3;
  uint4 e[MAXV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      e[i] = *reinterpret_cast<const uint4*>(edit + v * 8);
      const uint32_t w[4] = {e[i].x, e[i].y, e[i].z, e[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_half2(w[j]);
        s += t.x + t.y;
      }
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint32_t w[4] = {e[i].x, e[i].y, e[i].z, e[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_half2(w[j]);
        q += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint4 su = __ldg(reinterpret_cast<const uint4*>(sty + v * 8));
      const uint32_t w[4] = {e[i].x, e[i].y, e[i].z, e[i].w}, sw[4] = {su.x, su.y, su.z, su.w};
      const float4* stp = reinterpret_cast<const float4*>(st + (size_t)v * 16);  // 8 x (mean, std)
      float ms[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(stp + j);
        ms[4 * j] = t.x;
        ms[4 * j + 1] = t.y;
        ms[4 * j + 2] = t.z;
        ms[4 * j + 3] = t.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_half2(w[j]), sy = unpack_half2(sw[j]);
        const float a = beta * ((t.x - mean) * rstd * ms[4 * j + 1] + ms[4 * j]) + (1.0f - beta) * sy.x;
        const float b = beta * ((t.y - mean) * rstd * ms[4 * j + 3] + ms[4 * j + 2]) + (1.0f - beta) * sy.y;
        o[j] = pack_half2(a, b);
      }
      *reinterpret_cast<uint4*>(edit + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// one warp per (frame, token) of the edit branch
template <int MAXV>
__global__ void attn_shift_kernel(__half* __restrict__ qkv, int ld, int F, int N, int C, const float* __restrict__ stats,
                                  float alpha, float beta, float gamma, const float* __restrict__ abg) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= F * N) return;
  if (abg) {   // step parameters in device memory: the launch is the same for every DDIM step (CUDA-graph replay)
    alpha = __ldg(abg);
    beta = __ldg(abg + 1);
    gamma = __ldg(abg + 2);
  }
  const int lane = threadIdx.x & 31;
  const int f = tok / N;
  const size_t branch = (size_t)F * N * ld;
  __half* row0 = qkv + (size_t)tok * ld;   // content
  __half* row1 = row0 + branch;            // style
  __half* row2 = row1 + branch;            // edit
  const int nvec = C >> 3;
  // Q
  for (int v = lane; v < nvec; v += 32) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(row0 + v * 8));
    const uint4 b = *reinterpret_cast<const uint4*>(row2 + v * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_half2(aw[j]), y = unpack_half2(bw[j]);
      // the reference rounds the blend to fp16 before the gamma scaling (two in-place assignments)
      const float2 m = unpack_half2(pack_half2(alpha * x.x + (1.0f - alpha) * y.x, alpha * x.y + (1.0f - alpha) * y.y));
      o[j] = pack_half2(gamma * m.x, gamma * m.y);
    }
    *reinterpret_cast<uint4*>(row2 + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  const float* st = stats + (size_t)f * 2 * C * 2;  // [2C][2]: K columns then V columns
  shift_kv_row<MAXV>(row1 + C, row2 + C, st, C, lane, beta);
  shift_kv_row<MAXV>(row1 + 2 * C, row2 + 2 * C, st + (size_t)C * 2, C, lane, beta);
}

}  // namespace uv

using namespace uv;

extern "C" int64_t univst_attn_shift_workspace_bytes(int32_t F, int32_t C) {
  // partials [F][chunks][2C/8][16] + stats [F][2C][2]
  return ((int64_t)F * kStatChunks * (2 * C / 8) * 16 + (int64_t)F * 2 * C * 2) * sizeof(float);
}

static int attn_shift(void* QKV, int32_t ld, int32_t F, int32_t N, int32_t C, float alpha, float beta, float gamma,
                      const float* abg, void* workspace, void* stream) {
  UV_REQUIRE(QKV && workspace, "attn_shift: null pointer");
  UV_REQUIRE(F > 0 && N > 1 && C % 8 == 0 && C <= 2048 && ld % 8 == 0 && ld >= 3 * C, "attn_shift: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  __half* base = (__half*)QKV;
  const int nvec = 2 * C / 8;  // K|V columns
  UV_REQUIRE(nvec <= 1024, "attn_shift: too many channels");
  const int rows_par = nvec >= 256 ? 1 : 256 / nvec;
  int nchunks = (N + 31) / 32;
  if (nchunks > kStatChunks) nchunks = kStatChunks;
  const int rows_per_chunk = (N + nchunks - 1) / nchunks;
  float* partial = (float*)workspace;
  float* stats = partial + (size_t)F * kStatChunks * nvec * 16;
  const __half* styleKV = base + (size_t)F * N * ld + C;  // branch 1, column C
  const size_t sh = (size_t)rows_par * nvec * 16 * sizeof(float);
  if (sh > 48 * 1024) {
    static bool cfg = false;
    if (!cfg) {
      UV_CHECK_CUDA(cudaFuncSetAttribute(colstats_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      cfg = true;
    }
  }
  colstats_partial_kernel<<<dim3(nchunks, F), nvec * rows_par, sh, st>>>(styleKV, ld, (size_t)N * ld, N, nvec, rows_par,
                                                                      rows_per_chunk, partial);
  UV_CHECK_CUDA(cudaGetLastError());
  colstats_final_kernel<<<dim3((nvec * 8 + 255) / 256, F), 256, 0, st>>>(partial, nchunks, nvec, N, stats);
  UV_CHECK_CUDA(cudaGetLastError());
  const int warps = 8;
  const int blocks = (F * N + warps - 1) / warps;
  if (C <= 1024)
    attn_shift_kernel<4><<<blocks, warps * 32, 0, st>>>(base, ld, F, N, C, stats, alpha, beta, gamma, abg);
  else
    attn_shift_kernel<8><<<blocks, warps * 32, 0, st>>>(base, ld, F, N, C, stats, alpha, beta, gamma, abg);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_attn_shift_f16(void* QKV, int32_t ld, int32_t F, int32_t N, int32_t C, float alpha, float beta,
                                     float gamma, void* workspace, void* stream) {
  return attn_shift(QKV, ld, F, N, C, alpha, beta, gamma, nullptr, workspace, stream);
}

extern "C" int univst_attn_shift_dev_f16(void* QKV, int32_t ld, int32_t F, int32_t N, int32_t C, const float* abg,
                                         void* workspace, void* stream) {
  UV_REQUIRE(abg, "attn_shift_dev: null parameter pointer");
  return attn_shift(QKV, ld, F, N, C, 0.0f, 0.0f, 0.0f, abg, workspace, stream);
}

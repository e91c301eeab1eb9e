// HBM-bound normalisation kernels (channels-last fp16 activations, fp32 statistics).
//
//   GroupNorm   : statistics per (batch, group) over `rows` tokens x (C / groups) channels.  With rows = F*H*W
//                 and batch = branch this is the reference's GroupNorm on a 5-D tensor whose statistics span all
//                 frames (resnet.py:338,369; unet_3d_condition.py:439); with rows = H*W and batch = image it is
//                 the per-frame GroupNorm of the transformer (attention.py:121).  The input may be the channel
//                 concat of two tensors (unet_3d_blocks.py:523,618) -- the concat only ever exists as this
//                 kernel's normalised output.
//   LayerNorm   : per token over C (attention.py:290,312,329).
#include "host_util.h"
#include "ptx.cuh"

namespace uv {

static constexpr int kGnMaxChunks = 64;

__device__ __forceinline__ const uint4* gn_src(const __half* x1, const __half* x2, int C1, int C2, size_t row, int v) {
  // v-th 8-channel vector of the concatenated row
  const int c = v * 8;
  return (c < C1) ? reinterpret_cast<const uint4*>(x1 + row * C1 + c)
                  : reinterpret_cast<const uint4*>(x2 + row * C2 + (c - C1));
}

// grid (nchunks, NB); block = nvec * rows_par threads.  partial[b][chunk][group][2] = (sum, sumsq)
__global__ void gn_stats_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, int C1, int C2, int rows,
                                int groups, int nvec, int rows_par, int rows_per_chunk, float* __restrict__ partial) {
  extern __shared__ float sh[];  // [threads][8] per-thread pair sums, then [groups][2]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int C = C1 + C2, cpg = C / groups;
  const int v = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const int r_begin = chunk * rows_per_chunk, r_end = min(rows, r_begin + rows_per_chunk);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (int r = r_begin + rl; r < r_end; r += rows_par) {
    const uint4 u = __ldg(gn_src(x1, x2, C1, C2, (size_t)b * rows + r, v));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_half2(w[j]);
      s[j] += f.x + f.y;
      q[j] += f.x * f.x + f.y * f.y;
    }
  }
  // deterministic fold (no float atomics): thread (rl, v) parks its four channel-pair sums, then thread g adds up the
  // pairs of group g in a fixed order (cpg is even: a channel pair never straddles two groups)
  float* mine = sh + (size_t)threadIdx.x * 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mine[j] = s[j];
    mine[4 + j] = q[j];
  }
  __syncthreads();
  float* out = partial + ((size_t)b * gridDim.x + chunk) * groups * 2;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float gs = 0.0f, gq = 0.0f;
    const int p0 = g * (cpg / 2), p1 = p0 + cpg / 2;  // channel-pair range of the group
    for (int o = 0; o < rows_par; ++o)
      for (int pr = p0; pr < p1; ++pr) {
        const float* e = sh + ((size_t)o * nvec + (pr >> 2)) * 8;
        gs += e[pr & 3];
        gq += e[4 + (pr & 3)];
      }
    out[g * 2] = gs;
    out[g * 2 + 1] = gq;
  }
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// grid (row blocks, NB); each block first folds the chunk partials of its batch into mean / rstd.
__global__ void gn_apply_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, int C1, int C2, int rows,
                                int groups, int nvec, int nchunks, const float* __restrict__ partial,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps, int silu,
                                int rows_per_block, int stat_rows, __half* __restrict__ y) {
  extern __shared__ float sh[];  // [groups][2] = (mean, rstd)
  const int b = blockIdx.y;
  const int C = C1 + C2, cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.0f, q = 0.0f;
    const float* pp = partial + (size_t)b * nchunks * groups * 2 + g * 2;
    for (int c = 0; c < nchunks; ++c) {
      s += pp[(size_t)c * groups * 2];
      q += pp[(size_t)c * groups * 2 + 1];
    }
    const float n = (float)stat_rows * (float)cpg;  // rows the statistics were taken over (all ranks when sharded)
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.0f);
    sh[g * 2] = mean;
    sh[g * 2 + 1] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  const int total = (r_end - r_begin) * nvec;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int r = r_begin + idx / nvec, v = idx % nvec;
    const size_t row = (size_t)b * rows + r;
    const uint4 u = __ldg(gn_src(x1, x2, C1, C2, row, v));
    const uint4 gm = __ldg(reinterpret_cast<const uint4*>(gamma + v * 8));
    const uint4 bt = __ldg(reinterpret_cast<const uint4*>(beta + v * 8));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w}, gw[4] = {gm.x, gm.y, gm.z, gm.w}, bw[4] = {bt.x, bt.y, bt.z, bt.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (v * 8 + j * 2) / cpg;
      const float mean = sh[g * 2], rstd = sh[g * 2 + 1];
      const float2 f = unpack_half2(w[j]), ga = unpack_half2(gw[j]), be = unpack_half2(bw[j]);
      float a = (f.x - mean) * rstd * ga.x + be.x;
      float c = (f.y - mean) * rstd * ga.y + be.y;
      if (silu) {
        // the reference rounds the GroupNorm output to fp16 before the activation
        a = silu_f(__half2float(__float2half_rn(a)));
        c = silu_f(__half2float(__float2half_rn(c)));
      }
      o[j] = pack_half2(a, c);
    }
    *reinterpret_cast<uint4*>(y + row * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one warp per token row
__global__ void layernorm_kernel(const __half* __restrict__ x, int rows, int C, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, float eps, __half* __restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  constexpr int kMaxV = 8;  // C <= 2048
  uint4 buf[kMaxV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      buf[i] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)row * C + v * 8));
      const uint32_t w[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_half2(w[j]);
        s += f.x + f.y;
      }
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint32_t w[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_half2(w[j]);
        q += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint4 gm = __ldg(reinterpret_cast<const uint4*>(gamma + v * 8));
      const uint4 bt = __ldg(reinterpret_cast<const uint4*>(beta + v * 8));
      const uint32_t w[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w}, gw[4] = {gm.x, gm.y, gm.z, gm.w},
                     bw[4] = {bt.x, bt.y, bt.z, bt.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_half2(w[j]), ga = unpack_half2(gw[j]), be = unpack_half2(bw[j]);
        o[j] = pack_half2((f.x - mean) * rstd * ga.x + be.x, (f.y - mean) * rstd * ga.y + be.y);
      }
      *reinterpret_cast<uint4*>(y + (size_t)row * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace uv

using namespace uv;

extern "C" int64_t univst_groupnorm_workspace_bytes(int32_t NB, int32_t groups) {
  return (int64_t)NB * kGnMaxChunks * groups * 2 * sizeof(float);
}

extern "C" int univst_groupnorm_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                                    int32_t groups, const void* gamma, const void* beta, float eps, int32_t silu,
                                    void* Y, void* workspace, void* stream) {
  UV_REQUIRE(X1 && Y && gamma && beta && workspace, "groupnorm: null pointer");
  if (!X2) C2 = 0;
  const int C = C1 + C2;
  UV_REQUIRE(NB > 0 && rows > 0 && groups > 0 && C % groups == 0, "groupnorm: bad shape");
  UV_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && (C / groups) % 2 == 0, "groupnorm: channels %% 8, channels per group even");
  UV_REQUIRE(C / 8 <= 1024, "groupnorm: at most 8192 channels");
  const int nvec = C / 8;
  const int rows_par = nvec >= 256 ? 1 : 256 / nvec;
  const int threads = nvec * rows_par;
  int nchunks = (rows + 63) / 64;
  if (nchunks > kGnMaxChunks) nchunks = kGnMaxChunks;
  const int rows_per_chunk = (rows + nchunks - 1) / nchunks;
  cudaStream_t st = (cudaStream_t)stream;
  gn_stats_kernel<<<dim3(nchunks, NB), threads, threads * 8 * sizeof(float), st>>>(
      (const __half*)X1, (const __half*)X2, C1, C2, rows, groups, nvec, rows_par, rows_per_chunk, (float*)workspace);
  UV_CHECK_CUDA(cudaGetLastError());
  // apply: ~32 KiB of activations per block
  int rows_per_block = (16384 + C - 1) / C;
  if (rows_per_block < 1) rows_per_block = 1;
  const int row_blocks = (rows + rows_per_block - 1) / rows_per_block;
  gn_apply_kernel<<<dim3(row_blocks, NB), 256, groups * 2 * sizeof(float), st>>>(
      (const __half*)X1, (const __half*)X2, C1, C2, rows, groups, nvec, nchunks, (const float*)workspace,
      (const __half*)gamma, (const __half*)beta, eps, silu, rows_per_block, rows, (__half*)Y);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

namespace uv {
// sums[b][g][2] = sum over the chunk partials (fixed order)
__global__ void gn_fold_kernel(const float* __restrict__ partial, int nchunks, int groups, float* __restrict__ sums) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) {
    float s = 0.0f;
    for (int c = 0; c < nchunks; ++c) s += partial[((size_t)b * nchunks + c) * groups * 2 + i];
    sums[(size_t)b * groups * 2 + i] = s;
  }
}
}  // namespace uv

// Split form for frame-sharded execution: local (sum, sum of squares) per (batch, group) -> [all-reduce over ranks by
// the caller] -> apply with the global row count.
extern "C" int univst_groupnorm_stats_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                                          int32_t groups, float* sums, void* workspace, void* stream) {
  UV_REQUIRE(X1 && sums && workspace, "groupnorm_stats: null pointer");
  if (!X2) C2 = 0;
  const int C = C1 + C2;
  UV_REQUIRE(NB > 0 && rows > 0 && groups > 0 && C % groups == 0, "groupnorm_stats: bad shape");
  UV_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && (C / groups) % 2 == 0 && C / 8 <= 1024, "groupnorm_stats: bad channels");
  const int nvec = C / 8;
  const int rows_par = nvec >= 256 ? 1 : 256 / nvec;
  const int threads = nvec * rows_par;
  int nchunks = (rows + 63) / 64;
  if (nchunks > kGnMaxChunks) nchunks = kGnMaxChunks;
  const int rows_per_chunk = (rows + nchunks - 1) / nchunks;
  cudaStream_t st = (cudaStream_t)stream;
  gn_stats_kernel<<<dim3(nchunks, NB), threads, threads * 8 * sizeof(float), st>>>(
      (const __half*)X1, (const __half*)X2, C1, C2, rows, groups, nvec, rows_par, rows_per_chunk, (float*)workspace);
  UV_CHECK_CUDA(cudaGetLastError());
  gn_fold_kernel<<<NB, 64, 0, st>>>((const float*)workspace, nchunks, groups, sums);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_groupnorm_apply_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                                          int32_t groups, const float* sums, int64_t stat_rows, const void* gamma,
                                          const void* beta, float eps, int32_t silu, void* Y, void* stream) {
  UV_REQUIRE(X1 && Y && gamma && beta && sums, "groupnorm_apply: null pointer");
  if (!X2) C2 = 0;
  const int C = C1 + C2;
  UV_REQUIRE(NB > 0 && rows > 0 && groups > 0 && C % groups == 0 && stat_rows >= rows, "groupnorm_apply: bad shape");
  const int nvec = C / 8;
  int rows_per_block = (16384 + C - 1) / C;
  if (rows_per_block < 1) rows_per_block = 1;
  const int row_blocks = (rows + rows_per_block - 1) / rows_per_block;
  gn_apply_kernel<<<dim3(row_blocks, NB), 256, groups * 2 * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)X1, (const __half*)X2, C1, C2, rows, groups, nvec, 1, sums, (const __half*)gamma,
      (const __half*)beta, eps, silu, rows_per_block, (int)stat_rows, (__half*)Y);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_layernorm_f16(const void* X, int32_t rows, int32_t C, const void* gamma, const void* beta,
                                    float eps, void* Y, void* stream) {
  UV_REQUIRE(X && Y && gamma && beta, "layernorm: null pointer");
  UV_REQUIRE(rows > 0 && C % 8 == 0 && C <= 2048, "layernorm: C must be a multiple of 8, at most 2048");
  const int warps = 8;
  layernorm_kernel<<<(rows + warps - 1) / warps, warps * 32, 0, (cudaStream_t)stream>>>(
      (const __half*)X, rows, C, (const __half*)gamma, (const __half*)beta, eps, (__half*)Y);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

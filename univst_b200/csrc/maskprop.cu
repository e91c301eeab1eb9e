// Point-matching mask propagation (reference: src/mask_propagation.py:72-83):
//   aff[m, n] = exp(<tar_n, src_m> / T) on L2-normalised features; per target point n keep the entries >= its
//   topk-th largest (ties kept), normalise the kept column, segs_tar[:, n] = segs[:, kept] . aff[kept, n].
// The M x N affinity (246 MB for M = 15 k) is never materialised: one CTA owns 32 target points and streams the
// source points through shared memory twice -- pass 1 maintains each point's sorted top-k list across a warp
// (shuffle insertion), pass 2 recomputes the same scores bit-for-bit and accumulates the kept ones in ascending
// source order.  fp32 CUDA-core arithmetic on purpose: the kept-index set has to match the fp32 reference.
#include "host_util.h"
#include "ptx.cuh"

namespace uv {

static constexpr int kTR = 32;    // target rows per CTA
static constexpr int kSC = 128;   // source columns per chunk
static constexpr int kKC = 32;    // feature channels per smem step
static constexpr int kMaxClassRegs = 8;  // classes <= 256

// rows: y[r, :] = x[r, :] / max(||x[r, :]||, 1e-12)   (F.normalize(dim=1))
__global__ void normalize_rows_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) {
    const float v = x[(size_t)row * C + c];
    s += v * v;
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int c = lane; c < C; c += 32) y[(size_t)row * C + c] = x[(size_t)row * C + c] * inv;
}

// columns of x[C, M] -> y[M, C] = normalised columns (F.normalize(dim=0)), transposed for coalesced dot products
__global__ void normalize_cols_t_kernel(const float* __restrict__ x, int C, int M, float* __restrict__ y) {
  __shared__ float tile[32][33];
  __shared__ float inv_s[32];
  const int m0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  float s = 0.0f;
  for (int c = ty; c < C; c += 8) {
    const float v = (m0 + tx < M) ? x[(size_t)c * M + m0 + tx] : 0.0f;
    s += v * v;
  }
  tile[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += tile[i][tx];
    inv_s[tx] = 1.0f / fmaxf(sqrtf(t), 1e-12f);
  }
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int i = ty; i < 32; i += 8)
      tile[i][tx] = (c0 + i < C && m0 + tx < M) ? x[(size_t)(c0 + i) * M + m0 + tx] * inv_s[tx] : 0.0f;
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
      if (m0 + i < M && c0 + tx < C) y[(size_t)(m0 + i) * C + c0 + tx] = tile[tx][i];
    __syncthreads();
  }
}

// aff tile [kTR][kSC] of this CTA's targets against source chunk `chunk` -> smem S
__device__ __forceinline__ void aff_tile(const float* __restrict__ tar, const float* __restrict__ src, int N, int C, int M,
                                         int n0, int chunk, float inv_temp_unused, float temperature, float (*At)[kTR + 1],
                                         float (*Bs)[kSC + 4], float (*S)[kSC + 1]) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;  // 4 cols x 4 rows per thread
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const int m0 = chunk * kSC;
  for (int k0 = 0; k0 < C; k0 += kKC) {
    // At[k][r] = tar[n0 + r][k0 + k]; Bs[k][c] = src[m0 + c][k0 + k]
    for (int e = tid; e < kTR * kKC; e += 256) {
      const int r = e / kKC, k = e % kKC;
      At[k][r] = (n0 + r < N && k0 + k < C) ? tar[(size_t)(n0 + r) * C + k0 + k] : 0.0f;
    }
    for (int e = tid; e < kSC * kKC; e += 256) {
      const int c = e / kKC, k = e % kKC;
      Bs[k][c] = (m0 + c < M && k0 + k < C) ? src[(size_t)(m0 + c) * C + k0 + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kKC; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = At[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 32 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tx + 32 * j;
      // out-of-range source columns can never be selected
      S[ty * 4 + i][c] = (m0 + c < M) ? expf(__fdiv_rn(acc[i][j], temperature)) : -1.0f;
    }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
maskprop_kernel(const float* __restrict__ tar, const float* __restrict__ src, const float* __restrict__ segs, int N, int C,
                int M, int Ccls, float temperature, int topk, float* __restrict__ out, float* __restrict__ thr_out,
                int* __restrict__ kept_idx, int kept_cap) {
  __shared__ float At[kKC][kTR + 1];
  __shared__ float Bs[kKC][kSC + 4];
  __shared__ float S[kTR][kSC + 1];
  const int n0 = blockIdx.x * kTR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (M + kSC - 1) / kSC;

  // ---- pass 1: per-row sorted top-k list, lane i holds the i-th largest affinity seen so far
  float list[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) list[rr] = -INFINITY;
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    aff_tile(tar, src, N, C, M, n0, chunk, 0.0f, temperature, At, Bs, S);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int row = warp * 4 + rr;
      float cur = list[rr];
      float thr = __shfl_sync(0xffffffffu, cur, topk - 1);
#pragma unroll
      for (int t = 0; t < kSC / 32; ++t) {
        const float v = S[row][t * 32 + lane];
        unsigned cand = __ballot_sync(0xffffffffu, v > thr);
        while (cand) {
          const int srcl = __ffs(cand) - 1;
          cand &= cand - 1;
          const float x = __shfl_sync(0xffffffffu, v, srcl);
          if (x > thr) {  // warp-uniform (thr may have risen since the ballot)
            const unsigned ge = __ballot_sync(0xffffffffu, lane < topk && cur >= x);
            const int pos = __popc(ge);
            const float up = __shfl_up_sync(0xffffffffu, cur, 1);
            if (lane == pos) cur = x;
            else if (lane > pos && lane < topk) cur = up;
            thr = __shfl_sync(0xffffffffu, cur, topk - 1);
          }
        }
      }
      list[rr] = cur;
    }
    __syncthreads();
  }
  float thr4[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) thr4[rr] = __shfl_sync(0xffffffffu, list[rr], topk - 1);

  // ---- pass 2: same scores again; kept entries (aff >= threshold) accumulated in ascending source order
  float num[4][kMaxClassRegs], den[4];
  int cnt[4] = {0, 0, 0, 0};
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    den[rr] = 0.0f;
#pragma unroll
    for (int u = 0; u < kMaxClassRegs; ++u) num[rr][u] = 0.0f;
  }
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    aff_tile(tar, src, N, C, M, n0, chunk, 0.0f, temperature, At, Bs, S);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int row = warp * 4 + rr;
#pragma unroll
      for (int t = 0; t < kSC / 32; ++t) {
        const float v = S[row][t * 32 + lane];
        unsigned keep = __ballot_sync(0xffffffffu, v >= thr4[rr]);
        while (keep) {
          const int srcl = __ffs(keep) - 1;
          keep &= keep - 1;
          const float a = __shfl_sync(0xffffffffu, v, srcl);
          const int m = chunk * kSC + t * 32 + srcl;
          den[rr] += a;
          if (kept_idx && lane == 0 && n0 + row < N && cnt[rr] < kept_cap) kept_idx[(size_t)(n0 + row) * kept_cap + cnt[rr]] = m;
          ++cnt[rr];
#pragma unroll
          for (int u = 0; u < kMaxClassRegs; ++u) {
            const int c = lane + 32 * u;
            if (c < Ccls) num[rr][u] = fmaf(__ldg(segs + (size_t)c * M + m), a, num[rr][u]);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int n = n0 + warp * 4 + rr;
    if (n >= N) continue;
#pragma unroll
    for (int u = 0; u < kMaxClassRegs; ++u) {
      const int c = lane + 32 * u;
      if (c < Ccls) out[(size_t)c * N + n] = num[rr][u] / den[rr];
    }
    if (thr_out && lane == 0) thr_out[n] = thr4[rr];
    if (kept_idx && lane == 0)
      for (int e = cnt[rr]; e < kept_cap; ++e) kept_idx[(size_t)n * kept_cap + e] = -1;
  }
}

}  // namespace uv

using namespace uv;

extern "C" int64_t univst_maskprop_workspace_bytes(int32_t N, int32_t C, int32_t M) {
  return ((int64_t)N * C + (int64_t)M * C) * sizeof(float);
}

extern "C" int univst_maskprop_f32(const float* feat_tar, const float* feat_src, const float* segs, int32_t N, int32_t C,
                                   int32_t M, int32_t Ccls, float temperature, int32_t topk, float* segs_tar,
                                   float* thresholds, int32_t* kept_idx, int32_t kept_cap, void* workspace,
                                   void* stream) {
  UV_REQUIRE(feat_tar && feat_src && segs && segs_tar && workspace, "maskprop: null pointer");
  UV_REQUIRE(N > 0 && C > 0 && M >= topk && topk >= 1 && topk <= 32, "maskprop: need 1 <= topk <= 32 <= M");
  UV_REQUIRE(Ccls >= 1 && Ccls <= 32 * kMaxClassRegs, "maskprop: at most 256 classes");
  UV_REQUIRE(temperature > 0.0f, "maskprop: temperature must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  float* tar_n = (float*)workspace;
  float* src_n = tar_n + (size_t)N * C;
  normalize_rows_kernel<<<(N + 7) / 8, 256, 0, st>>>(feat_tar, N, C, tar_n);
  UV_CHECK_CUDA(cudaGetLastError());
  normalize_cols_t_kernel<<<(M + 31) / 32, 256, 0, st>>>(feat_src, C, M, src_n);
  UV_CHECK_CUDA(cudaGetLastError());
  maskprop_kernel<<<(N + kTR - 1) / kTR, 256, 0, st>>>(tar_n, src_n, segs, N, C, M, Ccls, temperature, topk, segs_tar,
                                                     thresholds, kept_idx, kept_cap);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

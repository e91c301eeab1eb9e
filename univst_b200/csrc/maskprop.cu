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
static constexpr int kSC = 256;   // source columns per chunk
static constexpr int kKC = 32;    // feature channels per smem step
static constexpr int kMaxClassRegs = 8;  // classes <= 256
static constexpr int kStageFloats = kKC * kTR + kKC * kSC;                       // one pipeline stage: At | Bs
static constexpr int kSmemBytes = (2 * kStageFloats + kTR * (kSC + 1)) * 4;      // 2 stages + the affinity tile

// Both operands are normalised once into K-MAJOR layouts ([C][ld], the point index contiguous), so that the tile loads
// of the affinity GEMM are contiguous 16-byte copies (cp.async) in global AND conflict-free in shared memory.
// rows of x[rows, C] -> y[C, ld] = normalised rows, transposed   (F.normalize(dim=1) of the target features)
__global__ void normalize_rows_t_kernel(const float* __restrict__ x, int rows, int C, int ld, float* __restrict__ y) {
  __shared__ float tile[32][33];
  __shared__ float inv_s[32];
  const int r0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  // norms: warp ty handles rows ty, ty + 8, ...
  for (int i = ty; i < 32; i += 8) {
    float s = 0.0f;
    if (r0 + i < rows)
      for (int c = tx; c < C; c += 32) {
        const float v = x[(size_t)(r0 + i) * C + c];
        s += v * v;
      }
    s = warp_sum(s);
    if (tx == 0) inv_s[i] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int i = ty; i < 32; i += 8)
      tile[i][tx] = (r0 + i < rows && c0 + tx < C) ? x[(size_t)(r0 + i) * C + c0 + tx] * inv_s[i] : 0.0f;
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
      if (c0 + i < C && r0 + tx < ld) y[(size_t)(c0 + i) * ld + r0 + tx] = (r0 + tx < rows) ? tile[tx][i] : 0.0f;
    __syncthreads();
  }
}

// columns of x[C, M] -> y[C, ld] = normalised columns, same orientation (F.normalize(dim=0) of the source features)
__global__ void normalize_cols_kernel(const float* __restrict__ x, int C, int M, int ld, float* __restrict__ y) {
  __shared__ float part[8][32];
  __shared__ float inv_s[32];
  const int m0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  float s = 0.0f;
  for (int c = ty; c < C; c += 8) {
    const float v = (m0 + tx < M) ? x[(size_t)c * M + m0 + tx] : 0.0f;
    s += v * v;
  }
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    inv_s[tx] = 1.0f / fmaxf(sqrtf(t), 1e-12f);
  }
  __syncthreads();
  if (m0 + tx < ld)
    for (int c = ty; c < C; c += 8) y[(size_t)c * ld + m0 + tx] = (m0 + tx < M) ? x[(size_t)c * M + m0 + tx] * inv_s[tx] : 0.0f;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// aff tile [kTR][kSC] of this CTA's targets against source chunk `chunk` -> smem S.
// 256 threads, 4 x 8 outputs each (rows ty*4.., columns tx*4.. and 128 + tx*4..): per k one broadcast LDS.128 of the
// targets, two conflict-free LDS.128 of the sources and 32 FFMAs; the K loop runs over a two-stage cp.async pipeline.
// Every accumulator is the plain sequential fma over k = 0 .. C-1, so pass 2 reproduces pass 1 bit for bit.
// tarT [C][ldn], srcT [C][ldm] (K-major, padded to multiples of 4 floats; rows beyond C do not exist -> C % 32 handled
// by zero-filling), ldn / ldm multiples of 4 and the padding columns are zero.
__device__ __forceinline__ void aff_tile(const float* __restrict__ tarT, const float* __restrict__ srcT, int N, int C, int M,
                                         int ldn, int ldm, int n0, int chunk, float temperature, float* stage,
                                         float (*S)[kSC + 1]) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  const int m0 = chunk * kSC;
  const int nsteps = (C + kKC - 1) / kKC;
  auto load_stage = [&](int step, int buf) {
    float* At = stage + buf * kStageFloats;          // [kKC][kTR]
    float* Bs = At + kKC * kTR;                      // [kKC][kSC]
    const int k0 = step * kKC;
    {  // targets: 32 k x 8 float4
      const int k = tid >> 3, r4 = (tid & 7) * 4;
      float* dst = At + k * kTR + r4;
      if (k0 + k < C && n0 + r4 < ldn) cp_async16(dst, tarT + (size_t)(k0 + k) * ldn + n0 + r4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {  // sources: 32 k x 64 float4
      const int e = tid + 256 * it;
      const int k = e >> 6, c4 = (e & 63) * 4;
      float* dst = Bs + k * kSC + c4;
      if (k0 + k < C && m0 + c4 < ldm) cp_async16(dst, srcT + (size_t)(k0 + k) * ldm + m0 + c4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  load_stage(0, 0);
  for (int step = 0; step < nsteps; ++step) {
    const int buf = step & 1;
    if (step + 1 < nsteps) {
      load_stage(step + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* At = stage + buf * kStageFloats;
    const float* Bs = At + kKC * kTR;
#pragma unroll 8
    for (int k = 0; k < kKC; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(At + k * kTR + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * kSC + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * kSC + 128 + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();   // the buffer just read is refilled by the next iteration's load
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j < 4 ? 0 : 128) + tx * 4 + (j & 3);
      // out-of-range source columns can never be selected
      S[ty * 4 + i][c] = (m0 + c < M) ? expf(__fdiv_rn(acc[i][j], temperature)) : -1.0f;
    }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
maskprop_kernel(const float* __restrict__ tar, const float* __restrict__ src, const float* __restrict__ segs, int N, int C,
                int M, int ldn, int ldm, int Ccls, float temperature, int topk, float* __restrict__ out,
                float* __restrict__ thr_out, int* __restrict__ kept_idx, int kept_cap) {
  extern __shared__ __align__(16) float mp_smem[];
  float* stage = mp_smem;                                                       // 2 x (At | Bs)
  float (*S)[kSC + 1] = reinterpret_cast<float (*)[kSC + 1]>(mp_smem + 2 * kStageFloats);
  const int n0 = blockIdx.x * kTR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (M + kSC - 1) / kSC;

  // ---- pass 1: per-row sorted top-k list, lane i holds the i-th largest affinity seen so far
  float list[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) list[rr] = -INFINITY;
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    aff_tile(tar, src, N, C, M, ldn, ldm, n0, chunk, temperature, stage, S);
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
    aff_tile(tar, src, N, C, M, ldn, ldm, n0, chunk, temperature, stage, S);
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

static inline int64_t pad4(int64_t v) { return (v + 3) & ~int64_t(3); }

extern "C" int64_t univst_maskprop_workspace_bytes(int32_t N, int32_t C, int32_t M) {
  return (pad4(N) * C + pad4(M) * C) * sizeof(float);
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
  UV_REQUIRE(((uintptr_t)workspace & 15) == 0, "maskprop: workspace must be 16-byte aligned");
  const int ldn = (int)pad4(N), ldm = (int)pad4(M);
  float* tar_n = (float*)workspace;                 // [C][ldn], K-major
  float* src_n = tar_n + (size_t)C * ldn;           // [C][ldm]
  normalize_rows_t_kernel<<<(ldn + 31) / 32, 256, 0, st>>>(feat_tar, N, C, ldn, tar_n);
  UV_CHECK_CUDA(cudaGetLastError());
  normalize_cols_kernel<<<(ldm + 31) / 32, 256, 0, st>>>(feat_src, C, M, ldm, src_n);
  UV_CHECK_CUDA(cudaGetLastError());
  static bool configured = false;
  if (!configured) {
    UV_CHECK_CUDA(cudaFuncSetAttribute(maskprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  maskprop_kernel<<<(N + kTR - 1) / kTR, 256, kSmemBytes, st>>>(tar_n, src_n, segs, N, C, M, ldn, ldm, Ccls, temperature,
                                                               topk, segs_tar, thresholds, kept_idx, kept_cap);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

// HBM-bound normalisation kernels (channels-last fp16 activations, fp32 statistics).
//
//   GroupNorm   : statistics per (batch, group) over `rows` tokens x (C / groups) channels.  With rows = F*H*W
//                 and batch = branch this is the reference's GroupNorm on a 5-D tensor whose statistics span all
//                 frames (resnet.py:338,369; unet_3d_condition.py:439); with rows = H*W and batch = image it is
//                 the per-frame GroupNorm of the transformer (attention.py:121).  The input may be the channel
//                 concat of two tensors (unet_3d_blocks.py:523,618) -- the concat only ever exists as this
//                 kernel's normalised output.
//   LayerNorm   : per token over C (attention.py:290,312,329).
#include "xrank.cuh"

namespace uv {

static constexpr int kGnMaxChunks = 512;
static constexpr int kGnUnroll = 4;   // rows in flight per thread

__device__ __forceinline__ const uint4* gn_src(const __half* x1, const __half* x2, int C1, int C2, size_t row, int v) {
  // v-th 8-channel vector of the concatenated row
  const int c = v * 8;
  return (c < C1) ? reinterpret_cast<const uint4*>(x1 + row * C1 + c)
                  : reinterpret_cast<const uint4*>(x2 + row * C2 + (c - C1));
}

// Thread mapping shared by the statistics and the apply kernel: block = nvec * rows_par threads, thread (rl, v) owns
// the v-th 8-channel vector of rows rl, rl + rows_par, ... of its row range -- a warp reads whole contiguous rows, the
// per-channel constants of a thread never change, and kGnUnroll independent 16-byte loads are in flight per thread.

// arrival counters of the statistics kernels (zero at module load, re-armed by the last block of every launch); the host
// hands consecutive launches different slots
static constexpr int kGnFusedChunks = 24;     // a fused fold adds at most this many chunk partials per entry
static constexpr int kGnMaxFusedNB = 255;     // per-entry counters [0, 255), the all-entries counter at [255]
__device__ unsigned g_gn_arrivals[64 * 256];

// grid (nchunks, NB).  partial[b][chunk][group][2] = (sum, sumsq); the LAST block to finish folds the chunk partials of every
// batch entry into sums[b][group][2] in a fixed order (deterministic, no float atomics) -- and, when the statistics span
// the rows of other ranks (world > 1: frame-sharded resnet.py:338,369), exchanges them through the ranks' control blocks
// (xrank.cuh) and adds the ranks' sums in rank order: one kernel where there used to be a statistics, a fold and a
// collective launch.
__global__ void gn_stats_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, int C1, int C2, int rows,
                                int groups, int nvec, int rows_par, int rows_per_chunk, float* __restrict__ partial,
                                float* __restrict__ sums, unsigned* __restrict__ arrivals, XrankPeers P, int rank, int world) {
  extern __shared__ float sh[];  // [threads][8] per-thread pair sums, then [groups][2]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int C = C1 + C2, cpg = C / groups;
  const int v = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const int r_begin = chunk * rows_per_chunk, r_end = min(rows, r_begin + rows_per_chunk);
  const uint4* src = gn_src(x1, x2, C1, C2, (size_t)b * rows, v);
  const size_t rstride = (size_t)((v * 8 < C1) ? C1 : C2) / 8;   // uint4 per row of the source this thread reads
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (int r = r_begin + rl; r < r_end; r += rows_par * kGnUnroll) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int rr = r + k * rows_par;
      u[k] = make_uint4(0, 0, 0, 0);
      if (rr < r_end) u[k] = __ldg(src + (size_t)rr * rstride);
    }
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_half2(w[j]);
        s[j] += f.x + f.y;
        q[j] = fmaf(f.x, f.x, fmaf(f.y, f.y, q[j]));
      }
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
  if (!arrivals) return;   // big tensors: a separate, wider fold kernel follows (gn_launch_stats)
  // ---- the last block of batch entry b folds b's chunk partials (<= kGnFusedChunks of them: one round of independent
  // loads, then a fixed-order sum); with statistics that span other ranks, the block that completes the last entry exchanges
  __shared__ int s_last, s_last_all;
  __threadfence();
  __syncthreads();
  const int NB = gridDim.y, nchunks = gridDim.x, G2 = groups * 2;
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(arrivals + b, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *reinterpret_cast<volatile unsigned*>(arrivals + b) = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int parts = blockDim.x >> 6;                 // blockDim >= 128: at least two interleaved partial sums per entry
  const int part = threadIdx.x >> 6, tt = threadIdx.x & 63;
  float* fold = sh;                                   // [parts][64]
  float* dst = (world > 1) ? partial + ((size_t)NB * nchunks) * G2 : sums;   // sharded: parked behind the partials first
  for (int i0 = 0; i0 < G2; i0 += 64) {
    const int i = i0 + tt;
    float v[kGnFusedChunks / 2];
#pragma unroll
    for (int k = 0; k < kGnFusedChunks / 2; ++k) {
      const int c = part + k * parts;
      v[k] = (part < parts && i < G2 && c < nchunks) ? __ldcg(partial + ((size_t)b * nchunks + c) * G2 + i) : 0.0f;
    }
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < kGnFusedChunks / 2; ++k) acc += v[k];
    if (part < parts) fold[part * 64 + tt] = acc;
    __syncthreads();
    if (part == 0 && i < G2) {
      float r = fold[tt];
      for (int k = 1; k < parts; ++k) r += fold[k * 64 + tt];
      dst[(size_t)b * G2 + i] = r;
    }
    __syncthreads();
  }
  if (world <= 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(arrivals + kGnMaxFusedNB, 1u);
    s_last_all = (t == (unsigned)NB - 1);
    if (s_last_all) *reinterpret_cast<volatile unsigned*>(arrivals + kGnMaxFusedNB) = 0;
  }
  __syncthreads();
  if (!s_last_all) return;
  __threadfence();
  uint32_t* ctl = P.ctl[rank];
  const uint32_t par = (*reinterpret_cast<volatile uint32_t*>(ctl + kXrEpoch) + 1) & 1u;
  const int n = NB * G2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = __ldcg(dst + i);
    for (int r = 0; r < world; ++r)
      reinterpret_cast<float*>(P.ctl[r] + kXrSlots)[((size_t)par * kXrankMaxRanks + rank) * kXrankSlotFloats + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < 32) xrank_sync_warp(P, rank, world);
  __syncthreads();
  const float* slots = reinterpret_cast<const float*>(ctl + kXrSlots) + (size_t)par * kXrankMaxRanks * kXrankSlotFloats;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.0f;
    for (int r = 0; r < world; ++r) acc += __ldcg(slots + (size_t)r * kXrankSlotFloats + i);
    sums[i] = acc;
  }
}

// Fold of a big tensor's chunk partials (hundreds of chunks: too long a dependent chain for the last block of the statistics
// kernel): sums[b][g][2] = sum over the chunk partials in a fixed order, 16 interleaved partial sums per entry combined by
// a fixed tree.  grid NB, block 16 * 64.
__global__ void gn_fold_kernel(const float* __restrict__ partial, int nchunks, int groups, float* __restrict__ sums) {
  __shared__ float sh[16][64];
  const int b = blockIdx.x;
  const int part = threadIdx.x >> 6, t = threadIdx.x & 63;
  for (int i0 = 0; i0 < groups * 2; i0 += 64) {
    const int i = i0 + t;
    float acc = 0.0f;
    if (i < groups * 2)
      for (int c = part; c < nchunks; c += 16) acc += partial[((size_t)b * nchunks + c) * groups * 2 + i];
    sh[part][t] = acc;
    __syncthreads();
    if (part == 0 && i < groups * 2) {
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = sh[k][t];
#pragma unroll
      for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int k = 0; k < w; ++k) v[k] += v[k + w];
      sums[(size_t)b * groups * 2 + i] = v[0];
    }
    __syncthreads();
  }
}

// The same for statistics that span the rows of other ranks: fold, store the NB x groups x 2 sums into slot [rank] of every
// rank, synchronise, add the slots in rank order.  One block of 1024 threads.
__global__ void gn_fold_xrank_kernel(const float* __restrict__ partial, int nchunks, int nchunks_stride, int NB, int groups,
                                     float* __restrict__ sums, XrankPeers P, int rank, int world) {
  __shared__ float sh[16][64];
  __shared__ float local[kXrankSlotFloats];
  uint32_t* mine = P.ctl[rank];
  const uint32_t par = (*reinterpret_cast<volatile uint32_t*>(mine + kXrEpoch) + 1) & 1u;
  const int part = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int G2 = groups * 2;
  for (int b = 0; b < NB; ++b)
    for (int i0 = 0; i0 < G2; i0 += 64) {
      const int i = i0 + t;
      float acc = 0.0f;
      if (i < G2)
        for (int c = part; c < nchunks; c += 16) acc += partial[((size_t)b * nchunks_stride + c) * G2 + i];
      sh[part][t] = acc;
      __syncthreads();
      if (part == 0 && i < G2) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = sh[k][t];
#pragma unroll
        for (int w = 8; w > 0; w >>= 1)
#pragma unroll
          for (int k = 0; k < w; ++k) v[k] += v[k + w];
        local[b * G2 + i] = v[0];
      }
      __syncthreads();
    }
  const int n = NB * G2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = local[i];
    for (int r = 0; r < world; ++r)
      reinterpret_cast<float*>(P.ctl[r] + kXrSlots)[((size_t)par * kXrankMaxRanks + rank) * kXrankSlotFloats + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < 32) xrank_sync_warp(P, rank, world);
  __syncthreads();
  const float* slots = reinterpret_cast<const float*>(mine + kXrSlots) + (size_t)par * kXrankMaxRanks * kXrankSlotFloats;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.0f;
    for (int r = 0; r < world; ++r) acc += __ldcg(slots + (size_t)r * kXrankSlotFloats + i);
    sums[i] = acc;
  }
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// grid (row blocks, NB), block = nvec * rows_par.  y = x * a + b with a = rstd * gamma, b = beta - mean * a held in
// registers for the thread's 8 channels.
__global__ void gn_apply_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, int C1, int C2, int rows,
                                int groups, int nvec, int rows_par, const float* __restrict__ sums,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps, int silu,
                                int rows_per_block, float stat_rows, __half* __restrict__ y) {
  const int b = blockIdx.y;
  const int C = C1 + C2, cpg = C / groups;
  const int v = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  float ca[8], cb[8];
  {
    const uint4 gm = __ldg(reinterpret_cast<const uint4*>(gamma + v * 8));
    const uint4 bt = __ldg(reinterpret_cast<const uint4*>(beta + v * 8));
    const uint32_t gw[4] = {gm.x, gm.y, gm.z, gm.w}, bw[4] = {bt.x, bt.y, bt.z, bt.w};
    const float n = stat_rows * (float)cpg;  // elements the statistics were taken over (all ranks when sharded)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (v * 8 + j * 2) / cpg;   // cpg is even: a channel pair never straddles two groups
      const float sm = __ldg(sums + ((size_t)b * groups + g) * 2), sq = __ldg(sums + ((size_t)b * groups + g) * 2 + 1);
      const float mean = sm / n;
      const float rstd = rsqrtf(fmaxf(sq / n - mean * mean, 0.0f) + eps);
      const float2 ga = unpack_half2(gw[j]), be = unpack_half2(bw[j]);
      ca[2 * j] = rstd * ga.x;
      ca[2 * j + 1] = rstd * ga.y;
      cb[2 * j] = fmaf(-mean, ca[2 * j], be.x);
      cb[2 * j + 1] = fmaf(-mean, ca[2 * j + 1], be.y);
    }
  }
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  const uint4* src = gn_src(x1, x2, C1, C2, (size_t)b * rows, v);
  const size_t rstride = (size_t)((v * 8 < C1) ? C1 : C2) / 8;
  uint4* dst = reinterpret_cast<uint4*>(y + (size_t)b * rows * C) + v;
  const size_t dstride = (size_t)C / 8;
  for (int r = r_begin + rl; r < r_end; r += rows_par * kGnUnroll) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int rr = r + k * rows_par;
      if (rr < r_end) u[k] = __ldg(src + (size_t)rr * rstride);
    }
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int rr = r + k * rows_par;
      if (rr >= r_end) break;
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_half2(w[j]);
        float a = fmaf(f.x, ca[2 * j], cb[2 * j]);
        float c = fmaf(f.y, ca[2 * j + 1], cb[2 * j + 1]);
        if (silu) {
          // the reference rounds the GroupNorm output to fp16 before the activation
          const float2 h = unpack_half2(pack_half2(a, c));
          a = silu_f(h.x);
          c = silu_f(h.y);
        }
        o[j] = pack_half2(a, c);
      }
      dst[(size_t)rr * dstride] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// One warp per token row, NV = ceil(C / 256) 16-byte vectors per lane; single pass over registers: sum and sum of
// squares share one shuffle tree.
// mod_rpg > 0 (adaLN modulation of the SD3 MMDiT blocks, diffusers AdaLayerNormZero / AdaLayerNormContinuous): the norm has
// no affine parameters; gamma / beta are per-SAMPLE rows [rows / mod_rpg, mod_ld] and y = norm(x) * (1 + gamma) + beta.
template <int NV>
__global__ void layernorm_kernel(const __half* __restrict__ x, int rows, int C, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, float eps, __half* __restrict__ y, int mod_rpg, int mod_ld) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float one = 0.0f;
  if (mod_rpg > 0) {
    gamma += (size_t)(row / mod_rpg) * mod_ld;
    beta += (size_t)(row / mod_rpg) * mod_ld;
    one = 1.0f;
  }
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  uint4 buf[NV];
  float s = 0.0f, q = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    buf[i] = make_uint4(0, 0, 0, 0);
    if (v < nvec) buf[i] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)row * C + v * 8));
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t w[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_half2(w[j]);
      s += f.x + f.y;
      q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const float mean = s / (float)C;
  const float rstd = rsqrtf(fmaxf(q / (float)C - mean * mean, 0.0f) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
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
        o[j] = pack_half2((f.x - mean) * rstd * (ga.x + one) + be.x, (f.y - mean) * rstd * (ga.y + one) + be.y);
      }
      *reinterpret_cast<uint4*>(y + (size_t)row * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

struct GnPlan {
  int nvec, rows_par, threads, nchunks, rows_per_chunk, rows_per_block, row_blocks;
  bool fused;   // the statistics kernel's last block folds (small tensors: few, fat chunks)
};
static GnPlan gn_plan(int C, int rows, int NB = 1) {
  GnPlan g;
  g.nvec = C / 8;
  g.rows_par = g.nvec >= 256 ? 1 : 256 / g.nvec;
  g.threads = g.nvec * g.rows_par;
  // statistics: about 32 row-steps of kGnUnroll rows per thread and chunk
  g.nchunks = (rows + g.rows_par * kGnUnroll * 8 - 1) / (g.rows_par * kGnUnroll * 8);
  if (g.nchunks > kGnMaxChunks) g.nchunks = kGnMaxChunks;
  if (g.nchunks < 1) g.nchunks = 1;
  // small statistics spans (<= 8 MB per batch entry: a frame shard, the deep levels, per-frame norms): at most 24 fat
  // chunks, folded by the entry's last block -- two launches per GroupNorm instead of three.  The plan is a function of
  // (C, rows) ALONE: the summation order of an entry's statistics must not depend on how many entries share the launch
  // (a one-branch call must reproduce the edit branch of a three-branch call bit for bit; a frame shard its frames).
  g.fused = (size_t)rows * C * 2 <= (8u << 20) && NB < kGnMaxFusedNB;
  if (g.fused && g.nchunks > kGnFusedChunks) g.nchunks = kGnFusedChunks;
  g.rows_per_chunk = (rows + g.nchunks - 1) / g.nchunks;
  // apply: a few unrolled steps per thread and block
  g.rows_per_block = g.rows_par * kGnUnroll * 4;
  g.row_blocks = (rows + g.rows_per_block - 1) / g.rows_per_block;
  return g;
}

static int gn_launch_stats(const void* X1, const void* X2, int C1, int C2, int NB, int rows, int groups, float* sums,
                           void* workspace, cudaStream_t st, const XrankPeers* peers = nullptr, int rank = 0, int world = 0) {
  const GnPlan g = gn_plan(C1 + C2, rows, NB);
  UV_REQUIRE(g.threads >= 128, "groupnorm: at least 128 threads per block (channels %% 8, rows_par) -- internal plan error");
  UV_REQUIRE(world <= 1 || NB * groups * 2 <= kXrankSlotFloats, "groupnorm: NB * groups * 2 exceeds the exchange slot (%d floats)",
             kXrankSlotFloats);
  static unsigned next_slot = 0;
  static unsigned* arrivals = nullptr;
  if (!arrivals) UV_CHECK_CUDA(cudaGetSymbolAddress((void**)&arrivals, g_gn_arrivals));
  const unsigned slot = ((next_slot++) & 63u) * 256u;
  XrankPeers none{};
  gn_stats_kernel<<<dim3(g.nchunks, NB), g.threads, g.threads * 8 * sizeof(float), st>>>(
      (const __half*)X1, (const __half*)X2, C1, C2, rows, groups, g.nvec, g.rows_par, g.rows_per_chunk, (float*)workspace, sums,
      g.fused ? arrivals + slot : nullptr, peers ? *peers : none, rank, world);
  UV_CHECK_CUDA(cudaGetLastError());
  if (!g.fused) {
    if (world > 1)
      gn_fold_xrank_kernel<<<1, 1024, 0, st>>>((const float*)workspace, g.nchunks, g.nchunks, NB, groups, sums, *peers, rank, world);
    else
      gn_fold_kernel<<<NB, 1024, 0, st>>>((const float*)workspace, g.nchunks, groups, sums);
    UV_CHECK_CUDA(cudaGetLastError());
  }
  return UNIVST_OK;
}

static int gn_launch_apply(const void* X1, const void* X2, int C1, int C2, int NB, int rows, int groups, const float* sums,
                           int64_t stat_rows, const void* gamma, const void* beta, float eps, int silu, void* Y,
                           cudaStream_t st) {
  const GnPlan g = gn_plan(C1 + C2, rows);
  gn_apply_kernel<<<dim3(g.row_blocks, NB), g.threads, 0, st>>>(
      (const __half*)X1, (const __half*)X2, C1, C2, rows, groups, g.nvec, g.rows_par, sums, (const __half*)gamma,
      (const __half*)beta, eps, silu, g.rows_per_block, (float)stat_rows, (__half*)Y);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

// entry points for the cross-rank form (xrank.cu): statistics + exchange / apply only / shape check / folded-sum location
int gn_stats_xrank(const void* X1, const void* X2, int C1, int C2, int NB, int rows, int groups, float* sums, void* workspace,
                   const XrankPeers& peers, int rank, int world, cudaStream_t st) {
  return gn_launch_stats(X1, X2, C1, C2, NB, rows, groups, sums, workspace, st, &peers, rank, world);
}
int gn_apply_launch(const void* X1, const void* X2, int C1, int C2, int NB, int rows, int groups, const float* sums,
                    int64_t stat_rows, const void* gamma, const void* beta, float eps, int silu, void* Y, cudaStream_t st) {
  return gn_launch_apply(X1, X2, C1, C2, NB, rows, groups, sums, stat_rows, gamma, beta, eps, silu, Y, st);
}
float* gn_sums_of(void* workspace, int NB, int groups) { return (float*)workspace + (size_t)NB * kGnMaxChunks * groups * 2; }

}  // namespace uv

using namespace uv;

// workspace: [NB][kGnMaxChunks][groups][2] chunk partials followed by [NB][groups][2] folded sums
extern "C" int64_t univst_groupnorm_workspace_bytes(int32_t NB, int32_t groups) {
  return (int64_t)NB * (kGnMaxChunks + 1) * groups * 2 * sizeof(float);
}

namespace uv {
int gn_check_shape(const void* X1, const void* X2, int32_t& C1, int32_t& C2, int32_t NB, int32_t rows, int32_t groups);
}
static int gn_check(const void* X1, const void* X2, int32_t& C1, int32_t& C2, int32_t NB, int32_t rows, int32_t groups) {
  return uv::gn_check_shape(X1, X2, C1, C2, NB, rows, groups);
}
int uv::gn_check_shape(const void* X1, const void* X2, int32_t& C1, int32_t& C2, int32_t NB, int32_t rows, int32_t groups) {
  if (!X2) C2 = 0;
  const int C = C1 + C2;
  UV_REQUIRE(NB > 0 && rows > 0 && groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: bad shape");
  UV_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && (C / groups) % 2 == 0, "groupnorm: channels %% 8, channels per group even");
  UV_REQUIRE(C / 8 <= 1024, "groupnorm: at most 8192 channels");
  return UNIVST_OK;
}

extern "C" int univst_groupnorm_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                                    int32_t groups, const void* gamma, const void* beta, float eps, int32_t silu,
                                    void* Y, void* workspace, void* stream) {
  UV_REQUIRE(X1 && Y && gamma && beta && workspace, "groupnorm: null pointer");
  int r = gn_check(X1, X2, C1, C2, NB, rows, groups);
  if (r) return r;
  float* sums = (float*)workspace + (size_t)NB * kGnMaxChunks * groups * 2;
  r = gn_launch_stats(X1, X2, C1, C2, NB, rows, groups, sums, workspace, (cudaStream_t)stream);
  if (r) return r;
  return gn_launch_apply(X1, X2, C1, C2, NB, rows, groups, sums, rows, gamma, beta, eps, silu, Y, (cudaStream_t)stream);
}

// Split form for frame-sharded execution: local (sum, sum of squares) per (batch, group) -> [all-reduce over ranks by
// the caller] -> apply with the global row count.
extern "C" int univst_groupnorm_stats_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                                          int32_t groups, float* sums, void* workspace, void* stream) {
  UV_REQUIRE(X1 && sums && workspace, "groupnorm_stats: null pointer");
  int r = gn_check(X1, X2, C1, C2, NB, rows, groups);
  if (r) return r;
  return gn_launch_stats(X1, X2, C1, C2, NB, rows, groups, sums, workspace, (cudaStream_t)stream);
}

extern "C" int univst_groupnorm_apply_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                                          int32_t groups, const float* sums, int64_t stat_rows, const void* gamma,
                                          const void* beta, float eps, int32_t silu, void* Y, void* stream) {
  UV_REQUIRE(X1 && Y && gamma && beta && sums, "groupnorm_apply: null pointer");
  int r = gn_check(X1, X2, C1, C2, NB, rows, groups);
  if (r) return r;
  UV_REQUIRE(stat_rows >= rows, "groupnorm_apply: stat_rows < rows");
  return gn_launch_apply(X1, X2, C1, C2, NB, rows, groups, sums, stat_rows, gamma, beta, eps, silu, Y, (cudaStream_t)stream);
}

static int layernorm_launch(const void* X, int32_t rows, int32_t C, const void* gamma, const void* beta, float eps, void* Y,
                            int mod_rpg, int mod_ld, void* stream);

extern "C" int univst_layernorm_f16(const void* X, int32_t rows, int32_t C, const void* gamma, const void* beta,
                                    float eps, void* Y, void* stream) {
  return layernorm_launch(X, rows, C, gamma, beta, eps, Y, 0, 0, stream);
}

extern "C" int univst_layernorm_modulate_f16(const void* X, int32_t rows, int32_t C, const void* scale, const void* shift,
                                             int32_t ld, int32_t rows_per_sample, float eps, void* Y, void* stream) {
  UV_REQUIRE(rows_per_sample > 0 && ld >= C && ld % 8 == 0 && ((uintptr_t)scale & 15) == 0 && ((uintptr_t)shift & 15) == 0,
             "layernorm_modulate: per-sample scale / shift rows must be 16-byte aligned with a row stride %% 8");
  return layernorm_launch(X, rows, C, scale, shift, eps, Y, rows_per_sample, ld, stream);
}

static int layernorm_launch(const void* X, int32_t rows, int32_t C, const void* gamma, const void* beta, float eps, void* Y,
                            int mod_rpg, int mod_ld, void* stream) {
  UV_REQUIRE(X && Y && gamma && beta, "layernorm: null pointer");
  UV_REQUIRE(rows > 0 && C % 8 == 0 && C <= 2048, "layernorm: C must be a multiple of 8, at most 2048");
  const int warps = 8;
  const dim3 grid((rows + warps - 1) / warps);
  const __half *x = (const __half*)X, *g = (const __half*)gamma, *b = (const __half*)beta;
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = (C / 8 + 31) / 32;
  switch (nv) {
    case 1: layernorm_kernel<1><<<grid, warps * 32, 0, st>>>(x, rows, C, g, b, eps, (__half*)Y, mod_rpg, mod_ld); break;
    case 2: layernorm_kernel<2><<<grid, warps * 32, 0, st>>>(x, rows, C, g, b, eps, (__half*)Y, mod_rpg, mod_ld); break;
    case 3: layernorm_kernel<3><<<grid, warps * 32, 0, st>>>(x, rows, C, g, b, eps, (__half*)Y, mod_rpg, mod_ld); break;
    case 4: layernorm_kernel<4><<<grid, warps * 32, 0, st>>>(x, rows, C, g, b, eps, (__half*)Y, mod_rpg, mod_ld); break;
    case 5: layernorm_kernel<5><<<grid, warps * 32, 0, st>>>(x, rows, C, g, b, eps, (__half*)Y, mod_rpg, mod_ld); break;
    default: layernorm_kernel<8><<<grid, warps * 32, 0, st>>>(x, rows, C, g, b, eps, (__half*)Y, mod_rpg, mod_ld); break;
  }
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

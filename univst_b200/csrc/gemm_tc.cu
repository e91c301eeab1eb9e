// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M, N_out] = epilogue( A[M, K] * W[N, K]^T )        fp16 operands, fp32 accumulation in TMEM
//
// One CTA owns one 128 x BN output tile.  Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM
// allocator + tcgen05.mma issuer (one lane), warps 2..9 = epilogue (TMEM -> registers -> global).
// Operand tiles are 128B-swizzled K-major [rows][64 halves]; the A tile of the convolution is a 4-D TMA
// box over the NHWC activation shifted by the filter tap, with out-of-bounds zero fill standing in for
// the padding -- no im2col buffer exists.  The skip-connection concat of the UNet up blocks is a second
// tensor map consumed by the same K loop.
//
// Reference call sites replaced: see include/univst_b200.h (univst_gemm_f16 / univst_conv3x3_f16).
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "host_util.h"
#include "ptx.cuh"

namespace uv {

static constexpr int kBM = 128;   // rows per tile (UMMA_M, cta_group::1)
static constexpr int kBK = 64;    // halves per k-block = one 128-byte swizzle row
static constexpr int kEpiWarps = 8;  // two per TMEM lane quadrant, each draining every other 32-column chunk
static constexpr int kThreads = 64 + 32 * kEpiWarps;

struct GemmParams {
  int M, N, N_out, BN;
  int num_kb;           // k-blocks of 64
  int mode;             // 0 = linear A[M,K]; 1 = conv taps over NHWC; 2 = conv, row-block tiles (any H, W)
  int kb_split;         // linear: first k-block served by A2; conv: channel blocks per tap served by A (cbs1)
  int K1;               // linear: K columns in A; conv: channels in A (C1)
  int cbs;              // conv: channel blocks per tap (cbs1 + cbs2)
  int Cin;              // conv: total input channels (weight K index = tap * Cin + channel)
  int HW, W;            // conv: output pixels per image, output width (== box width)
  int plane_stride;     // conv: images per parity plane (stride-2 input was rearranged into 4 planes)
  // mode 2: a tile is `gen_th` full-width rows of one image (`gen_rb` such row blocks per image) or, for images of
  // <= 128 pixels, `gen_bn` whole images; only its first gen_th * W * gen_bn rows are real (the rest are masked).
  // Rows wider than 128 pixels (gen_seg > 1) are cut into gen_seg segments of 128: a tile is one segment of one row.
  int gen_th, gen_rb, gen_bn, gen_tiles_m, gen_seg;
  int8_t tap_dy[9], tap_dx[9], tap_plane[9];
  int stages;
  int cluster;          // 2: CTA pairs share every weight tile (each loads half of it and multicasts), 1: independent CTAs
  // split-K (few output tiles, long reductions: the deep UNet levels of a frame shard): a work item is (tile, K slice);
  // every slice parks its fp32 accumulator in `part`, then each slice adds up its share of the tile in slice order
  // (deterministic) and runs the epilogue on it
  int splits, kb_per_split;
  float* part;          // [tile][slice][epilogue warp][chunk][8][32] float4 (split_epilogue)
  unsigned* cnt;        // [tile][kEpiWarps][2] parked / done counters, zero between launches
  // epilogue
  const __half* bias;
  const __half* rowvec;
  int rows_per_group;
  int rowvec_ld;
  int act;              // 1: SiLU applied to the fp16-rounded result
  const __half* residual;
  int ldr;
  const __half* bias2;
  int geglu;
  float out_scale;
  __half* D;
  int ldd;
};

// Exact-form (erf) GELU with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the fp16 rounding of
// the result): MUFU.RCP + MUFU.EX2 + a degree-5 Horner chain instead of libdevice's branchy erff.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// tanh-form GELU (diffusers FeedForward(activation_fn="gelu-approximate"), the SD3 MMDiT blocks):
// 0.5 x (1 + tanh(sqrt(2 / pi) (x + 0.044715 x^3))), tanh through one exponential
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
  const float e = __expf(2.0f * u);
  const float th = 1.0f - 2.0f / (e + 1.0f);     // tanh(u); e = inf -> 1, e = 0 -> -1
  return 0.5f * x * (1.0f + th);
}

// Per-warp staging tile in shared memory: 32 rows x 32 halves (64 B per row), the 16-byte piece index XOR-swizzled with
// (row >> 1) & 3 so that both access patterns below are bank-conflict free:
//   "own row"   : thread r touches row r, pieces 0..3 (the TMEM layout: one accumulator row per thread)
//   "coalesced" : lane l touches row (l >> 2) + 8 i, piece l & 3 -- one instruction covers 8 rows x 64 contiguous bytes
//                 of global memory instead of 32 rows x 16 bytes
static constexpr int kStageTileBytes = 32 * 64;
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, uint32_t row, uint32_t piece) {
  return base + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4);
}

struct EpiCtx {
  uint32_t stg;        // this warp's staging tile (shared address)
  uint32_t lane;
  int row0;            // first row of the warp's 32-row slab
  int M;
};

// coalesced residual fetch of one 32-column chunk: 4 x 16 B per lane (zeros where out of range)
__device__ __forceinline__ void residual_fetch(const GemmParams& p, const EpiCtx& e, int n, uint4 (&r)[4]) {
  const int col = n + (int)(e.lane & 3) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = e.row0 + (int)(e.lane >> 2) + 8 * i;
    r[i] = make_uint4(0, 0, 0, 0);
    if (row < e.M && col < p.N_out) r[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ldr + col));
  }
}

// own-row results (4 x 16 B) -> staging -> coalesced global stores
__device__ __forceinline__ void staged_store(const GemmParams& p, const EpiCtx& e, int n, const uint32_t (&o)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    st_shared_v4(stage_addr(e.stg, e.lane, g), o[g * 4], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]);
  __syncwarp();
  const int col = n + (int)(e.lane & 3) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t rl = (e.lane >> 2) + 8 * i;
    const int row = e.row0 + (int)rl;
    uint32_t w[4];
    ld_shared_v4(stage_addr(e.stg, rl, e.lane & 3), w);
    if (row < e.M && col < p.N_out)
      *reinterpret_cast<uint4*>(p.D + (size_t)row * p.ldd + col) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncwarp();   // the tile is rewritten by the next chunk
}

__device__ __forceinline__ void epilogue_rows(const GemmParams& p, uint32_t taddr, bool row_ok, const __half* rv,
                                              const __half* res, __half* drow, int n0, int tile_n, int chunk0,
                                              const EpiCtx& e, uint64_t* acc_full, uint32_t acc_parity) {
  const bool staged = (p.N_out & 7) == 0;   // whole 8-column groups only: every 16-byte piece is all in or all out
  uint4 rnext[4];
  // the first residual chunk is requested before the accumulator is even complete
  if (!p.geglu && staged && p.residual && n0 + chunk0 * 32 < p.N_out) residual_fetch(p, e, n0 + chunk0 * 32, rnext);
  mbar_wait(acc_full, acc_parity);
  tc_fence_after();
  if (!p.geglu) {
    for (int c0 = chunk0 * 32; c0 < p.BN; c0 += 64) {
      if (n0 + c0 >= p.N_out) break;  // warp-uniform
      uint32_t acc[32];
      tmem_ld32(taddr + c0, acc);
      uint32_t rw[16];
      if (staged && p.residual) {
        // residual: coalesced loads (issued one chunk ahead) -> staging -> own row
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_shared_v4(stage_addr(e.stg, (e.lane >> 2) + 8 * i, e.lane & 3), rnext[i].x, rnext[i].y, rnext[i].z, rnext[i].w);
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g) ld_shared_v4(stage_addr(e.stg, e.lane, g), *reinterpret_cast<uint32_t(*)[4]>(&rw[g * 4]));
        if (c0 + 64 < p.BN && n0 + c0 + 64 < p.N_out) residual_fetch(p, e, n0 + c0 + 64, rnext);
      }
      // the chunk's bias (the same 4 x 8 columns in every lane) is requested before the TMEM load is awaited: issued
      // right before use, each of these loads stalled the warp for an L1 / L2 round trip
      uint4 bq[4], vq[4];
      if (staged && p.bias) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          bq[g] = make_uint4(0, 0, 0, 0);
          if (n0 + c0 + g * 8 < p.N_out) bq[g] = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + c0 + g * 8));
        }
      }
      if (staged && rv) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          vq[g] = make_uint4(0, 0, 0, 0);
          if (n0 + c0 + g * 8 < p.N_out) vq[g] = __ldg(reinterpret_cast<const uint4*>(rv + n0 + c0 + g * 8));
        }
      }
      tc_wait_ld();
      if (staged) {
        uint32_t o[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + c0 + g * 8;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]);
          if (n < p.N_out) {
            if (p.bias) {
              const uint32_t bw[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = unpack_half2(bw[j]);
                v[2 * j] += f.x;
                v[2 * j + 1] += f.y;
              }
            }
            if (rv) {
              const uint32_t bw[4] = {vq[g].x, vq[g].y, vq[g].z, vq[g].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = unpack_half2(bw[j]);
                v[2 * j] += f.x;
                v[2 * j + 1] += f.y;
              }
            }
            if (p.residual) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = unpack_half2(rw[g * 4 + j]);
                v[2 * j] += f.x;
                v[2 * j + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) o[g * 4 + j] = pack_half2(v[2 * j] * p.out_scale, v[2 * j + 1] * p.out_scale);
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 y = unpack_half2(o[g * 4 + j]);
              o[g * 4 + j] = pack_half2(y.x / (1.0f + __expf(-y.x)), y.y / (1.0f + __expf(-y.y)));
            }
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 y = unpack_half2(o[g * 4 + j]);
              o[g * 4 + j] = pack_half2(gelu_tanh(y.x), gelu_tanh(y.y));
            }
          }
          if (p.bias2 && n < p.N_out) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias2 + n));
            const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f = unpack_half2(bw[j]);
              float2 y = unpack_half2(o[g * 4 + j]);
              o[g * 4 + j] = pack_half2(y.x + f.x, y.y + f.y);
            }
          }
        }
        staged_store(p, e, n0 + c0, o);
        continue;
      }
      // ragged N_out (conv_out: 4 channels): one thread per row, scalar tail
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n0 + c0 + g * 8;
        if (n >= p.N_out) break;
        for (int j = 0; j < 8 && n + j < p.N_out; ++j) {
          float x = __uint_as_float(acc[g * 8 + j]);
          if (p.bias) x += __half2float(p.bias[n + j]);
          if (rv) x += __half2float(rv[n + j]);
          if (res) x += __half2float(res[n + j]);
          __half y = __float2half_rn(x * p.out_scale);
          if (p.act == 1) {
            const float yf = __half2float(y);
            y = __float2half_rn(yf / (1.0f + __expf(-yf)));
          } else if (p.act == 2) {
            y = __float2half_rn(gelu_tanh(__half2float(y)));
          }
          if (p.bias2) y = __float2half_rn(__half2float(y) + __half2float(p.bias2[n + j]));
          drow[n + j] = y;
        }
      }
    }
  } else {
    // GEGLU: tile columns [0, BN/2) are values, [BN/2, BN) the matching gates (weights packed that way).
    const int half_bn = p.BN >> 1;
    const int no0 = tile_n * half_bn;
    for (int c0 = chunk0 * 32; c0 < half_bn; c0 += 64) {
      if (no0 + c0 >= p.N_out) break;
      uint32_t av[32], ag[32];
      tmem_ld32(taddr + c0, av);
      tmem_ld32(taddr + half_bn + c0, ag);
      tc_wait_ld();
      uint32_t o[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = c0 + g * 8;
        // fp32 all the way: value * gelu(gate) with one rounding at the end (the reference rounds the projection
        // and the GELU to fp16 first; skipping that is both cheaper and closer to the fp32 oracle)
        uint4 bv = make_uint4(0, 0, 0, 0), bg = make_uint4(0, 0, 0, 0);
        if (p.bias && no0 + c < p.N_out) {
          bv = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + c));
          bg = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + half_bn + c));
        }
        const uint32_t bvw[4] = {bv.x, bv.y, bv.z, bv.w}, bgw[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fv = unpack_half2(bvw[j]), fg = unpack_half2(bgw[j]);
          const float v0 = (__uint_as_float(av[g * 8 + 2 * j]) + fv.x) * gelu_erf(__uint_as_float(ag[g * 8 + 2 * j]) + fg.x);
          const float v1 =
              (__uint_as_float(av[g * 8 + 2 * j + 1]) + fv.y) * gelu_erf(__uint_as_float(ag[g * 8 + 2 * j + 1]) + fg.y);
          o[g * 4 + j] = pack_half2(v0, v1);
        }
      }
      staged_store(p, e, no0 + c0, o);   // GEGLU outputs are whole 8-column groups (N % 128 == 0)
    }
  }
}

// Split-K epilogue of one epilogue warp (sub-block j = its 32 accumulator rows x every other 32-column chunk).
//  1. park: the slice's fp32 accumulator chunks go to the workspace in a warp-coalesced layout -- chunk ci, 4-column group
//     jj, row lane -> float4 at ((ci * 8 + jj) * 32 + lane): every store / load instruction moves 512 contiguous bytes;
//  2. arrive on the sub-block's counter and wait until all S slices have parked it (all work items of a split launch are
//     co-resident: the host never creates more of them than there are SMs);
//  3. reduce-scatter: the (chunks x 8) float4 rows of the sub-block are dealt round-robin to the S slices; each slice adds
//     ITS rows over all slices in slice order (deterministic), applies the epilogue to the 4 columns x 32 rows it now
//     holds and stores them -- the reduction traffic is spread over all the CTAs of the launch instead of funnelling
//     S x 128 x BN floats through the one SM that happened to finish last;
//  4. a second counter tells when every slice is done reading; the last one re-arms both counters for the next launch.
__device__ __forceinline__ void split_epilogue(const GemmParams& p, uint32_t taddr, int tile, int slice, int S, int j,
                                               int chunk0, int q, int lane, int m0, int m_lim, int n0, uint64_t* acc_full,
                                               uint32_t acc_parity) {
  const int maxch = (p.BN + 63) >> 6;
  const int nch = (p.BN - chunk0 * 32 + 63) >> 6;                    // chunks this warp owns
  const size_t sub_floats = (size_t)maxch * 1024;                     // one sub-block of one slice
  float* tile_part = p.part + ((size_t)tile * S * kEpiWarps + j) * sub_floats;   // slice s at + s * kEpiWarps * sub_floats
  const size_t slice_stride = (size_t)kEpiWarps * sub_floats;
  mbar_wait(acc_full, acc_parity);
  tc_fence_after();
  {
    float4* mine = reinterpret_cast<float4*>(tile_part + (size_t)slice * slice_stride) + lane;
    for (int ci = 0; ci < nch; ++ci) {
      uint32_t acc[32];
      tmem_ld32(taddr + chunk0 * 32 + ci * 64, acc);
      tc_wait_ld();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        __stcg(mine + (ci * 8 + jj) * 32,
               make_float4(__uint_as_float(acc[4 * jj]), __uint_as_float(acc[4 * jj + 1]), __uint_as_float(acc[4 * jj + 2]),
                           __uint_as_float(acc[4 * jj + 3])));
    }
  }
  __threadfence();
  __syncwarp();
  unsigned* parked = p.cnt + ((size_t)tile * kEpiWarps + j) * 2;
  unsigned* done = parked + 1;
  if (lane == 0) {
    atomicAdd(parked, 1u);
    // (bounded: a launch whose slices are not all resident -- which the host never creates -- gives up after seconds
    // instead of hanging the GPU)
    for (unsigned spins = 0; *reinterpret_cast<volatile unsigned*>(parked) < (unsigned)S && spins < (1u << 23); ++spins) {
    }
    __threadfence();
  }
  __syncwarp();
  const int row = m0 + q * 32 + lane;
  const bool row_ok = row < m_lim;
  const __half* rv = (p.rowvec && row_ok) ? p.rowvec + (size_t)(row / p.rows_per_group) * p.rowvec_ld : nullptr;
  for (int vr = slice; vr < nch * 8; vr += S) {
    const float4* src = reinterpret_cast<const float4*>(tile_part) + (size_t)vr * 32 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int sl = 0;
    for (; sl + 4 <= S; sl += 4) {   // four loads in flight, added in slice order
      const float4 v0 = __ldcg(src + (size_t)sl * (slice_stride / 4));
      const float4 v1 = __ldcg(src + (size_t)(sl + 1) * (slice_stride / 4));
      const float4 v2 = __ldcg(src + (size_t)(sl + 2) * (slice_stride / 4));
      const float4 v3 = __ldcg(src + (size_t)(sl + 3) * (slice_stride / 4));
      acc.x = ((acc.x + v0.x) + v1.x) + v2.x + v3.x;
      acc.y = ((acc.y + v0.y) + v1.y) + v2.y + v3.y;
      acc.z = ((acc.z + v0.z) + v1.z) + v2.z + v3.z;
      acc.w = ((acc.w + v0.w) + v1.w) + v2.w + v3.w;
    }
    for (; sl < S; ++sl) {
      const float4 v = __ldcg(src + (size_t)sl * (slice_stride / 4));
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    const int col = n0 + chunk0 * 32 + (vr >> 3) * 64 + (vr & 7) * 4;
    if (!row_ok || col >= p.N_out) continue;
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    __half out[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (col + c >= p.N_out) break;
      float x = v[c];
      if (p.bias) x += __half2float(__ldg(p.bias + col + c));
      if (rv) x += __half2float(__ldg(rv + col + c));
      if (p.residual) x += __half2float(__ldg(p.residual + (size_t)row * p.ldr + col + c));
      __half y = __float2half_rn(x * p.out_scale);
      if (p.act == 1) {
        const float yf = __half2float(y);
        y = __float2half_rn(yf / (1.0f + __expf(-yf)));
      } else if (p.act == 2) {
        y = __float2half_rn(gelu_tanh(__half2float(y)));
      }
      if (p.bias2) y = __float2half_rn(__half2float(y) + __half2float(__ldg(p.bias2 + col + c)));
      out[c] = y;
    }
    __half* d = p.D + (size_t)row * p.ldd + col;
    if (col + 4 <= p.N_out && (p.ldd & 3) == 0) {
      *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(out);
    } else {
      for (int c = 0; c < 4 && col + c < p.N_out; ++c) d[c] = out[c];
    }
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == (unsigned)S - 1) {   // every slice has finished reading: re-arm for the next launch
      *reinterpret_cast<volatile unsigned*>(parked) = 0;
      *reinterpret_cast<volatile unsigned*>(done) = 0;
    }
  }
}

// Persistent kernel: CTA c processes tiles c, c + gridDim.x, ... (tile = tile_m * tiles_n + tile_n, so the N tiles
// of one A row-block run at the same time on neighbouring CTAs and share A through L2).  Two TMEM accumulators
// alternate between consecutive tiles: the epilogue warps drain tile i while the MMA warp already accumulates
// tile i + 1, and the TMA producer runs ahead across tile boundaries.
// CL = 2 (thread-block clusters of two): the pair works on two consecutive 128-row blocks of the SAME weight tile.  Each
// CTA fetches half of every B k-block and multicasts it into both shared memories, so the weight traffic out of L2 --
// the larger half of what these kernels pull through L2, which is what bounds the 128 x 160 / 224 tiles -- is halved.
// A stage may be refilled only when both CTAs' MMAs have consumed it: the commit that frees a stage is multicast too.
// SPLIT: the split-K variant (work items = (tile, K slice), split_epilogue) is a separate instantiation so that the plain
// kernel's epilogue keeps its register allocation (folded into one kernel it spilled: the K = 320 GEMMs ran 1.7x slower).
template <int CL, bool SPLIT = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB | B BN*128 B], then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = (uint32_t)p.BN * kBK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* stg = smem + (size_t)p.stages * stage_bytes;   // [kEpiWarps] staging tiles of the epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + kEpiWarps * kStageTileBytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int tiles_n = (p.N + p.BN - 1) / p.BN;
  const int tiles_m = (p.mode == 2) ? p.gen_tiles_m : (p.M + kBM - 1) / kBM;
  // work items: CL = 1: tiles (m, n), n fastest, one per CTA; CL = 2: pairs of row blocks (2 mm, 2 mm + 1) x n per cluster
  const uint32_t crank = (CL == 2) ? cluster_ctarank() : 0u;
  // (split-K: work item = tile * splits + slice, slice fastest -- the slices of a tile run side by side)
  const int S = SPLIT ? p.splits : 1;
  const int num_tiles = (CL == 2 ? (tiles_m + 1) / 2 : tiles_m) * tiles_n * S;
  const int tile_first = (CL == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = (CL == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto kb_begin = [&](int work) { return S == 1 ? 0 : (work % S) * p.kb_per_split; };
  auto kb_end = [&](int work) {
    if (S == 1) return p.num_kb;
    const int e = (work % S + 1) * p.kb_per_split;
    return e < p.num_kb ? e : p.num_kb;
  };
  auto tile_m0 = [&](int tile) {   // first output row of the tile; may be >= M for the odd tail
    const int mm = (tile / tiles_n) * CL + (int)crank;
    if (p.mode == 2) {
      if (p.gen_seg > 1) return (mm / p.gen_seg) * p.W + (mm % p.gen_seg) * kBM;   // (image, row) = mm / gen_seg: rows are consecutive
      return (mm / p.gen_rb) * p.gen_bn * p.HW + (mm % p.gen_rb) * p.gen_th * p.W;
    }
    return mm * kBM;
  };
  auto tile_mlim = [&](int m0) {    // one past the last real output row of the tile
    if (p.mode != 2) return p.M;
    if (p.gen_seg > 1) {                                    // a segment ends with its image row
      const int row_end = (m0 / p.W + 1) * p.W;
      return m0 + kBM < row_end ? m0 + kBM : row_end;
    }
    const int group_end = (m0 / p.HW + p.gen_bn) * p.HW;   // the row block must not run into the next image (group)
    int lim = m0 + p.gen_th * p.W * p.gen_bn;
    lim = lim < group_end ? lim : group_end;
    return lim < p.M ? lim : p.M;
  };
  // bytes one stage receives: in mode 2 the A box holds only the real rows
  const uint32_t tx_bytes = (p.mode == 2 && p.gen_seg == 1) ? (uint32_t)(p.gen_th * p.W * p.gen_bn) * (kBK * 2) + b_bytes : stage_bytes;

  // accumulator stride: BN rounded up to a power of two >= 32 (TMEM allocations are powers of two)
  uint32_t acc_cols = 32;
  while ((int)acc_cols < p.BN) acc_cols <<= 1;
  const uint32_t tmem_cols = acc_cols * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);   // one commit per CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 32 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();   // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      uint32_t phase = 0;
      int s = 0;
      for (int work = tile_first; work < num_tiles; work += tile_step) {
        const int tile = work / S;
        const int m0 = tile_m0(tile);
        const int n0 = (tile % tiles_n) * p.BN;
        int cn = 0, cy = 0, cx = 0;
        if (p.mode >= 1) {
          cn = m0 / p.HW;
          cy = (m0 % p.HW) / p.W;
          if (p.mode == 2) cx = m0 % p.W;   // non-zero only for the segments of rows wider than 128 pixels
        }
        for (int kb = kb_begin(work), kbe = kb_end(work); kb < kbe; ++kb) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          mbar_expect_tx(&full_bar[s], tx_bytes);
          int kw;  // K coordinate into the weight matrix
          if (p.mode == 0) {
            kw = kb * kBK;
            if (kb < p.kb_split)
              tma_load_2d(sa, &tmA, &full_bar[s], kb * kBK, m0);
            else
              tma_load_2d(sa, &tmA2, &full_bar[s], kb * kBK - p.K1, m0);
          } else {
            const int tap = kb / p.cbs;
            const int cb = kb - tap * p.cbs;
            const int x = cx + p.tap_dx[tap], y = cy + p.tap_dy[tap], n = cn + p.tap_plane[tap] * p.plane_stride;
            if (cb < p.kb_split) {
              kw = tap * p.Cin + cb * kBK;
              tma_load_4d(sa, &tmA, &full_bar[s], cb * kBK, x, y, n);
            } else {
              kw = tap * p.Cin + p.K1 + (cb - p.kb_split) * kBK;
              tma_load_4d(sa, &tmA2, &full_bar[s], (cb - p.kb_split) * kBK, x, y, n);
            }
          }
          if constexpr (CL == 2) {   // my half of the weight k-block, delivered to both CTAs
            const int half = p.BN >> 1;
            tma_load_2d_mcast(sb + (size_t)crank * half * (kBK * 2), &tmB, &full_bar[s], kw, n0 + (int)crank * half, 0x3);
          } else {
            tma_load_2d(sb, &tmB, &full_bar[s], kw, n0);
          }
          if (++s == p.stages) {
            s = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    // The whole warp runs the loop convergently; lane 0 probes the barriers (voted) and an elected lane issues each
    // tcgen05 instruction, so stage addresses and descriptors stay in uniform registers.  (A plain `if (lane == 0)`
    // region costs an ELECT / R2UR sequence of ~15 dependent instructions per UMMA -- more than a 128 x 160 x 16 MMA
    // takes -- which made every BN < 256 tile issue-bound.)
    const uint32_t idesc = make_idesc_f16(kBM, (uint32_t)p.BN, 0, 0);
    const uint32_t smem0 = smem_u32(smem);
    uint32_t phase = 0;
    int s = 0;
    int it = 0;
    for (int work = tile_first; work < num_tiles; work += tile_step, ++it) {
      const int a = it & 1;
      // wait until the epilogue has drained this accumulator (its (it / 2)-th use)
      mbar_wait_warp(&tmem_empty_bar[a], (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t acc = tmem_base + a * acc_cols;
      const int kb0 = kb_begin(work);
      for (int kb = kb0, kbe = kb_end(work); kb < kbe; ++kb) {
        mbar_wait_warp(&full_bar[s], phase);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
        const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(sa + a_bytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // advance 16 halves = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
          umma_f16_ss_elect(acc, desc_advance(da, k * 2), desc_advance(db, k * 2), idesc, ((kb - kb0) | k) ? 1u : 0u);
        }
        if constexpr (CL == 2) {
          if (elect_one()) tc_commit_mcast(&empty_bar[s], 0x3);   // the stage is free when BOTH CTAs' MMAs have retired
        } else {
          tc_commit_elect(&empty_bar[s]);  // frees the smem stage once these MMAs retire
        }
        if (++s == p.stages) {
          s = 0;
          phase ^= 1;
        }
      }
      tc_commit_elect(&tmem_full_bar[a]);
    }
  } else {
    // -------------------------------------------------- epilogue warps 2..9 (TMEM lane quadrant = warp % 4)
    const uint32_t q = warp & 3;
    const int chunk0 = (int)((warp - 2) >> 2);  // which half of the 32-column chunks this warp drains
    int it = 0;
    for (int work = tile_first; work < num_tiles; work += tile_step, ++it) {
      const int tile = work / S;
      const int a = it & 1;
      const int tile_n = tile % tiles_n;
      const int m0 = tile_m0(tile);
      const int n0 = tile_n * p.BN;
      const int row = m0 + (int)(q * 32 + lane);
      const uint32_t taddr = tmem_base + a * acc_cols + ((q * 32) << 16);
      const int m_lim = tile_mlim(m0);
      const bool row_ok = row < m_lim;
      const __half* rv = (p.rowvec && row_ok) ? p.rowvec + (size_t)(row / p.rows_per_group) * p.rowvec_ld : nullptr;
      const __half* res = (p.residual && row_ok) ? p.residual + (size_t)row * p.ldr : nullptr;
      __half* drow = p.D + (size_t)row * p.ldd;
      EpiCtx e;
      e.stg = smem_u32(stg) + (warp - 2) * kStageTileBytes;
      e.lane = lane;
      e.row0 = m0 + (int)(q * 32);
      e.M = m_lim;
      if constexpr (!SPLIT) {
        epilogue_rows(p, taddr, row_ok, rv, res, drow, n0, tile_n, chunk0, e, &tmem_full_bar[a], (uint32_t)((it >> 1) & 1));
      } else {
        split_epilogue(p, taddr, tile, work % S, S, (int)(warp - 2), chunk0, (int)q, (int)lane, m0, m_lim, n0,
                       &tmem_full_bar[a], (uint32_t)((it >> 1) & 1));
      }
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[a]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int pick_bn(int N, int geglu) {
  // experiments: UNIVST_BN_OVERRIDE="640:256,320:192" forces the tile width for a given N
  if (const char* e = getenv("UNIVST_BN_OVERRIDE")) {
    for (const char* q = e; *q;) {
      char* end;
      const long n = strtol(q, &end, 10);
      if (*end != ':') break;
      const long bn = strtol(end + 1, &end, 10);
      if (n == N && bn >= 32 && bn <= 256 && bn % 32 == 0 && !geglu) return (int)bn;
      q = (*end == ',') ? end + 1 : end;
      if (*end != ',') break;
    }
  }
  if (geglu) return (N % 256 == 0) ? 256 : 128;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 256) return (N + 31) & ~31;
  // Tile-width choice by a measured cost model: one 128 x BN tile costs about (BN + 375) units -- the A side of a tile
  // (TMA fill + operand reads of the 128-row block) is a large fixed cost, so few wide tiles beat many exact ones even
  // when the last tile is ragged.  Measured (profiles/r01_tile_width_sweep.txt): conv 640 -> 640 at 32 x 32: BN 160
  // (4 exact tiles) 377 us, 224 (3 tiles, 5 % padding) 316 us, 256 326 us, 128 436 us; GEMM N = 960: 160 220 us, 192 199 us.
  int best = 256;
  long best_cost = -1;
  for (int bn = 256; bn >= 128; bn -= 32) {   // multiples of 32: the epilogue drains whole 32-column chunks
    const long tiles = (N + bn - 1) / bn;
    const long cost = tiles * (bn + 375);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// CTA pairs with multicast weight tiles (UNIVST_GEMM_CLUSTER=1): at least two row blocks, tile width a multiple of 16
static int pick_cluster(int M, int BN) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("UNIVST_GEMM_CLUSTER");
    enabled = e ? atoi(e) : 0;
  }
  return (enabled && M > kBM && BN % 16 == 0) ? 2 : 1;
}

// split-K workspace: caller-owned device memory registered per stream (univst_gemm_set_workspace); no workspace, no split
static constexpr size_t kSplitCounterBytes = 256 * 1024;
struct SplitWs {
  cudaStream_t stream;
  void* ptr;
  size_t bytes;
};
static std::mutex g_ws_mutex;
static std::vector<SplitWs> g_ws;
static int g_splitk_max_tiles = 0;   // univst_gemm_tune: split the reduction when a launch has at most this many tiles

static void pick_splits(GemmParams& p, int num_tiles, cudaStream_t stream) {
  p.splits = 1;
  p.kb_per_split = p.num_kb;
  p.part = nullptr;
  p.cnt = nullptr;
  if (g_splitk_max_tiles <= 0 || p.cluster != 1 || p.geglu || num_tiles > g_splitk_max_tiles || p.num_kb < 16) return;
  int S = num_sms() / num_tiles;
  if (S > p.num_kb / 4) S = p.num_kb / 4;
  if (S > 32) S = 32;
  // below four slices the fixed cost of the exchange (park, two counters, reduce-scatter: ~10 us of dependent L2 round
  // trips) is not paid back (measured: 48 tiles x 3 slices 38.8 us vs 36 us unsplit; 5 tiles x 26 slices 26 vs 57 us)
  if (S < 4) return;
  SplitWs ws{};
  {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    for (const SplitWs& w : g_ws)
      if (w.stream == stream) ws = w;
  }
  if (!ws.ptr || (size_t)num_tiles * kEpiWarps * 2 * sizeof(unsigned) > kSplitCounterBytes) return;
  const size_t per_slice = (size_t)num_tiles * kEpiWarps * ((p.BN + 63) / 64) * 1024 * sizeof(float);
  const size_t room = ws.bytes - kSplitCounterBytes;
  if ((size_t)S * per_slice > room) S = (int)(room / per_slice);
  if (S < 4) return;
  const int kbs = (p.num_kb + S - 1) / S;
  p.kb_per_split = kbs;
  p.splits = (p.num_kb + kbs - 1) / kbs;   // no empty slice
  p.cnt = (unsigned*)ws.ptr;
  p.part = (float*)((uint8_t*)ws.ptr + kSplitCounterBytes);
}

static int launch(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, GemmParams& p,
                  cudaStream_t stream) {
  const uint32_t stage_bytes = kBM * kBK * 2 + (uint32_t)p.BN * kBK * 2;
  // one persistent CTA per SM: the smem ring is as deep as fits so that the producer prefetches across tiles
  const uint32_t budget = 200 * 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + kEpiWarps * kStageTileBytes + (2 * stages + 4) * sizeof(uint64_t) +
                      16 + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    UV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  const int tiles_m = (p.mode == 2) ? p.gen_tiles_m : (p.M + kBM - 1) / kBM, tiles_n = (p.N + p.BN - 1) / p.BN;
  if (p.cluster == 2) {
    p.splits = 1;
    p.kb_per_split = p.num_kb;
    const int pairs = ((tiles_m + 1) / 2) * tiles_n;
    const int max_pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (pairs < max_pairs ? pairs : max_pairs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    UV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2>, tmA, tmA2, tmB, p));
    return UNIVST_OK;
  }
  const int num_tiles = tiles_m * tiles_n;
  pick_splits(p, num_tiles, stream);
  const int num_work = num_tiles * p.splits;
  const int grid = num_work < num_sms() ? num_work : num_sms();
  if (p.splits > 1)
    gemm_tc_kernel<1, true><<<grid, kThreads, smem, stream>>>(tmA, tmA2, tmB, p);
  else
    gemm_tc_kernel<1><<<grid, kThreads, smem, stream>>>(tmA, tmA2, tmB, p);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

static void fill_epilogue(GemmParams& p, const univst_epilogue_t* ep) {
  p.bias = p.rowvec = p.residual = p.bias2 = nullptr;
  p.rows_per_group = 1;
  p.rowvec_ld = 0;
  p.act = 0;
  p.ldr = 0;
  p.geglu = 0;
  p.out_scale = 1.0f;
  if (ep) {
    p.bias = (const __half*)ep->bias;
    p.rowvec = (const __half*)ep->rowvec;
    p.rows_per_group = ep->rows_per_group > 0 ? ep->rows_per_group : 1;
    p.rowvec_ld = ep->rowvec_ld;
    p.act = ep->act;
    p.residual = (const __half*)ep->residual;
    p.ldr = ep->ldr;
    p.bias2 = (const __half*)ep->bias2;
    p.geglu = ep->geglu;
    p.out_scale = ep->out_scale != 0.0f ? ep->out_scale : 1.0f;
  }
}

}  // namespace uv

using namespace uv;

extern "C" int univst_gemm_set_workspace(void* ws, int64_t bytes, void* stream) {
  UV_REQUIRE(!ws || (bytes >= (int64_t)(kSplitCounterBytes + (1 << 20)) && ((uintptr_t)ws & 255) == 0),
             "gemm_set_workspace: at least 1.25 MiB, 256-byte aligned (or NULL to unregister)");
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (size_t i = 0; i < g_ws.size(); ++i)
    if (g_ws[i].stream == (cudaStream_t)stream) {
      g_ws.erase(g_ws.begin() + i);
      break;
    }
  if (ws) g_ws.push_back(SplitWs{(cudaStream_t)stream, ws, (size_t)bytes});
  return UNIVST_OK;
}

extern "C" int univst_gemm_tune(int32_t splitk_max_tiles) {
  g_splitk_max_tiles = splitk_max_tiles;
  return UNIVST_OK;
}

extern "C" int univst_gemm_f16(const void* A, int32_t lda, const void* A2, int32_t lda2, int32_t K1, const void* W,
                               int32_t M, int32_t N, int32_t K, void* D, int32_t ldd, const univst_epilogue_t* ep,
                               void* stream) {
  UV_REQUIRE(A && W && D && M > 0 && N > 0 && K > 0, "gemm: null pointer or empty shape");
  UV_REQUIRE(lda % 8 == 0 && K % 8 == 0 && ldd % 8 == 0, "gemm: lda, K, ldd must be multiples of 8 (16-byte rows)");
  UV_REQUIRE(!A2 || (K1 % kBK == 0 && K1 > 0 && K1 < K && lda2 % 8 == 0), "gemm: K1 must be a multiple of 64");
  GemmParams p{};
  fill_epilogue(p, ep);
  UV_REQUIRE(!p.geglu || N % 128 == 0, "gemm: GEGLU needs N %% 128 == 0");
  UV_REQUIRE(!p.residual || p.ldr % 8 == 0, "gemm: residual ld must be a multiple of 8");
  p.M = M;
  p.N = N;
  p.BN = pick_bn(N, p.geglu);
  p.N_out = p.geglu ? N / 2 : N;
  p.mode = 0;
  p.K1 = A2 ? K1 : K;
  p.num_kb = (K + kBK - 1) / kBK;
  p.kb_split = A2 ? K1 / kBK : p.num_kb;
  p.D = (__half*)D;
  p.ldd = ldd;

  CUtensorMap tmA, tmA2, tmB;
  {
    uint64_t dims[2] = {(uint64_t)p.K1, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {kBK, kBM};
    int r = make_tmap_f16(&tmA, A, 2, dims, str, box, true);
    if (r) return r;
  }
  if (A2) {
    uint64_t dims[2] = {(uint64_t)(K - K1), (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda2 * 2};
    uint32_t box[2] = {kBK, kBM};
    int r = make_tmap_f16(&tmA2, A2, 2, dims, str, box, true);
    if (r) return r;
  } else {
    tmA2 = tmA;
  }
  {
    p.cluster = pick_cluster(M, p.BN);
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {kBK, (uint32_t)(p.BN / p.cluster)};   // cluster: each CTA fetches half of the tile's rows
    int r = make_tmap_f16(&tmB, W, 2, dims, str, box, true);
    if (r) return r;
  }
  return launch(tmA, tmA2, tmB, p, (cudaStream_t)stream);
}

// kind: 0 = 3x3, stride 1, padding 1; 1 = 3x3, stride 2, padding 1 (parity planes); 2 = 3x3, stride 2, padding (0, 1, 0, 1)
// -- one zero row / column AFTER the image, as diffusers' Downsample2D(padding=0) pads (parity planes); 3 = 3x1 over the
// image rows only (the (3, 1, 1) temporal convolutions of the temporal VAE decoder: rows = frames, columns = pixels)
static int conv_taps(const void* X, const void* X2, int32_t NB, int32_t H, int32_t W, int32_t C1, int32_t C2, const void* Wt,
                     int32_t Cout, int kind, void* Y, int32_t ldy, const univst_epilogue_t* ep, void* stream) {
  const int stride = (kind == 1 || kind == 2) ? 2 : 1;
  const int ntaps = kind == 3 ? 3 : 9;
  UV_REQUIRE(X && Wt && Y && NB > 0 && Cout > 0 && C1 > 0, "conv3x3: null pointer or empty shape");
  UV_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && (ldy % 8 == 0 || Cout < 8), "conv3x3: channel counts and ldy must be multiples of 8");
  UV_REQUIRE(!X2 || C1 % kBK == 0, "conv3x3: with a second source, C1 must be a multiple of 64");
  UV_REQUIRE(stride == 1 || !X2, "conv3x3: stride 2 takes a single (parity-plane) source");
  // output geometry; for stride 2 the caller passes the 4 parity planes [4][NB][H/2][W/2][C] and H, W of the OUTPUT
  const int Ho = H, Wo = W;
  UV_REQUIRE(Ho > 0 && Wo > 0, "conv3x3: empty image");
  // power-of-two images (width <= 128) tile exactly into 128-row blocks of consecutive pixels (mode 1); anything else takes
  // row-block tiles
  const bool pow2 = (Wo & (Wo - 1)) == 0 && (Ho & (Ho - 1)) == 0 && Wo <= 128;
  GemmParams p{};
  fill_epilogue(p, ep);
  UV_REQUIRE(!p.geglu, "conv3x3: no GEGLU epilogue");
  const int Cin = C1 + (X2 ? C2 : 0);
  p.M = NB * Ho * Wo;
  p.N = Cout;
  p.N_out = Cout;
  p.BN = pick_bn(Cout, 0);
  p.mode = pow2 ? 1 : 2;
  p.K1 = C1;
  p.Cin = Cin;
  const int cbs1 = (C1 + kBK - 1) / kBK;
  const int cbs2 = X2 ? (C2 + kBK - 1) / kBK : 0;
  p.kb_split = cbs1;
  p.cbs = cbs1 + cbs2;
  p.num_kb = ntaps * p.cbs;
  p.HW = Ho * Wo;
  p.W = Wo;
  p.plane_stride = NB;
  if (kind == 3)
    for (int t = 0; t < 3; ++t) {
      p.tap_dy[t] = (int8_t)(t - 1);
      p.tap_dx[t] = 0;
      p.tap_plane[t] = 0;
    }
  for (int ky = 0; ky < 3 && kind != 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int t = ky * 3 + kx;
      if (stride == 1) {
        p.tap_dy[t] = (int8_t)(ky - 1);
        p.tap_dx[t] = (int8_t)(kx - 1);
        p.tap_plane[t] = 0;
      } else if (kind == 2) {
        // input row 2*oy + ky: ky=0 -> even row of pair oy; ky=1 -> odd row of pair oy; ky=2 -> even row of pair oy+1
        const int hp = (ky == 1) ? 1 : 0, wp = (kx == 1) ? 1 : 0;
        p.tap_dy[t] = (int8_t)(ky == 2 ? 1 : 0);
        p.tap_dx[t] = (int8_t)(kx == 2 ? 1 : 0);
        p.tap_plane[t] = (int8_t)(hp * 2 + wp);
      } else {
        // input row 2*oy + ky - 1: ky=0 -> odd row of pair oy-1; ky=1 -> even row of pair oy; ky=2 -> odd row of pair oy
        const int hp = (ky == 1) ? 0 : 1, wp = (kx == 1) ? 0 : 1;
        p.tap_dy[t] = (int8_t)(ky == 0 ? -1 : 0);
        p.tap_dx[t] = (int8_t)(kx == 0 ? -1 : 0);
        p.tap_plane[t] = (int8_t)(hp * 2 + wp);
      }
    }
  p.D = (__half*)Y;
  p.ldd = ldy;

  int bw = Wo;
  int bh = (Ho * Wo >= kBM) ? kBM / bw : Ho;
  int bn = kBM / (bw * bh);
  p.gen_seg = 1;
  if (!pow2 && Wo > kBM) {       // one 128-pixel segment of one row per tile
    bw = kBM;
    bh = 1;
    bn = 1;
    p.gen_seg = (Wo + kBM - 1) / kBM;
    p.gen_th = 1;
    p.gen_bn = 1;
    p.gen_rb = Ho * p.gen_seg;
    p.gen_tiles_m = NB * Ho * p.gen_seg;
  } else if (!pow2) {
    if (Ho * Wo <= kBM) {        // whole images: bn per tile
      bh = Ho;
      bn = kBM / (Ho * Wo);
      p.gen_rb = 1;
    } else {                     // bh full-width rows of one image
      bh = kBM / Wo;
      bn = 1;
      p.gen_rb = (Ho + bh - 1) / bh;
    }
    p.gen_th = bh;
    p.gen_bn = bn;
    p.gen_tiles_m = ((NB + bn - 1) / bn) * p.gen_rb;
  }
  const int planes = (stride == 2) ? 4 : 1;
  CUtensorMap tmA, tmA2, tmB;
  {
    uint64_t dims[4] = {(uint64_t)C1, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)NB * planes};
    uint64_t str[3] = {(uint64_t)C1 * 2, (uint64_t)C1 * Wo * 2, (uint64_t)C1 * Wo * Ho * 2};
    uint32_t box[4] = {kBK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    int r = make_tmap_f16(&tmA, X, 4, dims, str, box, true);
    if (r) return r;
  }
  if (X2) {
    uint64_t dims[4] = {(uint64_t)C2, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)C2 * 2, (uint64_t)C2 * Wo * 2, (uint64_t)C2 * Wo * Ho * 2};
    uint32_t box[4] = {kBK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    int r = make_tmap_f16(&tmA2, X2, 4, dims, str, box, true);
    if (r) return r;
  } else {
    tmA2 = tmA;
  }
  {
    p.cluster = pow2 ? pick_cluster(p.M, p.BN) : 1;
    uint64_t dims[2] = {(uint64_t)ntaps * Cin, (uint64_t)Cout};
    uint64_t str[1] = {(uint64_t)ntaps * Cin * 2};
    uint32_t box[2] = {kBK, (uint32_t)(p.BN / p.cluster)};
    int r = make_tmap_f16(&tmB, Wt, 2, dims, str, box, true);
    if (r) return r;
  }
  return launch(tmA, tmA2, tmB, p, (cudaStream_t)stream);
}

extern "C" int univst_conv3x3_f16(const void* X, const void* X2, int32_t NB, int32_t H, int32_t W, int32_t C1,
                                  int32_t C2, const void* Wt, int32_t Cout, int32_t stride, void* Y, int32_t ldy,
                                  const univst_epilogue_t* ep, void* stream) {
  UV_REQUIRE(stride == 1 || stride == 2, "conv3x3: stride must be 1 or 2");
  return conv_taps(X, X2, NB, H, W, C1, C2, Wt, Cout, stride == 2 ? 1 : 0, Y, ldy, ep, stream);
}

extern "C" int univst_conv3x3_s2_pad_after_f16(const void* X, int32_t NB, int32_t H, int32_t W, int32_t C, const void* Wt,
                                               int32_t Cout, void* Y, int32_t ldy, const univst_epilogue_t* ep,
                                               void* stream) {
  return conv_taps(X, nullptr, NB, H, W, C, 0, Wt, Cout, 2, Y, ldy, ep, stream);
}

extern "C" int univst_conv_temporal3_f16(const void* X, int32_t NB, int32_t F, int32_t HW, int32_t C, const void* Wt,
                                         int32_t Cout, void* Y, int32_t ldy, const univst_epilogue_t* ep, void* stream) {
  return conv_taps(X, nullptr, NB, F, HW, C, 0, Wt, Cout, 3, Y, ldy, ep, stream);
}

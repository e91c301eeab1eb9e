// Temporal self-attention of the AnimateDiff motion modules (backbones/animatediff/models/motion_module.py:276-336,
// VersatileAttention in "Temporal" mode): every pixel attends over the F frames of its clip, per head.
//
// Per (branch, pixel, head) this is a 16 x 16 x d problem -- 4 GFLOP per module at 64 x 64 against 0.5 GB of Q/K/V
// and output traffic, i.e. HBM-bound by two orders of magnitude, and the design is about bytes: one CTA per
// (branch, pixel) pulls the F rows [Q | K | V] of that pixel (3C contiguous halves each) with 16-byte loads, one warp
// per head works out of shared memory, and the F output rows leave as contiguous C-wide stores.  The reference's
// "(b f) d c -> (b d) f c" rearrange (and its inverse) never happens: it is the address arithmetic of the loads.
//
// Two kernels.  F <= 16 (every configuration of the reference: 16-frame clips): the per-head math is two tiny
// register-level matrix products (warp-level mma.sync m16n8k16 -- a 16-frame clip is exactly one fragment; tcgen05's
// 128-row tiles have nothing to hold on to here), ~100 instructions per head, so the kernel runs at memory speed.  The
// first version did the same math with scalar FMAs (lane = query frame) and was instruction-bound at 13 % of the HBM
// peak; it is kept for 16 < F <= 32.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace uv {

struct TAttnParams {
  const __half* QKV;   // [B*F*N, ld]: Q at column 0, K at C, V at 2C
  __half* O;           // [B*F*N, ldo]
  int ld, ldo;
  int F, N, H, d;
  float scale_log2;    // d^-0.5 * log2(e)
};

// FT: compile-time bound on the number of frames (16 or 32) so that the score row lives in registers.
template <int FT>
__global__ void __launch_bounds__(256) temporal_attention_kernel(const TAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int C = p.H * p.d;
  const int du = p.d >> 3;                 // 16-byte units per head row
  const int su = du | 1;                   // padded row stride (odd number of units: conflict-free 16-byte row reads)
  const int F = p.F;
  // shared layout: [which = q|k|v][head][frame][su] units of 16 bytes
  uint4* sm = reinterpret_cast<uint4*>(smem_raw);
  const int head_units = F * su;
  const int which_units = p.H * head_units;
  const int b = blockIdx.x / p.N, pix = blockIdx.x % p.N;
  const size_t row0 = (size_t)b * F * p.N + pix;     // row of frame 0; frame f is f * N rows further

  // ---- cooperative load: F rows of 3C halves
  const int row_units = 3 * C / 8;
  for (int v = threadIdx.x; v < F * row_units; v += blockDim.x) {
    const int f = v / row_units, u = v - f * row_units;
    const int col = u * 8;
    const int which = col / C, hc = col - which * C;
    const int h = hc / p.d, c = hc - h * p.d;
    const uint4 val = *reinterpret_cast<const uint4*>(p.QKV + (row0 + (size_t)f * p.N) * p.ld + col);
    sm[which * which_units + h * head_units + f * su + (c >> 3)] = val;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < p.H; h += (blockDim.x >> 5)) {
    uint4* q = sm + h * head_units;
    const uint4* k = sm + which_units + h * head_units;
    const uint4* v = sm + 2 * which_units + h * head_units;
    const int i = lane < F ? lane : F - 1;    // idle lanes shadow the last frame (no divergence, results unused)
    float s[FT];
#pragma unroll
    for (int j = 0; j < FT; ++j) s[j] = 0.0f;
    for (int u = 0; u < du; ++u) {
      const uint4 qv = q[i * su + u];
      const __half2* q2 = reinterpret_cast<const __half2*>(&qv);
      float qf[8];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const float2 t = __half22float2(q2[x]);
        qf[2 * x] = t.x;
        qf[2 * x + 1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < FT; ++j) {
        if (j < F) {
          const uint4 kv = k[j * su + u];   // same address in every lane: broadcast
          const __half2* k2 = reinterpret_cast<const __half2*>(&kv);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float2 t = __half22float2(k2[x]);
            s[j] = fmaf(qf[2 * x], t.x, s[j]);
            s[j] = fmaf(qf[2 * x + 1], t.y, s[j]);
          }
        }
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < FT; ++j)
      if (j < F) m = fmaxf(m, s[j]);
    float l = 0.0f;
#pragma unroll
    for (int j = 0; j < FT; ++j) {
      s[j] = j < F ? exp2f((s[j] - m) * p.scale_log2) : 0.0f;
      l += s[j];
    }
    const float inv_l = 1.0f / l;
    // the reference rounds the probabilities to fp16 before the bmm with V (get_attention_scores returns fp16)
#pragma unroll
    for (int j = 0; j < FT; ++j) s[j] = __half2float(__float2half_rn(s[j] * inv_l));
    for (int u = 0; u < du; ++u) {
      float acc[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) acc[x] = 0.0f;
#pragma unroll
      for (int j = 0; j < FT; ++j) {
        if (j < F) {
          const uint4 vv = v[j * su + u];
          const __half2* v2 = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float2 t = __half22float2(v2[x]);
            acc[2 * x] = fmaf(s[j], t.x, acc[2 * x]);
            acc[2 * x + 1] = fmaf(s[j], t.y, acc[2 * x + 1]);
          }
        }
      }
      if (lane < F)   // the q row of this lane is dead from here on: it takes the output
        q[i * su + u] = make_uint4(pack_half2(acc[0], acc[1]), pack_half2(acc[2], acc[3]), pack_half2(acc[4], acc[5]),
                                   pack_half2(acc[6], acc[7]));
    }
  }
  __syncthreads();

  // ---- cooperative store: F rows of C halves
  const int out_units = C / 8;
  for (int v = threadIdx.x; v < F * out_units; v += blockDim.x) {
    const int f = v / out_units, u = v - f * out_units;
    const int col = u * 8;
    const int h = col / p.d, c = col - h * p.d;
    *reinterpret_cast<uint4*>(p.O + (row0 + (size_t)f * p.N) * p.ldo + col) = sm[h * head_units + f * su + (c >> 3)];
  }
}

// ------------------------------------------------------------------------------------------------ F <= 16
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// Shared layout as above but with 16 frame rows per head and the head dim zero-padded to a multiple of 16 (the K
// extent of the score product): [which][head][16][su] units of 16 bytes, su odd.
// KC = ceil(d / 16) compile-time (3: d = 40, 5: d = 80, 10: d = 160, 1: d <= 16): the fragment loops unroll.
template <int KC>
__global__ void __launch_bounds__(256) temporal_attention_mma_kernel(const TAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int C = p.H * p.d;
  const int du = p.d >> 3;                 // real 16-byte units per head row
  constexpr int pu = KC * 2;               // padded units
  constexpr int su = pu | 1;               // row stride
  const int F = p.F;
  uint4* sm = reinterpret_cast<uint4*>(smem_raw);
  constexpr int head_units = 16 * su;
  const int which_units = p.H * head_units;
  const int b = blockIdx.x / p.N, pix = blockIdx.x % p.N;
  const size_t row0 = (size_t)b * F * p.N + pix;

  // ---- cooperative load; frames >= F and channels >= d are zero (they enter the products as exact zeros).
  // Thread (tx, ty): tx = 16-byte unit inside a frame row, ty = frame parity; the (q|k|v, head, unit) split of tx is
  // worked out once, and the 8 frame loads of a thread are all in flight before the first shared-memory store (the
  // first version interleaved one load with one store and three integer divisions each: 49 % ALU, long-scoreboard bound).
  {
    const int real_units = 3 * C / 8;                 // 16-byte units of a [Q | K | V] row
    const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
    for (int u0 = 0; u0 < real_units; u0 += 128) {
      const int u = u0 + tx;
      const bool active = u < real_units;
      const int col = u * 8;
      const int which = col / C, hc = col - which * C;
      const int h = hc / p.d, c = hc - h * p.d;
      uint4* dst = sm + which * which_units + h * head_units + (c >> 3);
      const __half* src = p.QKV + row0 * p.ld + col;
      uint4 val[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = 2 * i + ty;
        val[i] = make_uint4(0u, 0u, 0u, 0u);
        if (active && f < F) val[i] = *reinterpret_cast<const uint4*>(src + (size_t)f * p.N * p.ld);
      }
      if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[(2 * i + ty) * su] = val[i];
      }
    }
    // zero the padded units [du, pu) of every (which, head, frame) row
    if (pu > du) {
      const int pad = pu - du, rows = 3 * p.H * 16;
      for (int v = threadIdx.x; v < rows * pad; v += blockDim.x) {
        const int rw = v / pad, k = v - rw * pad;
        sm[rw * su + du + k] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gr = lane >> 2, gc = (lane & 3) * 2;   // fragment row / first column of this lane
  for (int h = warp; h < p.H; h += (blockDim.x >> 5)) {
    const uint32_t qb = smem_u32(sm + h * head_units);
    const uint32_t kb = smem_u32(sm + which_units + h * head_units);
    const uint32_t vb = smem_u32(sm + 2 * which_units + h * head_units);
    // S = Q K^T: 16 x 16, two n-tiles of 8 keys
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      const uint32_t col = (uint32_t)(kc * 16 + gc) * 2;
      uint32_t a[4];
      a[0] = lds32(qb + gr * (su * 16) + col);
      a[1] = lds32(qb + (gr + 8) * (su * 16) + col);
      a[2] = lds32(qb + gr * (su * 16) + col + 16);
      a[3] = lds32(qb + (gr + 8) * (su * 16) + col + 16);
      // B[k][n] = K[n][k]: key n = gr (+ 8 for the second tile), k = gc (+ 8)
      mma_16816(s0, a, lds32(kb + gr * (su * 16) + col), lds32(kb + gr * (su * 16) + col + 16));
      mma_16816(s1, a, lds32(kb + (gr + 8) * (su * 16) + col), lds32(kb + (gr + 8) * (su * 16) + col + 16));
    }
    // softmax over the 16 keys of rows gr (elements 0, 1) and gr + 8 (elements 2, 3); keys >= F are masked
    float pr[8] = {s0[0], s0[1], s1[0], s1[1], s0[2], s0[3], s1[2], s1[3]};   // row gr: keys gc, gc+1, gc+8, gc+9 | row gr+8
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const int key = gc + (x & 1) + ((x >> 1) & 1) * 8;
      if (key >= F) pr[x] = -INFINITY;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float m = fmaxf(fmaxf(pr[4 * r], pr[4 * r + 1]), fmaxf(pr[4 * r + 2], pr[4 * r + 3]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float l = 0.0f;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        pr[4 * r + x] = exp2f((pr[4 * r + x] - m) * p.scale_log2);
        l += pr[4 * r + x];
      }
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const float inv_l = 1.0f / l;
#pragma unroll
      for (int x = 0; x < 4; ++x) pr[4 * r + x] *= inv_l;
    }
    // P as the A operand of O = P V (fp16, as the reference's get_attention_scores returns it)
    uint32_t pa[4];
    pa[0] = pack_half2(pr[0], pr[1]);   // row gr,     keys gc, gc + 1
    pa[1] = pack_half2(pr[4], pr[5]);   // row gr + 8
    pa[2] = pack_half2(pr[2], pr[3]);   // row gr,     keys gc + 8, gc + 9
    pa[3] = pack_half2(pr[6], pr[7]);   // row gr + 8
    // O = P V in n-tiles of 8 channels; B[k][n] = V[frame k][channel n] via ldmatrix.trans (rows = frames)
#pragma unroll
    for (int nt = 0; nt < KC * 2; ++nt) {
      if (nt < du) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(vb + (lane & 15) * (su * 16) + nt * 16, b0, b1);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        mma_16816(o, pa, b0, b1);
        // the Q tile of this head is dead: it takes the output (row gr / gr + 8, channels nt * 8 + gc, + 1)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(qb + gr * (su * 16) + nt * 16 + gc * 2), "r"(pack_half2(o[0], o[1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(qb + (gr + 8) * (su * 16) + nt * 16 + gc * 2), "r"(pack_half2(o[2], o[3])) : "memory");
      }
    }
  }
  __syncthreads();

  // ---- cooperative store: F rows of C halves
  const int out_units = C / 8;
  for (int v = threadIdx.x; v < F * out_units; v += blockDim.x) {
    const int f = v / out_units, u = v - f * out_units;
    const int h = u / du, cu = u - h * du;
    *reinterpret_cast<uint4*>(p.O + (row0 + (size_t)f * p.N) * p.ldo + u * 8) = sm[h * head_units + f * su + cu];
  }
}

template <int KC>
static int launch_tattn_mma(const TAttnParams& p, int64_t grid, int threads, cudaStream_t st) {
  const size_t smem = (size_t)3 * p.H * 16 * ((KC * 2) | 1) * 16;
  UV_REQUIRE(smem <= 227 * 1024, "temporal_attention: %zu bytes of shared memory needed", smem);
  static bool configured = false;
  if (!configured) {
    UV_CHECK_CUDA(cudaFuncSetAttribute(temporal_attention_mma_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  temporal_attention_mma_kernel<KC><<<(unsigned)grid, threads, smem, st>>>(p);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

}  // namespace uv

using namespace uv;

extern "C" int univst_temporal_attention_f16(const void* QKV, int32_t ld, int32_t B, int32_t F, int32_t N, int32_t H,
                                             int32_t d, void* O, int32_t ldo, void* stream) {
  UV_REQUIRE(QKV && O, "temporal_attention: null pointer");
  UV_REQUIRE(B > 0 && F > 0 && N > 0 && H > 0 && d > 0, "temporal_attention: empty shape");
  UV_REQUIRE(F <= 32, "temporal_attention: at most 32 frames per clip (got %d)", F);
  UV_REQUIRE(d % 8 == 0, "temporal_attention: head dim must be a multiple of 8");
  UV_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ld >= 3 * H * d && ldo >= H * d, "temporal_attention: bad row strides");
  UV_REQUIRE(((uintptr_t)QKV | (uintptr_t)O) % 16 == 0, "temporal_attention: 16-byte alignment");
  TAttnParams p{};
  p.QKV = (const __half*)QKV;
  p.O = (__half*)O;
  p.ld = ld;
  p.ldo = ldo;
  p.F = F;
  p.N = N;
  p.H = H;
  p.d = d;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
  const size_t smem = (size_t)3 * H * F * ((d / 8) | 1) * 16;
  UV_REQUIRE(smem <= 227 * 1024, "temporal_attention: %zu bytes of shared memory needed", smem);
  const int threads = 32 * (H < 8 ? H : 8);
  const int64_t grid = (int64_t)B * N;
  UV_REQUIRE(grid < (1ll << 31), "temporal_attention: too many pixels");
  cudaStream_t st = (cudaStream_t)stream;
  static int force_scalar = -1;
  if (force_scalar < 0) {
    const char* e = getenv("UNIVST_TATTN_SCALAR");
    force_scalar = e ? (atoi(e) != 0) : 0;
  }
  if (F <= 16 && !force_scalar) {
    const int kc = (d + 15) / 16;
    switch (kc) {
      // (always 256 threads: the load phase maps threads to (unit, frame parity); idle warps skip the head loop)
      case 1: return launch_tattn_mma<1>(p, grid, 256, st);
      case 2: return launch_tattn_mma<2>(p, grid, 256, st);
      case 3: return launch_tattn_mma<3>(p, grid, 256, st);   // d = 40
      case 4: return launch_tattn_mma<4>(p, grid, 256, st);
      case 5: return launch_tattn_mma<5>(p, grid, 256, st);   // d = 80
      case 10: return launch_tattn_mma<10>(p, grid, 256, st); // d = 160
      default: break;                                             // other head dims: scalar kernel below
    }
  }
  if (F <= 16) {
    static bool configured = false;
    if (!configured) {
      UV_CHECK_CUDA(cudaFuncSetAttribute(temporal_attention_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
    temporal_attention_kernel<16><<<(unsigned)grid, threads, smem, st>>>(p);
  } else {
    static bool configured = false;
    if (!configured) {
      UV_CHECK_CUDA(cudaFuncSetAttribute(temporal_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = true;
    }
    temporal_attention_kernel<32><<<(unsigned)grid, threads, smem, st>>>(p);
  }
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

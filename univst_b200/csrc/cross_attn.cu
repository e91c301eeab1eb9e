// Cross-attention over a short context (the 77 CLIP tokens; diffusers Attention + AttnProcessor2_0 called at
// backbones/video_diffusion_sd/models/attention.py:316-323 and backbones/animatediff/models/attention.py:337-348).
//
// Per image and head this is N x 77 x d: 19 GFLOP per layer at 64 x 64 against 252 MB of Q / O traffic -- HBM-bound.
// Routed through the big fused-attention kernel it ran at 9x the HBM floor (366 us per 64 x 64 layer, 2.6 ms per UNet
// call): one KV tile per CTA means every CTA pays the full tcgen05 setup (512-column TMEM allocation, barrier init,
// TMA descriptor fetch) for a microsecond of math, and its 227 KB of shared memory allow one CTA per SM.  Here the
// whole K / V of a (branch, head) sits in a few KB of shared memory, a warp owns 16 query rows and does both products
// as register-level fragments (mma.sync m16n8k16 -- the score matrix of a warp is 16 x 80), the softmax is a single
// pass (no online rescale: all keys are in one tile), and many CTAs share an SM.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace uv {

static constexpr int kTilesPerCta = 2;   // query tiles of 128 rows walked by one CTA (K / V stay in shared memory)

struct XAttnParams {
  const __half* Q;     // [NI*N, ldq]
  const __half* K;     // [NIkv*Nkv, ldkv]
  const __half* V;
  __half* O;           // [NI*N, ldo]
  const int* kv_src;   // [NI]: K/V image (branch) of each query image
  int ldq, ldkv, ldo;
  int N, Nkv, H, d;
  float scale_log2;
};

__device__ __forceinline__ void xa_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float xa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t xa_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void xa_ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// KC = ceil(d / 16) (K extent of the score product), KT = ceil(Nkv / 16) (key tiles of 16; Nkv <= 16 KT).
// Shared: K [16 KT][su], V [16 KT][su], Q / O [128][su] in 16-byte units, su = (2 KC) | 1 (odd: conflict-free rows).
template <int KC, int KT>
__global__ void __launch_bounds__(256) cross_attention_kernel(const XAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int pu = KC * 2, su = pu | 1, keys = KT * 16;
  uint4* sK = reinterpret_cast<uint4*>(smem_raw);
  uint4* sV = sK + keys * su;
  uint4* sQ = sV + keys * su;
  const int du = p.d >> 3;
  // heads vary fastest over the grid: the CTAs that touch the same query rows (d of the C columns each) run at the
  // same time, so their partial-row reads and writes meet in L2.  A CTA keeps its K / V and walks kTilesPerCta query
  // tiles of 128 rows; the next tile's Q is fetched into registers while the current one is being computed.
  const int img = blockIdx.z, head = blockIdx.x;
  const int tile0 = blockIdx.y * kTilesPerCta;
  const int ntiles = (p.N + 127) / 128;
  const int kvimg = p.kv_src[img];
  const int tid = threadIdx.x;
  constexpr int kIt = (keys * pu + 255) / 256, qIt = (128 * pu + 255) / 256;
  uint4 qq[qIt];
  auto fetch_q = [&](int tile) {
    const int q0 = tile * 128;
#pragma unroll
    for (int i = 0; i < qIt; ++i) {
      const int v = tid + 256 * i, r = v / pu, cu = v - r * pu;
      qq[i] = make_uint4(0u, 0u, 0u, 0u);
      if (v < 128 * pu && tile < ntiles && q0 + r < p.N && cu < du)
        qq[i] = *reinterpret_cast<const uint4*>(p.Q + ((size_t)img * p.N + q0 + r) * p.ldq + head * p.d + cu * 8);
    }
  };
  // ---- K, V of this (branch, head): keys >= Nkv and channels >= d are zero (all loads in flight before the stores)
  {
    uint4 kk[kIt], vv[kIt];
#pragma unroll
    for (int i = 0; i < kIt; ++i) {
      const int v = tid + 256 * i, key = v / pu, cu = v - key * pu;
      kk[i] = vv[i] = make_uint4(0u, 0u, 0u, 0u);
      if (v < keys * pu && key < p.Nkv && cu < du) {
        const size_t off = ((size_t)kvimg * p.Nkv + key) * p.ldkv + head * p.d + cu * 8;
        kk[i] = *reinterpret_cast<const uint4*>(p.K + off);
        vv[i] = *reinterpret_cast<const uint4*>(p.V + off);
      }
    }
    fetch_q(tile0);
#pragma unroll
    for (int i = 0; i < kIt; ++i) {
      const int v = tid + 256 * i, key = v / pu, cu = v - key * pu;
      if (v < keys * pu) {
        sK[key * su + cu] = kk[i];
        sV[key * su + cu] = vv[i];
      }
    }
  }
  const int warp = tid >> 5, lane = tid & 31;
  const int gr = lane >> 2, gc = (lane & 3) * 2;
  const uint32_t qb = smem_u32(sQ + warp * 16 * su), kb = smem_u32(sK), vb = smem_u32(sV);
  constexpr uint32_t rs = su * 16;   // row stride in bytes

  for (int tile = tile0; tile < tile0 + kTilesPerCta && tile < ntiles; ++tile) {
    const int q0 = tile * 128;
#pragma unroll
    for (int i = 0; i < qIt; ++i) {
      const int v = tid + 256 * i, r = v / pu, cu = v - r * pu;
      if (v < 128 * pu) sQ[r * su + cu] = qq[i];
    }
    __syncthreads();
    if (tile + 1 < tile0 + kTilesPerCta) fetch_q(tile + 1);   // in flight during this tile's math

    // S = Q K^T: 16 rows x (2 KT) tiles of 8 keys
    float s[2 * KT][4];
#pragma unroll
    for (int t = 0; t < 2 * KT; ++t)
#pragma unroll
      for (int x = 0; x < 4; ++x) s[t][x] = 0.0f;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      const uint32_t col = (uint32_t)(kc * 16 + gc) * 2;
      uint32_t a[4];
      a[0] = xa_lds32(qb + gr * rs + col);
      a[1] = xa_lds32(qb + (gr + 8) * rs + col);
      a[2] = xa_lds32(qb + gr * rs + col + 16);
      a[3] = xa_lds32(qb + (gr + 8) * rs + col + 16);
#pragma unroll
      for (int t = 0; t < 2 * KT; ++t)
        xa_mma(s[t], a, xa_lds32(kb + (t * 8 + gr) * rs + col), xa_lds32(kb + (t * 8 + gr) * rs + col + 16));
    }
    // single-pass softmax over the keys of rows gr (elements 0, 1) and gr + 8 (elements 2, 3)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2 * KT; ++t) {
      if (t * 8 + 8 > p.Nkv) {   // only the tile(s) that straddle the end of the context need masking (warp-uniform)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int key = t * 8 + gc + (x & 1);
          if (key >= p.Nkv) s[t][x] = -INFINITY;
        }
      }
      m0 = fmaxf(m0, fmaxf(s[t][0], s[t][1]));
      m1 = fmaxf(m1, fmaxf(s[t][2], s[t][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.0f, l1 = 0.0f;
    const float nm0 = -m0 * p.scale_log2, nm1 = -m1 * p.scale_log2;   // one FFMA + one MUFU.EX2 per score
#pragma unroll
    for (int t = 0; t < 2 * KT; ++t) {
      s[t][0] = xa_ex2(fmaf(s[t][0], p.scale_log2, nm0));
      s[t][1] = xa_ex2(fmaf(s[t][1], p.scale_log2, nm0));
      s[t][2] = xa_ex2(fmaf(s[t][2], p.scale_log2, nm1));
      s[t][3] = xa_ex2(fmaf(s[t][3], p.scale_log2, nm1));
      l0 += s[t][0] + s[t][1];
      l1 += s[t][2] + s[t][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    // O = P V: the score fragments are the A operand (16 keys per step = two score tiles), V via ldmatrix.trans.
    // The Q rows of this warp are dead: they take the output.
    __syncwarp();
    uint32_t pa[KT][4];
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      pa[t][0] = pack_half2(s[2 * t][0], s[2 * t][1]);
      pa[t][1] = pack_half2(s[2 * t][2], s[2 * t][3]);
      pa[t][2] = pack_half2(s[2 * t + 1][0], s[2 * t + 1][1]);
      pa[t][3] = pack_half2(s[2 * t + 1][2], s[2 * t + 1][3]);
    }
#pragma unroll
    for (int nt = 0; nt < KC * 2; ++nt) {
      if (nt < du) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < KT; ++t) {
          uint32_t b0, b1;
          xa_ldmatrix_x2_trans(vb + (t * 16 + (lane & 15)) * rs + nt * 16, b0, b1);
          xa_mma(o, pa[t], b0, b1);
        }
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(qb + gr * rs + nt * 16 + gc * 2), "r"(pack_half2(o[0] * i0, o[1] * i0)) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(qb + (gr + 8) * rs + nt * 16 + gc * 2), "r"(pack_half2(o[2] * i1, o[3] * i1)) : "memory");
      }
    }
    __syncthreads();
    // ---- cooperative store of the 128 output rows (d halves each)
    for (int v = tid; v < 128 * du; v += 256) {
      const int r = v / du, cu = v - r * du;
      if (q0 + r < p.N)
        *reinterpret_cast<uint4*>(p.O + ((size_t)img * p.N + q0 + r) * p.ldo + head * p.d + cu * 8) = sQ[r * su + cu];
    }
    __syncthreads();   // the Q / O tile is refilled by the next iteration
  }
}

template <int KC, int KT>
static int launch_xattn(const XAttnParams& p, int NI, cudaStream_t st) {
  constexpr int su = (KC * 2) | 1;
  const size_t smem = (size_t)(2 * KT * 16 + 128) * su * 16;
  UV_REQUIRE(smem <= 227 * 1024, "cross_attention: %zu bytes of shared memory needed", smem);
  static bool configured = false;
  if (!configured) {
    UV_CHECK_CUDA(cudaFuncSetAttribute(cross_attention_kernel<KC, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int ntiles = (p.N + 127) / 128;
  dim3 grid(p.H, (ntiles + kTilesPerCta - 1) / kTilesPerCta, NI);
  cross_attention_kernel<KC, KT><<<grid, 256, smem, st>>>(p);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

}  // namespace uv

using namespace uv;

extern "C" int univst_cross_attention_supported(int32_t d, int32_t Nkv) {
  const int kc = (d + 15) / 16;
  return (d % 8 == 0 && Nkv > 0 && Nkv <= 80 && (kc == 3 || kc == 4 || kc == 5 || kc == 10 || kc == 1 || kc == 2)) ? 1 : 0;
}

extern "C" int univst_cross_attention_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv, int32_t NI,
                                          int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv, const int32_t* kv_src,
                                          void* O, int32_t ldo, void* stream) {
  UV_REQUIRE(Q && K && V && O && kv_src, "cross_attention: null pointer");
  UV_REQUIRE(NI > 0 && NIkv > 0 && H > 0 && N > 0, "cross_attention: empty shape");
  UV_REQUIRE(univst_cross_attention_supported(d, Nkv), "cross_attention: head dim %d / %d context tokens not supported "
             "(head dims up to 80 or 160, at most 80 tokens); use univst_sc_attention_f16", d, Nkv);
  UV_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0, "cross_attention: row strides must be multiples of 8");
  UV_REQUIRE(((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) % 16 == 0, "cross_attention: 16-byte alignment");
  XAttnParams p{};
  p.Q = (const __half*)Q;
  p.K = (const __half*)K;
  p.V = (const __half*)V;
  p.O = (__half*)O;
  p.kv_src = kv_src;
  p.ldq = ldq;
  p.ldkv = ldkv;
  p.ldo = ldo;
  p.N = N;
  p.Nkv = Nkv;
  p.H = H;
  p.d = d;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
  cudaStream_t st = (cudaStream_t)stream;
  const int kc = (d + 15) / 16;
  switch (kc) {   // 80 keys = 5 tiles of 16 (77 CLIP tokens)
    case 1: return launch_xattn<1, 5>(p, NI, st);
    case 2: return launch_xattn<2, 5>(p, NI, st);
    case 3: return launch_xattn<3, 5>(p, NI, st);    // d = 40
    case 4: return launch_xattn<4, 5>(p, NI, st);    // d = 64
    case 5: return launch_xattn<5, 5>(p, NI, st);    // d = 80
    default: return launch_xattn<10, 5>(p, NI, st);  // d = 160
  }
}

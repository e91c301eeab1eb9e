// Fused sparse-causal attention for sm_100a: O = softmax(Q K^T / sqrt(d)) V with the K/V sequence of an image
// being the concatenation of the K/V of up to three *source frames* (previous / self / first).  The concat
// is never materialised: each KV tile is a TMA box addressed through a per-image source table, so the same
// kernel serves the patched attn1 of the reference (pnp_utils.py:59-92, KV = [prev, first]), the stock
// SparseCausalAttention (models/attention.py:384-420, KV = [prev, self, first]), plain self-attention and the
// 77-token cross-attention (attention.py:316-323; one shared K/V "image").
//
// Structure (FlashAttention-4 style, one CTA per (q-tile group, head, image)):
//   warp 0        TMA producer: Q tiles once, K and V tiles through two 2-deep rings
//   warp 1 (+ one extra warp per additional query tile)  tcgen05.mma issuers, one per query tile:
//                 S = Q K^T into a TMEM slot, O += P V with P read from shared memory
//   warps 2..     softmax groups of 128 threads, one thread per query row: TMEM -> registers, running max with
//                 lazy (thresholded) rescale of the TMEM-resident O, exp2, fp16 P tile written 128B-swizzled
//   NQ = 2 / 4: that many query tiles per CTA share the K/V ring and keep 2 / 4 softmax warps per scheduler;
//   NQ = 1: one query tile whose two S slots alternate between consecutive KV tiles.
//   An S slot is released as soon as its tile sits in registers, so the next Q K^T of that query tile runs under the
//   current tile's exponentials (the MUFU pipe, not the tensor pipe, bounds d = 40).
// Head dims that are not multiples of 64 (SD-1.5: 40 / 80 / 160) are zero-filled by TMA out-of-bounds
// handling; nothing is padded in global memory.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace uv {

struct AttnParams {
  int NI;        // images (query side)
  int H;         // heads
  int d;         // true head dim
  int N;         // query tokens per image
  int Nkv;       // key tokens per source image
  int nsrc;      // sources per image
  const int* kv_src;  // [NI][nsrc] image index into the K/V tensors
  __half* O;     // [NI*N][ldo]
  int ldo;
  float scale_log2;  // d^-0.5 * log2(e)
};

static constexpr uint32_t kOBase = 256;     // TMEM column of the first O accumulator
static constexpr float kRescaleThreshold = 8.0f;  // log2 units: P <= 2^8 before a forced rescale
static constexpr int kDefaultVariant = 0;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max relative error 7.5e-5 -- below the
// half-ulp of the fp16 P it feeds): a quarter of the exponentials take this path so that the MUFU pipe, which bounds
// attention at head dim 40, is relieved (the FlashAttention-4 trick).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: nearest integer of x lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);    // fractional part in [-0.5, 0.5]
  float p = fmaf(0.0551716574f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int NQ, int BKV>
struct AttnCfg {
  static constexpr int kDepth = (NQ == 1) ? 2 : 1;      // S / P slots per query tile
  static constexpr int kSlots = NQ * kDepth;              // S slots
  static constexpr int kPSlots = NQ * 2;                  // P tiles are double-buffered per query tile
  static constexpr int kThreads = 64 + 128 * NQ + 32 * (NQ - 1);
  static constexpr int kQChunkBytes = 128 * 128;        // [128 rows][64 halves]
  static constexpr int kKVChunkBytes = BKV * 128;       // [BKV rows][64 halves]
  static constexpr int kPBytes = 128 * BKV * 2;         // [128 rows][BKV halves] as BKV/64 swizzled chunks
  static constexpr int kBarriers = 40;
  static_assert(kSlots * BKV <= 256, "S slots must fit below the O accumulators");
};

// Work decomposition: query tile q of the CTA streams KV tiles j = 0..T-1.  Tile (q, j) uses S/P slot
// q * kDepth + j % kDepth for the (j / kDepth)-th time.  Every query tile has its own MMA-issuing thread and its own
// softmax group, so the tiles only meet at the K/V ring.
template <int NQ, int BKV, int POLY>
__global__ void __launch_bounds__(AttnCfg<NQ, BKV>::kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using L = AttnCfg<NQ, BKV>;
  constexpr int D = L::kDepth, NS = L::kSlots;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int dch = (p.d + 63) >> 6;                    // 64-column chunks of the head dim
  const int dpad = (p.d + 15) & ~15;                  // head dim rounded to the UMMA K / N granularity
  const uint32_t q_bytes = (uint32_t)dch * L::kQChunkBytes;        // one Q tile
  const uint32_t kv_bytes = (uint32_t)dch * L::kKVChunkBytes;      // one K (or V) tile
  uint8_t* sQ = smem;                                  // [NQ][dch][128][64]
  uint8_t* sK = sQ + NQ * q_bytes;                     // [2][dch][BKV][64]
  uint8_t* sV = sK + 2 * kv_bytes;                     // [2][dch][BKV][64]
  uint8_t* sP = sV + 2 * kv_bytes;                     // [NQ * 2][BKV/64][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + L::kPSlots * L::kPBytes);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* k_empty = bars + 3;       // 2
  uint64_t* v_full = bars + 5;        // 2
  uint64_t* v_empty = bars + 7;       // 2
  uint64_t* s_full = bars + 9;        // NS (<= 4)
  uint64_t* s_free = bars + 13;       // NS
  uint64_t* p_full = bars + 17;       // NQ * 2 (<= 8)
  uint64_t* o_done = bars + 25;       // NQ * 2: P V of tile (q, j) commits to o_done[2 q + (j & 1)]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 33);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int qt0 = blockIdx.x * NQ;     // first 128-row query tile of this CTA
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int tps = (p.Nkv + BKV - 1) / BKV;   // KV tiles per source
  const int T = p.nsrc * tps;                // KV tiles in total

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], NQ);   // one tcgen05.commit per MMA issuer
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], NQ);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 128);
    }
    for (int s = 0; s < L::kPSlots; ++s) {
      mbar_init(&p_full[s], 128);
      mbar_init(&o_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t opad = (uint32_t)dpad;
  const bool is_mma = (warp == 1) || (warp >= 2 + 4 * NQ);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, NQ * q_bytes);
      for (int q = 0; q < NQ; ++q)
        for (int c = 0; c < dch; ++c)
          tma_load_4d(sQ + q * q_bytes + c * L::kQChunkBytes, &tmQ, q_full, c * 64, head, (qt0 + q) * 128, img);
      const int* src = p.kv_src + (size_t)img * p.nsrc;
      uint32_t phase = 0;
      for (int j = 0; j < T; ++j) {
        const int s = j & 1;
        const int simg = src[j / tps];
        const int row0 = (j % tps) * BKV;
        mbar_wait(&k_empty[s], phase ^ 1);
        mbar_expect_tx(&k_full[s], kv_bytes);
        for (int c = 0; c < dch; ++c)
          tma_load_4d(sK + s * kv_bytes + c * L::kKVChunkBytes, &tmK, &k_full[s], c * 64, head, row0, simg);
        mbar_wait(&v_empty[s], phase ^ 1);
        mbar_expect_tx(&v_full[s], kv_bytes);
        for (int c = 0; c < dch; ++c)
          tma_load_4d(sV + s * kv_bytes + c * L::kKVChunkBytes, &tmV, &v_full[s], c * 64, head, row0, simg);
        if (s == 1) phase ^= 1;
      }
    }
  } else if (is_mma) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer of query tile q
      const int q = (warp == 1) ? 0 : (int)(warp - (2 + 4 * NQ)) + 1;
      const uint32_t idesc_s = make_idesc_f16(128, BKV, 0, 0);      // S = Q K^T : both K-major
      const uint32_t idesc_o = make_idesc_f16(128, opad, 0, 1);     // O += P V  : P K-major, V MN-major
      auto issue_s = [&](int j) {
        const int ks = j & 1, slot = q * D + j % D;
        mbar_wait(&k_full[ks], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ + q * q_bytes);
        const uint32_t ka = smem_u32(sK + ks * kv_bytes);
        int step = 0;
        for (int c = 0; c < dch; ++c) {
          const int nk = min(64, dpad - c * 64) >> 4;
          const uint64_t da = make_smem_desc_sw128(qa + c * L::kQChunkBytes, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(ka + c * L::kKVChunkBytes, 16, 1024);
          for (int k = 0; k < nk; ++k, ++step)
            umma_f16_ss(tmem_base + slot * BKV, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_s, step ? 1u : 0u);
        }
        tc_commit(&k_empty[ks]);
        tc_commit(&s_full[slot]);
      };
      auto issue_pv = [&](int j) {
        const int vs = j & 1, slot = q * 2 + (j & 1);        // P slot, used for the (j / 2)-th time
        mbar_wait(&p_full[slot], (uint32_t)((j >> 1) & 1));
        mbar_wait(&v_full[vs], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint32_t pa = smem_u32(sP + slot * L::kPBytes);
        const uint32_t va = smem_u32(sV + vs * kv_bytes);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A: P chunk (k / 4), 32-byte step inside the swizzle row.  B: V, 16 keys = 2 KiB further down;
          // the next 64 head-dim columns are a whole chunk away (LBO).
          const uint64_t da = make_smem_desc_sw128(pa + (k >> 2) * (128 * 128) + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(va + k * 2048, L::kKVChunkBytes, 1024);
          umma_f16_ss(tmem_base + kOBase + q * opad, da, db, idesc_o, (j | k) ? 1u : 0u);
        }
        tc_commit(&v_empty[vs]);
        tc_commit(&o_done[slot]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      // The tile's slots are filled up front; afterwards a slot is refilled with the next scores that map to it as
      // soon as the softmax group has pulled the current tile into registers (s_free), i.e. *during* that tile's
      // exponentials -- a group never waits for the tensor pipe between two tiles.
      for (int j = 0; j < D && j < T; ++j) issue_s(j);
      for (int j = 0; j < T; ++j) {
        if (j + D < T) {
          mbar_wait(&s_free[q * D + j % D], (uint32_t)((j / D) & 1));
          tc_fence_after();
          issue_s(j + D);
        }
        issue_pv(j);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax group g = query tile g
    const int g = (int)(warp - 2) >> 2;
    const uint32_t quad = warp & 3;                     // TMEM lane quadrant this warp may touch
    const uint32_t r = quad * 32 + lane;                // row inside the 128-row tile
    const uint32_t lane_off = (quad * 32) << 16;
    const uint32_t o_addr = tmem_base + kOBase + g * opad + lane_off;
    float m_used = -INFINITY;   // max baked into O and l
    float l = 0.0f;
    for (int j = 0; j < T; ++j) {
      const int slot = g * D + j % D;
      mbar_wait(&s_full[slot], (uint32_t)((j / D) & 1));
      tc_fence_after();
      // S tile -> registers: all loads in flight, one wait
      uint32_t sraw[BKV];
      {
        const uint32_t sa = tmem_base + slot * BKV + lane_off;
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) tmem_ld32(sa + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sraw[c * 32]));
        tc_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&s_free[slot]);  // the slot may be overwritten with the next scores from here on
      const int valid = min(BKV, p.Nkv - (j % tps) * BKV);
      if (valid < BKV) {
#pragma unroll
        for (int x = 0; x < BKV; ++x)
          if (x >= valid) sraw[x] = 0xff800000u;  // -inf
      }
      // row max of the raw scores: 8 independent chains (a single chain is BKV dependent FMNMX)
      float mx8[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) mx8[x] = __uint_as_float(sraw[x]);
#pragma unroll
      for (int x = 8; x < BKV; ++x) mx8[x & 7] = fmaxf(mx8[x & 7], __uint_as_float(sraw[x]));
      const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                             fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]))) * p.scale_log2;
      const bool need = mx > m_used + kRescaleThreshold;   // first tile: m_used = -inf -> always true
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, mx);
        if (j > 0) {
          // rare path: O is rescaled in TMEM, so the previous P V of this query tile must have landed
          mbar_wait(&o_done[g * 2 + ((j - 1) & 1)], (uint32_t)(((j - 1) >> 1) & 1));
          tc_fence_after();
          const float alpha = ex2_approx(m_used - m_new);   // lanes that did not need it: alpha <= 1, harmless
          l *= alpha;
          for (uint32_t c = 0; c < opad; c += 16) {
            uint32_t t[16];
            tmem_ld16(o_addr + c, t);
            tc_wait_ld();
#pragma unroll
            for (int x = 0; x < 16; ++x) t[x] = __float_as_uint(__uint_as_float(t[x]) * alpha);
            tmem_st16(o_addr + c, t);
          }
          tc_wait_st();
        }
        m_used = m_new;
      }
      // P = 2^(s * scale - m_used): one FFMA + one MUFU.EX2 per element, fp16, written as swizzled K-major
      // chunks (chunk kc holds keys [64 kc, 64 kc + 64)); the row sum runs in 4 independent chains
      const int pslot = g * 2 + (j & 1);
      if (j >= 2) {  // the P buffer was last read by the P V of tile j - 2: long done, the wait is (almost) free
        mbar_wait(&o_done[pslot], (uint32_t)(((j - 2) >> 1) & 1));
      }
      uint8_t* prow = sP + pslot * L::kPBytes + r * 128;
      float ls4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const float neg_m = -m_used;
#pragma unroll
      for (int c16 = 0; c16 < BKV / 8; ++c16) {
        uint32_t w[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float a0 = fmaf(__uint_as_float(sraw[c16 * 8 + 2 * x]), p.scale_log2, neg_m);
          const float a1 = fmaf(__uint_as_float(sraw[c16 * 8 + 2 * x + 1]), p.scale_log2, neg_m);
          const float e0 = ex2_approx(a0);
          // POLY = 1: every 4th exponential off the MUFU pipe; POLY = 2: every 2nd
          const float e1 = (POLY == 2 || (POLY == 1 && (x & 1))) ? ex2_poly(a1) : ex2_approx(a1);
          ls4[x] += e0 + e1;
          w[x] = pack_half2(e0, e1);
        }
        const int kc = c16 >> 3, cc = c16 & 7;
        *reinterpret_cast<uint4*>(prow + kc * (128 * 128) + ((cc ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[pslot]);
    }
    // ------------------------------------------------------------------ epilogue: O / l -> global
    if (T >= 2) mbar_wait(&o_done[g * 2 + ((T - 2) & 1)], (uint32_t)(((T - 2) >> 1) & 1));
    mbar_wait(&o_done[g * 2 + ((T - 1) & 1)], (uint32_t)(((T - 1) >> 1) & 1));
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = (qt0 + g) * 128 + (int)r;
    const bool row_ok = qrow < p.N;
    __half* orow = p.O + ((size_t)img * p.N + qrow) * p.ldo + head * p.d;
    for (uint32_t c = 0; c < opad; c += 16) {
      uint32_t t[16];
      tmem_ld16(o_addr + c, t);
      tc_wait_ld();
      if (!row_ok) continue;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        const int col = (int)c + h8 * 8;
        if (col < p.d) {
          uint32_t w[4];
#pragma unroll
          for (int x = 0; x < 4; ++x)
            w[x] = pack_half2(__uint_as_float(t[h8 * 8 + 2 * x]) * inv_l, __uint_as_float(t[h8 * 8 + 2 * x + 1]) * inv_l);
          *reinterpret_cast<uint4*>(orow + col) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NQ, int BKV, int POLY>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                       cudaStream_t stream) {
  using L = AttnCfg<NQ, BKV>;
  const int dch = (p.d + 63) / 64;
  const size_t smem = (size_t)NQ * dch * L::kQChunkBytes + 4 * (size_t)dch * L::kKVChunkBytes +
                      (size_t)L::kPSlots * L::kPBytes + L::kBarriers * sizeof(uint64_t) + 1024;
  UV_REQUIRE(smem <= 227 * 1024, "attention: tile configuration needs %zu bytes of shared memory", smem);
  UV_REQUIRE(256 + NQ * ((p.d + 15) & ~15) <= 512, "attention: O accumulators do not fit into TMEM");
  static bool configured = false;
  if (!configured) {
    UV_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<NQ, BKV, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
    configured = true;
  }
  dim3 grid((p.N + 128 * NQ - 1) / (128 * NQ), p.H, p.NI);
  attention_tc_kernel<NQ, BKV, POLY><<<grid, L::kThreads, smem, stream>>>(tq, tk, tv, p);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

}  // namespace uv

using namespace uv;

extern "C" int univst_sc_attention_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv,
                                       int32_t NI, int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv,
                                       const int32_t* kv_src, int32_t nsrc, void* O, int32_t ldo, void* stream) {
  UV_REQUIRE(Q && K && V && O && kv_src, "sc_attention: null pointer");
  UV_REQUIRE(NI > 0 && NIkv > 0 && H > 0 && N > 0 && Nkv > 0 && nsrc > 0, "sc_attention: empty shape");
  UV_REQUIRE(d % 8 == 0 && d >= 8 && d <= 192, "sc_attention: head dim must be a multiple of 8 in [8, 192]");
  UV_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0, "sc_attention: row strides must be multiples of 8");
  UV_REQUIRE(((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) % 16 == 0, "sc_attention: 16-byte alignment");
  AttnParams p{};
  p.NI = NI;
  p.H = H;
  p.d = d;
  p.N = N;
  p.Nkv = Nkv;
  p.nsrc = nsrc;
  p.kv_src = kv_src;
  p.O = (__half*)O;
  p.ldo = ldo;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);

  // tile configuration: d <= 64 -> variant from UNIVST_ATTN_VARIANT (0: 2 query tiles x 128 keys, 1: + polynomial exp2,
  // 2: 4 query tiles x 64 keys, 3: + polynomial exp2, 4: 2 x 128 with half of the exp2 polynomial); 64 < d <= 128 -> 2 x 64; d > 128 -> 1 x 64
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("UNIVST_ATTN_VARIANT");
    variant = e ? atoi(e) : kDefaultVariant;
    if (variant < 0 || variant > 4) variant = kDefaultVariant;
  }
  const int bkv = (d <= 64 && variant < 2) ? 128 : 64;
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)H, (uint64_t)N, (uint64_t)NI};
    uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ldq * 2, (uint64_t)N * ldq * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    int r = make_tmap_f16(&tq, Q, 4, dims, str, box, true);
    if (r) return r;
  }
  {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)H, (uint64_t)Nkv, (uint64_t)NIkv};
    uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ldkv * 2, (uint64_t)Nkv * ldkv * 2};
    uint32_t box[4] = {64, 1, (uint32_t)bkv, 1};
    int r = make_tmap_f16(&tk, K, 4, dims, str, box, true);
    if (r) return r;
    r = make_tmap_f16(&tv, V, 4, dims, str, box, true);
    if (r) return r;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (d <= 64) {
    switch (variant) {
      case 0: return launch_attn<2, 128, 0>(tq, tk, tv, p, st);
      case 1: return launch_attn<2, 128, 1>(tq, tk, tv, p, st);
      case 2: return launch_attn<4, 64, 0>(tq, tk, tv, p, st);
      case 3: return launch_attn<4, 64, 1>(tq, tk, tv, p, st);
      default: return launch_attn<2, 128, 2>(tq, tk, tv, p, st);
    }
  }
  if (d <= 128) return launch_attn<2, 64, 0>(tq, tk, tv, p, st);
  return launch_attn<1, 64, 0>(tq, tk, tv, p, st);
}

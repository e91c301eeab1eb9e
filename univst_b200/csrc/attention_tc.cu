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
//   RS variants: the softmax row sums are computed by the tensor pipe (P times a block of ones into 16 extra accumulator
//   columns) instead of 128 FADDs per thread and tile.
//   Repeated K/V sources of an image (frame 0: [prev, self, first] = [0, 0, 0]) are streamed once with log2(count)
//   added to their scores -- exact, and 1/16 fewer tiles per clip.
// Head dims that are not multiples of 64 (SD-1.5: 40 / 80 / 160) are zero-filled by TMA out-of-bounds
// handling; nothing is padded in global memory.
#include <stdlib.h>

#include <type_traits>

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
  int stagger;       // start offset between the softmax groups of a CTA, clocks
  int dedupe;        // collapse repeated source images of a row into one pass with a log2(multiplicity) score bias
  // frame-sharded execution: sources >= NI live in PEER memory (NVLink P2P, read tile by tile by the TMA producer --
  // the K/V halo "exchange" is the attention kernel's own loads).  Source NI + (bank - 1) * bankB + b is image
  // b * bankFl + (bank == 1 ? bankFl - 1 : 0) of bank 1 (previous rank's buffer) / bank 2 (rank 0's buffer).
  int bankB;         // images per remote bank (branches); 0 = no remote banks
  int bankFl;        // frames per branch in a peer's buffer
  // joint attention (SD3): sources >= NIkv name image (src - NIkv) of a SECOND K/V tensor (bank 1) with its own token
  // count -- the text tokens appended to the image keys (video_diffusion_sd3/pnp_utils.py:100-102)
  int NIkv;          // images in the first K/V tensor
  int Nkv2;          // tokens per image of the second tensor; 0 = no second tensor
};

// K / V tensor maps: bank 0 = this rank's buffer, 1 = the previous rank's, 2 = rank 0's (same shape and strides)
struct KVMaps {
  CUtensorMap k[3];
  CUtensorMap v[3];
};

__device__ __forceinline__ void resolve_source(const AttnParams& p, int src, int& bank, int& image) {
  bank = 0;
  image = src;
  if (p.Nkv2 > 0) {
    if (src >= p.NIkv) {
      bank = 1;
      image = src - p.NIkv;
    }
  } else if (p.bankB > 0 && src >= p.NI) {
    const int r = src - p.NI;
    bank = 1 + r / p.bankB;
    image = (r - (bank - 1) * p.bankB) * p.bankFl + (bank == 1 ? p.bankFl - 1 : 0);
  }
}

// tokens of source i (the second K/V tensor has its own count) and its number of BKV-token tiles
__device__ __forceinline__ int src_tokens(const AttnParams& p, int image) {
  return (p.Nkv2 > 0 && image >= p.NIkv) ? p.Nkv2 : p.Nkv;
}

// Source list of one image with repeated entries collapsed.  softmax over [K_a, K_a, K_b] equals softmax over
// [K_a, K_b] with the scores of K_a raised by ln 2, so a source that occurs c times is streamed once with log2(c)
// added to its (log2-domain) scores: frame 0 of a clip ([prev, self, first] = [0, 0, 0]) costs one pass instead of
// three, frame 1 two instead of three.  Every role of the CTA derives the same list from the table row.
static constexpr int kMaxSrc = 4;
struct SrcList {
  int n;
  int img[kMaxSrc];
  int ntok[kMaxSrc];   // tokens of the source (the second K/V tensor of the joint attention has its own count)
  float bias[kMaxSrc];
};
__device__ __forceinline__ SrcList load_sources(const AttnParams& p, int img) {
  SrcList L;
  const int* row = p.kv_src + (size_t)img * p.nsrc;
  int cnt[kMaxSrc];
  L.n = 0;
#pragma unroll
  for (int u = 0; u < kMaxSrc; ++u) {
    L.img[u] = 0;
    cnt[u] = 0;
  }
#pragma unroll
  for (int s = 0; s < kMaxSrc; ++s) {
    if (s < p.nsrc) {
      const int v = row[s];
      bool found = false;
      if (p.dedupe) {
#pragma unroll
        for (int u = 0; u < kMaxSrc; ++u)
          if (u < L.n && L.img[u] == v && !found) {
            ++cnt[u];
            found = true;
          }
      }
      if (!found) {
#pragma unroll
        for (int u = 0; u < kMaxSrc; ++u)
          if (u == L.n) {
            L.img[u] = v;
            cnt[u] = 1;
          }
        ++L.n;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kMaxSrc; ++u) {
    L.bias[u] = cnt[u] == 2 ? 1.0f : cnt[u] == 3 ? 1.5849625007f : cnt[u] == 4 ? 2.0f : 0.0f;
    L.ntok[u] = src_tokens(p, L.img[u]);
  }
  return L;
}
__device__ __forceinline__ int src_img(const SrcList& L, int i) {
  return i == 0 ? L.img[0] : i == 1 ? L.img[1] : i == 2 ? L.img[2] : L.img[3];
}
__device__ __forceinline__ int src_ntok(const SrcList& L, int i) {
  return i == 0 ? L.ntok[0] : i == 1 ? L.ntok[1] : i == 2 ? L.ntok[2] : L.ntok[3];
}
__device__ __forceinline__ float src_bias(const SrcList& L, int i) {
  return i == 0 ? L.bias[0] : i == 1 ? L.bias[1] : i == 2 ? L.bias[2] : L.bias[3];
}

static constexpr uint32_t kOBase = 256;     // TMEM column of the first O accumulator
static constexpr float kRescaleThreshold = 8.0f;  // log2 units: P <= 2^8 before a forced rescale
static constexpr int kDefaultVariant = 19;

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

// Two of them per call on packed fp32 (FADD2 / FFMA2): 2 FMNMX + 3 FADD2-class + 3 FFMA2 + 2 IMAD for two exponentials, i.e.
// five issue slots per exponential against eight MUFU-pipe cycles per warp instruction.
__device__ __forceinline__ void ex2_poly2(float& e0, float& e1) {
  const float x0 = fmaxf(e0, -126.0f), x1 = fmaxf(e1, -126.0f);
  const unsigned long long x = pack_f32x2(x0, x1);
  const unsigned long long t = add2(x, pack_f32x2(12582912.0f, 12582912.0f));
  const unsigned long long u = add2(t, pack_f32x2(-12582912.0f, -12582912.0f));
  const unsigned long long f = fma2(u, pack_f32x2(-1.0f, -1.0f), x);
  unsigned long long q = fma2(pack_f32x2(0.0551716574f, 0.0551716574f), f, pack_f32x2(0.2426111251f, 0.2426111251f));
  q = fma2(q, f, pack_f32x2(0.6932609677f, 0.6932609677f));
  q = fma2(q, f, pack_f32x2(0.9999280572f, 0.9999280572f));
  float p0, p1, t0, t1;
  unpack_f32x2(q, p0, p1);
  unpack_f32x2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int NQ, int BKV>
struct AttnCfg {
  static constexpr int kDepth = (NQ == 1) ? 2 : 1;      // S slots per query tile
  static constexpr int kSlots = NQ * kDepth;              // S slots
  static constexpr int kPSlots = NQ * 2;                  // P tiles are double-buffered per query tile
  static constexpr int kThreads = 64 + 128 * NQ + 32 * (NQ - 1);
  static constexpr int kQChunkBytes = 128 * 128;        // [128 rows][64 halves]
  static constexpr int kKVChunkBytes = BKV * 128;       // [BKV rows][64 halves]
  static constexpr int kPBytes = 128 * BKV * 2;         // [128 rows][BKV halves] as BKV/64 swizzled chunks
  static constexpr int kBarriers = 40;
  static constexpr bool kConvergentIssue = BKV == 64;   // see the MMA issuer
  static_assert(kSlots * BKV <= 256, "S slots must fit below the O accumulators");
};

// Work decomposition: query tile q of the CTA streams KV tiles j = 0..T-1.  Tile (q, j) uses S/P slot
// q * kDepth + j % kDepth for the (j / kDepth)-th time.  Every query tile has its own MMA-issuing thread and its own
// softmax group, so the tiles only meet at the K/V ring.
// JT = 1: joint attention -- sources may belong to a second K/V tensor with its own token count (ragged tiles per source)
template <int NQ, int BKV, int POLY, int RS, int JT>
__global__ void __launch_bounds__(AttnCfg<NQ, BKV>::kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ KVMaps kvm, const AttnParams p) {
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
  // RS: the softmax row sums come out of the tensor pipe as P x 1 (a 16-key x 64-column block of fp16 ones, laid out
  // like a V tile, multiplied into 16 extra accumulator columns behind O) instead of 128 FADDs per thread and tile
  // RS == 2: P never touches shared memory -- the softmax threads store it to tensor memory (tcgen05.st) and P V reads
  // its A operand from there (no st.shared, no proxy fence, a third of the shared-memory traffic of the kernel)
  uint8_t* sOnes = sP + (RS == 2 ? 0 : L::kPSlots * L::kPBytes);   // [16][64] ones (RS only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + (RS ? 2048 : 0));
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
  if (RS && threadIdx.x == 0) {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if ((uint32_t)(reinterpret_cast<uint8_t*>(bars + L::kBarriers) - smem_raw) > dyn) {
      printf("univst_b200: attention shared-memory window too small / misaligned\n");
      __trap();
    }
  }

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int qt0 = blockIdx.x * NQ;     // first 128-row query tile of this CTA
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const SrcList SL = load_sources(p, img);
  const int tps = (p.Nkv + BKV - 1) / BKV;   // KV tiles per source of the first tensor
  auto tiles_of = [&](int si) { return JT ? (src_ntok(SL, si) + BKV - 1) / BKV : tps; };   // KV tiles of source si
  int T = SL.n * tps;                        // KV tiles in total
  if constexpr (JT) {
    T = 0;
#pragma unroll
    for (int si = 0; si < kMaxSrc; ++si)
      if (si < SL.n) T += (SL.ntok[si] + BKV - 1) / BKV;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&kvm.k[0]);
    tma_prefetch_desc(&kvm.v[0]);
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
  if (RS && warp == 2) {
    for (int i = lane; i < 2048 / 16; i += 32) st_shared_v4(smem_u32(sOnes) + i * 16, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t opad = (uint32_t)dpad;
  const uint32_t ostride = opad + (RS ? 16u : 0u);     // TMEM columns per query tile: O, then the row sums
  // TMEM map: S slots from column 0; RS == 2: one fp16 P tile per query tile (BKV / 2 columns) at 256, accumulators behind
  constexpr uint32_t kPT = 256;
  const uint32_t obase = (RS == 2) ? kPT + NQ * (BKV / 2) : kOBase;
  const bool is_mma = (warp == 1) || (warp >= 2 + 4 * NQ);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, NQ * q_bytes);
      for (int q = 0; q < NQ; ++q)
        for (int c = 0; c < dch; ++c)
          tma_load_4d(sQ + q * q_bytes + c * L::kQChunkBytes, &tmQ, q_full, c * 64, head, (qt0 + q) * 128, img);
      // K and V rings are fed independently (non-blocking polls): a K slot frees as soon as the Q K^T that read it has
      // run, a V slot only after the P V one softmax later, so a single in-order K, V, K, V ... stream would hold every
      // K load back behind the wait for a V slot and leave the score MMA starved by the TMA latency.
      int jk = 0, sik = 0, jtk = 0;   // next K tile, its source index and tile inside the source
      int jv = 0, siv = 0, jtv = 0;
      int tk = tiles_of(0), tv = tk;  // tiles of the current K / V source
      while (jk < T || jv < T) {
        bool progress = false;
        if (jk < T && (jk < 2 || mbar_test_wait(&k_empty[jk & 1], (uint32_t)(((jk >> 1) & 1) ^ 1)))) {
          const int s = jk & 1;
          mbar_expect_tx(&k_full[s], kv_bytes);
          int bank, image;
          resolve_source(p, src_img(SL, sik), bank, image);
          for (int c = 0; c < dch; ++c)
            tma_load_4d(sK + s * kv_bytes + c * L::kKVChunkBytes, &kvm.k[bank], &k_full[s], c * 64, head, jtk * BKV, image);
          if (++jtk == tk) {
            jtk = 0;
            tk = tiles_of(++sik);
          }
          ++jk;
          progress = true;
        }
        if (jv < T && (jv < 2 || mbar_test_wait(&v_empty[jv & 1], (uint32_t)(((jv >> 1) & 1) ^ 1)))) {
          const int s = jv & 1;
          mbar_expect_tx(&v_full[s], kv_bytes);
          int bank, image;
          resolve_source(p, src_img(SL, siv), bank, image);
          for (int c = 0; c < dch; ++c)
            tma_load_4d(sV + s * kv_bytes + c * L::kKVChunkBytes, &kvm.v[bank], &v_full[s], c * 64, head, jtv * BKV, image);
          if (++jtv == tv) {
            jtv = 0;
            tv = tiles_of(++siv);
          }
          ++jv;
          progress = true;
        }
        if (!progress) __nanosleep(64);
      }
    }
  } else if (is_mma) {
    // ---------------------------------------------------------------- MMA issuer of query tile q
    // Two issue styles (AttnCfg::kConvergentIssue), chosen by measurement.  Convergent (below): the whole warp runs the
    // loop, lane 0 probes the barriers (voted) and an elected lane issues each tcgen05 instruction -- the shortest
    // turn-around from "scores consumed" to "next scores issued", which is what bounds the 64-key tiles (head dims
    // 80 / 160: 442 -> 416 us and 124 -> 105 us per layer).  Single lane (further down): only lane 0 runs the loop;
    // measured faster for the MUFU-bound 128-key tiles of head dim 40 (4.9 vs 6.0-6.5 ms per layer).
    static_assert(NQ <= 2, "one MMA-issuing warp per query tile: warp 1 and warp 2 + 4 NQ");
    auto mma_role = [&](auto qc) {
    constexpr int q = decltype(qc)::value;   // compile-time: everything derived from it stays in uniform registers
    const uint32_t idesc_s = make_idesc_f16(128, BKV, 0, 0);      // S = Q K^T : both K-major
    const uint32_t idesc_o = make_idesc_f16(128, opad, 0, 1);     // O += P V  : P K-major, V MN-major
    const uint32_t tmem_o = tmem_base + obase + q * ostride;
    const uint32_t idesc_l = make_idesc_f16(128, 16, 0, 1);
    const uint64_t od = make_smem_desc_sw128(smem_u32(sOnes), 16, 1024);
    // Descriptor words are built once; per UMMA only the 14-bit start-address field of the low word moves (it cannot
    // carry out of the field: shared addresses are below 256 KiB).  The loop over KV tiles is unrolled by two so that
    // the ring stage / P buffer of a tile is a compile-time constant.
    const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ + q * q_bytes), 16, 1024);
    const uint64_t kd0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t vd0 = make_smem_desc_sw128(smem_u32(sV), L::kKVChunkBytes, 1024);
    const uint64_t pd0 = make_smem_desc_sw128(smem_u32(sP + (q * 2) * L::kPBytes), 16, 1024);
    const uint32_t kv16 = kv_bytes >> 4;
    auto issue_s = [&](int j, auto parc) {
      constexpr int par = decltype(parc)::value;          // j & 1
      const int slot = q * D + (D == 2 ? par : 0);
      mbar_wait_warp(&k_full[par], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint64_t kd = desc_advance(kd0, par * kv16);
      const uint32_t tmem_s = tmem_base + slot * BKV;
      int step = 0;
      for (int c = 0; c < dch; ++c) {
        const int nk = min(64, dpad - c * 64) >> 4;
        for (int k = 0; k < nk; ++k, ++step)
          umma_f16_ss_elect(tmem_s, desc_advance(qd, (c * L::kQChunkBytes + k * 32) >> 4),
                            desc_advance(kd, (c * L::kKVChunkBytes + k * 32) >> 4), idesc_s, step ? 1u : 0u);
      }
      tc_commit_elect(&k_empty[par]);
      tc_commit_elect(&s_full[slot]);
    };
    auto issue_pv = [&](int j, auto parc) {
      constexpr int par = decltype(parc)::value;          // j & 1: V ring stage and P buffer
      constexpr int slot = q * 2 + par;
      mbar_wait_warp(&p_full[slot], (uint32_t)((j >> 1) & 1));
      mbar_wait_warp(&v_full[par], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint64_t pd = desc_advance(pd0, par * (L::kPBytes >> 4));
      const uint64_t vd = desc_advance(vd0, par * kv16);
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k) {
        // A: P chunk (k / 4), 32-byte step inside the swizzle row.  B: V, 16 keys = 2 KiB further down;
        // the next 64 head-dim columns are a whole chunk away (LBO).
        umma_f16_ss_elect(tmem_o, desc_advance(pd, ((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4),
                          desc_advance(vd, (k * 2048) >> 4), idesc_o, (j | k) ? 1u : 0u);
        if constexpr (RS)
          umma_f16_ss_elect(tmem_o + opad, desc_advance(pd, ((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4), od, idesc_l,
                            (j | k) ? 1u : 0u);
      }
      tc_commit_elect(&v_empty[par]);
      tc_commit_elect(&o_done[slot]);
    };
    using P0 = std::integral_constant<int, 0>;
    using P1 = std::integral_constant<int, 1>;
    mbar_wait_warp(q_full, 0);
    tc_fence_after();
    // The tile's slots are filled up front; afterwards a slot is refilled with the next scores that map to it as
    // soon as the softmax group has pulled the current tile into registers (s_free), i.e. *during* that tile's
    // exponentials -- a group never waits for the tensor pipe between two tiles.
    issue_s(0, P0{});
    if (D == 2 && T > 1) issue_s(1, P1{});
    auto step = [&](int j, auto parc) {
      constexpr int par = decltype(parc)::value;
      if (j + D < T) {
        mbar_wait_warp(&s_free[q * D + (D == 2 ? par : 0)], (uint32_t)((j / D) & 1));
        tc_fence_after();
        if constexpr (D == 2) issue_s(j + 2, parc);
        else issue_s(j + 1, std::integral_constant<int, 1 - par>{});
      }
      issue_pv(j, parc);
    };
    for (int j = 0; j < T; j += 2) {
      step(j, P0{});
      if (j + 1 < T) step(j + 1, P1{});
    }
    };
    if constexpr (L::kConvergentIssue) {
      if (warp == 1) {
        mma_role(std::integral_constant<int, 0>{});
      } else {
        if constexpr (NQ > 1) mma_role(std::integral_constant<int, 1>{});
      }
    } else {
      if (lane == 0) {
        // ---------------------------------------------------------------- MMA issuer of query tile q
        const int q = (warp == 1) ? 0 : (int)(warp - (2 + 4 * NQ)) + 1;
        const uint32_t idesc_s = make_idesc_f16(128, BKV, 0, 0);      // S = Q K^T : both K-major
        const uint32_t idesc_o = make_idesc_f16(128, opad, 0, 1);     // O += P V  : P K-major, V MN-major
        const uint32_t idesc_l = make_idesc_f16(128, 16, 0, 1);       // l += P 1
        const uint64_t od = make_smem_desc_sw128(smem_u32(sOnes), 16, 1024);
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
            if constexpr (RS == 2) {
              const uint32_t ta = tmem_base + kPT + q * (BKV / 2) + k * 8;   // 16 keys = 8 packed columns
              umma_f16_ts(tmem_base + obase + q * ostride, ta, db, idesc_o, (j | k) ? 1u : 0u);
              umma_f16_ts(tmem_base + obase + q * ostride + opad, ta, od, idesc_l, (j | k) ? 1u : 0u);
            } else {
              umma_f16_ss(tmem_base + obase + q * ostride, da, db, idesc_o, (j | k) ? 1u : 0u);
              if constexpr (RS) umma_f16_ss(tmem_base + obase + q * ostride + opad, da, od, idesc_l, (j | k) ? 1u : 0u);
            }
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
    }
  } else {
    // ------------------------------------------------------------------ softmax group g = query tile g
    // One thread per query row.  The S tile is pulled into registers in one go and its TMEM slot released at once
    // (the next Q K^T of this query tile runs under this tile's exponentials).  The tile is then consumed in 32-column
    // chunks: the max of chunk c + 1 is computed in the same basic block as the exponentials of chunk c, so FMNMX and
    // MUFU work are mixed instead of alternating in long phases, and a chunk is checked against the running (lazily
    // updated) reference maximum right before its exponentials.
    // The softmax groups of a CTA start staggered by about half a tile (p.stagger clocks): two warps of a scheduler
    // that run in lockstep leave the MUFU pipe idle whenever both are between their exponentials; out of phase, one
    // warp's loads / max / barrier work hides under the other's exponentials.
    const int g = (int)(warp - 2) >> 2;
    const uint32_t quad = warp & 3;                     // TMEM lane quadrant this warp may touch
    const uint32_t r = quad * 32 + lane;                // row inside the 128-row tile
    const uint32_t lane_off = (quad * 32) << 16;
    const uint32_t o_addr = tmem_base + obase + g * ostride + lane_off;
    const uint32_t p_taddr = tmem_base + kPT + g * (BKV / 2) + lane_off;   // RS == 2: this row's P tile in TMEM
    const uint32_t prow0 = smem_u32(sP) + r * 128;
    const uint32_t swz = r & 7;
    constexpr int NC = BKV / 32;
    float m_used = -INFINITY;   // max baked into O and l
    float l = 0.0f;
    int jt = 0;                 // tile index inside the current source image
    int si = 0;                 // current source
    float bias = SL.bias[0];    // log2 multiplicity of the current source
    int stok = JT ? SL.ntok[0] : p.Nkv, stiles = (stok + BKV - 1) / BKV;   // its tokens / tiles
    for (int j = 0; j < T; ++j) {
      const int slot = g * D + j % D;
      const int pslot = g * 2 + (j & 1);
      const uint32_t prow = prow0 + pslot * L::kPBytes;
      mbar_wait(&s_full[slot], (uint32_t)((j / D) & 1));
      if (j == 0 && g > 0 && p.stagger > 0) {   // counted from the arrival of the first scores, not from the launch
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)g * p.stagger) {
        }
      }
      tc_fence_after();
      // S tile -> registers: all loads in flight, one wait
      uint32_t sraw[BKV];
      {
        const uint32_t sa = tmem_base + slot * BKV + lane_off;
#pragma unroll
        for (int c = 0; c < NC; ++c) tmem_ld32(sa + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sraw[c * 32]));
        tc_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&s_free[slot]);  // the slot may be overwritten with the next scores from here on
      const int valid = stok - jt * BKV;   // >= BKV except on the ragged last tile of a source
      const float tbias = bias;
      if (++jt == stiles) {
        jt = 0;
        bias = src_bias(SL, ++si);
        if constexpr (JT) {
          stok = src_ntok(SL, si);
          stiles = (stok + BKV - 1) / BKV;
        }
      }
      if (valid < BKV) {
#pragma unroll
        for (int x = 0; x < BKV; ++x)
          if (x >= valid) sraw[x] = 0xff800000u;  // -inf
      }
      if (RS != 2 && j >= 2) {  // the P buffer was last read by the P V of tile j - 2: long done, the wait is (almost) free
        mbar_wait(&o_done[pslot], (uint32_t)(((j - 2) >> 1) & 1));
      }
      auto chunk_max = [&](int c) {
        float mx4[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) mx4[x] = __uint_as_float(sraw[c * 32 + x]);
#pragma unroll
        for (int x = 4; x < 32; ++x) mx4[x & 3] = fmaxf(mx4[x & 3], __uint_as_float(sraw[c * 32 + x]));
        return fmaf(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), p.scale_log2, tbias);
      };
      float ls4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      float mx_next = chunk_max(0);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float mx = mx_next;
        const bool need = mx > m_used + kRescaleThreshold;   // very first chunk: m_used = -inf -> always true
        if (__any_sync(0xffffffffu, need)) {
          // rare path: move the reference maximum.  Everything already scaled by 2^-m_used follows: the row sum, the
          // P chunks of this tile that sit in shared memory, and the O accumulator in TMEM (for which the previous
          // P V of this query tile must have landed).
          const float m_new = fmaxf(m_used, mx);
          const float alpha = (m_new == m_used) ? 1.0f : ex2_approx(m_used - m_new);
          if constexpr (!RS) {
            l *= alpha;
#pragma unroll
            for (int x = 0; x < 4; ++x) ls4[x] *= alpha;
          }
          if (c > 0) {
            const __half2 a2 = __float2half2_rn(alpha);
            if constexpr (RS == 2) {
              tc_wait_st();   // this tile's earlier P stores must have landed before they are read back
#pragma unroll 1
              for (int pc = 0; pc < c; ++pc) {   // chunks of 32 keys = 16 packed columns already stored for this tile
                uint32_t w[16];
                tmem_ld16(p_taddr + pc * 16, w);
                tc_wait_ld();
#pragma unroll
                for (int x = 0; x < 16; ++x) {
                  __half2 h = __hmul2(*reinterpret_cast<__half2*>(&w[x]), a2);
                  w[x] = *reinterpret_cast<uint32_t*>(&h);
                }
                tmem_st16(p_taddr + pc * 16, w);
              }
            } else {
#pragma unroll 1
              for (int pc = 0; pc < c * 4; ++pc) {
                const uint32_t addr = prow + (pc >> 3) * (128 * 128) + ((((uint32_t)pc & 7) ^ swz) << 4);
                uint32_t w[4];
                ld_shared_v4(addr, w);
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  __half2 h = __hmul2(*reinterpret_cast<__half2*>(&w[x]), a2);
                  w[x] = *reinterpret_cast<uint32_t*>(&h);
                }
                st_shared_v4(addr, w[0], w[1], w[2], w[3]);
              }
            }
          }
          if (j > 0) {
            mbar_wait(&o_done[g * 2 + ((j - 1) & 1)], (uint32_t)(((j - 1) >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (uint32_t oc = 0; oc < ostride; oc += 16) {   // O and (RS) the row sums behind it
              uint32_t t[16];
              tmem_ld16(o_addr + oc, t);
              tc_wait_ld();
#pragma unroll
              for (int x = 0; x < 16; ++x) t[x] = __float_as_uint(__uint_as_float(t[x]) * alpha);
              tmem_st16(o_addr + oc, t);
            }
            tc_wait_st();
          }
          m_used = m_new;
        }
        // P = 2^(s * scale - m_used): one FFMA + one MUFU.EX2 per element, fp16, written as swizzled K-major
        // chunks (chunk kc holds keys [64 kc, 64 kc + 64)); the row sum runs in 4 independent chains
        const float neg_m = tbias - m_used;
        if (c + 1 < NC) mx_next = chunk_max(c + 1);
        // (all 32 exponentials of the chunk are issued back to back before the first one is consumed: the MUFU
        // latency is then paid once per chunk instead of once per few elements, which is what lets two warps keep
        // the pipe busy)
        float e[32];
        if constexpr (POLY == 6) {   // packed FFMA2: two scores per issue slot
#pragma unroll
          for (int x = 0; x < 32; x += 2)
            fma2_bcast(e[x], e[x + 1], __uint_as_float(sraw[c * 32 + x]), __uint_as_float(sraw[c * 32 + x + 1]), p.scale_log2, neg_m);
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) e[x] = fmaf(__uint_as_float(sraw[c * 32 + x]), p.scale_log2, neg_m);
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          // POLY = 1: every 4th exponential off the MUFU pipe; 4: every 8th; 0 / 6: none (6 = packed FFMA2 score scaling)
          const bool poly = (POLY == 1 && (x & 3) == 3) || (POLY == 4 && (x & 7) == 7);
          e[x] = poly ? ex2_poly(e[x]) : ex2_approx(e[x]);
        }
        if constexpr (RS == 2) {
          uint32_t w[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) w[x] = pack_half2(e[2 * x], e[2 * x + 1]);
          if (c == 0 && j >= 1) {   // the single P tile was last read by the P V of tile j - 1
            mbar_wait(&o_done[g * 2 + ((j - 1) & 1)], (uint32_t)(((j - 1) >> 1) & 1));
            tc_fence_after();
          }
          tmem_st16(p_taddr + c * 16, w);
        } else {
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            uint32_t w[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              if constexpr (!RS) ls4[x] += e[q8 * 8 + 2 * x] + e[q8 * 8 + 2 * x + 1];
              w[x] = pack_half2(e[q8 * 8 + 2 * x], e[q8 * 8 + 2 * x + 1]);
            }
            const int pc = c * 4 + q8;   // 16-byte piece of the P row
            st_shared_v4(prow + (pc >> 3) * (128 * 128) + ((((uint32_t)pc & 7) ^ swz) << 4), w[0], w[1], w[2], w[3]);
          }
        }
      }
      l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      if constexpr (RS == 2) tc_wait_st();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[pslot]);
    }
    // ------------------------------------------------------------------ epilogue: O / l -> global
    if (T >= 2) mbar_wait(&o_done[g * 2 + ((T - 2) & 1)], (uint32_t)(((T - 2) >> 1) & 1));
    mbar_wait(&o_done[g * 2 + ((T - 1) & 1)], (uint32_t)(((T - 1) >> 1) & 1));
    tc_fence_after();
    if constexpr (RS) {
      uint32_t t[16];
      tmem_ld16(o_addr + opad, t);
      tc_wait_ld();
      l = __uint_as_float(t[0]);
    }
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

// ------------------------------------------------------------------------------------------------------------------
// Split-row variant for head dims <= 48 (SD-1.5's 64 x 64 level, the layer that dominates the clip): two query tiles x
// 128 keys as above, but every query row is shared by TWO softmax threads, each owning 64 of the 128 key columns of a
// tile with its OWN running maximum, its own P chunk, and its own O / row-sum accumulator in TMEM (O_half += P_half
// V_half).  The two halves are independent online softmaxes over disjoint key subsets and are merged once, in the
// epilogue: O = (w_a O_a + w_b O_b) / (w_a l_a + w_b l_b), w_x = 2^(m_x - max(m_a, m_b)).  No per-tile exchange is needed, and
// the CTA runs 16 softmax warps (4 per scheduler) instead of 8: with one row per thread the MUFU pipe sat idle ~40 % of
// the time because two warps per scheduler cannot cover each other's TMEM loads, maxima, packs and barrier waits.
// TMEM: S 2 x 128 columns, then 4 accumulators of (dpad + 16) columns: 256 + 4 x 64 = 512 at d = 40.
template <int POLY, int CI>
__global__ void __launch_bounds__(608, 1)   // 19 warps -> 96 registers (the allocation unit is 16 per thread: 104 does not fit)
attention_tc_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ KVMaps kvm, const AttnParams p) {
  constexpr int NQ = 2, BKV = 128;
  constexpr uint32_t kTile = 128 * 128;     // bytes of a [128][64]-half tile chunk
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int dpad = (p.d + 15) & ~15;
  uint8_t* sQ = smem;                         // [2][128][64]
  uint8_t* sK = sQ + NQ * kTile;              // [2][128][64]
  uint8_t* sV = sK + 2 * kTile;               // [2][128][64]
  uint8_t* sP = sV + 2 * kTile;               // [2 q][2 buffers][2 halves][128][64]
  uint8_t* sOnes = sP + 8 * kTile;            // [16][64] ones
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + 2048);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* k_empty = bars + 3;       // 2
  uint64_t* v_full = bars + 5;        // 2
  uint64_t* v_empty = bars + 7;       // 2
  uint64_t* s_full = bars + 9;        // 2
  uint64_t* s_free = bars + 11;       // 2
  uint64_t* p_full = bars + 13;       // 4
  uint64_t* o_done = bars + 17;       // 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  if (threadIdx.x == 0) {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if ((uint32_t)(reinterpret_cast<uint8_t*>(bars + 24) - smem_raw) > dyn) {
      printf("univst_b200: attention shared-memory window too small / misaligned\n");
      __trap();
    }
  }
  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int qt0 = blockIdx.x * NQ;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const SrcList SL = load_sources(p, img);
  const int tps = (p.Nkv + BKV - 1) / BKV;   // every source has Nkv tokens here (no second K/V tensor: see the launcher)
  const int T = SL.n * tps;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&kvm.k[0]);
    tma_prefetch_desc(&kvm.v[0]);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], NQ);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], NQ);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 256);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&p_full[s], 256);
      mbar_init(&o_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp == 2) {
    for (int i = lane; i < 2048 / 16; i += 32) st_shared_v4(smem_u32(sOnes) + i * 16, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t opad = (uint32_t)dpad;
  const uint32_t ostride = opad + 16u;          // O, then the row sums
  const bool is_mma = (warp == 1) || (warp == 18);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer (K and V rings polled independently)
      mbar_expect_tx(q_full, NQ * kTile);
      for (int q = 0; q < NQ; ++q) tma_load_4d(sQ + q * kTile, &tmQ, q_full, 0, head, (qt0 + q) * 128, img);
      int jk = 0, sik = 0, jtk = 0;
      int jv = 0, siv = 0, jtv = 0;
      while (jk < T || jv < T) {
        bool progress = false;
        if (jk < T && (jk < 2 || mbar_test_wait(&k_empty[jk & 1], (uint32_t)(((jk >> 1) & 1) ^ 1)))) {
          const int s = jk & 1;
          mbar_expect_tx(&k_full[s], kTile);
          int bank, image;
          resolve_source(p, src_img(SL, sik), bank, image);
          tma_load_4d(sK + s * kTile, &kvm.k[bank], &k_full[s], 0, head, jtk * BKV, image);
          if (++jtk == tps) {
            jtk = 0;
            ++sik;
          }
          ++jk;
          progress = true;
        }
        if (jv < T && (jv < 2 || mbar_test_wait(&v_empty[jv & 1], (uint32_t)(((jv >> 1) & 1) ^ 1)))) {
          const int s = jv & 1;
          mbar_expect_tx(&v_full[s], kTile);
          int bank, image;
          resolve_source(p, src_img(SL, siv), bank, image);
          tma_load_4d(sV + s * kTile, &kvm.v[bank], &v_full[s], 0, head, jtv * BKV, image);
          if (++jtv == tps) {
            jtv = 0;
            ++siv;
          }
          ++jv;
          progress = true;
        }
        if (!progress) __nanosleep(64);
      }
    }
  } else if (is_mma) {
    // CI = 1: the whole warp runs the role convergently (lane 0 probes the barriers, an elected lane issues), so that
    // descriptors live in uniform registers; CI = 0: lane 0 alone.  The issuers share their schedulers with four softmax
    // warps each, so their instruction count matters.
    auto wait = [&](uint64_t* bar, uint32_t par) {
      if constexpr (CI) mbar_wait_warp(bar, par);
      else mbar_wait(bar, par);
    };
    auto mma = [&](uint32_t dt, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
      if constexpr (CI) umma_f16_ss_elect(dt, da, db, idesc, accum);
      else umma_f16_ss(dt, da, db, idesc, accum);
    };
    auto commit = [&](uint64_t* bar) {
      if constexpr (CI) tc_commit_elect(bar);
      else tc_commit(bar);
    };
    if (CI || lane == 0) {
      // ---------------------------------------------------------------- MMA issuer of query tile q
      const int q = (warp == 1) ? 0 : 1;
      const uint32_t idesc_s = make_idesc_f16(128, BKV, 0, 0);      // S = Q K^T : both K-major
      const uint32_t idesc_o = make_idesc_f16(128, opad, 0, 1);     // O_half += P_half V_half : V MN-major
      const uint32_t idesc_l = make_idesc_f16(128, 16, 0, 1);       // l_half += P_half 1
      const uint64_t od = make_smem_desc_sw128(smem_u32(sOnes), 16, 1024);
      const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ + q * kTile), 16, 1024);
      const uint64_t kd0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t vd0 = make_smem_desc_sw128(smem_u32(sV), kTile, 1024);
      const uint64_t pd0 = make_smem_desc_sw128(smem_u32(sP + q * 4 * kTile), 16, 1024);
      const int nk = dpad >> 4;
      auto issue_s = [&](int j) {
        const int ks = j & 1;
        wait(&k_full[ks], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint64_t kd = desc_advance(kd0, ks * (kTile >> 4));
        for (int k = 0; k < nk; ++k)
          mma(tmem_base + q * BKV, desc_advance(qd, k * 2), desc_advance(kd, k * 2), idesc_s, k ? 1u : 0u);
        commit(&k_empty[ks]);
        commit(&s_full[q]);
      };
      auto issue_pv = [&](int j) {
        const int vs = j & 1, slot = q * 2 + (j & 1);
        wait(&p_full[slot], (uint32_t)((j >> 1) & 1));
        wait(&v_full[vs], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint64_t pd = desc_advance(pd0, (j & 1) * ((2 * kTile) >> 4));
        const uint64_t vd = desc_advance(vd0, vs * (kTile >> 4));
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const int hf = k >> 2;                                     // which half of the keys / which accumulator
          const uint32_t acc = tmem_base + kOBase + (q * 2 + hf) * ostride;
          const uint64_t da = desc_advance(pd, (hf * kTile + (k & 3) * 32) >> 4);
          const uint64_t db = desc_advance(vd, (k * 2048) >> 4);
          const uint32_t accum = (j | (k & 3)) ? 1u : 0u;
          mma(acc, da, db, idesc_o, accum);
          mma(acc + opad, da, od, idesc_l, accum);
        }
        commit(&v_empty[vs]);
        commit(&o_done[slot]);
      };
      wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) {
          wait(&s_free[q], (uint32_t)(j & 1));
          tc_fence_after();
          issue_s(j + 1);
        }
        issue_pv(j);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: group g = query tile, half hf = key half
    const int sw = (int)warp - 2;                       // 0 .. 15
    const int g = sw >> 3;
    const int hf = (sw >> 2) & 1;
    const uint32_t quad = warp & 3;                     // TMEM lane quadrant this warp may touch
    const uint32_t r = quad * 32 + lane;                // row inside the 128-row tile
    const uint32_t lane_off = (quad * 32) << 16;
    const uint32_t acc_addr = tmem_base + kOBase + (g * 2 + hf) * ostride + lane_off;
    const uint32_t prow0 = smem_u32(sP) + hf * kTile + r * 128;
    const uint32_t swz = r & 7;
    float m_used = -INFINITY;   // max baked into this half's O and l
    int jt = 0, si = 0;
    float bias = SL.bias[0];
    for (int j = 0; j < T; ++j) {
      const int pslot = g * 2 + (j & 1);
      const uint32_t prow = prow0 + pslot * 2 * kTile;
      mbar_wait(&s_full[g], (uint32_t)(j & 1));
      tc_fence_after();
      uint32_t sraw[64];
      {
        const uint32_t sa = tmem_base + g * BKV + hf * 64 + lane_off;
        tmem_ld32(sa, *reinterpret_cast<uint32_t(*)[32]>(&sraw[0]));
        tmem_ld32(sa + 32, *reinterpret_cast<uint32_t(*)[32]>(&sraw[32]));
        tc_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&s_free[g]);
      const int valid = p.Nkv - jt * BKV - hf * 64;   // valid columns of this half (>= 64 except on a ragged last tile)
      const float tbias = bias;
      if (++jt == tps) {
        jt = 0;
        bias = src_bias(SL, ++si);
      }
      if (valid < 64) {
#pragma unroll
        for (int x = 0; x < 64; ++x)
          if (x >= valid) sraw[x] = 0xff800000u;  // -inf
      }
      if (j >= 2) mbar_wait(&o_done[pslot], (uint32_t)(((j - 2) >> 1) & 1));   // P buffer free (P V of tile j - 2)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float mx4[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) mx4[x] = __uint_as_float(sraw[c * 32 + x]);
#pragma unroll
        for (int x = 4; x < 32; ++x) mx4[x & 3] = fmaxf(mx4[x & 3], __uint_as_float(sraw[c * 32 + x]));
        const float mx = fmaf(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), p.scale_log2, tbias);
        const bool need = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rare path: move this half's reference maximum; its P pieces of this tile and its accumulators follow
          const float m_new = fmaxf(m_used, mx);
          const float alpha = (m_new == m_used) ? 1.0f : ex2_approx(m_used - m_new);
          if (c > 0) {
            const __half2 a2 = __float2half2_rn(alpha);
#pragma unroll 1
            for (int pc = 0; pc < 4; ++pc) {
              const uint32_t addr = prow + ((((uint32_t)pc & 7) ^ swz) << 4);
              uint32_t w[4];
              ld_shared_v4(addr, w);
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                __half2 h = __hmul2(*reinterpret_cast<__half2*>(&w[x]), a2);
                w[x] = *reinterpret_cast<uint32_t*>(&h);
              }
              st_shared_v4(addr, w[0], w[1], w[2], w[3]);
            }
          }
          if (j > 0) {
            mbar_wait(&o_done[g * 2 + ((j - 1) & 1)], (uint32_t)(((j - 1) >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (uint32_t oc = 0; oc < ostride; oc += 16) {
              uint32_t t[16];
              tmem_ld16(acc_addr + oc, t);
              tc_wait_ld();
#pragma unroll
              for (int x = 0; x < 16; ++x) t[x] = __float_as_uint(__uint_as_float(t[x]) * alpha);
              tmem_st16(acc_addr + oc, t);
            }
            tc_wait_st();
          }
          m_used = m_new;
        }
        // a half with no valid key so far keeps m_used = -inf: its scores are all -inf and must map to P = 0, not NaN
        const float neg_m = (m_used == -INFINITY) ? 0.0f : tbias - m_used;
        float e[32];
        if constexpr (POLY >= 6) {
#pragma unroll
          for (int x = 0; x < 32; x += 2)
            fma2_bcast(e[x], e[x + 1], __uint_as_float(sraw[c * 32 + x]), __uint_as_float(sraw[c * 32 + x + 1]), p.scale_log2, neg_m);
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) e[x] = fmaf(__uint_as_float(sraw[c * 32 + x]), p.scale_log2, neg_m);
        }
        if constexpr (POLY >= 7) {   // packed polynomial exponentials for 1/4 (7), 1/8 (8) or 1/2 (10) of the pairs
#pragma unroll
          for (int x = 0; x < 32; x += 2) {
            const bool poly = (POLY == 7 && (x & 7) == 6) || (POLY == 8 && (x & 15) == 14) || (POLY == 10 && (x & 3) == 2);
            if (poly) {
              ex2_poly2(e[x], e[x + 1]);
            } else {
              e[x] = ex2_approx(e[x]);
              e[x + 1] = ex2_approx(e[x + 1]);
            }
          }
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) {
            const bool poly = (POLY == 1 && (x & 3) == 3) || (POLY == 4 && (x & 7) == 7);
            e[x] = poly ? ex2_poly(e[x]) : ex2_approx(e[x]);
          }
        }
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          uint32_t w[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) w[x] = pack_half2(e[q8 * 8 + 2 * x], e[q8 * 8 + 2 * x + 1]);
          const int pc = c * 4 + q8;   // 16-byte piece of this half's 128-byte P row
          st_shared_v4(prow + ((((uint32_t)pc & 7) ^ swz) << 4), w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[pslot]);
    }
    // ------------------------------------------------------------------ epilogue: merge the halves, O / l -> global
    if (T >= 2) mbar_wait(&o_done[g * 2 + ((T - 2) & 1)], (uint32_t)(((T - 2) >> 1) & 1));
    mbar_wait(&o_done[g * 2 + ((T - 1) & 1)], (uint32_t)(((T - 1) >> 1) & 1));
    tc_fence_after();
    float l_own;
    {
      uint32_t t[16];
      tmem_ld16(acc_addr + opad, t);
      tc_wait_ld();
      l_own = __uint_as_float(t[0]);
    }
    // exchange (m, l) with the thread that owns the other half of this row, through the (dead) Q tile of this group
    float2* xch = reinterpret_cast<float2*>(sQ + g * kTile);          // [2 halves][128 rows]
    xch[hf * 128 + r] = make_float2(m_used, l_own);
    asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
    const float2 other = xch[(1 - hf) * 128 + r];
    const float m_all = fmaxf(m_used, other.x);
    const float w_own = (m_used == -INFINITY) ? 0.0f : ex2_approx(m_used - m_all);
    const float w_oth = (other.x == -INFINITY) ? 0.0f : ex2_approx(other.x - m_all);
    const float inv_l = 1.0f / (w_own * l_own + w_oth * other.y);
    const float s_own = w_own * inv_l, s_oth = w_oth * inv_l;
    const uint32_t oth_addr = tmem_base + kOBase + (g * 2 + (1 - hf)) * ostride + lane_off;
    const int qrow = (qt0 + g) * 128 + (int)r;
    const bool row_ok = qrow < p.N;
    __half* orow = p.O + ((size_t)img * p.N + qrow) * p.ldo + head * p.d;
    // 16-column chunks alternate between the two threads of a row
    for (uint32_t c = (uint32_t)hf * 16; c < opad; c += 32) {
      uint32_t ta[16], tb[16];
      tmem_ld16(acc_addr + c, ta);
      tmem_ld16(oth_addr + c, tb);
      tc_wait_ld();
      if (!row_ok) continue;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        const int col = (int)c + h8 * 8;
        if (col < p.d) {
          uint32_t w[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float v0 = __uint_as_float(ta[h8 * 8 + 2 * x]) * s_own + __uint_as_float(tb[h8 * 8 + 2 * x]) * s_oth;
            const float v1 = __uint_as_float(ta[h8 * 8 + 2 * x + 1]) * s_own + __uint_as_float(tb[h8 * 8 + 2 * x + 1]) * s_oth;
            w[x] = pack_half2(v0, v1);
          }
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

template <int POLY, int CI = 0>
static int launch_attn_split(const CUtensorMap& tq, const KVMaps& kvm, const AttnParams& p, cudaStream_t stream) {
  UV_REQUIRE(p.d <= 48, "attention (split rows): head dim <= 48");
  const size_t smem = 227 * 1024;   // 14 tiles of 16 KiB + ones + barriers: 704 B of slack for the 1 KiB alignment
  static bool configured = false;
  if (!configured) {
    UV_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_split_kernel<POLY, CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  dim3 grid((p.N + 255) / 256, p.H, p.NI);
  attention_tc_split_kernel<POLY, CI><<<grid, 608, smem, stream>>>(tq, kvm, p);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

template <int NQ, int BKV, int POLY, int RS = 0, int JT = 0>
static int launch_attn(const CUtensorMap& tq, const KVMaps& kvm, const AttnParams& p, cudaStream_t stream) {
  using L = AttnCfg<NQ, BKV>;
  const int dch = (p.d + 63) / 64;
  size_t smem = (size_t)NQ * dch * L::kQChunkBytes + 4 * (size_t)dch * L::kKVChunkBytes +
                      (RS == 2 ? 0 : (size_t)L::kPSlots * L::kPBytes) + (RS ? 2048 : 0) + L::kBarriers * sizeof(uint64_t) +
                      1024;   // alignment slack
  // RS at 2 x 128: 224 KiB of tiles + 2 KiB ones + barriers leave 704 B of slack; the kernel traps if the dynamic window
  // turns out to be less aligned than that (it is 1 KiB aligned in practice)
  if (RS && smem > 227 * 1024) smem = 227 * 1024;
  UV_REQUIRE(smem <= 227 * 1024, "attention: tile configuration needs %zu bytes of shared memory", smem);
  UV_REQUIRE((RS == 2 ? 256 + NQ * (BKV / 2) : 256) + NQ * (((p.d + 15) & ~15) + (RS ? 16 : 0)) <= 512,
             "attention: O accumulators do not fit into TMEM");
  static_assert(RS != 2 || (NQ == 2 && BKV == 128), "P in TMEM: 2 query tiles x 128 keys only");
  static bool configured = false;
  if (!configured) {
    UV_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<NQ, BKV, POLY, RS, JT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
    configured = true;
  }
  dim3 grid((p.N + 128 * NQ - 1) / (128 * NQ), p.H, p.NI);
  attention_tc_kernel<NQ, BKV, POLY, RS, JT><<<grid, L::kThreads, smem, stream>>>(tq, kvm, p);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

}  // namespace uv

using namespace uv;

static int g_variant = -1, g_dedupe = -1, g_stagger = -2;

extern "C" int univst_attention_tune(int32_t variant, int32_t dedupe, int32_t stagger) {
  g_variant = (variant < 0 || variant > 22) ? -1 : variant;   // -1: back to the environment / built-in default
  g_dedupe = dedupe < 0 ? -1 : (dedupe != 0);
  g_stagger = stagger < 0 ? -2 : stagger;
  return UNIVST_OK;
}

// Shared body of the two entry points.  Kb / Vb: K / V base pointers of bank 0 (local), 1 (previous rank), 2 (rank 0);
// banks 1 and 2 may be null (no remote sources).
static int sc_attention_impl(const void* Q, int32_t ldq, const void* const* Kb, const void* const* Vb, int32_t ldkv,
                             int32_t NI, int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv, const int32_t* kv_src,
                             int32_t nsrc, void* O, int32_t ldo, int32_t bankB, int32_t bankFl, void* stream,
                             int32_t ldkv2 = 0, int32_t NIkv2 = 0, int32_t Nkv2 = 0) {
  const void* K = Kb[0];
  const void* V = Vb[0];
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
  p.bankB = bankB;
  p.bankFl = bankFl;
  p.NIkv = NIkv;
  p.Nkv2 = Nkv2;

  // tile configuration: d <= 64 -> variant from univst_attention_tune / UNIVST_ATTN_VARIANT (see the switches below).
  // Default 19: d <= 48 -> split rows (two softmax threads per query row, 16 softmax warps) with the row sums on the
  // tensor pipe, packed FFMA2 score scaling, every exp2 on the MUFU pipe and convergent (elected-lane) MMA issue;
  // 48 < d <= 64 -> the same without the row split (variant 16).  Measured per 64x64 layer, same run: 4.36 ms (19),
  // 4.41 (17 = 19 with single-lane MMA issue), 4.59 (16), 4.61 (13 = 17 without FFMA2),
  // 4.70 (9 = 16 without FFMA2), 4.77 (0: row sums as FADDs), 5.27 (1: a quarter of the exp2 as polynomials -- the
  // extra FMA / ALU instructions cost more issue slots than the MUFU relief returns).  The kernel is bound by issue
  // slots around the MUFU work, so every instruction removed from the softmax loop shows.
  // 64 < d <= 128 -> 2 x 64; d > 128 -> 1 x 64
  int& variant = g_variant;
  if (variant < 0) {
    const char* e = getenv("UNIVST_ATTN_VARIANT");
    variant = e ? atoi(e) : kDefaultVariant;
    if (variant < 0 || variant > 22) variant = kDefaultVariant;
  }
  int& dedupe = g_dedupe;
  if (dedupe < 0) {
    const char* e = getenv("UNIVST_ATTN_DEDUPE");
    dedupe = e ? (atoi(e) != 0) : 1;
  }
  UV_REQUIRE(nsrc <= kMaxSrc, "sc_attention: at most %d K/V sources per image", kMaxSrc);
  p.dedupe = dedupe;
  const int bkv = d <= 64 ? 128 : 64;   // must match the launch_attn<> picked below
  int& stagger = g_stagger;
  if (stagger == -2) {
    const char* e = getenv("UNIVST_ATTN_STAGGER");
    stagger = e ? atoi(e) : -1;
  }
  p.stagger = stagger >= 0 ? stagger : 0;   // measured neutral (0 .. 2500 clocks): off by default
  CUtensorMap tq;
  KVMaps kvm;
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
    for (int b = 0; b < 3; ++b) {   // absent banks alias the local buffer (never addressed: bankB = 0 or the table has no such source)
      uint64_t bdims[4] = {dims[0], dims[1], dims[2], dims[3]}, bstr[3] = {str[0], str[1], str[2]};
      if (b == 1 && Nkv2 > 0) {     // the second K/V tensor of the joint attention has its own geometry
        bdims[2] = (uint64_t)Nkv2;
        bdims[3] = (uint64_t)NIkv2;
        bstr[1] = (uint64_t)ldkv2 * 2;
        bstr[2] = (uint64_t)Nkv2 * ldkv2 * 2;
      }
      int r = make_tmap_f16(&kvm.k[b], Kb[b] ? Kb[b] : K, 4, bdims, bstr, box, true);
      if (r) return r;
      r = make_tmap_f16(&kvm.v[b], Vb[b] ? Vb[b] : V, 4, bdims, bstr, box, true);
      if (r) return r;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (Nkv2 > 0) {   // joint attention (second K/V tensor with its own token count): the ragged-source instantiations
    if (d <= 64) return launch_attn<2, 128, 6, 1, 1>(tq, kvm, p, st);
    if (d <= 128) return launch_attn<2, 64, 0, 0, 1>(tq, kvm, p, st);
    return launch_attn<1, 64, 0, 0, 1>(tq, kvm, p, st);
  }
  if (d <= 64 && (int64_t)Nkv * nsrc <= 128 && variant == kDefaultVariant)   // one KV tile (cross-attention): the
    return launch_attn<2, 128, 0>(tq, kvm, p, st);                        // row-sum MMA is pure overhead
  if (variant == 18 && d <= 48) return launch_attn<2, 128, 6, 2>(tq, kvm, p, st);   // 16 + P through tensor memory
  if ((variant == 16 || (variant >= 13 && d > 48)) && d <= 64)
    return launch_attn<2, 128, 6, 1>(tq, kvm, p, st);   // variant 9 + packed FFMA2 score scaling
  if (d <= 48 && variant >= 13) {   // split rows: two softmax threads per query row (head dim 40)
    switch (variant) {
      case 13: return launch_attn_split<0>(tq, kvm, p, st);
      case 17: return launch_attn_split<6>(tq, kvm, p, st);   // + packed FFMA2 score scaling
      case 19: return launch_attn_split<6, 1>(tq, kvm, p, st);   // + convergent (whole-warp, elected-lane) MMA issue
      case 14: return launch_attn_split<4>(tq, kvm, p, st);   // + 1/8 of the exp2 as polynomials
      case 20: return launch_attn_split<7, 1>(tq, kvm, p, st);   // 19 + 1/4 of the exp2 as packed (FFMA2) polynomials
      case 21: return launch_attn_split<8, 1>(tq, kvm, p, st);   // 19 + 1/8
      case 22: return launch_attn_split<10, 1>(tq, kvm, p, st);  // 19 + 1/2
      default: return launch_attn_split<1>(tq, kvm, p, st);   // + 1/4
    }
  }
  if (d <= 64) {
    switch (variant) {
      case 0: return launch_attn<2, 128, 0>(tq, kvm, p, st);
      case 1: return launch_attn<2, 128, 1>(tq, kvm, p, st);
      // (2 - 6, 8, 11: retired experiments -- 64-key tiles for small head dims, 1/2, 3/8 and all-polynomial exp2; all slower)
      case 7: return launch_attn<2, 128, 1, 1>(tq, kvm, p, st);   // row sums on the tensor pipe + 1/4 polynomial
      case 9: return launch_attn<2, 128, 0, 1>(tq, kvm, p, st);   // row sums on the tensor pipe, all exp2 on MUFU
      case 10: return launch_attn<2, 128, 4, 1>(tq, kvm, p, st);  // row sums on the tensor pipe, 1/8 polynomial
      case 12: return launch_attn<2, 128, 4, 0>(tq, kvm, p, st);  // 1/8 polynomial
      default: return launch_attn<2, 128, 0, 1>(tq, kvm, p, st);  // (13+ with 48 < d <= 64: no split-row kernel)
    }
  }
  if (d <= 128) return launch_attn<2, 64, 0>(tq, kvm, p, st);
  return launch_attn<1, 64, 0>(tq, kvm, p, st);
}

extern "C" int univst_sc_attention_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv, int32_t NI,
                                       int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv, const int32_t* kv_src,
                                       int32_t nsrc, void* O, int32_t ldo, void* stream) {
  const void* Kb[3] = {K, nullptr, nullptr};
  const void* Vb[3] = {V, nullptr, nullptr};
  return sc_attention_impl(Q, ldq, Kb, Vb, ldkv, NI, NIkv, H, d, N, Nkv, kv_src, nsrc, O, ldo, 0, 0, stream);
}

extern "C" int univst_sc_attention_sharded_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv,
                                               int32_t NI, int32_t H, int32_t d, int32_t N, const int32_t* kv_src,
                                               int32_t nsrc, void* O, int32_t ldo, const void* K_prev, const void* V_prev,
                                               const void* K_first, const void* V_first, int32_t B, int32_t Fl,
                                               void* stream) {
  UV_REQUIRE(B > 0 && Fl > 0 && B * Fl == NI, "sc_attention_sharded: NI must be B * Fl");
  UV_REQUIRE((((uintptr_t)K_prev | (uintptr_t)V_prev | (uintptr_t)K_first | (uintptr_t)V_first) % 16) == 0,
             "sc_attention_sharded: 16-byte alignment");
  const void* Kb[3] = {K, K_prev, K_first};
  const void* Vb[3] = {V, V_prev, V_first};
  return sc_attention_impl(Q, ldq, Kb, Vb, ldkv, NI, NI, H, d, N, N, kv_src, nsrc, O, ldo, B, Fl, stream);
}

extern "C" int univst_joint_attention_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv, int32_t NI,
                                          int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv, const void* K2,
                                          const void* V2, int32_t ldkv2, int32_t NIkv2, int32_t Nkv2, const int32_t* kv_src,
                                          int32_t nsrc, void* O, int32_t ldo, void* stream) {
  UV_REQUIRE(K2 && V2 && NIkv2 > 0 && Nkv2 > 0 && ldkv2 % 8 == 0, "joint_attention: bad second K/V tensor");
  UV_REQUIRE((((uintptr_t)K2 | (uintptr_t)V2) % 16) == 0, "joint_attention: 16-byte alignment");
  const void* Kb[3] = {K, K2, nullptr};
  const void* Vb[3] = {V, V2, nullptr};
  return sc_attention_impl(Q, ldq, Kb, Vb, ldkv, NI, NIkv, H, d, N, Nkv, kv_src, nsrc, O, ldo, 0, 0, stream, ldkv2, NIkv2, Nkv2);
}

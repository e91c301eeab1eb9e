// Device-side cross-rank synchronisation over peer-mapped control blocks (see xrank.cu for the protocol).
#pragma once
#include "host_util.h"
#include "ptx.cuh"

namespace uv {

static constexpr int kXrankMaxRanks = 16;
static constexpr int kXrankSlotFloats = 1024;          // payload of the small exchange: NB * groups * 2 floats
static constexpr unsigned long long kXrankTimeoutNs = 10000000000ull;

// word offsets inside the control block
static constexpr int kXrFlags = 0;                       // [16] flags[src] = last epoch signalled by rank src (peers write)
static constexpr int kXrEpoch = 16;                      // synchronisations completed (local)
static constexpr int kXrDone = 17;                       // block-arrival counter of multi-block kernels (local)
static constexpr int kXrError = 18;                      // sticky: 1 + rank that was waited for when a wait timed out
static constexpr int kXrWaitNs = 20;                     // (64-bit, words 20-21) nanoseconds spent waiting for peers so far
static constexpr int kXrSyncs = 22;                      // synchronisations completed (diagnostic twin of the epoch)
static constexpr int kXrSlots = 32;                      // float slots[2][16][kXrankSlotFloats]
static constexpr size_t kXrankCtlWords = kXrSlots + 2ull * kXrankMaxRanks * kXrankSlotFloats;

struct XrankPeers {
  uint32_t* ctl[kXrankMaxRanks];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// one 16-byte store to a multicast address: the NVSwitch writes it into the same offset of every rank's buffer
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One full warp: advance the epoch, signal every peer, wait for every peer.  All data stores of the kernel must have been
// ordered before the call (xrank_kernel_tail below).  Returns with the peers' stores of this step visible to the warp.
__device__ __forceinline__ void xrank_sync_warp(const XrankPeers& P, int rank, int world) {
  uint32_t* mine = P.ctl[rank];
  const int lane = threadIdx.x & 31;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(mine + kXrEpoch) + 1;
  __syncwarp();
  if (lane == 0) *reinterpret_cast<volatile uint32_t*>(mine + kXrEpoch) = e;
  __threadfence_system();
  unsigned long long waited = 0;
  if (lane < world && lane != rank) {
    st_release_sys(P.ctl[lane] + kXrFlags + rank, e);
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(mine + kXrFlags + lane) - e) < 0) {
      if (global_timer_ns() - t0 > kXrankTimeoutNs) {
        *reinterpret_cast<volatile uint32_t*>(mine + kXrError) = 1u + (uint32_t)lane;
        break;
      }
    }
    waited = global_timer_ns() - t0;
  }
  // diagnostics: the longest wait of this synchronisation (= how late the slowest peer was) and a counter
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, waited, o);
    waited = other > waited ? other : waited;
  }
  if (lane == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(mine + kXrWaitNs), waited);
    atomicAdd(mine + kXrSyncs, 1u);
  }
  __syncwarp();
}

// Tail of a (multi-block) kernel whose threads stored into peer memory: the last block to arrive synchronises with the
// other ranks.  Every thread of every block calls it; returns true in the block that did the synchronisation (after it).
__device__ __forceinline__ bool xrank_kernel_tail(const XrankPeers& P, int rank, int world) {
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  uint32_t* mine = P.ctl[rank];
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(mine + kXrDone, 1u);
    s_last = (t == gridDim.x * gridDim.y - 1);
    if (s_last) *reinterpret_cast<volatile uint32_t*>(mine + kXrDone) = 0;
  }
  __syncthreads();
  if (!s_last) return false;
  if (threadIdx.x < 32) xrank_sync_warp(P, rank, world);
  __syncthreads();
  return true;
}

static inline int fill_peers(XrankPeers& P, void* const* ctl, int rank, int world, const char* what) {
  UV_REQUIRE(ctl && world >= 1 && world <= kXrankMaxRanks && rank >= 0 && rank < world, "%s: bad rank / world (up to 16 ranks)", what);
  for (int r = 0; r < kXrankMaxRanks; ++r) P.ctl[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    UV_REQUIRE(ctl[r] && ((uintptr_t)ctl[r] & 15) == 0, "%s: null / misaligned control block of rank %d", what, r);
    P.ctl[r] = (uint32_t*)ctl[r];
  }
  return UNIVST_OK;
}

}  // namespace uv

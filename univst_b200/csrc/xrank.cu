// Cross-rank synchronisation and small exchanges of the frame-sharded UNet, written against peer-mapped
// ("symmetric") memory over NVLink instead of NCCL launches:
//
//   * every rank owns one control block (univst_xrank_ctl_bytes(), zero-initialised) that all ranks have mapped;
//   * a synchronisation is "store my next epoch into flags[me] of every peer, spin until every flags[peer] of my own block
//     has reached that epoch".  The epoch counter lives in DEVICE memory and is advanced by the kernel itself, so a launch
//     has no per-call host argument: a CUDA graph that contains it can be replayed;
//   * every rank executes the same sequence of synchronisations, so one counter serves all call sites;
//   * the synchronisation is the TAIL of the kernel that produced the data (K/V halo push, GroupNorm partial sums,
//     frames <-> pixels exchange, noise-prediction gather): stores -> fence -> (last block) signal -> wait.  When the
//     kernel retires, the peers' data for this step has landed in local memory and is visible to the next kernel on
//     the stream -- no separate barrier launch, no collective library call;
//   * a wait that does not complete within kXrankTimeoutNs sets the sticky `error` word and returns, so that a rank
//     that died cannot hang the others' GPUs (the host checks the word; results are invalid once it is set).
#include "xrank.cuh"

namespace uv {

__global__ void xrank_barrier_kernel(XrankPeers P, int rank, int world) { xrank_sync_warp(P, rank, world); }

// ------------------------------------------------------------------------------------------------------------
// Block copies into peer memory + synchronisation: K/V halo of the frame-sharded attn1 (my last frame -> the next rank's
// "previous frame" bank; rank 0: the clip's first frame -> every rank's "first frame" bank) and the gather of the
// noise prediction (my frames -> their place in every rank's full-clip buffer).
struct PushDesc {
  const __half* src;
  __half* mc;                   // multicast address of the destination (NVSwitch replicates one store to every rank) or null
  __half* dst[kXrankMaxRanks];
  long long src_blk, dst_blk;   // rows between consecutive blocks
  int ld_src, ld_dst, nblk, rows, cols;
};
struct PushArgs {
  PushDesc d[2];
  int n;
};

__global__ void xrank_push_kernel(PushArgs a, XrankPeers P, int rank, int world) {
  constexpr int U = 4;   // 16-byte pieces in flight per thread: the loads of a batch are issued before its (posted) stores
  for (int k = 0; k < a.n; ++k) {
    const PushDesc& d = a.d[k];
    const int vpr = d.cols >> 3;
    const long long total = (long long)d.nblk * d.rows * vpr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
      uint4 val[U];
      long long off[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        off[u] = -1;
        if (i < total) {
          const int v = (int)(i % vpr);
          const long long rr = i / vpr;
          const int row = (int)(rr % d.rows), blk = (int)(rr / d.rows);
          val[u] = __ldg(reinterpret_cast<const uint4*>(d.src + (blk * d.src_blk + row) * d.ld_src + v * 8));
          off[u] = (blk * d.dst_blk + row) * d.ld_dst + v * 8;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (off[u] < 0) continue;
        if (d.mc) {
          multimem_st_v4(d.mc + off[u], val[u]);
        } else {
          for (int r = 0; r < world; ++r)
            if (d.dst[r]) *reinterpret_cast<uint4*>(d.dst[r] + off[u]) = val[u];
        }
      }
    }
  }
  xrank_kernel_tail(P, rank, world);
}

}  // namespace uv

using namespace uv;

extern "C" int64_t univst_xrank_ctl_bytes(void) { return (int64_t)(kXrankCtlWords * sizeof(uint32_t)); }
extern "C" int32_t univst_xrank_slot_floats(void) { return kXrankSlotFloats; }

extern "C" int univst_xrank_barrier(void* const* ctl, int32_t rank, int32_t world, void* stream) {
  XrankPeers P;
  int r = fill_peers(P, ctl, rank, world, "xrank_barrier");
  if (r) return r;
  xrank_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, rank, world);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_xrank_push_f16(const univst_push_t* pushes, int32_t npush, void* const* ctl, int32_t rank,
                                     int32_t world, void* stream) {
  XrankPeers P;
  int r = fill_peers(P, ctl, rank, world, "xrank_push");
  if (r) return r;
  UV_REQUIRE(npush >= 0 && npush <= 2 && (npush == 0 || pushes), "xrank_push: at most two block copies per launch");
  PushArgs a{};
  size_t total = 0;
  for (int k = 0; k < npush; ++k) {
    const univst_push_t& p = pushes[k];
    UV_REQUIRE(p.src && p.nblk > 0 && p.rows > 0 && p.cols > 0, "xrank_push: empty block copy");
    UV_REQUIRE(p.cols % 8 == 0 && p.ld_src % 8 == 0 && p.ld_dst % 8 == 0 && ((uintptr_t)p.src & 15) == 0,
               "xrank_push: columns / row strides must be multiples of 8 halves, 16-byte aligned source");
    PushDesc& d = a.d[a.n];
    bool any = false;
    for (int q = 0; q < kXrankMaxRanks; ++q) {
      d.dst[q] = q < world ? (__half*)p.dst[q] : nullptr;
      UV_REQUIRE(((uintptr_t)d.dst[q] & 15) == 0, "xrank_push: destinations must be 16-byte aligned");
      any |= d.dst[q] != nullptr;
    }
    if (!any && !p.mc_dst) continue;
    d.src = (const __half*)p.src;
    d.mc = (__half*)p.mc_dst;
    UV_REQUIRE(((uintptr_t)d.mc & 15) == 0, "xrank_push: the multicast destination must be 16-byte aligned");
    d.src_blk = p.src_blk_rows;
    d.dst_blk = p.dst_blk_rows;
    d.ld_src = p.ld_src;
    d.ld_dst = p.ld_dst;
    d.nblk = p.nblk;
    d.rows = p.rows;
    d.cols = p.cols;
    total += (size_t)p.nblk * p.rows * (p.cols / 8);
    ++a.n;
  }
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  xrank_push_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a, P, rank, world);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

namespace uv {
// norm.cu
int gn_stats_xrank(const void* X1, const void* X2, int C1, int C2, int NB, int rows, int groups, float* sums, void* workspace,
                   const XrankPeers& peers, int rank, int world, cudaStream_t st);
int gn_apply_launch(const void* X1, const void* X2, int C1, int C2, int NB, int rows, int groups, const float* sums,
                    int64_t stat_rows, const void* gamma, const void* beta, float eps, int silu, void* Y, cudaStream_t st);
int gn_check_shape(const void* X1, const void* X2, int32_t& C1, int32_t& C2, int32_t NB, int32_t rows, int32_t groups);
float* gn_sums_of(void* workspace, int NB, int groups);
}  // namespace uv

extern "C" int univst_groupnorm_xrank_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                                          int32_t groups, const void* gamma, const void* beta, float eps, int32_t silu,
                                          void* Y, void* workspace, void* const* ctl, int32_t rank, int32_t world,
                                          void* stream) {
  UV_REQUIRE(X1 && Y && gamma && beta && workspace, "groupnorm_xrank: null pointer");
  XrankPeers P;
  int r = fill_peers(P, ctl, rank, world, "groupnorm_xrank");
  if (r) return r;
  r = gn_check_shape(X1, X2, C1, C2, NB, rows, groups);
  if (r) return r;
  cudaStream_t st = (cudaStream_t)stream;
  float* sums = gn_sums_of(workspace, NB, groups);
  // statistics -> (last block) fold + store into every rank's slot + synchronise + add the slots in rank order
  r = gn_stats_xrank(X1, X2, C1, C2, NB, rows, groups, sums, workspace, P, rank, world, st);
  if (r) return r;
  return gn_apply_launch(X1, X2, C1, C2, NB, rows, groups, sums, (int64_t)rows * world, gamma, beta, eps, silu, Y, st);
}

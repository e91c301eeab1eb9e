// Sliding-window flow-warp smoothing of decoded frames (reference: src/cal_optica_flow.py:20-46 and the window loop
// of backbones/video_diffusion_sd/pipelines/stable_diffusion.py:725-751), one coalesced bilinear-gather kernel per
// key frame.  Integer / byte work, bit-exact with the reference's NumPy + cv2.remap arithmetic:
//   occlusion  : || ((p + fwd(p)) + bwd(p)) - p ||_2 > 1.5 in fp32 with the reference's operation order (no FMA)
//   warp       : cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) on uint8: coordinates rounded to 1/32 px (half to even),
//                exact 15-bit bilinear weights, (sum + 2^14) >> 15
//   blend      : occluded pixels keep the key frame; the key frame becomes trunc(mean of itself and its <= 4 warped
//                neighbours) IN PLACE (later keys read already-smoothed neighbours, as the reference does)
// The optical flows are inputs (RAFT is a third-party network whose weights are unavailable offline).
#include "host_util.h"

namespace uv {

struct WarpNeighbours {
  int n;
  int idx[4];
  const float2* fwd[4];  // flow key -> neighbour, [H][W] (x, y)
  const float2* bwd[4];  // flow neighbour -> key
};

__device__ __forceinline__ int tap_u8(const uint8_t* img, int H, int W, int y, int x, int c) {
  return (y >= 0 && y < H && x >= 0 && x < W) ? (int)img[((size_t)y * W + x) * 3 + c] : 0;
}

__global__ void flow_warp_key_kernel(uint8_t* __restrict__ frames, int H, int W, int key, WarpNeighbours nb,
                                     float threshold) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int y = p / W, x = p - y * W;
  const size_t fs = (size_t)H * W * 3;
  uint8_t* kf = frames + (size_t)key * fs + (size_t)p * 3;
  const int k0 = kf[0], k1 = kf[1], k2 = kf[2];
  int a0 = k0, a1 = k1, a2 = k2;
  const float xf = (float)x, yf = (float)y;
  for (int i = 0; i < nb.n; ++i) {
    const float2 f = __ldg(nb.fwd[i] + p), b = __ldg(nb.bwd[i] + p);
    const float mx = __fadd_rn(xf, f.x), my = __fadd_rn(yf, f.y);
    const float ex = __fsub_rn(__fadd_rn(mx, b.x), xf), ey = __fsub_rn(__fadd_rn(my, b.y), yf);
    const float err = __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
    if (err > threshold) {
      a0 += k0, a1 += k1, a2 += k2;
      continue;
    }
    const int sx = __float2int_rn(__fmul_rn(mx, 32.0f)), sy = __float2int_rn(__fmul_rn(my, 32.0f));
    int ix = sx >> 5, iy = sy >> 5;
    const int fx = sx & 31, fy = sy & 31;
    ix = max(-32768, min(32767, ix));
    iy = max(-32768, min(32767, iy));
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    const uint8_t* img = frames + (size_t)nb.idx[i] * fs;
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s = tap_u8(img, H, W, iy, ix, c) * w00 + tap_u8(img, H, W, iy, ix + 1, c) * w01 +
                    tap_u8(img, H, W, iy + 1, ix, c) * w10 + tap_u8(img, H, W, iy + 1, ix + 1, c) * w11;
      v[c] = min(255, max(0, (s + (1 << 14)) >> 15));
    }
    a0 += v[0], a1 += v[1], a2 += v[2];
  }
  const float wgt = (float)(nb.n + 1);
  kf[0] = (uint8_t)(int)__fdiv_rn((float)a0, wgt);
  kf[1] = (uint8_t)(int)__fdiv_rn((float)a1, wgt);
  kf[2] = (uint8_t)(int)__fdiv_rn((float)a2, wgt);
}

__global__ void mask_select_kernel(const uint8_t* __restrict__ keep_mask, const uint8_t* __restrict__ orig,
                                   const uint8_t* __restrict__ est, size_t npix, uint8_t* __restrict__ out) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    const uint8_t* s = keep_mask[p] ? orig : est;
    out[p * 3] = s[p * 3];
    out[p * 3 + 1] = s[p * 3 + 1];
    out[p * 3 + 2] = s[p * 3 + 2];
  }
}

}  // namespace uv

using namespace uv;

extern "C" int univst_flow_warp_key_u8(void* frames, int32_t F, int32_t H, int32_t W, int32_t key, int32_t n_neighbours,
                                       const int32_t* neighbour_idx, const void* const* fwd_flows,
                                       const void* const* bwd_flows, float threshold, void* stream) {
  UV_REQUIRE(frames && F > 0 && H > 0 && W > 0 && key >= 0 && key < F, "flow_warp_key: bad frame arguments");
  UV_REQUIRE(n_neighbours >= 0 && n_neighbours <= 4, "flow_warp_key: at most 4 neighbours (window radius 2)");
  WarpNeighbours nb{};
  nb.n = n_neighbours;
  for (int i = 0; i < n_neighbours; ++i) {
    UV_REQUIRE(neighbour_idx[i] >= 0 && neighbour_idx[i] < F && neighbour_idx[i] != key && fwd_flows[i] && bwd_flows[i],
               "flow_warp_key: bad neighbour %d", i);
    nb.idx[i] = neighbour_idx[i];
    nb.fwd[i] = (const float2*)fwd_flows[i];
    nb.bwd[i] = (const float2*)bwd_flows[i];
  }
  const int threads = 256;
  flow_warp_key_kernel<<<(H * W + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>((uint8_t*)frames, H, W, key,
                                                                                             nb, threshold);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

extern "C" int univst_mask_select_u8(const uint8_t* keep_mask, const void* orig, const void* est, int64_t npix, void* out,
                                     void* stream) {
  UV_REQUIRE(keep_mask && orig && est && out && npix > 0, "mask_select: bad arguments");
  size_t blocks = ((size_t)npix + 255) / 256;
  if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
  mask_select_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(keep_mask, (const uint8_t*)orig, (const uint8_t*)est,
                                                                   (size_t)npix, (uint8_t*)out);
  UV_CHECK_CUDA(cudaGetLastError());
  return UNIVST_OK;
}

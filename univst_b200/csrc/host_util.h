// Host-side helpers shared by the launchers: error codes, TMA tensor-map construction through the
// driver entry point (no link-time dependency on libcuda), cached device properties.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/univst_b200.h"

namespace uv {

#define UV_CHECK_CUDA(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      uv::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return UNIVST_ERR_CUDA;                                                                    \
    }                                                                                            \
  } while (0)

#define UV_REQUIRE(cond, ...)                \
  do {                                       \
    if (!(cond)) {                           \
      uv::set_last_error(__VA_ARGS__);       \
      return UNIVST_ERR_INVALID;             \
    }                                        \
  } while (0)

void set_last_error(const char* fmt, ...);
const char* last_error();
int num_sms();
int max_smem_optin();
int cc_major();

// Encode a tiled fp16 tensor map. dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of
// dims[i+1] (rank-1 entries). swizzle128: inner box must be 64 halves (128 B).
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128);

}  // namespace uv

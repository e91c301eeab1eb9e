// Host-side helpers: thread-local error string, device-property cache, TMA tensor-map encoding
// through the driver entry point (so the library has no link-time dependency on libcuda).
#include "host_util.h"

#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace uv {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static int g_num_sms = 0, g_smem_optin = 0, g_cc_major = 0;
static std::once_flag g_prop_once;
static void load_props() {
  std::call_once(g_prop_once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&g_cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  });
}
int num_sms() {
  load_props();
  return g_num_sms > 0 ? g_num_sms : 148;
}
int max_smem_optin() {
  load_props();
  return g_smem_optin;
}
int cc_major() {
  load_props();
  return g_cc_major;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128) {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  });
  if (!g_encode) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver");
    return UNIVST_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u stride0 %llu)",
                   (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                   (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0],
                   rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0,
                   (unsigned long long)(rank > 1 ? gstr[0] : 0));
    return UNIVST_ERR_CUDA;
  }
  return UNIVST_OK;
}

}  // namespace uv

extern "C" {
int univst_abi_version(void) { return UNIVST_ABI_VERSION; }
const char* univst_last_error(void) { return uv::last_error(); }
int univst_device_check(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    uv::set_last_error("no CUDA device visible");
    return UNIVST_ERR_NO_DEVICE;
  }
  if (uv::cc_major() != 10) {
    uv::set_last_error("device is compute capability %d.x; this library contains sm_100a code only", uv::cc_major());
    return UNIVST_ERR_NO_DEVICE;
  }
  return UNIVST_OK;
}
}

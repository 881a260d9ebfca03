#include "host_common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

// Default of MA_PDL (programmatic dependent launch of the chained kernels); see DESIGN.md section 5 for the A/B numbers.
#ifndef MA_PDL_DEFAULT
#define MA_PDL_DEFAULT 1
#endif

namespace ma {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap(out, base, MA_BF16, rank, dims, strides_bytes, box, 128);
}

int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return MA_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("tensor map base %p is not 16-byte aligned", base);
    return MA_ERR_INVALID;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16 != 0) {
      set_last_error("tensor map stride[%d]=%llu bytes is not a multiple of 16", i, (unsigned long long)strides_bytes[i]);
      return MA_ERR_INVALID;
    }
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                               : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dtype == MA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                  const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u stride0 %llu)", (int)r,
                   rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                   rank > 1 ? box[1] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return MA_ERR_CUDA;
  }
  return MA_OK;
}

static int g_pdl_mode = -1;

bool pdl_enabled() {
  int& mode = g_pdl_mode;
  if (mode < 0) {
    const char* e = getenv("MA_PDL");
    mode = e ? (e[0] != '0') : MA_PDL_DEFAULT;
  }
  return mode != 0;
}

static int g_stream_k_mode = -1;

bool stream_k_enabled() {
  int& mode = g_stream_k_mode;
  if (mode < 0) {
    const char* e = getenv("MA_GEMM_STREAMK");
    mode = e ? (e[0] != '0') : 1;
  }
  return mode != 0;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MA_MAX_DEVICES) return 0;
  return dev;
}

int device_sm_count() {
  static int sms[MA_MAX_DEVICES] = {};  // per device ordinal: a process may drive several GPUs
  const int dev = current_device();
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

}  // namespace ma

extern "C" const char* ma_last_error(void) { return ma::g_last_error; }

extern "C" int ma_abi_version(void) { return MA_ABI_VERSION; }

extern "C" int ma_set_pdl(int enabled) {
  const int before = ma::pdl_enabled() ? 1 : 0;
  if (enabled >= 0) ma::g_pdl_mode = enabled ? 1 : 0;
  return before;
}

extern "C" int ma_set_stream_k(int enabled) {
  const int before = ma::stream_k_enabled() ? 1 : 0;
  if (enabled >= 0) ma::g_stream_k_mode = enabled ? 1 : 0;
  return before;
}

extern "C" int ma_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MA_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MA_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return MA_OK;
}

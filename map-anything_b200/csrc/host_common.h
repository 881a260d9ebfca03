// Host-side helpers shared by every launcher of the C-ABI library:
// status codes + last-error string, TMA tensor-map encoding through the driver entry point
// (resolved at run time with cudaGetDriverEntryPoint, so the .so has no link-time dependency
// on libcuda and can be dlopen'ed on a GPU-less box for the symbol-export test).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mapanything_b200.h"

namespace ma {

void set_last_error(const char* fmt, ...);

#define MA_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ma::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MA_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define MA_REQUIRE(cond, ...)                  \
  do {                                         \
    if (!(cond)) {                             \
      ma::set_last_error(__VA_ARGS__);         \
      return MA_ERR_INVALID;                   \
    }                                          \
  } while (0)

// Encodes a bf16 tiled tensor map with 128-byte swizzle. dims/strides innermost first;
// strides_bytes[i] is the stride of dimension i+1 (dimension 0 is contiguous).
// Returns MA_OK or an error code (message in ma_last_error()).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

// General form: dtype = MA_BF16 / MA_F32, swizzle_bytes in {0, 32, 64, 128}.
int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes);

// Per-device state (SM count, the large-dynamic-shared-memory opt-in of each kernel) is cached per device ordinal: the
// same process may run the model on several GPUs (model.to("cuda:1")).
constexpr int MA_MAX_DEVICES = 64;
int current_device();
int device_sm_count();

// Programmatic dependent launch of the chained hot-path kernels (GEMM, attention, LayerNorm): MA_PDL=0/1, see ptx.cuh.
bool pdl_enabled();
bool stream_k_enabled();

// cudaLaunchKernelEx wrapper; with pdl = true the kernel may start before the previous kernel of the stream has
// completed, and MUST execute pdl_wait() before its first global memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ma

// diffute_b200 — internal host-side declarations shared by the .cu files and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/diffute_b200.h"

namespace dfu {

// last error text for dfu_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define DFU_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      dfu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFU_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define DFU_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      dfu::set_error(__VA_ARGS__);    \
      return DFU_ERR_INVALID;         \
    }                                 \
  } while (0)

// Tiled fp16 tensor map with 128-byte swizzle. dims/strides innermost first; strides in BYTES for dims 1..rank-1.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

int num_sms();

// 1 unless DFU_PDL=0: launch with the programmatic-stream-serialization attribute (kernels call pdl_wait()).
int pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace dfu

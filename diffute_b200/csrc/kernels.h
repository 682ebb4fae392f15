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

int num_sms();  // of the CURRENT device (cached per device)

// true the first time it is called for (slot, current device): guards per-device one-time setup such as
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), which a second device in the same process needs again
enum : int { ONCE_GEMM_ATTR = 0, ONCE_ATTN_ATTR, ONCE_GEMV_ATTR, ONCE_CONV_OUT_ATTR, ONCE_GEMM2_ATTR, ONCE_GN_SMEM_ATTR, kOnceSlots = 8 };
bool first_use_on_device(int slot);

// 1 unless DFU_PDL=0: launch with the programmatic-stream-serialization attribute (kernels call pdl_wait()).
int pdl_enabled();

// cluster_x > 1 launches thread-block clusters of that many consecutive CTAs along x
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.numAttrs = 1;
  if (cluster_x > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  return launch_kc(kernel, grid, block, smem, stream, 1, static_cast<Args&&>(args)...);
}

}  // namespace dfu

// diffute_b200 — host utilities: error text, tensor-map encoding through the driver entry point.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "kernels.h"

namespace dfu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  // libcuda is reached through the runtime: no link-time dependency on the driver library.
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return DFU_ERR_DRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tensor map base %p not 16-byte aligned", base);
    return DFU_ERR_INVALID;
  }
  // Encoded maps are cached by (base, rank, dims, strides, box): the engine's buffers are static, so after the first
  // step every launch finds its 128-byte descriptors here instead of calling into the driver 2-4 times.
  static std::mutex mu;
  static std::unordered_map<std::string, CUtensorMap> cache;
  std::string key(reinterpret_cast<const char*>(&base), sizeof(base));
  key.append(reinterpret_cast<const char*>(&rank), sizeof(rank));
  key.append(reinterpret_cast<const char*>(gdim), sizeof(cuuint64_t) * rank);
  key.append(reinterpret_cast<const char*>(gstr), sizeof(cuuint64_t) * (rank - 1));
  key.append(reinterpret_cast<const char*>(gbox), sizeof(cuuint32_t) * rank);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return DFU_OK;
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                   gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r,
              rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
              (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), gbox[0],
              rank > 1 ? gbox[1] : 0, rank > 2 ? gbox[2] : 0, rank > 3 ? gbox[3] : 0);
    return DFU_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 16384) cache.clear();  // bounded: shapes / buffers of a long-lived process may change
    cache.emplace(std::move(key), *out);
  }
  return DFU_OK;
}

int pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFU_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

// Per-DEVICE caches: one process may drive several GPUs (the header promises "re-entrant per stream"), and both the SM
// count and cudaFuncSetAttribute are properties of the current device, not of the process.
constexpr int kMaxDevices = 64;

int num_sms() {
  static int n[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  if (n[dev] > 0) return n[dev];
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  n[dev] = v;
  return v;
}

bool first_use_on_device(int slot) {
  static bool done[kOnceSlots][kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || slot < 0 || slot >= kOnceSlots) return true;
  if (done[slot][dev]) return false;
  done[slot][dev] = true;
  return true;
}

}  // namespace dfu

extern "C" {
int dfu_version(void) { return 100; }
const char* dfu_last_error(void) { return dfu::get_error(); }
int dfu_num_sms(void) { return dfu::num_sms(); }
}

// diffute_b200 — load-time weight packing: ONE launch repacks every parameter of a network from the diffusers
// state-dict layout (fp32, [Cout, Cin, kh, kw] / [N, K]) into the layouts the kernels read (K-major fp16 hi/lo
// planes with tap-major K, GEGLU 16/16 row interleave, q|k|v and time_emb_proj stacks, fp32 small-conv layouts,
// folded shortcut biases).  Not on the sampling path; it exists so that loading a checkpoint is a handful of kernel
// launches instead of several per tensor.
#include "common.cuh"
#include "kernels.h"

namespace dfu {

// element e of job j  ->  (row, k) with k = tap * cin + ci in the DESTINATION order; source is torch's
// [rows][cin][taps] (OIHW flattened), i.e. k_src = ci * taps + tap.
__global__ void __launch_bounds__(256) pack_weights_kernel(const DfuPackJob* __restrict__ jobs,
                                                           const long long* __restrict__ prefix, int njobs,
                                                           long long total) {
  for (long long u = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; u < total;
       u += static_cast<long long>(gridDim.x) * blockDim.x) {
    int lo = 0, hi = njobs - 1;  // last job whose first element is <= u
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= u) lo = mid; else hi = mid - 1;
    }
    const DfuPackJob J = jobs[lo];
    const long long e = u - prefix[lo];
    const int K = J.taps * J.cin;
    const int row = static_cast<int>(e / K);
    const int k = static_cast<int>(e - static_cast<long long>(row) * K);
    const int tap = k / J.cin, ci = k - tap * J.cin;
    const size_t si = (static_cast<size_t>(row) * J.cin + ci) * J.taps + tap;
    float v = J.src[si];
    if (J.src2) v += J.src2[si];
    int drow = row;
    if (J.geglu) {  // rows [a_0..a_{n-1}, g_0..g_{n-1}] -> blocks of 32: 16 value rows, then the matching 16 gate rows
      const int n = J.rows >> 1;
      const int r = row < n ? row : row - n;
      drow = (r >> 4) * 32 + (r & 15) + (row < n ? 0 : 16);
    }
    if (J.mode == 1) {  // transposed fp32 [cin*taps (source k order)][rows]: conv_small_in weights
      static_cast<float*>(J.dst)[static_cast<size_t>(ci * J.taps + tap) * J.dst_ld + J.dst_row0 + drow] = v;
      continue;
    }
    const size_t di = static_cast<size_t>(J.dst_row0 + drow) * J.dst_ld + k;
    if (J.planes == 0) {
      static_cast<float*>(J.dst)[di] = v;
    } else {
      __half* d = static_cast<__half*>(J.dst);
      const __half h = __float2half_rn(v);
      d[di] = h;
      if (J.planes > 1) d[di + J.plane_stride] = __float2half_rn(v - __half2float(h));
    }
  }
}

}  // namespace dfu

extern "C" int dfu_pack_weights(const DfuPackJob* jobs_dev, const int64_t* prefix_dev, int njobs, int64_t total,
                                void* stream) {
  using namespace dfu;
  DFU_REQUIRE(jobs_dev && prefix_dev && njobs > 0 && total > 0, "pack_weights: empty job list");
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 32;
  if (blocks > cap) blocks = cap;
  pack_weights_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      jobs_dev, reinterpret_cast<const long long*>(prefix_dev), njobs, total);
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

// diffute_b200 — definitions shared by the two tcgen05 contraction kernels (gemm.cu: one tile per CTA, split-K
// clusters; gemm2.cu: persistent CTA pairs, cta_group::2, double-buffered TMEM): operand-group tables, epilogue
// parameters and the per-quad fused epilogue helpers.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace dfu {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 12;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;  // 16 KiB smem slot for A (box may fill fewer rows)

struct GroupDev {
  int a_mode, ntaps, nchunks, a_plane, b_plane, kb_per_pass;
  int b_static;  // B of this group is a constant (weights): may be fetched before the producer kernel completes
  int8_t dn[9], dy[9], dx[9];
};

// batched matrix products (DfuGemm.batch > 1): 128-row m-tiles never straddle two batches
struct BatchDev {
  int tiles_per_batch;  // 0: not batched
  int rows_per_batch;   // m / batch (output rows of one product)
  int a_batch_rows, b_batch_rows;
};
// m-tile index -> first output row, first A row, B row offset, end of the valid output rows
__device__ __forceinline__ void batch_coords(const BatchDev& bt, int tm, int M, int& m0, int& a_row0, int& b_off, int& m_end) {
  if (bt.tiles_per_batch > 0) {
    const int b = tm / bt.tiles_per_batch, ti = tm - b * bt.tiles_per_batch;
    m0 = b * bt.rows_per_batch + ti * 128;
    a_row0 = b * bt.a_batch_rows + ti * 128;
    b_off = b * bt.b_batch_rows;
    m_end = (b + 1) * bt.rows_per_batch;
  } else {
    m0 = a_row0 = tm * 128;
    b_off = 0;
    m_end = M;
  }
}

struct EpiParams {
  int M, N;
  int epi;
  int act;  // 1: exact-erf GELU on the finished value
  float alpha;
  const float* bias;
  const float* rowvec;
  int rowvec_ld, rows_per_sample;
  const float* residual;
  int ldr;
  float* out_f32;
  int ldo;
  __half* out_f16;
  int ldh;
  int out_planes;
  long long out_plane_stride;
};

struct GemmKernelParams {
  int block_n, tiles_m, tiles_n, splits, stages, total_kb;
  int ngroups, npass;
  GroupDev g[2];
  int conv, B, H, W, bw, bh, bn, tiles_x, tiles_y;
  uint32_t a_tx_bytes[2];  // bytes one A box delivers (per group)
  uint32_t b_tx_bytes;
  uint32_t tmem_cols;
  float* ws;
  BatchDev bt;
  int two_prod;           // 1: weight tiles and activation tiles are issued by two different warps
  int cluster;         // > 1: the `splits` K-slices of a tile form a thread-block cluster and reduce through DSMEM
  unsigned int* sync;  // grid-barrier words (zero between launches); non-null => fused split-K second stage
  EpiParams e;
};

// ---------------------------------------------------------------------------------------------
// shared epilogue, one 4-column quad of output row m at a time so that consecutive lanes touch consecutive 16-byte
// pieces of a row (coalesced residual loads and output stores).
//   v = alpha*acc + bias[n] + rowvec[sample(m), n] + residual[m, n]  ->  fp32 | fp16 hi/lo planes
//   GEGLU: out[m, n/2] = (a + bias_a) * gelu_erf(g + bias_g), a/g = value / gate quads 16 columns apart
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_f16x4(__half* dst, float4 v, bool lo_plane, long long plane_stride) {
  __align__(8) __half h[4];
  h[0] = __float2half_rn(v.x); h[1] = __float2half_rn(v.y); h[2] = __float2half_rn(v.z); h[3] = __float2half_rn(v.w);
  *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(h);
  if (lo_plane) {
    __align__(8) __half l[4];
    l[0] = __float2half_rn(v.x - __half2float(h[0]));
    l[1] = __float2half_rn(v.y - __half2float(h[1]));
    l[2] = __float2half_rn(v.z - __half2float(h[2]));
    l[3] = __float2half_rn(v.w - __half2float(h[3]));
    *reinterpret_cast<uint2*>(dst + plane_stride) = *reinterpret_cast<const uint2*>(l);
  }
}

// (the epilogue is instruction-issue bound — ~2000 cycles per 32-column chunk round at 16 epilogue warps per SM, measured
// with scripts/trace_step.py — so everything per-row is hoisted by the callers and alpha == 1 costs nothing)
__device__ __forceinline__ float4 epi_affine(const EpiParams& e, int m, int n, float4 v, const float* sbias = nullptr,
                                             int n_tile0 = 0) {
  if (e.alpha != 1.0f) {
    v.x *= e.alpha; v.y *= e.alpha; v.z *= e.alpha; v.w *= e.alpha;
  }
  if (e.bias) {
    const float4 t = sbias ? *reinterpret_cast<const float4*>(sbias + (n - n_tile0))
                           : __ldg(reinterpret_cast<const float4*>(e.bias + n));
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  if (e.rowvec) {
    // one sample per 128-row tile is the common case (rows_per_sample >= 128): the division is then a compare
    const int smp = (m < e.rows_per_sample) ? 0 : m / e.rows_per_sample;
    const float4 t = __ldg(reinterpret_cast<const float4*>(e.rowvec + static_cast<size_t>(smp) * e.rowvec_ld + n));
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  return v;
}

// residual already fetched by the caller (t = 0 when there is none)
__device__ __forceinline__ void epi_quad_res(const EpiParams& e, int m, int n, float4 v, float4 t,
                                             const float* sbias = nullptr, int n_tile0 = 0) {
  v = epi_affine(e, m, n, v, sbias, n_tile0);
  v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  if (e.act) {
    v.x = gelu_erf_f(v.x); v.y = gelu_erf_f(v.y); v.z = gelu_erf_f(v.z); v.w = gelu_erf_f(v.w);
  }
  if (e.epi == DFU_EPI_F32) {
    *reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(m) * e.ldo + n) = v;
  } else {
    store_f16x4(e.out_f16 + static_cast<size_t>(m) * e.ldh + n, v, e.out_planes > 1, e.out_plane_stride);
  }
}

__device__ __forceinline__ void epi_quad(const EpiParams& e, int m, int n, float4 v) {
  v = epi_affine(e, m, n, v);
  if (e.residual) {
    const float4 t = *reinterpret_cast<const float4*>(e.residual + static_cast<size_t>(m) * e.ldr + n);
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  if (e.act) {
    v.x = gelu_erf_f(v.x); v.y = gelu_erf_f(v.y); v.z = gelu_erf_f(v.z); v.w = gelu_erf_f(v.w);
  }
  if (e.epi == DFU_EPI_F32) {
    *reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(m) * e.ldo + n) = v;
  } else {
    store_f16x4(e.out_f16 + static_cast<size_t>(m) * e.ldh + n, v, e.out_planes > 1, e.out_plane_stride);
  }
}

// n_a = packed column of the value quad (the gate quad sits at n_a + 16); output column = block*16 + offset
__device__ __forceinline__ void epi_geglu_quad(const EpiParams& e, int m, int n_a, float4 a, float4 g,
                                               const float* sbias = nullptr, int n_tile0 = 0) {
  a = epi_affine(e, m, n_a, a, sbias, n_tile0);
  g = epi_affine(e, m, n_a + 16, g, sbias, n_tile0);
  float4 o;
  o.x = a.x * gelu_erf_f(g.x); o.y = a.y * gelu_erf_f(g.y); o.z = a.z * gelu_erf_f(g.z); o.w = a.w * gelu_erf_f(g.w);
  const int n_out = (n_a >> 5) * 16 + (n_a & 15);
  store_f16x4(e.out_f16 + static_cast<size_t>(m) * e.ldh + n_out, o, e.out_planes > 1, e.out_plane_stride);
}

constexpr int kStageLd = 36;                          // floats per staged row (16-byte aligned, conflict-free)
constexpr int kStageFloats = 32 * kStageLd;           // per epilogue warp

// host side (gemm.cu): geometry of the 128-row conv m-tiles, tensor-map encoding of one operand group
struct Plan {
  int block_n, splits, stages, tiles_m, tiles_n, total_kb;
  int bw, bh, bn, tiles_x, tiles_y;
  size_t ws_bytes;
  size_t smem_bytes;
  int pair;  // 1: run on gemm2 (persistent CTA pairs); block_n is then the pair's N tile and splits == 1
};
int plan_gemm(const DfuGemm* d, Plan* pl);
int encode_group(const DfuGemm* d, const DfuGemmOperand& o, const Plan& pl, int b_box_rows, CUtensorMap* mA,
                 CUtensorMap* mB, uint32_t* a_tx);
void fill_epi_params(const DfuGemm* d, EpiParams& e);
void fill_batch_dev(const DfuGemm* d, BatchDev& bt);
void fill_group_dev(const DfuGemmOperand& o, GroupDev& G);
int run_gemm2(const DfuGemm* d, const Plan& pl, cudaStream_t stream);  // gemm2.cu

}  // namespace dfu

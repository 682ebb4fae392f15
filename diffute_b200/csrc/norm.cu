// diffute_b200 — memory-bound normalisation / cast kernels (GroupNorm+SiLU, LayerNorm, fp32 -> fp16 operand casts).
//
// All activations are NHWC fp32 (the residual stream); these kernels read them once with 128-bit loads and write
// the 16-bit (hi / lo) operand planes the tcgen05 contraction core consumes, so the fp32 -> fp16 cast, the
// activation, the skip concat, the nearest-2x upsample and the stride-2 space-to-depth split never cost a pass
// of their own.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace dfu {

// =============================================================================================
// GroupNorm statistics: per (sample, pixel-chunk, group) partial sum / sum of squares
//   grid (chunks, B), block (C/4, TY).  Thread (tx, ty) owns channel quad tx for pixels ty, ty+TY, ...
// =============================================================================================
constexpr int kGnPixPerThread = 4;

struct GnSrc {
  const float* src0;
  const float* src1;  // optional second tensor of a channel concat ([h, skip] in the up blocks)
  int C0, C1;         // channels of each source
  int HW;
};

__device__ __forceinline__ float4 gn_load(const GnSrc& s, int b, int pix, int cq) {
  const int c = cq * 4;
  if (c < s.C0) return *reinterpret_cast<const float4*>(s.src0 + (static_cast<size_t>(b) * s.HW + pix) * s.C0 + c);
  return *reinterpret_cast<const float4*>(s.src1 + (static_cast<size_t>(b) * s.HW + pix) * s.C1 + (c - s.C0));
}

__global__ void gn_stats_kernel(GnSrc s, int groups, int pix_per_cta, float2* __restrict__ partial) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_GN_STATS);
  pdl_wait();
  DFU_TR_MARK(6);
  extern __shared__ float sm[];  // [TY][2][C] per-row-of-threads channel sums, reduced in a fixed order (deterministic)
  const int C = s.C0 + s.C1;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, s.HW);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float sx[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
  // <= kGnPixPerThread pixels per thread, all loads issued before the first use (memory-level parallelism)
  float4 v[kGnPixPerThread];
#pragma unroll
  for (int j = 0; j < kGnPixPerThread; ++j) {
    const int p = p0 + threadIdx.y + j * blockDim.y;
    v[j] = (p < p1) ? gn_load(s, b, p, threadIdx.x) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < kGnPixPerThread; ++j) {
    sx[0] += v[j].x; sq[0] += v[j].x * v[j].x;
    sx[1] += v[j].y; sq[1] += v[j].y * v[j].y;
    sx[2] += v[j].z; sq[2] += v[j].z * v[j].z;
    sx[3] += v[j].w; sq[3] += v[j].w * v[j].w;
  }
  const int c = threadIdx.x * 4;
  float* mine = sm + static_cast<size_t>(threadIdx.y) * 2 * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mine[c + i] = sx[i];
    mine[C + c + i] = sq[i];
  }
  __syncthreads();
  if (tid < groups) {
    float a = 0.f, q = 0.f;
    for (int ty = 0; ty < static_cast<int>(blockDim.y); ++ty) {
      const float* row = sm + static_cast<size_t>(ty) * 2 * C;
      for (int i = 0; i < cpg; ++i) {
        a += row[tid * cpg + i];
        q += row[C + tid * cpg + i];
      }
    }
    __stcg(&partial[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * groups + tid], make_float2(a, q));
    __threadfence();
  }
  DFU_TR_END();
}

// per-(sample, group) mean / rstd from the chunk partials (large maps: thousands of chunks).  One block per
// (group, sample); each thread sums a strided subset with four loads in flight, then a fixed-order tree (deterministic).
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float2* __restrict__ partial, int nchunks, int groups, double count, float eps,
                   float2* __restrict__ stats) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_GN_FINALIZE);
  pdl_wait();
  DFU_TR_MARK(6);
  __shared__ double s_sum[256], s_sq[256];
  const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float2* pp = partial + static_cast<size_t>(b) * nchunks * groups + g;
  double sum = 0.0, sq = 0.0;
  int k = tid;
  for (; k + 768 < nchunks; k += 1024) {
    const float2 t0 = pp[static_cast<size_t>(k) * groups];
    const float2 t1 = pp[static_cast<size_t>(k + 256) * groups];
    const float2 t2 = pp[static_cast<size_t>(k + 512) * groups];
    const float2 t3 = pp[static_cast<size_t>(k + 768) * groups];
    sum = (((sum + t0.x) + t1.x) + t2.x) + t3.x;
    sq = (((sq + t0.y) + t1.y) + t2.y) + t3.y;
  }
  for (; k < nchunks; k += 256) {
    const float2 t = pp[static_cast<size_t>(k) * groups];
    sum += t.x;
    sq += t.y;
  }
  s_sum[tid] = sum;
  s_sq[tid] = sq;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      s_sum[tid] += s_sum[tid + o];
      s_sq[tid] += s_sq[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double mean = s_sum[0] / count;
    double var = s_sq[0] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[b * groups + g] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + eps)));
  }
}

// =============================================================================================
// GroupNorm apply (+SiLU) -> fp16 operand planes (and optionally an fp32 copy and a raw-cast operand)
// =============================================================================================
struct GnApply {
  GnSrc s;
  int groups;
  int pix_per_cta;
  const float2* stats;    // [B][groups] (mean, rstd) from gn_finalize_kernel, or null:
  const float2* partial;  // ... then every CTA combines the `nchunks` chunk partials itself (few chunks: cheaper
  int nchunks;            //     than a third launch)
  double count;
  const float* gamma;
  const float* beta;
  float eps;
  int silu;
  __half* out16;          // [planes][B][HW][C] or null
  int planes;
  long long plane_stride;
  float* out32;           // optional fp32 [B][HW][C] (feeds the tiny fp32 output convs)
  __half* raw16;          // optional raw (un-normalised) cast of the same input, same layout as out16
};

__device__ __forceinline__ void store_split4(__half* dst, long long plane_stride, int planes, float4 v) {
  __align__(8) __half h[4];
  h[0] = __float2half_rn(v.x); h[1] = __float2half_rn(v.y); h[2] = __float2half_rn(v.z); h[3] = __float2half_rn(v.w);
  *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(h);
  if (planes > 1) {
    __align__(8) __half l[4];
    l[0] = __float2half_rn(v.x - __half2float(h[0]));
    l[1] = __float2half_rn(v.y - __half2float(h[1]));
    l[2] = __float2half_rn(v.z - __half2float(h[2]));
    l[3] = __float2half_rn(v.w - __half2float(h[3]));
    *reinterpret_cast<uint2*>(dst + plane_stride) = *reinterpret_cast<const uint2*>(l);
  }
}

__global__ void __launch_bounds__(1024, 1) gn_apply_kernel(GnApply a) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_GN_APPLY);
  __shared__ float s_mean[64], s_rstd[64];
  const int C = a.s.C0 + a.s.C1;
  const int cpg = C / a.groups;
  const int b = blockIdx.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int c = threadIdx.x * 4;
  // affine parameters are weights (HBM every step): fetched before waiting for the producer kernel
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
  pdl_wait();
  DFU_TR_MARK(6);
  // the activations do not depend on the statistics: request them first so both latencies overlap
  const int p0 = blockIdx.x * a.pix_per_cta;
  const int p1 = min(p0 + a.pix_per_cta, a.s.HW);
  float4 vin[kGnPixPerThread];
#pragma unroll
  for (int j = 0; j < kGnPixPerThread; ++j) {
    const int p = p0 + threadIdx.y + j * blockDim.y;
    vin[j] = (p < p1) ? gn_load(a.s, b, p, threadIdx.x) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (a.stats) {
    if (tid < a.groups) {
      const float2 t = a.stats[b * a.groups + tid];
      s_mean[tid] = t.x;
      s_rstd[tid] = t.y;
    }
  } else {
    // same arithmetic and order as gn_finalize_kernel: 8 threads per group, fixed-order double accumulation
    const int g = tid >> 3, sub = tid & 7;
    if (g < a.groups) {
      double sum = 0.0, sq = 0.0;
      const float2* pp = a.partial + static_cast<size_t>(b) * a.nchunks * a.groups + g;
      int k = sub;
      for (; k + 24 < a.nchunks; k += 32) {  // four independent loads in flight per thread, fixed summation order
        const float2 t0 = pp[static_cast<size_t>(k) * a.groups];
        const float2 t1 = pp[static_cast<size_t>(k + 8) * a.groups];
        const float2 t2 = pp[static_cast<size_t>(k + 16) * a.groups];
        const float2 t3 = pp[static_cast<size_t>(k + 24) * a.groups];
        sum = (((sum + t0.x) + t1.x) + t2.x) + t3.x;
        sq = (((sq + t0.y) + t1.y) + t2.y) + t3.y;
      }
      for (; k < a.nchunks; k += 8) {
        const float2 t = pp[static_cast<size_t>(k) * a.groups];
        sum += t.x;
        sq += t.y;
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      if (sub == 0) {
        const double mean = sum / a.count;
        double var = sq / a.count - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = static_cast<float>(mean);
        s_rstd[g] = static_cast<float>(1.0 / sqrt(var + a.eps));
      }
    }
  }
  __syncthreads();
  DFU_TR_MARK(7);
  float mu[4], rs[4];
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
  const float be[4] = {be4.x, be4.y, be4.z, be4.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int g = (c + i) / cpg;
    mu[i] = s_mean[g];
    rs[i] = s_rstd[g];
  }
#pragma unroll
  for (int j = 0; j < kGnPixPerThread; ++j) {
    const int p = p0 + threadIdx.y + j * blockDim.y;
    if (p >= p1) break;
    const float4 v = vin[j];
    float4 y;
    y.x = (v.x - mu[0]) * rs[0] * ga[0] + be[0];
    y.y = (v.y - mu[1]) * rs[1] * ga[1] + be[1];
    y.z = (v.z - mu[2]) * rs[2] * ga[2] + be[2];
    y.w = (v.w - mu[3]) * rs[3] * ga[3] + be[3];
    if (a.silu) {
      y.x = silu_f(y.x); y.y = silu_f(y.y); y.z = silu_f(y.z); y.w = silu_f(y.w);
    }
    const size_t off = (static_cast<size_t>(b) * a.s.HW + p) * C + c;
    if (a.out16) store_split4(a.out16 + off, a.plane_stride, a.planes, y);
    if (a.out32) *reinterpret_cast<float4*>(a.out32 + off) = y;
    if (a.raw16) store_split4(a.raw16 + off, a.plane_stride, a.planes, v);
  }
  DFU_TR_END();
}

// =============================================================================================
// Cluster GroupNorm: ONE launch, the input read ONCE, no grid-wide synchronisation.  The statistics of a group only
// involve that group's channels, so the work is cut by CHANNELS first: a "unit" is the smallest run of groups whose
// channel count is a multiple of 4 (1, 2 or 4 groups; q quads per pixel), and one thread-block cluster of S CTAs owns
// one (sample, unit): CTA `rank` takes a slab of pixels, every thread keeps its <= ITEMS quads in registers, the
// per-group sums are reduced warp -> CTA (shared memory, fixed order, double) -> cluster (distributed shared memory,
// rank order) and the normalisation is applied from the registers.  Replaces gn_stats + gn_apply (two launches, two
// reads, ~10 us per GroupNorm in the captured UNet step) wherever ITEMS <= 16 quads per thread fit.
// =============================================================================================
constexpr int kGnMaxU = 4;
template <int ITEMS>
__global__ void __launch_bounds__(512) gn_cluster_kernel(GnApply a, int U, int q, int TY, int S) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_GN_FUSED);
  __shared__ float s_w[16][kGnMaxU][2];
  __shared__ __align__(16) double s_cta[kGnMaxU][2];
  __shared__ float s_mean[kGnMaxU], s_rstd[kGnMaxU];
  const int C = a.s.C0 + a.s.C1;
  const int cpg = C / a.groups;
  const int rank = static_cast<int>(cluster_ctarank());
  const int unit = blockIdx.x / S;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int tx = tid % q, ty = tid / q;  // ty >= TY: spare threads of the last warp (no pixels, zero partials)
  const int cq = unit * q + tx;
  const int c = cq * 4;
  int gs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) gs[k] = (c + k) / cpg - unit * U;
  // affine parameters are weights: fetched before waiting for the producer kernel
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
  pdl_wait();
  DFU_TR_MARK(6);
  const int ppc = (a.s.HW + S - 1) / S;
  const int p0 = rank * ppc;
  const int p1 = min(p0 + ppc, a.s.HW);
  float4 v[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int p = p0 + ty + i * TY;
    v[i] = (ty < TY && p < p1) ? gn_load(a.s, b, p, cq) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float sx[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    sx[0] += v[i].x; sq[0] += v[i].x * v[i].x;
    sx[1] += v[i].y; sq[1] += v[i].y * v[i].y;
    sx[2] += v[i].z; sq[2] += v[i].z * v[i].z;
    sx[3] += v[i].w; sq[3] += v[i].w * v[i].w;
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int u = 0; u < kGnMaxU; ++u) {
    if (u < U) {  // uniform
      float gsum = 0.f, gsq = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (gs[k] == u) {
          gsum += sx[k];
          gsq += sq[k];
        }
      gsum = warp_sum(gsum);
      gsq = warp_sum(gsq);
      if (lane == 0) {
        s_w[warp][u][0] = gsum;
        s_w[warp][u][1] = gsq;
      }
    }
  }
  __syncthreads();
  if (tid < U) {
    double sd = 0.0, qd = 0.0;
    for (int w = 0; w < nwarps; ++w) {
      sd += s_w[w][tid][0];
      qd += s_w[w][tid][1];
    }
    s_cta[tid][0] = sd;
    s_cta[tid][1] = qd;
  }
  cluster_sync_all();  // every CTA's partial is in its shared memory
  if (tid < U) {
    double sd = 0.0, qd = 0.0;
    const uint32_t loc = smem_u32(&s_cta[tid][0]);
    double ps[8], pq[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {  // all remote loads in flight before the first add (S <= 8)
      const uint32_t ra = dsmem_addr(loc, static_cast<uint32_t>(r < S ? r : 0));
      ps[r] = ld_dsmem_f64(ra);
      pq[r] = ld_dsmem_f64(ra + 8);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {  // rank order: deterministic
      if (r < S) {
        sd += ps[r];
        qd += pq[r];
      }
    }
    const double mean = sd / a.count;
    double var = qd / a.count - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + a.eps));
  }
  cluster_arrive();  // "I have read my peers' partials"; the matching wait is the last statement of the kernel
  __syncthreads();
  DFU_TR_MARK(7);
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
  const float be[4] = {be4.x, be4.y, be4.z, be4.w};
  float mu[4], sc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mu[k] = s_mean[gs[k]];
    sc[k] = s_rstd[gs[k]] * ga[k];
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int p = p0 + ty + i * TY;
    if (ty < TY && p < p1) {
      const float4 x = v[i];
      float4 y;
      y.x = (x.x - mu[0]) * sc[0] + be[0];
      y.y = (x.y - mu[1]) * sc[1] + be[1];
      y.z = (x.z - mu[2]) * sc[2] + be[2];
      y.w = (x.w - mu[3]) * sc[3] + be[3];
      if (a.silu) {
        y.x = silu_f(y.x); y.y = silu_f(y.y); y.z = silu_f(y.z); y.w = silu_f(y.w);
      }
      const size_t off = (static_cast<size_t>(b) * a.s.HW + p) * C + c;
      if (a.out16) store_split4(a.out16 + off, a.plane_stride, a.planes, y);
      if (a.out32) *reinterpret_cast<float4*>(a.out32 + off) = y;
      if (a.raw16) store_split4(a.raw16 + off, a.plane_stride, a.planes, x);
    }
  }
  DFU_TR_END();
  cluster_wait();  // nobody leaves while a peer may still read its partial
}

// =============================================================================================
// Cluster GroupNorm with the slab staged in SHARED memory (cp.async, 16 bytes per request, no register cost): the same
// one-launch / one-read scheme as gn_cluster_kernel for slabs beyond 16 quads per thread — batched UNet levels (8 x 64 x
// 64 x 320 ran as five waves of register-resident clusters at 1.9 TB/s) and mid-sized VAE maps that fell back to the
// two-launch path.  A thread walks its quads at a stride that keeps its channel quad fixed (blockDim = TY * q), so the
// per-group bookkeeping is identical; the second pass reads the slab back from shared memory instead of registers.
// =============================================================================================
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(640) gn_cluster_smem_kernel(GnApply a, int U, int q, int TY, int S) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_GN_FUSED);
  extern __shared__ __align__(16) float4 slab[];  // [pixel - p0][q]
  __shared__ float s_w[20][kGnMaxU][2];
  __shared__ __align__(16) double s_cta[kGnMaxU][2];
  __shared__ float s_mean[kGnMaxU], s_rstd[kGnMaxU];
  const int C = a.s.C0 + a.s.C1;
  const int cpg = C / a.groups;
  const int rank = static_cast<int>(cluster_ctarank());
  const int unit = blockIdx.x / S;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int tx = tid % q, ty = tid / q;  // ty >= TY: spare threads of the last warp
  const int cq = unit * q + tx;
  const int c = cq * 4;
  int gs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) gs[k] = (c + k) / cpg - unit * U;
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
  pdl_wait();
  DFU_TR_MARK(6);
  const int ppc = (a.s.HW + S - 1) / S;
  const int p0 = rank * ppc;
  const int p1 = min(p0 + ppc, a.s.HW);
  if (ty < TY) {
    const float* src = (c < a.s.C0) ? a.s.src0 + static_cast<size_t>(b) * a.s.HW * a.s.C0 + c
                                    : a.s.src1 + static_cast<size_t>(b) * a.s.HW * a.s.C1 + (c - a.s.C0);
    const int ld = (c < a.s.C0) ? a.s.C0 : a.s.C1;
    for (int p = p0 + ty; p < p1; p += TY) cp_async16(&slab[(p - p0) * q + tx], src + static_cast<size_t>(p) * ld);
  }
  cp_async_wait_all();
  float sx[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
  if (ty < TY) {
    for (int p = p0 + ty; p < p1; p += TY) {  // (own copies only: no barrier needed before reading them back)
      const float4 v = slab[(p - p0) * q + tx];
      sx[0] += v.x; sq[0] += v.x * v.x;
      sx[1] += v.y; sq[1] += v.y * v.y;
      sx[2] += v.z; sq[2] += v.z * v.z;
      sx[3] += v.w; sq[3] += v.w * v.w;
    }
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int u = 0; u < kGnMaxU; ++u) {
    if (u < U) {
      float gsum = 0.f, gsq = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (gs[k] == u) {
          gsum += sx[k];
          gsq += sq[k];
        }
      gsum = warp_sum(gsum);
      gsq = warp_sum(gsq);
      if (lane == 0) {
        s_w[warp][u][0] = gsum;
        s_w[warp][u][1] = gsq;
      }
    }
  }
  __syncthreads();
  if (tid < U) {
    double sd = 0.0, qd = 0.0;
    for (int w = 0; w < nwarps; ++w) {
      sd += s_w[w][tid][0];
      qd += s_w[w][tid][1];
    }
    s_cta[tid][0] = sd;
    s_cta[tid][1] = qd;
  }
  cluster_sync_all();
  if (tid < U) {
    double sd = 0.0, qd = 0.0;
    const uint32_t loc = smem_u32(&s_cta[tid][0]);
    double ps[8], pq[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint32_t ra = dsmem_addr(loc, static_cast<uint32_t>(r < S ? r : 0));
      ps[r] = ld_dsmem_f64(ra);
      pq[r] = ld_dsmem_f64(ra + 8);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < S) {
        sd += ps[r];
        qd += pq[r];
      }
    }
    const double mean = sd / a.count;
    double var = qd / a.count - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + a.eps));
  }
  cluster_arrive();
  __syncthreads();
  DFU_TR_MARK(7);
  if (ty < TY) {
    const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
    const float be[4] = {be4.x, be4.y, be4.z, be4.w};
    float mu[4], sc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mu[k] = s_mean[gs[k]];
      sc[k] = s_rstd[gs[k]] * ga[k];
    }
    for (int p = p0 + ty; p < p1; p += TY) {
      const float4 x = slab[(p - p0) * q + tx];
      float4 y;
      y.x = (x.x - mu[0]) * sc[0] + be[0];
      y.y = (x.y - mu[1]) * sc[1] + be[1];
      y.z = (x.z - mu[2]) * sc[2] + be[2];
      y.w = (x.w - mu[3]) * sc[3] + be[3];
      if (a.silu) {
        y.x = silu_f(y.x); y.y = silu_f(y.y); y.z = silu_f(y.z); y.w = silu_f(y.w);
      }
      const size_t off = (static_cast<size_t>(b) * a.s.HW + p) * C + c;
      if (a.out16) store_split4(a.out16 + off, a.plane_stride, a.planes, y);
      if (a.out32) *reinterpret_cast<float4*>(a.out32 + off) = y;
      if (a.raw16) store_split4(a.raw16 + off, a.plane_stride, a.planes, x);
    }
  }
  DFU_TR_END();
  cluster_wait();
}

// =============================================================================================
// LayerNorm over the channel dim of [M, C] tokens -> fp16 operand planes.  One warp per token, two-pass in registers.
// =============================================================================================
constexpr int kLnMaxQuads = 10;  // C <= 1280

// QPL = float4 per lane per token (3: C <= 384, 5: C <= 640, 10: C <= 1280) keeps the register count proportional to the
// row length (64 resident warps per SM for the 320-wide level); TOK = tokens per warp, loaded together: at batch 8 the
// kernel is bandwidth-bound and needs the bytes in flight, at batch 1 it is latency-bound and TOK = 1 keeps the grid wide.
template <int QPL, int TOK>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ out16, int planes,
                 long long plane_stride, float* __restrict__ out32) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_LAYERNORM);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int C4 = C >> 2;
  // gamma / beta are weights (they stream from HBM every step): copy them to shared memory BEFORE waiting for the
  // producer kernel, so that their latency overlaps its tail instead of sitting on the critical path after the
  // reductions (shared memory, not registers: 80 more live registers would halve the occupancy)
  __shared__ __align__(16) float s_g[QPL * 128], s_b[QPL * 128];
  for (int qd = threadIdx.x; qd < C4; qd += blockDim.x) {
    reinterpret_cast<float4*>(s_g)[qd] = __ldg(reinterpret_cast<const float4*>(gamma) + qd);
    reinterpret_cast<float4*>(s_b)[qd] = __ldg(reinterpret_cast<const float4*>(beta) + qd);
  }
  __syncthreads();
  pdl_wait();
  DFU_TR_MARK(6);
  const int tok0 = warp * TOK;
  float4 v[TOK][QPL];
#pragma unroll
  for (int t = 0; t < TOK; ++t) {
    const float4* row = reinterpret_cast<const float4*>(x + static_cast<size_t>(tok0 + t) * C);
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (tok0 + t < M && q < C4) v[t][i] = row[q];
    }
  }
#pragma unroll
  for (int t = 0; t < TOK; ++t) {
    if (tok0 + t >= M) break;  // (warp-uniform)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (q < C4) s += v[t][i].x + v[t][i].y + v[t][i].z + v[t][i].w;
    }
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (q < C4) {
        const float a = v[t][i].x - mean, b = v[t][i].y - mean, c = v[t][i].z - mean, d = v[t][i].w - mean;
        ss += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    const size_t base = static_cast<size_t>(tok0 + t) * C;
#pragma unroll
    for (int i = 0; i < QPL; ++i) {
      const int q = lane + 32 * i;
      if (q < C4) {
        const float4 g = reinterpret_cast<const float4*>(s_g)[q];
        const float4 b = reinterpret_cast<const float4*>(s_b)[q];
        float4 y;
        y.x = (v[t][i].x - mean) * rstd * g.x + b.x;
        y.y = (v[t][i].y - mean) * rstd * g.y + b.y;
        y.z = (v[t][i].z - mean) * rstd * g.z + b.z;
        y.w = (v[t][i].w - mean) * rstd * g.w + b.w;
        if (out16) store_split4(out16 + base + q * 4, plane_stride, planes, y);
        if (out32) *reinterpret_cast<float4*>(out32 + base + q * 4) = y;
      }
    }
  }
  DFU_TR_END();
}

// =============================================================================================
// ViT patch embedding operand: NCHW fp32 -> fp16 planes [B * (1 + nP)][C*P*P], one zero row (CLS slot) per sample.
// One thread per 4 consecutive kx of an output row: a 16-byte read, an 8-byte write per plane.
// =============================================================================================
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ x, int B, int C, int H, int W, int P, __half* __restrict__ out, int planes,
                long long plane_stride) {
  pdl_trigger();
  pdl_wait();
  const int gw = W / P, gh = H / P, K = C * P * P, K4 = K >> 2, rows = 1 + gw * gh;
  const long long total = static_cast<long long>(B) * rows * K4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kq = static_cast<int>(i % K4);
    const long long t = i / K4;
    const int r = static_cast<int>(t % rows);
    const int b = static_cast<int>(t / rows);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r > 0) {
      const int k = kq * 4, c = k / (P * P), rem = k - c * P * P, ky = rem / P, kx = rem - ky * P;
      const int py = (r - 1) / gw, px = (r - 1) - py * gw;
      v = *reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * C + c) * H + py * P + ky) * W + px * P + kx);
    }
    store_split4(out + static_cast<size_t>(t) * K + kq * 4, plane_stride, planes, v);
  }
}

// =============================================================================================
// fp32 NHWC -> fp16 operand casts.  mode 0: same geometry; 1: nearest 2x upsample; 2: space-to-depth (stride-2 conv)
// =============================================================================================
__global__ void __launch_bounds__(256)
cast_kernel(const float* __restrict__ x, int B, int H, int W, int C, int mode, __half* __restrict__ out, int planes,
            long long plane_stride) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_CAST);
  pdl_wait();
  DFU_TR_MARK(6);
  const int C4 = C >> 2;
  const long long total = static_cast<long long>(B) * H * W * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(i % C4);
    long long t = i / C4;
    const int xw = static_cast<int>(t % W);
    t /= W;
    const int y = static_cast<int>(t % H);
    const int b = static_cast<int>(t / H);
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    if (mode == 0) {
      store_split4(out + i * 4, plane_stride, planes, v);
    } else if (mode == 1) {
      const int H2 = 2 * H, W2 = 2 * W;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t o = ((static_cast<size_t>(b) * H2 + 2 * y + dy) * W2 + 2 * xw + dx) * C + cq * 4;
          store_split4(out + o, plane_stride, planes, v);
        }
    } else {
      const int Hh = H >> 1, Wh = W >> 1;
      const int par = (y & 1) * 2 + (xw & 1);
      const size_t o = (((static_cast<size_t>(par) * B + b) * Hh + (y >> 1)) * Wh + (xw >> 1)) * C + cq * 4;
      store_split4(out + o, plane_stride, planes, v);
    }
  }
  DFU_TR_END();
}

static int ew_grid(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const int cap = (num_sms() > 0 ? num_sms() : 148) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace dfu

DFU_TRACE_SETTER(dfu_trace_set_norm)

using namespace dfu;

// block (C/4, ty), ty rows of threads each walking kGnPixPerThread pixels -> ppc pixels per CTA
static void gn_geometry(int B, int HW, int C, int* ty, int* ppc, int* chunks) {
  const int C4 = C / 4;
  int t = 512 / C4;
  if (t < 1) t = 1;
  if (t > 16) t = 16;
  if (t > HW) t = HW;
  auto nchunks = [&](int tt) { return (HW + tt * kGnPixPerThread - 1) / (tt * kGnPixPerThread); };
  // a grid a little over one CTA per SM runs as two waves (measured: 171 CTAs at 64x64x320): grow the CTA, up to
  // 1024 threads, until the grid fits one wave
  const int sms = num_sms() > 0 ? num_sms() : 148;
  while (static_cast<long long>(nchunks(t)) * B > sms && static_cast<long long>(nchunks(t)) * B <= sms + sms / 2 &&
         (t + 1) * C4 <= 1024 && t < 16 && t < HW)
    ++t;
  *ty = t;
  *ppc = t * kGnPixPerThread;
  *chunks = nchunks(t);
}

// Number of S-CTA clusters of gn_cluster_kernel<...> (T threads) the device can hold at once; cached per (S, T, variant).
static int gn_cluster_capacity(int S, int T, int variant) {
  static int cache[4][17][3];  // [log2 S][T / 32][variant], 0 = unknown, -1 = query failed
  int ls = 0;
  while ((1 << ls) < S) ++ls;
  const int ti = T / 32;
  if (ls > 3 || ti > 16) return -1;
  int& c = cache[ls][ti][variant];
  if (c != 0) return c;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S * 64, 1, 1);
  cfg.blockDim = dim3(T, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(S);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e;
  if (variant == 0) e = cudaOccupancyMaxActiveClusters(&n, gn_cluster_kernel<4>, &cfg);
  else if (variant == 1) e = cudaOccupancyMaxActiveClusters(&n, gn_cluster_kernel<8>, &cfg);
  else e = cudaOccupancyMaxActiveClusters(&n, gn_cluster_kernel<16>, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = -1;
  }
  c = n > 0 ? n : -1;
  return c;
}

// resident clusters of gn_cluster_smem_kernel for (S, T threads, dynamic smem); cached in a small table
static int gn_smem_capacity(int S, int T, size_t smem) {
  struct Key { int S, T; size_t smem; int cap; };
  static Key cache[64];
  static int n = 0;
  for (int i = 0; i < n; ++i)
    if (cache[i].S == S && cache[i].T == T && cache[i].smem == smem) return cache[i].cap;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S * 64, 1, 1);
  cfg.blockDim = dim3(T, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(S);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int c = 0;
  if (cudaOccupancyMaxActiveClusters(&c, gn_cluster_smem_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    c = -1;
  }
  if (c <= 0) c = -1;
  if (n < 64) cache[n++] = Key{S, T, smem, c};
  return c;
}

extern "C" {

size_t dfu_groupnorm_workspace(int B, int HW, int C, int groups) {
  int ty, ppc, chunks;
  gn_geometry(B, HW, C, &ty, &ppc, &chunks);
  return static_cast<size_t>(B) * (chunks + 1) * groups * sizeof(float2);
}

int dfu_groupnorm(const float* src0, int C0, const float* src1, int C1, int B, int HW, int groups, const float* gamma,
                  const float* beta, float eps, int silu, void* out16, int planes, int64_t plane_stride, float* out32,
                  void* raw16, void* workspace, size_t workspace_bytes, void* sync_words, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = C0 + C1;
  DFU_REQUIRE(src0 && C0 > 0 && C0 % 4 == 0 && C1 % 4 == 0 && (C1 == 0 || src1), "groupnorm: bad sources");
  DFU_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: C=%d groups=%d", C, groups);
  DFU_REQUIRE(C / 4 <= 1024, "groupnorm: C=%d too wide", C);
  DFU_REQUIRE(out16 || out32, "groupnorm: no output");
  const size_t need = dfu_groupnorm_workspace(B, HW, C, groups);
  if (!workspace || workspace_bytes < need) {
    set_error("groupnorm: workspace %zu < %zu", workspace_bytes, need);
    return DFU_ERR_WORKSPACE;
  }
  int ty, ppc, chunks;
  gn_geometry(B, HW, C, &ty, &ppc, &chunks);
  const int C4 = C / 4;
  dim3 block(C4, ty), grid(chunks, B);
  GnSrc s{src0, src1, C0, C1, HW};
  // ---- single-launch cluster path (see gn_cluster_kernel) whenever <= 16 quads per thread fit ----
  static const bool cluster_ok = !(getenv("DFU_GN_CLUSTER") && getenv("DFU_GN_CLUSTER")[0] == '0');
  const int cpg = C / groups;
  const int U = (cpg % 4 == 0) ? 1 : (cpg % 2 == 0 ? 2 : 4);
  if (cluster_ok && groups % U == 0 && (U * cpg) % 4 == 0 && (U * cpg) / 4 <= 256) {
    const int q = U * cpg / 4;
    const int nunits = groups / U;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    int bestS = 0, bestTY = 0, bestItems = 0;
    double best = 1e30;
    (void)sms;
    static const int max_items = getenv("DFU_GN_MAX_ITEMS") ? atoi(getenv("DFU_GN_MAX_ITEMS")) : 16;
    for (int S = 8; S >= 1; S >>= 1) {
      const int ppcS = (HW + S - 1) / S;
      if (ppcS * (S - 1) >= HW && S > 1) continue;  // an empty rank
      for (int T = 256; T <= 512; T += 256) {
        const int TYc = T / q;
        if (TYc < 1) continue;
        const int items = (ppcS + TYc - 1) / TYc;
        if (items > max_items) continue;
        const long long clusters = static_cast<long long>(nunits) * B;
        // clusters that can be resident at once (GPC granularity: e.g. only ~16 clusters of 8 fit a B200)
        const int cap = gn_cluster_capacity(S, ((q * TYc + 31) / 32) * 32, items <= 4 ? 0 : (items <= 8 ? 1 : 2));
        if (cap <= 0) continue;
        const double waves = static_cast<double>((clusters + cap - 1) / cap);
        const double cost = waves * (items + 6) * (T == 512 ? 1.1 : 1.0);
        if (cost < best) {
          best = cost; bestS = S; bestTY = TYc; bestItems = items;
        }
      }
    }
    if (const char* force = getenv("DFU_GN_FORCE")) {  // diagnostics: "S,T" overrides the choice (scripts/bench_gn.py)
      int fs = 0, ft = 0;
      if (sscanf(force, "%d,%d", &fs, &ft) == 2 && fs >= 1 && ft >= 32) {
        const int ppcS = (HW + fs - 1) / fs;
        const int TYc = ft / q;
        if (TYc >= 1 && (ppcS + TYc - 1) / TYc <= 16) {
          bestS = fs; bestTY = TYc; bestItems = (ppcS + TYc - 1) / TYc;
        }
      }
    }
    // ---- shared-memory-staged variant: when the register-resident clusters would run as several waves (batched levels)
    // or do not fit at all (maps that used to fall back to two launches) ----
    static const bool smem_ok = !(getenv("DFU_GN_SMEM") && getenv("DFU_GN_SMEM")[0] == '0');
    static const bool smem_force = getenv("DFU_GN_SMEM") && getenv("DFU_GN_SMEM")[0] == '2';  // experiments
    const long long nclusters = static_cast<long long>(nunits) * B;
    double reg_us = 1e30;  // rough cost of the register path: ~7 us per wave of clusters (measured)
    if (bestS > 0) reg_us = best / (bestItems + 6) * 7.0;
    // (measured at batch 1 too: the staged variant is equal or faster — 64x64x320 7.0 -> 6.2 us, 16x16x1280 3.7 -> 2.9 us —
    // except for the 15-quad units of the 960 / 1920-channel concats)
    if (smem_ok && (bestS == 0 || reg_us >= 14.0 || smem_force || q <= 10)) {
      if (first_use_on_device(ONCE_GN_SMEM_ATTR))
        DFU_CHECK_CUDA(cudaFuncSetAttribute(gn_cluster_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int sS = 0, sTY = 0;
      size_t sBytes = 0;
      double sBest = 1e30;
      for (int S = 8; S >= 1; S >>= 1) {
        const int ppcS = (HW + S - 1) / S;
        if (ppcS * (S - 1) >= HW && S > 1) continue;
        const size_t bytes = static_cast<size_t>(ppcS) * q * 16;
        if (bytes > 200 * 1024) continue;
        int TYc = 512 / q;
        if (TYc > ppcS) TYc = ppcS;
        if (TYc < 1) continue;
        const int T = ((q * TYc + 31) / 32) * 32;
        if (T > 640) continue;
        const int cap = gn_smem_capacity(S, T, bytes);
        if (cap <= 0) continue;
        const double waves = static_cast<double>((nclusters + cap - 1) / cap);
        const double us = waves * (3.0 + static_cast<double>(bytes) / (30.0 * 1024));  // load + two passes over the slab
        if (us < sBest) {
          sBest = us; sS = S; sTY = TYc; sBytes = bytes;
        }
      }
      (void)reg_us;  // (measured: once the register path needs two waves the staged slab wins whenever it fits)
      if (sS > 0) {
        GnApply f;
        f.s = s; f.groups = groups; f.stats = nullptr; f.partial = nullptr; f.nchunks = 0; f.pix_per_cta = 0;
        f.count = static_cast<double>(cpg) * HW;
        f.gamma = gamma; f.beta = beta; f.eps = eps; f.silu = silu;
        f.out16 = static_cast<__half*>(out16); f.planes = planes; f.plane_stride = plane_stride;
        f.out32 = out32; f.raw16 = static_cast<__half*>(raw16);
        const int T = ((q * sTY + 31) / 32) * 32;
        DFU_CHECK_CUDA(launch_kc(gn_cluster_smem_kernel, dim3(sS * nunits, B), dim3(T), sBytes, stream, sS, f, U, q, sTY, sS));
        DFU_CHECK_CUDA(cudaGetLastError());
        return DFU_OK;
      }
    }
    if (bestS > 0) {
      GnApply f;
      f.s = s; f.groups = groups; f.stats = nullptr; f.partial = nullptr; f.nchunks = 0; f.pix_per_cta = 0;
      f.count = static_cast<double>(cpg) * HW;
      f.gamma = gamma; f.beta = beta; f.eps = eps; f.silu = silu;
      f.out16 = static_cast<__half*>(out16); f.planes = planes; f.plane_stride = plane_stride;
      f.out32 = out32; f.raw16 = static_cast<__half*>(raw16);
      const int T = ((q * bestTY + 31) / 32) * 32;
      dim3 cgrid(bestS * nunits, B);
      cudaError_t e;
      if (bestItems <= 4)
        e = launch_kc(gn_cluster_kernel<4>, cgrid, dim3(T), 0, stream, bestS, f, U, q, bestTY, bestS);
      else if (bestItems <= 8)
        e = launch_kc(gn_cluster_kernel<8>, cgrid, dim3(T), 0, stream, bestS, f, U, q, bestTY, bestS);
      else
        e = launch_kc(gn_cluster_kernel<16>, cgrid, dim3(T), 0, stream, bestS, f, U, q, bestTY, bestS);
      DFU_CHECK_CUDA(e);
      DFU_CHECK_CUDA(cudaGetLastError());
      return DFU_OK;
    }
  }
  const size_t stats_smem = static_cast<size_t>(ty) * 2 * C * sizeof(float);
  // Two-launch path (maps too large for the cluster kernel).  Built, measured in the captured step and removed: a
  // single-launch grid-barrier variant (15.9 vs 12.3 us at 64x64x320: the barrier costs more than the launch it
  // saves) and letting the last-arriving statistics CTA finalise (mean, rstd) (0.95 vs 0.69 ms of GroupNorm per UNet
  // step: a serial tail on the critical path).
  DFU_CHECK_CUDA(launch_k(gn_stats_kernel, dim3(grid), dim3(block), stats_smem, stream, s, groups, ppc, static_cast<float2*>(workspace)));
  GnApply a;
  a.s = s;
  a.groups = groups;
  a.pix_per_cta = ppc;
  a.partial = static_cast<const float2*>(workspace);
  a.nchunks = chunks;
  a.count = static_cast<double>(C / groups) * HW;
  a.stats = nullptr;
  const bool inline_finalize = chunks <= 256 && static_cast<int>(block.x * block.y) >= 8 * groups;
  if (!inline_finalize) {
    float2* stats = static_cast<float2*>(workspace) + static_cast<size_t>(B) * chunks * groups;
    DFU_CHECK_CUDA(launch_k(gn_finalize_kernel, dim3(groups, B), dim3(256), 0, stream, static_cast<const float2*>(workspace), chunks, groups, a.count, eps, stats));
    a.stats = stats;
  }
  a.gamma = gamma;
  a.beta = beta;
  a.eps = eps;
  a.silu = silu;
  a.out16 = static_cast<__half*>(out16);
  a.planes = planes;
  a.plane_stride = plane_stride;
  a.out32 = out32;
  a.raw16 = static_cast<__half*>(raw16);
  DFU_CHECK_CUDA(launch_k(gn_apply_kernel, dim3(grid), dim3(block), 0, stream, a));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_patchify_f16(const float* x, int B, int C, int H, int W, int P, void* out16, int planes, int64_t plane_stride,
                     void* stream_) {
  DFU_REQUIRE(x && out16 && P > 0 && P % 4 == 0 && H % P == 0 && W % P == 0 && W % 4 == 0, "patchify: H=%d W=%d P=%d", H, W, P);
  const long long total = static_cast<long long>(B) * (1 + (H / P) * (W / P)) * (C * P * P / 4);
  DFU_CHECK_CUDA(launch_k(patchify_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream_), x, B, C, H, W, P, static_cast<__half*>(out16), planes, plane_stride));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, void* out16,
                  int planes, int64_t plane_stride, float* out32, void* stream_) {
  DFU_REQUIRE(C % 4 == 0 && C / 4 <= 32 * kLnMaxQuads, "layernorm: C=%d unsupported", C);
  DFU_REQUIRE(out16 || out32, "layernorm: no output");
  const int warps_per_block = 8;
  // two tokens per warp once the grid is several waves deep (batched levels: bandwidth-bound, wants bytes in flight)
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int tok = (M >= sms * 8 * 8) ? 2 : 1;
  const int blocks = (M + warps_per_block * tok - 1) / (warps_per_block * tok);
  const int qpl = C / 4 <= 96 ? 3 : (C / 4 <= 160 ? 5 : 10);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  __half* o16 = static_cast<__half*>(out16);
#define DFU_LN_LAUNCH(Q, T) launch_k(layernorm_kernel<Q, T>, dim3(blocks), dim3(warps_per_block * 32), 0, st, x, M, C, gamma, beta, eps, o16, planes, plane_stride, out32)
  cudaError_t e;
  if (qpl == 3) e = tok == 2 ? DFU_LN_LAUNCH(3, 2) : DFU_LN_LAUNCH(3, 1);
  else if (qpl == 5) e = tok == 2 ? DFU_LN_LAUNCH(5, 2) : DFU_LN_LAUNCH(5, 1);
  else e = tok == 2 ? DFU_LN_LAUNCH(10, 2) : DFU_LN_LAUNCH(10, 1);
#undef DFU_LN_LAUNCH
  DFU_CHECK_CUDA(e);
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_cast_f16(const float* x, int B, int H, int W, int C, int mode, void* out16, int planes, int64_t plane_stride,
                 void* stream_) {
  DFU_REQUIRE(C % 4 == 0, "cast: C=%d", C);
  DFU_REQUIRE(mode >= 0 && mode <= 2, "cast: mode");
  DFU_REQUIRE(!(mode == 2 && ((H | W) & 1)), "cast: space-to-depth needs even H, W");
  const long long total = static_cast<long long>(B) * H * W * (C / 4);
  DFU_CHECK_CUDA(launch_k(cast_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream_), x, B, H, W, C, mode, static_cast<__half*>(out16), planes, plane_stride));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}
}

// diffute_b200 — small fp32 CUDA-core kernels around the tensor-core path: timestep embedding, batched GEMV
// (time-embedding MLP and the 22 time_emb_proj layers in one launch), the few-channel edge convolutions
// (conv_in 9->320, conv_out 320->4 with the scheduler step fused, VAE 3->128 / 128->3 / 4->512 / 512->8 and the
// 1x1 quant convs), row softmax for the single-head VAE attention and the scheduler updates.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace dfu {

// ---------------------------------------------------------------------------------------------
// diffusers `Timesteps`: emb[b, :half] = cos(t * f_i), emb[b, half:] = sin(t * f_i)  (flip_sin_to_cos=True)
// f_i = exp(-ln(10000) * i / (half - freq_shift))
// ---------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, int B, int dim, int flip_sin_to_cos,
                                      float freq_shift, float* __restrict__ out) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_TEMB);
  pdl_wait();
  DFU_TR_MARK(6);
  const int half = dim / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * half; i += gridDim.x * blockDim.x) {
    const int b = i / half, k = i % half;
    const float f = expf(-9.210340371976184f * static_cast<float>(k) / (static_cast<float>(half) - freq_shift));
    const float arg = t[b] * f;
    float s, c;
    sincosf(arg, &s, &c);
    float* o = out + static_cast<size_t>(b) * dim;
    if (flip_sin_to_cos) {
      o[k] = c;
      o[half + k] = s;
    } else {
      o[k] = s;
      o[half + k] = c;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// out[b, n] = act_out( bias[n] + sum_k W[n, k] * act_in(x[b, k]) ), fp32 weights streamed once with 128-bit loads.
// One warp per output column n; B <= 16 rows are accumulated together so W is read exactly once.
// ---------------------------------------------------------------------------------------------
constexpr int kGemvMaxB = 16;

__global__ void __launch_bounds__(256)
gemv_kernel(const float* __restrict__ x, int B, int K, int ldx, const float* __restrict__ W,
            const float* __restrict__ bias, int N, int silu_in, int silu_out, float* __restrict__ out, int ldo) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_GEMV);
  pdl_wait();
  DFU_TR_MARK(6);
  extern __shared__ float xs[];  // [B][K] activated input
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    const int b = i / K, k = i % K;
    float v = x[static_cast<size_t>(b) * ldx + k];
    xs[i] = silu_in ? silu_f(v) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int n = blockIdx.x * warps + (threadIdx.x >> 5); n < N; n += gridDim.x * warps) {
    float acc[kGemvMaxB];
#pragma unroll
    for (int b = 0; b < kGemvMaxB; ++b) acc[b] = 0.f;
    const float4* w = reinterpret_cast<const float4*>(W + static_cast<size_t>(n) * K);
    for (int q = lane; q < K / 4; q += 32) {
      const float4 wv = __ldg(w + q);
#pragma unroll
      for (int b = 0; b < kGemvMaxB; ++b) {
        if (b < B) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + b * K + q * 4);
          acc[b] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < kGemvMaxB; ++b) {
      if (b < B) {
        float v = warp_sum(acc[b]);
        if (lane == 0) {
          v += bias ? bias[n] : 0.f;
          out[static_cast<size_t>(b) * ldo + n] = silu_out ? silu_f(v) : v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Few-input-channel 3x3 / 1x1 conv: NCHW fp32 in (Cin <= 16, optionally gathered from up to 3 tensors, i.e. the
// cat([latents, mask, masked_latents]) of app.ipynb:811 is never materialised) -> NHWC fp32 out.  pad = k/2.
// ---------------------------------------------------------------------------------------------
struct SmallInSrc {
  const float* p[3];
  int c[3];
  long long bstride[3];  // elements between samples (0 broadcasts one sample to the batch)
  int nhwc;              // 1: sources are NHWC [B,H,W,c] instead of NCHW
};

// output pixels per CTA: 32, or 16 when 32 would leave SMs without a second CTA (the 64 x 64 UNet input at batch 1)

template <int kSmallInPix>
__global__ void __launch_bounds__(512)
conv_small_in_kernel(SmallInSrc s, int B, int H, int W, int Cin, int ksz, const float* __restrict__ wt,
                     const float* __restrict__ bias, int Cout, float pre_scale, float* __restrict__ out) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_CONV_IN);
  // block: kSmallInPix consecutive output pixels x all Cout (one thread per output channel: coalesced NHWC stores);
  // input patches staged in smem as [K][pixel] so that four pixels come with one 16-byte broadcast read; weights are
  // stored [Cin*k*k][Cout] so that consecutive lanes read consecutive addresses.
  extern __shared__ __align__(16) float sm[];  // [K][kSmallInPix]
  const int kk = ksz * ksz;
  const int K = Cin * kk;
  const int pad = ksz / 2;
  const int co = threadIdx.x;
  // the first nine taps' weights and the bias are constants: requested before waiting for the producer of the input
  float wv[9];
#pragma unroll
  for (int u = 0; u < 9; ++u) wv[u] = (co < Cout && u < K) ? __ldg(wt + static_cast<size_t>(u) * Cout + co) : 0.f;
  const float b0 = (bias && co < Cout) ? __ldg(bias + co) : 0.f;
  pdl_wait();
  DFU_TR_MARK(6);
  const int pix0 = blockIdx.x * kSmallInPix;  // pixel indices fit 32 bits (checked by the host): no 64-bit divisions
  const int npix = B * H * W;
#pragma unroll 4  // independent gathers: let four loads be in flight per thread instead of one
  for (int i = threadIdx.x; i < kSmallInPix * K; i += blockDim.x) {
    const int k = i / kSmallInPix, pl = i % kSmallInPix;  // consecutive threads -> consecutive pixels (coalesced NCHW)
    const int pix = pix0 + pl;
    float v = 0.f;
    if (pix < npix) {
      const int c = k / kk, t = k % kk;
      const int ky = t / ksz, kx = t % ksz;
      const int x = pix % W;
      const int yb = pix / W;
      const int y = yb % H;
      const int b = yb / H;
      const int iy = y + ky - pad, ix = x + kx - pad;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        int cc = c, si = 0;
        if (cc >= s.c[0]) { cc -= s.c[0]; si = 1; if (cc >= s.c[1]) { cc -= s.c[1]; si = 2; } }
        const long long o = s.nhwc ? (static_cast<long long>(iy) * W + ix) * s.c[si] + cc
                                   : (static_cast<long long>(cc) * H + iy) * W + ix;
        v = s.p[si][b * s.bstride[si] + o] * pre_scale;
      }
    }
    sm[i] = v;
  }
  __syncthreads();
  if (co < Cout) {
    float acc[kSmallInPix];
#pragma unroll
    for (int j = 0; j < kSmallInPix; ++j) acc[j] = b0;
    for (int k0 = 0; k0 < K; k0 += 9) {
      // the next nine weights are requested before this group's FMAs: their L2 latency hides behind ~300 instructions
      float wn[9];
#pragma unroll
      for (int u = 0; u < 9; ++u)
        wn[u] = (k0 + 9 + u < K) ? __ldg(wt + static_cast<size_t>(k0 + 9 + u) * Cout + co) : 0.f;
#pragma unroll
      for (int u = 0; u < 9; ++u) {
        if (k0 + u < K) {
          const float4* row = reinterpret_cast<const float4*>(sm + (k0 + u) * kSmallInPix);
#pragma unroll
          for (int j4 = 0; j4 < kSmallInPix / 4; ++j4) {
            const float4 x = row[j4];
            acc[4 * j4 + 0] = fmaf(wv[u], x.x, acc[4 * j4 + 0]);
            acc[4 * j4 + 1] = fmaf(wv[u], x.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(wv[u], x.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(wv[u], x.w, acc[4 * j4 + 3]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 9; ++u) wv[u] = wn[u];
    }
#pragma unroll
    for (int j = 0; j < kSmallInPix; ++j)
      if (pix0 + j < npix) out[static_cast<size_t>(pix0 + j) * Cout + co] = acc[j];
  }
  DFU_TR_END();
}

// ---------------------------------------------------------------------------------------------
// Counter-based Gaussian noise for the ancestral DDPM step (app.ipynb:816 draws it with torch.randn on the device):
// Philox4x32-10 (Salmon et al. 2011; key = 64-bit seed, counter = (element index lo, hi, step, 0)) -> two 24-bit
// uniforms in (0, 1) -> Box-Muller cosine branch.  A pure function of (seed, step, element): no state, no extra launch,
// the same stream whatever the tiling.  Restated in numpy and checked against the Random123 known-answer vectors.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__device__ __forceinline__ float philox_normal(uint32_t seed_lo, uint32_t seed_hi, uint32_t step, unsigned long long idx,
                                               uint32_t* bits = nullptr) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), step, 0u, seed_lo, seed_hi, r);
  if (bits) {
    bits[0] = r[0];
    bits[1] = r[1];
  }
  const float u1 = (static_cast<float>(r[0] >> 8) + 0.5f) * 5.9604644775390625e-8f;  // (k + 0.5) / 2^24 in (0, 1)
  const float u2 = (static_cast<float>(r[1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__global__ void __launch_bounds__(256) philox_normal_kernel(uint32_t seed_lo, uint32_t seed_hi, uint32_t step, long long n,
                                                            float* __restrict__ out, uint32_t* __restrict__ bits) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint32_t b[2];
    const float z = philox_normal(seed_lo, seed_hi, step, static_cast<unsigned long long>(i), b);
    if (out) out[i] = z;
    if (bits) {
      bits[2 * i] = b[0];
      bits[2 * i + 1] = b[1];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Few-output-channel conv: NHWC fp32 in (Cin multiple of 128... any multiple of 4) -> NCHW fp32 out (Cout <= 8),
// optional fused second 1x1 conv on the result (VAE quant_conv), optional fused scheduler update
// (x' = cx * x + ce * eps) so the UNet's conv_out writes the next latents directly.
// One warp per output pixel; lanes split the input channels.
// ---------------------------------------------------------------------------------------------
struct SmallOut {
  const float* x;      // [B,H,W,Cin]
  int B, H, W, Cin, ksz, Cout;
  const float* w;      // [Cout][ksz*ksz][Cin]  (repacked: channels innermost)
  const float* bias;
  const float* w2;     // optional [Cout2][Cout] 1x1 applied afterwards
  const float* b2;
  int Cout2;
  float* out;          // NCHW [B, Cout(2), H, W]
  // fused scheduler step (optional): sample/prev are NCHW [B,Cout,H,W]
  const float* sample;
  float* prev;
  const float* coef;   // device [2] = {cx, ce}; with `seed`: [4] = {cx, ce, sigma, step (uint32 bits)}
  const uint32_t* seed;  // optional device [2]: ancestral noise  prev += sigma * N(0, 1)[seed, step, element]
};

template <bool kSmemW>
__global__ void __launch_bounds__(256, kSmemW ? 2 : 4) conv_small_out_kernel(SmallOut p) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_CONV_OUT);
  // The packed weights ([Cout][k*k][Cin], up to ~150 KB) are constants: each CTA copies them to shared memory ONCE,
  // before waiting for the producer of the input, and its eight warps then walk pixels with a grid stride (was: one
  // warp per pixel re-reading all weights through L1 — 592 us for the 128->3 decoder output conv at 512x512).
  extern __shared__ __align__(16) float sw[];
  const int kk = p.ksz * p.ksz;
  const int wq = p.Cout * kk * (p.Cin >> 2);  // float4 count
  if (kSmemW) {
    for (int i = threadIdx.x; i < wq; i += blockDim.x)
      reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(p.w) + i);
    __syncthreads();
  }
  // (compile-time choice: a run-time select between the shared and the global pointer made every weight read a
  // generic-address load with 64-bit arithmetic — with the 64-bit pixel divisions, 2083 instructions per pixel, ncu)
  const float* wbase = kSmemW ? sw : p.w;
  pdl_wait();
  DFU_TR_MARK(6);
  const int lane = threadIdx.x & 31;
  const int npix = p.B * p.H * p.W;  // < 2^31 (checked by the host): 32-bit index arithmetic
  const int nwarps = static_cast<int>((gridDim.x * blockDim.x) >> 5);
  const int pad = p.ksz / 2;
  const int C4 = p.Cin >> 2;
  for (int pix = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5); pix < npix; pix += nwarps) {
    const int x = pix % p.W;
    const int yb = pix / p.W;
    const int y = yb % p.H;
    const int b = yb / p.H;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if constexpr (kSmemW) {
      // (many pixels per warp, few warps per SM) all taps of a channel quad are requested before the first FMA: one
      // memory latency per quad, not one per tap
      const float* xrow[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int iy = y + t / p.ksz - pad, ix = x + t % p.ksz - pad;
        const bool ok = t < kk && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        xrow[t] = ok ? p.x + ((static_cast<size_t>(b) * p.H + iy) * p.W + ix) * p.Cin : nullptr;
      }
      for (int q = lane; q < C4; q += 32) {
        float4 xv[9];
#pragma unroll
        for (int t = 0; t < 9; ++t)
          xv[t] = xrow[t] ? reinterpret_cast<const float4*>(xrow[t])[q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          if (t < kk) {
#pragma unroll
            for (int co = 0; co < 8; ++co) {
              if (co < p.Cout) {
                const float4 wv = *(reinterpret_cast<const float4*>(wbase + (static_cast<size_t>(co) * kk + t) * p.Cin) + q);
                acc[co] += wv.x * xv[t].x + wv.y * xv[t].y + wv.z * xv[t].z + wv.w * xv[t].w;
              }
            }
          }
        }
      }
    } else {
      // (one pixel per warp, every warp of the grid resident at once: latency is hidden by occupancy, so the loop
      // stays small — 64 registers; the batched form above needs 128 and ran this shape in two waves, 48 vs 11 us)
      for (int t = 0; t < kk; ++t) {
        const int iy = y + t / p.ksz - pad, ix = x + t % p.ksz - pad;
        if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
        const float4* xr = reinterpret_cast<const float4*>(p.x + ((static_cast<size_t>(b) * p.H + iy) * p.W + ix) * p.Cin);
        for (int q = lane; q < C4; q += 32) {
          const float4 xv = xr[q];
#pragma unroll
          for (int co = 0; co < 8; ++co) {
            if (co < p.Cout) {
              const float4 wv = __ldg(reinterpret_cast<const float4*>(p.w + (static_cast<size_t>(co) * kk + t) * p.Cin) + q);
              acc[co] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int co = 0; co < 8; ++co) acc[co] = warp_sum(acc[co]);
    if (lane == 0) {
      float v[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) v[co] = (co < p.Cout) ? acc[co] + (p.bias ? p.bias[co] : 0.f) : 0.f;
      int nout = p.Cout;
      float o[8];
      if (p.w2) {
        nout = p.Cout2;
        for (int j = 0; j < p.Cout2; ++j) {
          float a = p.b2 ? p.b2[j] : 0.f;
          for (int co = 0; co < p.Cout; ++co) a += p.w2[j * p.Cout + co] * v[co];
          o[j] = a;
        }
      } else {
#pragma unroll
        for (int co = 0; co < 8; ++co) o[co] = v[co];
      }
      const size_t hw = static_cast<size_t>(p.H) * p.W;
      const size_t base = static_cast<size_t>(b) * nout * hw + static_cast<size_t>(y) * p.W + x;
      for (int j = 0; j < nout; ++j) {
        if (p.out) p.out[base + j * hw] = o[j];
        if (p.prev) {
          float v = p.coef[0] * p.sample[base + j * hw] + p.coef[1] * o[j];
          if (p.seed) {  // DDPMScheduler.step: + sqrt(variance) * noise (sigma = 0 at the last step)
            const float sigma = p.coef[2];
            if (sigma != 0.f)
              v += sigma * philox_normal(p.seed[0], p.seed[1], __float_as_uint(p.coef[3]), base + j * hw);
          }
          p.prev[base + j * hw] = v;
        }
      }
    }
  }
  DFU_TR_END();
}

// ---------------------------------------------------------------------------------------------
// The UNet's conv_out (3x3, Cout = 4) register-blocked: a warp owns FOUR consecutive pixels of a row, a lane owns channel
// quads q = lane, lane + 32, ...; per input row it loads the six columns the four pixels' three horizontal taps touch
// (6 x LDG.128) and the 12 weight quads (LDS.128 from the CTA's shared copy) for 192 FMAs -- the one-pixel-per-warp
// kernel above issues a load per 4-16 FMAs and ran at 8 % of the fp32 peak (125 us at batch 8).  The 16 per-lane partial
// sums (4 pixels x 4 outputs) are reduced across the warp by halving (16 shuffles instead of 80).
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// 1x1, few input channels (Cin <= 64): one THREAD per pixel, weights broadcast from shared memory.  The warp-per-pixel
// kernels above leave 24 of 32 lanes idle at Cin = 32 (the channel-selection conv behind the VAE decoder's tensor-core
// conv_out, 262 144 pixels: 256 us); here a pixel is 8 x LDG.128 + Cout x 32 FMA in one thread and the NCHW stores of a
// warp are contiguous.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1x1_pixel_kernel(SmallOut p) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_CONV_OUT);
  __shared__ __align__(16) float sw[8 * 64];
  __shared__ float sb[8];
  for (int i = threadIdx.x; i < p.Cout * p.Cin; i += blockDim.x) sw[i] = __ldg(p.w + i);
  if (threadIdx.x < p.Cout) sb[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
  __syncthreads();
  pdl_wait();
  DFU_TR_MARK(6);
  const int npix = p.B * p.H * p.W;
  const int hw = p.H * p.W;
  const int C4 = p.Cin >> 2;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const float4* xr = reinterpret_cast<const float4*>(p.x + static_cast<size_t>(pix) * p.Cin);
    float acc[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) acc[co] = 0.f;
    for (int q = 0; q < C4; ++q) {
      const float4 v = xr[q];
#pragma unroll
      for (int co = 0; co < 8; ++co) {
        if (co < p.Cout) {
          const float4 wv = *reinterpret_cast<const float4*>(sw + co * p.Cin + 4 * q);
          acc[co] += wv.x * v.x + wv.y * v.y + wv.z * v.z + wv.w * v.w;
        }
      }
    }
    const int b = pix / hw, r = pix - b * hw;
#pragma unroll
    for (int co = 0; co < 8; ++co)
      if (co < p.Cout) p.out[(static_cast<size_t>(b) * p.Cout + co) * hw + r] = acc[co] + sb[co];
  }
  DFU_TR_END();
}

template <int PX>  // pixels of a row per warp
__global__ void __launch_bounds__(256, 2) conv_out4_kernel(SmallOut p) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_CONV_OUT);
  extern __shared__ __align__(16) float sw[];  // [4][9][Cin]
  const int C4 = p.Cin >> 2;
  for (int i = threadIdx.x; i < 36 * C4; i += blockDim.x)
    reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(p.w) + i);
  __syncthreads();
  pdl_wait();
  DFU_TR_MARK(6);
  const int lane = threadIdx.x & 31;
  const int W4 = p.W / PX;
  const int ngroups = p.B * p.H * W4;
  constexpr int NP = PX * 4;  // partial sums per lane
  const int nwarps = static_cast<int>((gridDim.x * blockDim.x) >> 5);
  const size_t hw = static_cast<size_t>(p.H) * p.W;
  for (int grp = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5); grp < ngroups; grp += nwarps) {
    const int x0 = (grp % W4) * PX;
    const int yb = grp / W4;
    const int y = yb % p.H;
    const int b = yb / p.H;
    float acc[NP];  // [pixel][output]
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = 0.f;
    for (int q = lane; q < C4; q += 32) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= p.H) continue;  // (warp-uniform)
        const float4* row = reinterpret_cast<const float4*>(p.x + (static_cast<size_t>(b) * p.H + iy) * p.W * p.Cin) + q;
        float4 in[PX + 2];
#pragma unroll
        for (int c = 0; c < PX + 2; ++c) {
          const int ix = x0 - 1 + c;
          in[c] = (ix >= 0 && ix < p.W) ? row[static_cast<size_t>(ix) * C4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int co = 0; co < 4; ++co) {
            const float4 wv = *(reinterpret_cast<const float4*>(sw + (co * 9 + ky * 3 + kx) * p.Cin) + q);
#pragma unroll
            for (int px = 0; px < PX; ++px) {
              const float4 v = in[px + kx];
              acc[px * 4 + co] += wv.x * v.x + wv.y * v.y + wv.z * v.z + wv.w * v.w;
            }
          }
        }
      }
    }
    // reduction by halving: each step with mask m = 16, 8, ... halves the number of sums a lane carries, until it holds
    // ONE of the NP sums (index = its top log2(NP) lane bits) over part of the warp; the remaining masks are plain adds
    int m = 16;
#pragma unroll
    for (int n = NP / 2; n >= 1; n >>= 1, m >>= 1) {
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const float keep = up ? acc[i + n] : acc[i];
        const float send = up ? acc[i] : acc[i + n];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
    float total = acc[0];
#pragma unroll
    for (; m >= 1; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
    constexpr int kShift = PX == 4 ? 1 : 2;  // 5 - log2(NP)
    if ((lane & ((1 << kShift) - 1)) == 0) {
      const int idx = lane >> kShift;
      const int px = idx >> 2, co = idx & 3;
      const float o = total + (p.bias ? p.bias[co] : 0.f);
      const size_t base = (static_cast<size_t>(b) * 4 + co) * hw + static_cast<size_t>(y) * p.W + x0 + px;
      if (p.out) p.out[base] = o;
      if (p.prev) {
        float v = p.coef[0] * p.sample[base] + p.coef[1] * o;
        if (p.seed) {
          const float sigma = p.coef[2];
          if (sigma != 0.f) v += sigma * philox_normal(p.seed[0], p.seed[1], __float_as_uint(p.coef[3]), base);
        }
        p.prev[base] = v;
      }
    }
  }
  DFU_TR_END();
}

// ---------------------------------------------------------------------------------------------
// elementwise: y = a*x + b*e (+ c*n)   (DDIM eta=0: c = 0;  DDPM: x0/x form pre-collapsed on the host)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
axpbypcz_kernel(const float* __restrict__ x, const float* __restrict__ e, const float* __restrict__ n, float a, float b,
                float c, float* __restrict__ y, long long total) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_MISC);
  pdl_wait();
  DFU_TR_MARK(6);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = a * x[i] + b * e[i];
    if (n) v += c * n[i];
    y[i] = v;
  }
}

// general scheduler update:  y = p0 * clamp?(a0*x + a1*m, -1, 1) + d0*x + d1*m + s*n
// (DDIM / DDPM step for epsilon-, v- and sample-prediction with optional clip_sample)
__global__ void __launch_bounds__(256)
sched_step_kernel(const float* __restrict__ x, const float* __restrict__ m, const float* __restrict__ n, float a0,
                  float a1, float p0, float d0, float d1, float sn, int clip, float* __restrict__ y,
                  float* __restrict__ x0_out, long long total) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_MISC);
  pdl_wait();
  DFU_TR_MARK(6);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float xv = x[i], mv = m[i];
    float x0 = a0 * xv + a1 * mv;
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float v = p0 * x0 + d0 * xv + d1 * mv;
    if (n) v += sn * n[i];
    y[i] = v;
    if (x0_out) x0_out[i] = x0;
  }
}

// y[b, i] = ca[b] * x[b, i] + cb[b] * e[b, i]   (add_noise / get_velocity with per-sample timesteps)
__global__ void __launch_bounds__(256)
axpby_rows_kernel(const float* __restrict__ x, const float* __restrict__ e, const float* __restrict__ ca,
                  const float* __restrict__ cb, float* __restrict__ y, long long per_row, long long total) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_MISC);
  pdl_wait();
  DFU_TR_MARK(6);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / per_row;
    y[i] = ca[b] * x[i] + cb[b] * e[i];
  }
}

// DiagonalGaussian sample from NCHW moments [B, 2*Cz, h, w]: z = (mean + exp(0.5*clamp(logvar,-30,20)) * eps) * scale
__global__ void __launch_bounds__(256)
gaussian_sample_kernel(const float* __restrict__ moments, const float* __restrict__ eps, int B, int Cz, int HW,
                       float scale, float* __restrict__ z) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_MISC);
  pdl_wait();
  DFU_TR_MARK(6);
  const long long total = static_cast<long long>(B) * Cz * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / (static_cast<long long>(Cz) * HW);
    const long long r = i % (static_cast<long long>(Cz) * HW);
    const float mean = moments[b * 2 * Cz * HW + r];
    float v = mean;
    if (eps) {
      float lv = moments[b * 2 * Cz * HW + static_cast<long long>(Cz) * HW + r];
      lv = fminf(fmaxf(lv, -30.f), 20.f);
      v += expf(0.5f * lv) * eps[i];
    }
    z[i] = v * scale;
  }
}

// ---------------------------------------------------------------------------------------------
// row softmax: fp32 scores [rows, n] (scaled by `scale`) -> fp16 operand planes [planes][rows][ldp]
// (single-head d=512 VAE attention; the UNet uses the fused flash kernel in attn.cu)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, int rows, int n, int lds, float scale, __half* __restrict__ p16,
                    int ldp, int planes, long long plane_stride) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_MISC);
  pdl_wait();
  DFU_TR_MARK(6);
  const int row = blockIdx.x;
  if (row >= rows) return;
  __shared__ float red[32];
  const float* sr = s + static_cast<size_t>(row) * lds;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, sr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sum += expf((sr[i] - mx) * scale);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.0f / sum;
  __half* pr = p16 + static_cast<size_t>(row) * ldp;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = expf((sr[i] - mx) * scale) * inv;
    const __half h = __float2half_rn(v);
    pr[i] = h;
    if (planes > 1) pr[plane_stride + i] = __float2half_rn(v - __half2float(h));
  }
}

// fp16 [planes][rows][cols] -> transposed [planes][cols][rows] (V^T for the VAE attention's P*V GEMM)
__global__ void transpose_f16_kernel(const __half* __restrict__ in, int rows, int cols, int ld_in, long long in_plane,
                                     __half* __restrict__ out, long long out_plane) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_MISC);
  pdl_wait();
  DFU_TR_MARK(6);
  __shared__ __half tile[32][33];
  const __half* src = in + blockIdx.z * in_plane;
  __half* dst = out + blockIdx.z * out_plane;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[static_cast<size_t>(r) * ld_in + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[static_cast<size_t>(c) * rows + r] = tile[threadIdx.x][j];
  }
}

static int ew_grid2(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const int cap = (num_sms() > 0 ? num_sms() : 148) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace dfu

DFU_TRACE_SETTER(dfu_trace_set_misc)

using namespace dfu;

extern "C" {

int dfu_timestep_embedding(const float* t, int B, int dim, int flip_sin_to_cos, float freq_shift, float* out,
                           void* stream) {
  DFU_REQUIRE(B > 0 && dim > 0 && dim % 2 == 0, "timestep_embedding: B=%d dim=%d", B, dim);
  DFU_CHECK_CUDA(launch_k(timestep_embed_kernel, dim3(ew_grid2(static_cast<long long>(B) * dim / 2, 128)), dim3(128), 0, static_cast<cudaStream_t>(stream), t, B, dim, flip_sin_to_cos, freq_shift, out));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_gemv(const float* x, int B, int K, int ldx, const float* W, const float* bias, int N, int silu_in,
             int silu_out, float* out, int ldo, void* stream) {
  DFU_REQUIRE(B >= 1 && B <= kGemvMaxB, "gemv: B=%d (max %d)", B, kGemvMaxB);
  DFU_REQUIRE(K % 4 == 0, "gemv: K=%d", K);
  const size_t smem = static_cast<size_t>(B) * K * sizeof(float);
  DFU_REQUIRE(smem <= 96 * 1024, "gemv: B*K too large");
  if (first_use_on_device(ONCE_GEMV_ATTR)) {
    DFU_CHECK_CUDA(cudaFuncSetAttribute(gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  int blocks = (N + 7) / 8;
  const int cap = (num_sms() > 0 ? num_sms() : 148) * 4;
  if (blocks > cap) blocks = cap;
  DFU_CHECK_CUDA(launch_k(gemv_kernel, dim3(blocks), dim3(256), smem, static_cast<cudaStream_t>(stream), x, B, K, ldx, W, bias, N, silu_in, silu_out, out, ldo));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_conv_small_in(const float* src0, int c0, int64_t bstride0, const float* src1, int c1, int64_t bstride1,
                      const float* src2, int c2, int64_t bstride2, int nhwc, int B, int H, int W, int ksz,
                      const float* w, const float* bias, int Cout, float pre_scale, float* out, void* stream) {
  const int Cin = c0 + c1 + c2;
  DFU_REQUIRE(src0 && Cin > 0 && Cin <= 16 && (ksz == 1 || ksz == 3), "conv_small_in: Cin=%d ksz=%d", Cin, ksz);
  SmallInSrc s;
  s.p[0] = src0; s.p[1] = src1; s.p[2] = src2;
  s.c[0] = c0; s.c[1] = c1; s.c[2] = c2;
  s.bstride[0] = bstride0; s.bstride[1] = bstride1; s.bstride[2] = bstride2;
  s.nhwc = nhwc;
  const long long npix = static_cast<long long>(B) * H * W;
  DFU_REQUIRE(npix < (1LL << 31), "conv_small_in: %lld pixels exceed the 32-bit index range", npix);
  static const int pix_env = getenv("DFU_CONV_IN_PIX") ? atoi(getenv("DFU_CONV_IN_PIX")) : 0;
  const int sms_ = num_sms() > 0 ? num_sms() : 148;
  const int pix_cta = pix_env ? pix_env : ((npix + 31) / 32 < 2LL * sms_ ? 16 : 32);
  const int blocks = static_cast<int>((npix + pix_cta - 1) / pix_cta);
  const size_t smem = static_cast<size_t>(pix_cta) * Cin * ksz * ksz * sizeof(float);
  DFU_REQUIRE(Cout >= 1 && Cout <= 512, "conv_small_in: Cout=%d (one thread per output channel, max 512)", Cout);
  const int threads = ((Cout + 31) / 32) * 32;
  DFU_CHECK_CUDA(launch_k(pix_cta == 16 ? conv_small_in_kernel<16> : conv_small_in_kernel<32>, dim3(blocks), dim3(threads), smem,
                          static_cast<cudaStream_t>(stream), s, B, H, W, Cin, ksz, w, bias, Cout, pre_scale, out));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_conv_small_out(const float* x, int B, int H, int W, int Cin, int ksz, const float* w, const float* bias,
                       int Cout, const float* w2, const float* b2, int Cout2, float* out, const float* sample,
                       float* prev, const float* coef, const uint32_t* seed, void* stream) {
  DFU_REQUIRE(Cout >= 1 && Cout <= 8 && Cin % 4 == 0 && (ksz == 1 || ksz == 3), "conv_small_out: Cout=%d Cin=%d", Cout,
              Cin);
  DFU_REQUIRE(!w2 || (Cout2 >= 1 && Cout2 <= 8), "conv_small_out: Cout2=%d", Cout2);
  DFU_REQUIRE(out || prev, "conv_small_out: no output");
  DFU_REQUIRE(!prev || (sample && coef), "conv_small_out: fused step needs sample and coef");
  SmallOut p;
  p.x = x; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.ksz = ksz; p.Cout = Cout;
  p.w = w; p.bias = bias; p.w2 = w2; p.b2 = b2; p.Cout2 = Cout2; p.out = out;
  p.sample = sample; p.prev = prev; p.coef = coef; p.seed = seed;
  const long long npix = static_cast<long long>(B) * H * W;
  const size_t wbytes = static_cast<size_t>(Cout) * ksz * ksz * Cin * sizeof(float);
  // staging the weights pays only when a CTA then walks many pixels (VAE maps); the 64x64 UNet output conv keeps
  // one warp per pixel reading the weights through L1 (measured 11 us vs 19 us with the 46 KB copy per CTA)
  static const int force_smem = getenv("DFU_CONV_OUT_SMEM") ? atoi(getenv("DFU_CONV_OUT_SMEM")) : -1;  // experiments
  const long long smem_pix = getenv("DFU_CONV_OUT_SMEM_PIX") ? atoll(getenv("DFU_CONV_OUT_SMEM_PIX")) : 64LL * 8;
  int w_in_smem = wbytes <= 200 * 1024 && npix >= smem_pix * (num_sms() > 0 ? num_sms() : 148);
  if (force_smem >= 0 && wbytes <= 200 * 1024) w_in_smem = force_smem;
  if (first_use_on_device(ONCE_CONV_OUT_ATTR)) {
    DFU_CHECK_CUDA(cudaFuncSetAttribute(conv_small_out_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFU_CHECK_CUDA(cudaFuncSetAttribute(conv_out4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  DFU_REQUIRE(npix < (1LL << 31), "conv_small_out: %lld pixels exceed the 32-bit index range", npix);
  if (ksz == 1 && Cin <= 64 && !w2 && !prev && out && npix >= 65536) {  // large maps only: tiny ones are launch-bound anyway
    long long nb = (npix + 255) / 256;
    const long long capp = 8LL * (num_sms() > 0 ? num_sms() : 148);
    if (nb > capp) nb = capp;
    DFU_CHECK_CUDA(launch_k(conv1x1_pixel_kernel, dim3(static_cast<unsigned>(nb)), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
    DFU_CHECK_CUDA(cudaGetLastError());
    return DFU_OK;
  }
  static const bool fast4 = !(getenv("DFU_CONV_OUT4") && getenv("DFU_CONV_OUT4")[0] == '0');
  if (fast4 && ksz == 3 && Cout == 4 && !w2 && W % 4 == 0 && wbytes <= 100 * 1024) {
    // register-blocked UNet conv_out: one warp per four pixels, at most two CTAs (weights in shared memory) per SM
    // (two pixels per warp for small maps was measured: 12.4 vs 12.6 us at batch 1, slower from batch 2 on)
    const long long cap4 = 2LL * (num_sms() > 0 ? num_sms() : 148);
    const long long groups = npix / 4;
    long long nb = (groups + 7) / 8;
    if (nb > cap4) nb = cap4;
    DFU_CHECK_CUDA(launch_k(conv_out4_kernel<4>, dim3(static_cast<unsigned>(nb)), dim3(256), wbytes,
                            static_cast<cudaStream_t>(stream), p));
    DFU_CHECK_CUDA(cudaGetLastError());
    return DFU_OK;
  }
  // eight pixels per CTA pass; enough CTAs for every SM, few enough that the weight copy is amortised over many pixels
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const long long per_sm = wbytes > 100 * 1024 ? 1 : 2;
  long long blocks = (npix + 7) / 8;
  if (w_in_smem && blocks > sms * per_sm) blocks = sms * per_sm;
  DFU_REQUIRE(npix < (1LL << 31), "conv_small_out: %lld pixels exceed the 32-bit index range", npix);
  if (w_in_smem)
    DFU_CHECK_CUDA(launch_k(conv_small_out_kernel<true>, dim3(static_cast<unsigned>(blocks)), dim3(256), wbytes, static_cast<cudaStream_t>(stream), p));
  else
    DFU_CHECK_CUDA(launch_k(conv_small_out_kernel<false>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_philox_normal(uint64_t seed, uint32_t step, int64_t n, float* out, uint32_t* bits, void* stream) {
  DFU_REQUIRE(n > 0 && (out || bits), "philox_normal: nothing to write");
  philox_normal_kernel<<<ew_grid2(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), step, n, out, bits);
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_axpbypcz(const float* x, const float* e, const float* n, float a, float b, float c, float* y, int64_t total,
                 void* stream) {
  DFU_CHECK_CUDA(launch_k(axpbypcz_kernel, dim3(ew_grid2(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, e, n, a, b, c, y, total));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_scheduler_step(const float* x, const float* m, const float* n, float a0, float a1, float p0, float d0,
                       float d1, float sn, int clip, float* y, float* x0_out, int64_t total, void* stream) {
  DFU_CHECK_CUDA(launch_k(sched_step_kernel, dim3(ew_grid2(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, m, n, a0, a1, p0, d0, d1, sn, clip, y, x0_out, total));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_axpby_rows(const float* x, const float* e, const float* ca, const float* cb, float* y, int B,
                   int64_t per_row, void* stream) {
  const long long total = static_cast<long long>(B) * per_row;
  DFU_CHECK_CUDA(launch_k(axpby_rows_kernel, dim3(ew_grid2(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, e, ca, cb, y, per_row, total));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_gaussian_sample(const float* moments, const float* eps, int B, int Cz, int HW, float scale, float* z,
                        void* stream) {
  DFU_CHECK_CUDA(launch_k(gaussian_sample_kernel, dim3(ew_grid2(static_cast<long long>(B) * Cz * HW, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), moments, eps, B, Cz, HW, scale, z));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_softmax_rows(const float* s, int rows, int n, int lds, float scale, void* p16, int ldp, int planes,
                     int64_t plane_stride, void* stream) {
  DFU_CHECK_CUDA(launch_k(softmax_rows_kernel, dim3(rows), dim3(256), 0, static_cast<cudaStream_t>(stream), s, rows, n, lds, scale, static_cast<__half*>(p16), ldp, planes, plane_stride));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

int dfu_transpose_f16(const void* in, int planes, int rows, int cols, int ld_in, int64_t in_plane, void* out,
                      int64_t out_plane, void* stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, planes), block(32, 8);
  DFU_CHECK_CUDA(launch_k(transpose_f16_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream), static_cast<const __half*>(in), rows, cols, ld_in, in_plane, static_cast<__half*>(out), out_plane));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}
}

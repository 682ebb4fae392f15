// diffute_b200 — tcgen05 contraction core (linear / 1x1 / 3x3 implicit-GEMM conv) for sm_100a.
//
// One CTA computes one 128 x block_n output tile over a K-range (split-K slice):
//   warp 0   : TMA producer  (cp.async.bulk.tensor 2-D for matrices / weights, 4-D NHWC boxes for conv taps;
//              the halo and the stride-2 / asymmetric padding come from TMA out-of-bounds zero fill)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, fp32 accumulate in TMEM)
//   warps 2-9: epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks): tcgen05.ld TMEM ->
//              registers -> smem transpose -> fused bias / time-embedding / residual / GEGLU / fp16-split stores,
//              coalesced 128-bit wide.
// Operands live in shared memory in the 128-byte-swizzled K-major layout that TMA writes and the UMMA
// descriptors read; a `stages`-deep full/empty mbarrier ring connects producer and issuer.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_shared.cuh"

namespace dfu {


// split-K second stage: every thread owns one output quad, sums the `splits` fp32 partials in slice order
// (deterministic), four independent 16-byte loads in flight at a time, then runs the fused epilogue.
__device__ __forceinline__ float4 sum_partials(const float* __restrict__ src, size_t plane, int splits) {
  // four independent 16-byte loads in flight per thread, summed in slice order.  (Eight in flight was measured slower:
  // the register cost halves the occupancy of this latency-bound kernel — 231 vs 206 us per UNet step.)
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    const float4 t0 = __ldcg(reinterpret_cast<const float4*>(src + (s + 0) * plane));
    const float4 t1 = __ldcg(reinterpret_cast<const float4*>(src + (s + 1) * plane));
    const float4 t2 = __ldcg(reinterpret_cast<const float4*>(src + (s + 2) * plane));
    const float4 t3 = __ldcg(reinterpret_cast<const float4*>(src + (s + 3) * plane));
    v.x = (((v.x + t0.x) + t1.x) + t2.x) + t3.x;
    v.y = (((v.y + t0.y) + t1.y) + t2.y) + t3.y;
    v.z = (((v.z + t0.z) + t1.z) + t2.z) + t3.z;
    v.w = (((v.w + t0.w) + t1.w) + t2.w) + t3.w;
  }
  for (; s < splits; ++s) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(src + s * plane));
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  return v;
}

// one output quad of the split-K second stage: sum the partials in slice order, then the fused epilogue
__device__ __forceinline__ void reduce_quad(const float* __restrict__ ws, int splits, const EpiParams& e, long long idx) {
  const bool geglu = e.epi == DFU_EPI_GEGLU;
  const int qpr = geglu ? e.N / 8 : e.N / 4;
  const size_t plane = static_cast<size_t>(e.M) * e.N;
  const int m = static_cast<int>(idx / qpr);
  const int qi = static_cast<int>(idx % qpr);
  if (geglu) {
    const int n_a = (qi >> 2) * 32 + (qi & 3) * 4;
    const float* src = ws + static_cast<size_t>(m) * e.N + n_a;
    epi_geglu_quad(e, m, n_a, sum_partials(src, plane, splits), sum_partials(src + 16, plane, splits));
  } else {
    const int n = qi * 4;
    epi_quad(e, m, n, sum_partials(ws + static_cast<size_t>(m) * e.N + n, plane, splits));
  }
}

// ---------------------------------------------------------------------------------------------
// (A push variant — every CTA storing its rows into the owner's shared memory with st.shared::cluster, local reduction
// afterwards — was built and measured in the captured step: 21.9 vs 19.0 us for the level-0 3x3 convs; distributed
// shared memory moves ~5-20 B/clk per SM in either direction, so the pull form with S loads in flight stays.)
// cluster split-K second stage (epilogue threads of one CTA): rows [split*128/S, (split+1)*128/S) of the tile, summed
// over the S partial tiles parked in the cluster's shared memories.  All S remote loads of a quad (and its residual)
// are in flight before the first add: the loop costs DSMEM bandwidth, not S serial ~200-cycle round trips.
// ---------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void cluster_reduce(const GemmKernelParams& p, const uint8_t* smem, int split, int n_tile0,
                                               int m0, int x0, int y0, int img0, int m_end) {
  // Only the rows the tile really has are reduced, split evenly over the S CTAs (a 64-pixel map fills half a tile: with
  // a fixed 128 / S rows per CTA half of the cluster would idle through the reduction).
  const int tile_rows = p.conv ? p.bw * p.bh * p.bn : min(kBlockM, m_end - m0);
  const int rows_per = (tile_rows + S - 1) / S;
  constexpr int CH = S < 8 ? S : 8;  // remote loads in flight per thread (registers: 16 in flight would spill)
  const bool geglu = p.e.epi == DFU_EPI_GEGLU;
  const int qpr = geglu ? p.block_n / 8 : p.block_n / 4;
  const int ldp = p.block_n + 4;
  const uint32_t part0 = smem_u32(smem);
  uint32_t peer[S];
#pragma unroll
  for (int q = 0; q < S; ++q) peer[q] = dsmem_addr(part0, static_cast<uint32_t>(q));
  const int te = threadIdx.x - 64;
  for (int idx = te; idx < rows_per * qpr; idx += kEpiThreads) {
    const int rr = split * rows_per + idx / qpr;
    if (rr >= tile_rows) break;
    const int qi = idx % qpr;
    const int col = geglu ? (qi >> 2) * 32 + (qi & 3) * 4 : qi * 4;
    int m;
    bool valid;
    if (p.conv) {
      const int ix = rr % p.bw;
      const int t = rr / p.bw;
      const int iy = t % p.bh;
      const int in = t / p.bh;
      const int x = x0 + ix, y = y0 + iy, img = img0 + in;
      valid = (in < p.bn) && (x < p.W) && (y < p.H) && (img < p.B);
      m = (img * p.H + y) * p.W + x;
    } else {
      m = m0 + rr;
      valid = m < m_end;
    }
    if (!valid) continue;
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!geglu && p.e.residual)
      rs = *reinterpret_cast<const float4*>(p.e.residual + static_cast<size_t>(m) * p.e.ldr + n_tile0 + col);
    const uint32_t off = static_cast<uint32_t>(rr * ldp + col) * 4u;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), g = a;
#pragma unroll
    for (int q0 = 0; q0 < S; q0 += CH) {  // slice order => deterministic
      float4 t[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) t[q] = ld_dsmem_f4(peer[q0 + q] + off);
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        a.x += t[q].x; a.y += t[q].y; a.z += t[q].z; a.w += t[q].w;
      }
    }
    if (geglu) {
#pragma unroll
      for (int q0 = 0; q0 < S; q0 += CH) {
        float4 u[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) u[q] = ld_dsmem_f4(peer[q0 + q] + off + 64);
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          g.x += u[q].x; g.y += u[q].y; g.z += u[q].z; g.w += u[q].w;
        }
      }
      epi_geglu_quad(p.e, m, n_tile0 + col, a, g);
    } else {
      epi_quad_res(p.e, m, n_tile0 + col, a, rs);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ GemmKernelParams p) {
  pdl_trigger();
  DFU_TR_SHARED_DECL();
  DFU_TR_SHARED_BEGIN(TR_GEMM | (p.splits << 8) | (p.cluster << 16) | (p.block_n << 20));
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[256];  // bias[n_tile0 .. n_tile0 + block_n)

  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // FP16X2 (npass = 3): a stage holds the hi AND lo tiles of both operands, fetched once, and the three products
  // hi*hi + lo*hi + hi*lo are issued from it — 4 tile loads per k-block instead of 6 (these contractions are bound by
  // L2 -> shared-memory traffic: arithmetic intensity x1.5)
  const uint32_t nplane = p.npass == 3 ? 2u : 1u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * 128u;
  const uint32_t stage_bytes = nplane * (kABytes + b_bytes);
  // (through a shuffle: provably warp-uniform, so the producer / issuer loops below are uniform control flow and ptxas
  // keeps TMA / MMA operands in uniform registers instead of an ELECT + R2UR.BROADCAST waterfall per instruction)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates -------------------------------------------------------------------
  int bid = blockIdx.x;
  int tn, tm, split;
  if (p.cluster > 1) {  // the K-slices of one tile are consecutive CTAs = one cluster
    split = bid % p.cluster;
    bid /= p.cluster;
    tn = bid % p.tiles_n;
    tm = bid / p.tiles_n;
  } else {
    tn = bid % p.tiles_n;
    bid /= p.tiles_n;
    tm = bid % p.tiles_m;
    split = bid / p.tiles_m;
  }
  const int kb0 = static_cast<int>(static_cast<long long>(p.total_kb) * split / p.splits);
  const int kb1 = static_cast<int>(static_cast<long long>(p.total_kb) * (split + 1) / p.splits);
  const int n_tile0 = tn * p.block_n;
  int m0, a_row0, b_off, m_end, x0 = 0, y0 = 0, img0 = 0;
  batch_coords(p.bt, tm, p.e.M, m0, a_row0, b_off, m_end);
  if (p.conv) {
    int t = tm;
    x0 = (t % p.tiles_x) * p.bw;
    t /= p.tiles_x;
    y0 = (t % p.tiles_y) * p.bh;
    img0 = (t / p.tiles_y) * p.bn;
  }

  // ---- one-time setup ---------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (p.ngroups > 1) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) DFU_TR_SHARED_MARK(5);
  // the prologue above overlapped the previous kernel; every role waits for it (griddepcontrol.wait) right before its
  // first access to data that kernel produced — the TMA producer only after it has requested the weight tiles

  // k-block -> loads (shared by the producer warps)
  const int kbg0 = p.g[0].kb_per_pass;
  // k-block -> (group, tap, channel chunk) and the loads of that k-block (hi [+ lo] tile of each operand)
  auto issue = [&](int kb, int stage, bool load_a, bool load_b) {
    int r = kb, gi = 0;
    if (r >= kbg0) {
      r -= kbg0;
      gi = 1;
    }
    const GroupDev& G = p.g[gi];
    const int tap = r / G.nchunks;
    const int chunk = r - tap * G.nchunks;
    const CUtensorMap* mA = gi ? &tmA1 : &tmA0;
    const CUtensorMap* mB = gi ? &tmB1 : &tmB0;
    uint8_t* sA = smem + stage * stage_bytes;
    uint8_t* sB = sA + nplane * kABytes;
    if (load_b) {  // the k-block's first load also arms the barrier with the bytes of ALL its tiles
      mbar_arrive_expect_tx(&full_bar[stage], nplane * (p.a_tx_bytes[gi] + p.b_tx_bytes));
      for (uint32_t pl = 0; pl < nplane; ++pl)
        tma_load_2d(sB + pl * b_bytes, mB, &full_bar[stage], (tap * G.nchunks + chunk) * kBlockK,
                    n_tile0 + b_off + static_cast<int>(pl) * G.b_plane);
    }
    if (load_a) {
      for (uint32_t pl = 0; pl < nplane; ++pl) {
        const int a_sel = static_cast<int>(pl) * G.a_plane;
        if (G.a_mode == 0) {
          tma_load_2d(sA + pl * kABytes, mA, &full_bar[stage], (tap * G.nchunks + chunk) * kBlockK, a_row0 + a_sel);
        } else {
          tma_load_4d(sA + pl * kABytes, mA, &full_bar[stage], chunk * kBlockK, x0 + G.dx[tap], y0 + G.dy[tap],
                      img0 + G.dn[tap] + a_sel);
        }
      }
    }
  };
  const bool pre = p.g[0].b_static && (p.ngroups == 1 || p.g[1].b_static);
  const int npre = pre ? min(p.stages, kb1 - kb0) : 0;
  // two producer lanes (p.two_prod): warp 0 streams the weight tiles (and arms the barriers), the first epilogue warp —
  // idle during the main loop — streams the activation tiles: a single issuing lane sustains only ~55 GB/s of TMA
  // traffic (~1 box row per 3.5 cycles), two lanes run in parallel
  const bool two_prod = p.two_prod != 0;

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the k-blocks, one elected lane issues =============
    {
      // Weights do not depend on the previous kernel: the B tiles of the first ring-full of k-blocks are requested
      // BEFORE griddepcontrol.wait, so their HBM latency (and, for the weight-streaming deep levels, a good part of
      // the streaming itself) overlaps the producer kernel's tail.
      if (elect_one())
        for (int i = 0; i < npre; ++i) issue(kb0 + i, i, false, true);
      __syncwarp();
      pdl_wait();  // activations (A) are valid from here on
      if (lane == 0) DFU_TR_SHARED_MARK(6);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (kb - kb0 < npre) {
          if (!two_prod && elect_one()) issue(kb, stage, true, false);
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (elect_one()) issue(kb, stage, !two_prod, true);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp in uniform control flow, one elected lane issues ============
    {
      const uint32_t idesc = umma_idesc_f16(kBlockM, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);  // (probing back to back instead of backing off: measured, no difference)
        tc_fence_after();
        if (lane == 0 && kb == kb0) DFU_TR_SHARED_MARK(7);
        const uint32_t sA = smem_u32(smem + stage * stage_bytes);
        const uint32_t sB = sA + nplane * kABytes;
        const int nprod = p.npass == 3 ? 3 : 1;
        if (elect_one()) {
          for (int ps = 0; ps < nprod; ++ps) {  // hi*hi [, lo*hi, hi*lo] from the same stage
            const uint64_t adesc = umma_desc_sw128(sA + (ps == 1 ? kABytes : 0u));
            const uint64_t bdesc = umma_desc_sw128(sB + (ps == 2 ? b_bytes : 0u));
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the >>4 address field
              umma_f16_ss(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                          (kb > kb0 || ps > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees this smem slot when the MMAs above have read it
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma_commit(&tmem_full_bar);  // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue warps (2..9) =============================================================
    const int q = warp & 3;         // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;  // which half of the 32-column chunks this warp handles
    const int r = q * 32 + lane;
    int m;
    bool valid;
    if (p.conv) {
      const int ix = r % p.bw;
      const int t = r / p.bw;
      const int iy = t % p.bh;
      const int in = t / p.bh;
      const int x = x0 + ix, y = y0 + iy, img = img0 + in;
      valid = (in < p.bn) && (x < p.W) && (y < p.H) && (img < p.B);
      m = (img * p.H + y) * p.W + x;
    } else {
      m = m0 + r;
      valid = m < m_end;
    }
    // ---- everything that does not need the accumulator happens while the main loop runs ----
    const bool direct = p.splits == 1;
    const bool geglu = direct && p.e.epi == DFU_EPI_GEGLU;
    const bool has_res = direct && !geglu && p.e.residual != nullptr;
    // the tile's bias slice is a weight: copied to shared memory before the wait (its first touch is an HBM miss)
    if (warp == 2 && p.e.bias != nullptr && lane * 8 < p.block_n) {
      *reinterpret_cast<float4*>(s_bias + lane * 8) = __ldg(reinterpret_cast<const float4*>(p.e.bias + n_tile0 + lane * 8));
      *reinterpret_cast<float4*>(s_bias + lane * 8 + 4) = __ldg(reinterpret_cast<const float4*>(p.e.bias + n_tile0 + lane * 8 + 4));
    }
    // the 32 rows of this warp are the same for every column chunk: (m, valid) of the rows this lane stores, once
    const int vmask = valid ? 1 : 0;
    int mrs[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = geglu ? (it & 3) * 8 + (lane >> 2) : it * 4 + (lane >> 3);
      const int mr = __shfl_sync(0xffffffffu, m, row);
      const int vr = __shfl_sync(0xffffffffu, vmask, row);
      mrs[it] = vr ? mr : -1;
    }
    const int cq = lane & 7;
    // sample index of the warp's rows (time-embedding row of the epilogue): uniform in all but tiny feature maps
    int smp0 = 0;
    bool one_sample = true;
    if (p.e.rowvec != nullptr) {
      const int smp = valid ? m / p.e.rows_per_sample : -1;
      smp0 = __reduce_max_sync(0xffffffffu, smp);
      one_sample = __all_sync(0xffffffffu, smp < 0 || smp == smp0);
      if (smp0 < 0) smp0 = 0;
    }
    pdl_wait();  // residual / rowvec are activations, and the outputs may alias buffers earlier kernels still read
    asm volatile("bar.sync 1, 256;" ::: "memory");  // s_bias visible to the eight epilogue warps
    if (warp == 2 && p.e.rowvec != nullptr && lane * 8 < p.block_n)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.e.rowvec + static_cast<size_t>((valid ? m : 0) / p.e.rows_per_sample) * p.e.rowvec_ld + n_tile0 + lane * 8));
    // residual quads are requested one chunk ahead (the first chunk's during the main loop): their L2 latency never
    // sits between the accumulator read and the stores
    // (one register buffer: each quad is re-requested for the next chunk right after it has been consumed)
    float4 res[8];
    auto fetch_res1 = [&](int c, int it) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_res && (c + cq * 4 < p.block_n) && mrs[it] >= 0)
        t = *reinterpret_cast<const float4*>(p.e.residual + static_cast<size_t>(mrs[it]) * p.e.ldr + n_tile0 + c + cq * 4);
      return t;
    };
#pragma unroll
    for (int it = 0; it < 8; ++it) res[it] = fetch_res1(cg * 32, it);
    if (two_prod && warp == 2) {  // activation-tile producer (see above); done long before the accumulator is
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (kb - kb0 >= npre) mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) issue(kb, stage, true, false);
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    mbar_wait_sleep(&tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) DFU_TR_SHARED_MARK(8);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // TMEM gives each thread one output row; a per-warp smem transpose turns that into row-contiguous 16-byte
    // quads per lane so global traffic is coalesced (4 full 128-byte lines per warp instruction).
    // (the operand ring is idle once tmem_full has fired, so its first 36 KiB double as the staging area)
    float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * kStageFloats;
#pragma unroll 1  // (one copy of the epilogue body: unrolled x4 the kernel outgrew the instruction cache)
    for (int c = cg * 32; c < p.block_n; c += 64) {
      uint32_t raw[32];
      tmem_ld32(taddr + static_cast<uint32_t>(c), raw);
      tmem_ld_wait();
      if (threadIdx.x == 64 && c == 0) DFU_TR_SHARED_MARK(12);
      const int ncol = p.block_n - c;  // columns of this chunk inside the tile (block_n may end mid-chunk)
      if (p.cluster > 1) {
        // cluster split-K: park this slice's fp32 partial tile in OWN shared memory (row r at r * (block_n + 4))
        float* prow = reinterpret_cast<float*>(smem) + static_cast<size_t>(r) * (p.block_n + 4) + c;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (4 * i < ncol)
            *reinterpret_cast<float4*>(prow + 4 * i) =
                make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                            __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
        continue;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stage + lane * kStageLd + 4 * i) =
            make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]), __uint_as_float(raw[4 * i + 2]),
                        __uint_as_float(raw[4 * i + 3]));
      __syncwarp();
      if (threadIdx.x == 64 && c == 0) DFU_TR_SHARED_MARK(14);
      const int n = n_tile0 + c;
      // per-chunk constants of this lane's column quad(s): bias (+ the time-embedding row when the warp's rows belong
      // to one sample) are added with one FFMA per element; nothing per-row is recomputed inside the loops
      const float alpha = p.e.alpha;
      if (geglu && p.e.rowvec == nullptr) {
        const int cq4 = lane & 3;
        float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
        if (p.e.bias) {
          ba = *reinterpret_cast<const float4*>(s_bias + c + cq4 * 4);
          bg = *reinterpret_cast<const float4*>(s_bias + c + 16 + cq4 * 4);
        }
        const int n_out = ((n + cq4 * 4) >> 5) * 16 + ((n + cq4 * 4) & 15);
        const bool lo_plane = p.e.out_planes > 1;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = it * 8 + (lane >> 2);
          const float4 a = *reinterpret_cast<const float4*>(stage + row * kStageLd + cq4 * 4);
          const float4 g = *reinterpret_cast<const float4*>(stage + row * kStageLd + 16 + cq4 * 4);
          if (mrs[it] >= 0) {
            float4 o;
            o.x = fmaf(a.x, alpha, ba.x) * gelu_erf_f(fmaf(g.x, alpha, bg.x));
            o.y = fmaf(a.y, alpha, ba.y) * gelu_erf_f(fmaf(g.y, alpha, bg.y));
            o.z = fmaf(a.z, alpha, ba.z) * gelu_erf_f(fmaf(g.z, alpha, bg.z));
            o.w = fmaf(a.w, alpha, ba.w) * gelu_erf_f(fmaf(g.w, alpha, bg.w));
            store_f16x4(p.e.out_f16 + static_cast<size_t>(mrs[it]) * p.e.ldh + n_out, o, lo_plane, p.e.out_plane_stride);
          }
        }
      } else if (geglu) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = it * 8 + (lane >> 2), cq4 = lane & 3;
          const float4 a = *reinterpret_cast<const float4*>(stage + row * kStageLd + cq4 * 4);
          const float4 g = *reinterpret_cast<const float4*>(stage + row * kStageLd + 16 + cq4 * 4);
          if (mrs[it] >= 0) epi_geglu_quad(p.e, mrs[it], n + cq4 * 4, a, g, s_bias, n_tile0);
        }
      } else if (!direct) {
        const bool qok = cq * 4 < ncol;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3);
          const float4 v = *reinterpret_cast<const float4*>(stage + row * kStageLd + cq * 4);
          if (qok && mrs[it] >= 0)
            __stcg(reinterpret_cast<float4*>(p.ws + (static_cast<size_t>(split) * p.e.M + mrs[it]) * p.e.N + n + cq * 4), v);
        }
      } else {
        const bool qok = cq * 4 < ncol;
        float4 aq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qok && p.e.bias) aq = *reinterpret_cast<const float4*>(s_bias + c + cq * 4);
        if (qok && p.e.rowvec && one_sample) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.e.rowvec + static_cast<size_t>(smp0) * p.e.rowvec_ld + n + cq * 4));
          aq.x += t.x; aq.y += t.y; aq.z += t.z; aq.w += t.w;
        }
        const bool per_row_vec = p.e.rowvec != nullptr && !one_sample;
        const bool f32_out = p.e.epi == DFU_EPI_F32;
        const bool lo_plane = p.e.out_planes > 1;
        const bool act = p.e.act != 0;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3);
          const float4 v = *reinterpret_cast<const float4*>(stage + row * kStageLd + cq * 4);
          if (qok && mrs[it] >= 0) {
            float4 o;
            o.x = fmaf(v.x, alpha, aq.x) + res[it].x;
            o.y = fmaf(v.y, alpha, aq.y) + res[it].y;
            o.z = fmaf(v.z, alpha, aq.z) + res[it].z;
            o.w = fmaf(v.w, alpha, aq.w) + res[it].w;
            if (per_row_vec) {  // a warp whose rows span two samples (tiny feature maps): the row's own vector
              const float4 t = __ldg(reinterpret_cast<const float4*>(
                  p.e.rowvec + static_cast<size_t>(mrs[it] / p.e.rows_per_sample) * p.e.rowvec_ld + n + cq * 4));
              o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            if (act) {
              o.x = gelu_erf_f(o.x); o.y = gelu_erf_f(o.y); o.z = gelu_erf_f(o.z); o.w = gelu_erf_f(o.w);
            }
            if (f32_out)
              *reinterpret_cast<float4*>(p.e.out_f32 + static_cast<size_t>(mrs[it]) * p.e.ldo + n + cq * 4) = o;
            else
              store_f16x4(p.e.out_f16 + static_cast<size_t>(mrs[it]) * p.e.ldh + n + cq * 4, o, lo_plane, p.e.out_plane_stride);
          }
          if (it == 0 && threadIdx.x == 64 && c == 0) DFU_TR_SHARED_MARK(15);
          if (c + 64 < p.block_n) res[it] = fetch_res1(c + 64, it);
        }
      }
      if (threadIdx.x == 64 && c == 0) DFU_TR_SHARED_MARK(13);
    }
    if (threadIdx.x == 64) DFU_TR_SHARED_MARK(9);
  }

  if (p.cluster > 1) {
    // ---- cluster split-K second stage: every CTA of the cluster reduces 128/S rows of the tile over all S partial
    // tiles (read through distributed shared memory, slice order => deterministic) and runs the fused epilogue ----
    cluster_sync_all();
    if (warp >= 2) {
      switch (p.cluster) {
        case 2: cluster_reduce<2>(p, smem, split, n_tile0, m0, x0, y0, img0, m_end); break;
        case 4: cluster_reduce<4>(p, smem, split, n_tile0, m0, x0, y0, img0, m_end); break;
        case 8: cluster_reduce<8>(p, smem, split, n_tile0, m0, x0, y0, img0, m_end); break;
        default: cluster_reduce<16>(p, smem, split, n_tile0, m0, x0, y0, img0, m_end); break;
      }
    }
    cluster_sync_all();  // nobody leaves while a peer may still read its partial tile
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DFU_TR_SHARED_END();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Small outputs (the deep UNet levels: 64..256 rows x 1280 columns, 10-20 slices): one thread per quad walks its slices in
// ~5 dependent rounds of L2 latency on a nearly empty GPU (4.5 us for 6.5 MB, measured).  Here T = 2/4/8 lanes share a
// quad: lane j loads slices j, j+T, ... (all in flight at once), adds them in slice order and the T partial sums are
// combined by a fixed xor-shuffle tree — one L2 round trip instead of five, still bit-identical run to run.
template <int T>
__device__ __forceinline__ float4 sum_partials_team(const float* __restrict__ src, size_t plane, int splits, int sub) {
  constexpr int kMaxPer = 8;
  float4 t[kMaxPer];
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) {
    const int sl = sub + i * T;
    t[i] = (sl < splits) ? __ldcg(reinterpret_cast<const float4*>(src + static_cast<size_t>(sl) * plane))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 v = t[0];
#pragma unroll
  for (int i = 1; i < kMaxPer; ++i) {
    v.x += t[i].x; v.y += t[i].y; v.z += t[i].z; v.w += t[i].w;
  }
#pragma unroll
  for (int o = T >> 1; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  return v;
}

template <int T>
__global__ void __launch_bounds__(256) splitk_reduce_team_kernel(const float* __restrict__ ws, int splits, EpiParams e) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_SPLITK_REDUCE);
  pdl_wait();
  DFU_TR_MARK(6);
  const bool geglu = e.epi == DFU_EPI_GEGLU;
  const int qpr = geglu ? e.N / 8 : e.N / 4;
  const size_t plane = static_cast<size_t>(e.M) * e.N;
  const long long total = static_cast<long long>(e.M) * qpr;            // quads; T consecutive lanes per quad
  const long long teams = (static_cast<long long>(gridDim.x) * blockDim.x) / T;
  const int sub = threadIdx.x % T;
  for (long long idx = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) / T; idx - (idx % (32 / T)) < total;
       idx += teams) {  // (whole warps iterate together: the shuffles need every lane)
    const bool live = idx < total;
    const long long id = live ? idx : total - 1;
    const int m = static_cast<int>(id / qpr);
    const int qi = static_cast<int>(id % qpr);
    if (geglu) {
      const int n_a = (qi >> 2) * 32 + (qi & 3) * 4;
      const float* src = ws + static_cast<size_t>(m) * e.N + n_a;
      const float4 a = sum_partials_team<T>(src, plane, splits, sub);
      const float4 g = sum_partials_team<T>(src + 16, plane, splits, sub);
      if (live && sub == 0) epi_geglu_quad(e, m, n_a, a, g);
    } else {
      const int n = qi * 4;
      const float4 v = sum_partials_team<T>(ws + static_cast<size_t>(m) * e.N + n, plane, splits, sub);
      if (live && sub == 0) epi_quad(e, m, n, v);
    }
  }
  DFU_TR_END();
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, EpiParams e) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_SPLITK_REDUCE);
  pdl_wait();
  DFU_TR_MARK(6);
  const long long total = static_cast<long long>(e.M) * (e.epi == DFU_EPI_GEGLU ? e.N / 8 : e.N / 4);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x)
    reduce_quad(ws, splits, e, idx);
  DFU_TR_END();
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// shared memory of the pair kernel: ring + 8 x 4.5 KiB epilogue staging + alignment slack
static size_t pair_smem(int stages, size_t stage_bytes) { return stages * stage_bytes + 8 * kStageFloats * 4 + 1024; }

int plan_gemm(const DfuGemm* d, Plan* pl) {
  pl->pair = 0;
  DFU_REQUIRE(d->m > 0 && d->n > 0, "gemm: empty problem m=%d n=%d", d->m, d->n);
  DFU_REQUIRE(d->ngroups == 1 || d->ngroups == 2, "gemm: ngroups must be 1 or 2");
  DFU_REQUIRE(d->npass == 1 || d->npass == 3, "gemm: npass must be 1 or 3");
  DFU_REQUIRE(d->n % 32 == 0, "gemm: n=%d must be a multiple of 32", d->n);
  int total_kb = 0;
  for (int g = 0; g < d->ngroups; ++g) {
    const DfuGemmOperand& o = d->g[g];
    DFU_REQUIRE(o.ntaps >= 1 && o.ntaps <= 9, "gemm: ntaps=%d", o.ntaps);
    DFU_REQUIRE(o.k_per_tap > 0 && o.k_per_tap % kBlockK == 0, "gemm: k_per_tap=%d must be a multiple of 64",
                o.k_per_tap);
    DFU_REQUIRE(o.a_mode == 0 || (o.a_mode == 1 && d->conv), "gemm: image operand needs conv=1");
    DFU_REQUIRE(!(o.a_mode == 0 && o.ntaps != 1), "gemm: matrix operand must have ntaps=1");
    DFU_REQUIRE(!(o.a_mode == 1 && o.a_c != o.k_per_tap), "gemm: a_c must equal k_per_tap");
    total_kb += o.ntaps * (o.k_per_tap / kBlockK);  // npass = 3 shares the k-block's tiles among its three products
  }
  pl->total_kb = total_kb;
  if (d->conv) {
    DFU_REQUIRE(d->m == d->B * d->H * d->W, "gemm: conv m=%d != B*H*W", d->m);
    pl->bw = d->W < 128 ? d->W : 128;
    pl->bh = 1;
    pl->bn = 1;
    if (d->W < 128) {
      int bh = 128 / d->W;
      if (bh > d->H) bh = d->H;
      pl->bh = bh;
      if (bh == d->H) {
        int bn = 128 / (d->H * d->W);
        if (bn < 1) bn = 1;
        if (bn > d->B) bn = d->B;
        pl->bn = bn;
      }
    }
    pl->tiles_x = (d->W + pl->bw - 1) / pl->bw;
    pl->tiles_y = (d->H + pl->bh - 1) / pl->bh;
    pl->tiles_m = pl->tiles_x * pl->tiles_y * ((d->B + pl->bn - 1) / pl->bn);
  } else {
    pl->bw = pl->bh = pl->bn = pl->tiles_x = pl->tiles_y = 0;
    pl->tiles_m = (d->m + kBlockM - 1) / kBlockM;
    if (d->batch > 1) {
      DFU_REQUIRE(d->m % d->batch == 0 && d->ngroups == 1 && d->g[0].a_mode == 0, "gemm: batch=%d needs one matrix operand group and m %% batch == 0", d->batch);
      pl->tiles_m = d->batch * ((d->m / d->batch + kBlockM - 1) / kBlockM);
    }
  }
  const int sms = num_sms() > 0 ? num_sms() : 148;
  DFU_REQUIRE(d->kernel >= 0 && d->kernel <= 2, "gemm: kernel=%d (0 auto, 1 tile-per-CTA, 2 persistent pairs)", d->kernel);
  // ---- persistent CTA-pair kernel (gemm2.cu): explicit, or automatic when the 256-row pair tiles alone fill the GPU
  {
    int pbn = 0;
    if (d->kernel == 2) {
      pbn = d->block_n;
      if (pbn <= 0) {
        const int cands[] = {256, 192, 160, 128, 96, 80, 64, 32};
        for (int c : cands)
          if (d->n % c == 0 && (d->epi != DFU_EPI_GEGLU || c % 32 == 0)) { pbn = c; break; }
      }
      DFU_REQUIRE(d->splits <= 1, "gemm: the pair kernel has no split-K (splits=%d)", d->splits);
    } else if (d->kernel == 0 && d->block_n <= 0 && d->splits <= 0 &&
               !(getenv("DFU_GEMM_AUTO_PAIR") && getenv("DFU_GEMM_AUTO_PAIR")[0] == '0')) {
      const int cands[] = {256, 192, 160, 128};
      for (int c : cands)
        if (d->n % c == 0 && (d->epi != DFU_EPI_GEGLU || c % 32 == 0)) { pbn = c; break; }
      const long long pair_tiles = static_cast<long long>((pl->tiles_m + 1) / 2) * (pbn ? d->n / pbn : 0);
      if (pair_tiles < (sms / 2) * 2 || total_kb < 4) pbn = 0;  // under two waves of pairs: the tile-per-CTA kernel
    }
    if (pbn > 0) {
      DFU_REQUIRE(pbn % 16 == 0 && pbn >= 32 && pbn <= 256 && d->n % pbn == 0, "gemm: bad pair block_n=%d for n=%d", pbn, d->n);
      DFU_REQUIRE(d->epi != DFU_EPI_GEGLU || pbn % 32 == 0, "gemm: GEGLU needs block_n %% 32 == 0, got %d", pbn);
      pl->pair = 1;
      pl->block_n = pbn;
      pl->tiles_n = d->n / pbn;
      pl->splits = 1;
      pl->ws_bytes = 0;
      const size_t stage_bytes = (d->npass == 3 ? 2 : 1) * (kABytes + static_cast<size_t>(pbn) * 64);
      int stages = d->stages > 0 ? d->stages : kMaxStages;
      if (stages > kMaxStages) stages = kMaxStages;
      while (stages > 2 && pair_smem(stages, stage_bytes) > 226 * 1024) --stages;
      pl->stages = stages;
      pl->smem_bytes = pair_smem(stages, stage_bytes);
      DFU_REQUIRE(pl->smem_bytes <= 226 * 1024, "gemm: pair smem %zu too large", pl->smem_bytes);
      return DFU_OK;
    }
  }
  int bn_ = d->block_n;
  int splits = d->splits;
  if (bn_ <= 0 || splits <= 0) {
    // Cost model (SM cycles).  Per K=16 step a 128 x bn tile costs max(tensor, smem): the tensor pipe needs bn/2
    // cycles, the SS-mode operand fetch (4 KiB of A + 32*bn bytes of B at 128 B/clk) 32 + bn/4 — so tiles narrower
    // than 128 columns are operand-bandwidth bound and parallelism is bought with split-K instead.  CTAs beyond one
    // per SM serialise on the tensor pipe; a split adds a partial-tile round trip through L2 and a second launch.
    const int cands[] = {256, 160, 128, 96, 80, 64, 32};
    const int scand[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32};
    double best = 1e30;
    int best_bn = 0, best_s = 1;
    for (int c : cands) {
      if (d->n % c != 0) continue;
      if (d->block_n > 0 && c != d->block_n) continue;
      if (d->epi == DFU_EPI_GEGLU && c % 32 != 0) continue;
      const int tiles = pl->tiles_m * (d->n / c);
      const double per_k16 = (c / 2.0 > 32.0 + c / 4.0) ? c / 2.0 : 32.0 + c / 4.0;
      for (int sp : scand) {
        if (d->splits > 0 && sp != d->splits) continue;
        if (sp > 1 && total_kb / sp < 2) continue;
        const int ctas = tiles * sp;
        const int kb = (total_kb + sp - 1) / sp;
        const double cta = 2500.0 + kb * 4.0 * per_k16 * d->npass + (c / 32) * 450.0;  // prologue/fill + main loop + epilogue
        const double rounds = static_cast<double>((ctas + sms - 1) / sms);
        double t = rounds * cta;
        if (sp > 1) {
          const double bytes = 2.0 * sp * static_cast<double>(d->m) * d->n * 4.0;  // partials written and re-read
          t += 4000.0 + bytes / 1500.0;                                             // launch + ~3 TB/s of L2
        }
        if (t < best) {
          best = t;
          best_bn = c;
          best_s = sp;
        }
      }
    }
    DFU_REQUIRE(best_bn > 0, "gemm: no tiling for n=%d (block_n=%d splits=%d)", d->n, d->block_n, d->splits);
    bn_ = best_bn;
    splits = best_s;
  }
  DFU_REQUIRE(bn_ % 16 == 0 && bn_ >= 16 && bn_ <= 256 && d->n % bn_ == 0, "gemm: bad block_n=%d for n=%d", bn_, d->n);
  DFU_REQUIRE(d->epi != DFU_EPI_GEGLU || bn_ % 32 == 0, "gemm: GEGLU needs block_n %% 32 == 0, got %d", bn_);
  pl->block_n = bn_;
  pl->tiles_n = d->n / bn_;
  DFU_REQUIRE(splits >= 1, "gemm: bad splits=%d", splits);
  if (splits > total_kb) splits = total_kb;  // (tables tuned when a 3-pass K loop was three times as long)
  pl->splits = splits;
  pl->ws_bytes = splits > 1 ? static_cast<size_t>(splits) * d->m * d->n * sizeof(float) : 0;
  int stages = d->stages;
  const size_t stage_bytes = (d->npass == 3 ? 2 : 1) * (kABytes + static_cast<size_t>(bn_) * 128);
  if (stages <= 0) {
    // ~1.4 us of TMA latency (measured, scripts/trace_step.py) at >= 64 B/clk per SM wants ~100 KiB in flight; stay
    // under half the SM so that the NEXT kernel's CTA can be co-resident and prefetch its weights (PDL)
    stages = static_cast<int>((110 * 1024) / stage_bytes);
    if (stages < 3) stages = 3;
  }
  {
    const int kb_cta = (total_kb + splits - 1) / splits;
    if (stages > kb_cta) stages = kb_cta < 2 ? 2 : kb_cta;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  while (stages > 2 && stages * stage_bytes + 1024 > 224 * 1024) --stages;
  pl->stages = stages;
  pl->smem_bytes = stages * stage_bytes + 1024;  // >= kStageBytes: the epilogue staging reuses the ring
  DFU_REQUIRE(pl->smem_bytes <= 224 * 1024, "gemm: smem %zu too large", pl->smem_bytes);
  return DFU_OK;
}

int encode_group(const DfuGemm* d, const DfuGemmOperand& o, const Plan& pl, int b_box_rows, CUtensorMap* mA,
                 CUtensorMap* mB, uint32_t* a_tx) {
  int rc;
  if (o.a_mode == 0) {
    uint64_t dims[2] = {static_cast<uint64_t>(o.ntaps) * o.k_per_tap, static_cast<uint64_t>(o.a_rows)};
    uint64_t str[1] = {static_cast<uint64_t>(o.a_ld) * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    rc = make_tmap_f16(mA, o.a, 2, dims, str, box);
    *a_tx = kABytes;
  } else {
    uint64_t dims[4] = {static_cast<uint64_t>(o.a_c), static_cast<uint64_t>(o.a_w), static_cast<uint64_t>(o.a_h),
                        static_cast<uint64_t>(o.a_rows)};
    uint64_t str[3] = {static_cast<uint64_t>(o.a_c) * 2, static_cast<uint64_t>(o.a_c) * o.a_w * 2,
                       static_cast<uint64_t>(o.a_c) * o.a_w * o.a_h * 2};
    uint32_t box[4] = {kBlockK, static_cast<uint32_t>(pl.bw), static_cast<uint32_t>(pl.bh),
                       static_cast<uint32_t>(pl.bn)};
    rc = make_tmap_f16(mA, o.a, 4, dims, str, box);
    *a_tx = static_cast<uint32_t>(pl.bw * pl.bh * pl.bn) * kBlockK * 2;
  }
  if (rc) return rc;
  uint64_t bdims[2] = {static_cast<uint64_t>(o.ntaps) * o.k_per_tap, static_cast<uint64_t>(o.b_rows)};
  uint64_t bstr[1] = {static_cast<uint64_t>(o.b_ld) * 2};
  uint32_t bbox[2] = {kBlockK, static_cast<uint32_t>(b_box_rows)};
  DFU_REQUIRE(o.b_ld >= o.ntaps * o.k_per_tap && o.b_ld % 8 == 0, "gemm: b_ld=%d < ntaps*k_per_tap", o.b_ld);
  return make_tmap_f16(mB, o.b, 2, bdims, bstr, bbox);
}

static long long g_stats[6] = {0, 0, 0, 0, 0, 0};  // launches, split launches, fused second stages, reduce launches,
                                                  // pair-kernel launches, last grid

void fill_group_dev(const DfuGemmOperand& o, GroupDev& G) {
  G.a_mode = o.a_mode;
  G.ntaps = o.ntaps;
  G.nchunks = o.k_per_tap / kBlockK;
  G.a_plane = o.a_plane;
  G.b_plane = o.b_plane;
  G.kb_per_pass = o.ntaps * G.nchunks;
  G.b_static = o.b_static;
  for (int t = 0; t < 9; ++t) {
    G.dn[t] = o.tap_dn[t];
    G.dy[t] = o.tap_dy[t];
    G.dx[t] = o.tap_dx[t];
  }
}

void fill_batch_dev(const DfuGemm* d, BatchDev& bt) {
  bt.tiles_per_batch = 0;
  bt.rows_per_batch = d->m;
  bt.a_batch_rows = bt.b_batch_rows = 0;
  if (d->batch > 1 && !d->conv) {
    bt.rows_per_batch = d->m / d->batch;
    bt.tiles_per_batch = (bt.rows_per_batch + kBlockM - 1) / kBlockM;
    bt.a_batch_rows = d->a_batch_rows;
    bt.b_batch_rows = d->b_batch_rows;
  }
}

void fill_epi_params(const DfuGemm* d, EpiParams& e) {
  e.M = d->m;
  e.N = d->n;
  e.epi = d->epi;
  e.act = d->act;
  e.alpha = d->alpha;
  e.bias = d->bias;
  e.rowvec = d->rowvec;
  e.rowvec_ld = d->rowvec_ld;
  e.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  e.residual = d->residual;
  e.ldr = d->ldr;
  e.out_f32 = d->out_f32;
  e.ldo = d->ldo;
  e.out_f16 = static_cast<__half*>(d->out_f16);
  e.ldh = d->ldh;
  e.out_planes = d->out_planes > 0 ? d->out_planes : 1;
  e.out_plane_stride = d->out_plane_stride;
}

static int run_gemm(const DfuGemm* d, cudaStream_t stream) {
  Plan pl;
  int rc = plan_gemm(d, &pl);
  if (rc) return rc;
  if (pl.pair) {
    DFU_REQUIRE(d->epi >= 0 && d->epi <= 2, "gemm: bad epi");
    if (d->epi == DFU_EPI_F32) DFU_REQUIRE(d->out_f32 && d->ldo % 4 == 0, "gemm: out_f32/ldo");
    if (d->epi != DFU_EPI_F32) DFU_REQUIRE(d->out_f16 && d->ldh % 8 == 0, "gemm: out_f16/ldh");
    if (d->residual) DFU_REQUIRE(d->ldr % 4 == 0, "gemm: ldr");
    if (d->rowvec) DFU_REQUIRE(d->rows_per_sample > 0 && d->rowvec_ld % 4 == 0, "gemm: rowvec");
    g_stats[0]++;
    g_stats[4]++;
    return run_gemm2(d, pl, stream);
  }
  if (pl.splits > 1) {
    if (!d->workspace || d->workspace_bytes < pl.ws_bytes) {
      set_error("gemm: split-K needs %zu workspace bytes, got %zu", pl.ws_bytes, d->workspace_bytes);
      return DFU_ERR_WORKSPACE;
    }
  }
  DFU_REQUIRE(d->epi >= 0 && d->epi <= 2, "gemm: bad epi");
  DFU_REQUIRE(d->act == 0 || (d->act == 1 && d->epi != DFU_EPI_GEGLU), "gemm: act=%d (GEGLU has its own gate)", d->act);
  if (d->epi == DFU_EPI_F32) DFU_REQUIRE(d->out_f32 && d->ldo % 4 == 0, "gemm: out_f32/ldo");
  if (d->epi != DFU_EPI_F32) DFU_REQUIRE(d->out_f16 && d->ldh % 8 == 0, "gemm: out_f16/ldh");
  if (d->residual) DFU_REQUIRE(d->ldr % 4 == 0, "gemm: ldr");
  if (d->rowvec) DFU_REQUIRE(d->rows_per_sample > 0 && d->rowvec_ld % 4 == 0, "gemm: rowvec");

  CUtensorMap mA[2], mB[2];
  GemmKernelParams p;
  memset(&p, 0, sizeof(p));
  for (int g = 0; g < d->ngroups; ++g) {
    rc = encode_group(d, d->g[g], pl, pl.block_n, &mA[g], &mB[g], &p.a_tx_bytes[g]);
    if (rc) return rc;
    fill_group_dev(d->g[g], p.g[g]);
  }
  if (d->ngroups == 1) {
    mA[1] = mA[0];
    mB[1] = mB[0];
  }
  p.block_n = pl.block_n;
  p.tiles_m = pl.tiles_m;
  p.tiles_n = pl.tiles_n;
  p.splits = pl.splits;
  p.stages = pl.stages;
  p.total_kb = pl.total_kb;
  p.ngroups = d->ngroups;
  p.npass = d->npass;
  p.conv = d->conv;
  p.B = d->B;
  p.H = d->H;
  p.W = d->W;
  p.bw = pl.bw;
  p.bh = pl.bh;
  p.bn = pl.bn;
  p.tiles_x = pl.tiles_x;
  p.tiles_y = pl.tiles_y;
  p.b_tx_bytes = static_cast<uint32_t>(pl.block_n) * kBlockK * 2;
  fill_batch_dev(d, p.bt);
  // (measured: 84 -> 74 us for a one-CTA-per-SM conv with a 6-deep ring, -1 % on the contraction class of the step)
  static const int two_prod = getenv("DFU_GEMM_2PROD") ? atoi(getenv("DFU_GEMM_2PROD")) : 1;
  p.two_prod = two_prod;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(pl.block_n)) cols <<= 1;
  p.tmem_cols = cols;
  p.ws = static_cast<float*>(d->workspace);
  EpiParams& e = p.e;
  fill_epi_params(d, e);

  if (first_use_on_device(ONCE_GEMM_ATTR)) {
    DFU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    DFU_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  const int grid = pl.tiles_m * pl.tiles_n * pl.splits;
  static const bool cluster_ok = !(getenv("DFU_SPLITK_CLUSTER") && getenv("DFU_SPLITK_CLUSTER")[0] == '0');
  p.cluster = 0;
  // (16 CTAs per cluster is above the portable limit of 8: the kernel opts in once per device, below)
  if (cluster_ok && (pl.splits == 2 || pl.splits == 4 || pl.splits == 8 || pl.splits == 16)) {
    // the partial tile (128 x (block_n + 4) fp32) must fit in the idle operand ring
    const size_t part = static_cast<size_t>(kBlockM) * (pl.block_n + 4) * 4;
    if (part <= pl.smem_bytes - 1024) p.cluster = pl.splits;
  }
  // (a grid-barrier variant of the second stage was built and measured on B200: the barrier, ~4-5 us, cost more than
  // the PDL-overlapped reduce launch it saved, so it is gone; `sync_words` is accepted and ignored)
  p.sync = nullptr;
  g_stats[0]++;
  if (pl.splits > 1) g_stats[1]++;
  if (p.sync) g_stats[2]++;
  if (pl.splits > 1 && !p.sync && !p.cluster) g_stats[3]++;
  if (p.cluster) g_stats[2]++;
  DFU_CHECK_CUDA(launch_kc(gemm_tc_kernel, dim3(grid), dim3(kGemmThreads), pl.smem_bytes, stream, p.cluster > 1 ? p.cluster : 1, mA[0], mB[0], mA[1], mB[1], p));
  DFU_CHECK_CUDA(cudaGetLastError());
  if (pl.splits > 1 && p.sync == nullptr && !p.cluster) {
    const long long total = static_cast<long long>(d->m) * (d->n / (d->epi == DFU_EPI_GEGLU ? 8 : 4));
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 8;
    if (blocks > cap) blocks = cap;
    // few quads (deep levels): several lanes per quad so that all slices of a quad are in flight at once
    int team = 1;
    while (team < 8 && total * team * 2 <= cap * 256 && team * 2 <= pl.splits && (pl.splits + team * 2 - 1) / (team * 2) >= 2) team *= 2;
    if (team > 1 && (pl.splits + team - 1) / team <= 8) {
      const long long tb = (total * team + 255) / 256;
      if (team == 2) DFU_CHECK_CUDA(launch_k(splitk_reduce_team_kernel<2>, dim3(static_cast<unsigned>(tb)), dim3(256), 0, stream, p.ws, pl.splits, e));
      else if (team == 4) DFU_CHECK_CUDA(launch_k(splitk_reduce_team_kernel<4>, dim3(static_cast<unsigned>(tb)), dim3(256), 0, stream, p.ws, pl.splits, e));
      else DFU_CHECK_CUDA(launch_k(splitk_reduce_team_kernel<8>, dim3(static_cast<unsigned>(tb)), dim3(256), 0, stream, p.ws, pl.splits, e));
      return DFU_OK;
    }
    DFU_CHECK_CUDA(launch_k(splitk_reduce_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, p.ws, pl.splits, e));
  }
  return DFU_OK;
}

}  // namespace dfu

DFU_TRACE_SETTER(dfu_trace_set_gemm)

extern "C" {
int dfu_gemm(const DfuGemm* desc, void* stream) {
  if (!desc) {
    dfu::set_error("gemm: null descriptor");
    return DFU_ERR_INVALID;
  }
  return dfu::run_gemm(desc, static_cast<cudaStream_t>(stream));
}
void dfu_gemm_stats(int64_t* out) {
  for (int i = 0; i < 6; ++i) out[i] = dfu::g_stats[i];
}
int dfu_gemm_plan(const DfuGemm* desc, int32_t* out) {
  dfu::Plan pl;
  if (!desc || !out) return DFU_ERR_INVALID;
  int rc = dfu::plan_gemm(desc, &pl);
  if (rc) return rc;
  out[0] = pl.block_n; out[1] = pl.splits; out[2] = pl.stages; out[3] = pl.tiles_m; out[4] = pl.tiles_n;
  out[5] = pl.total_kb; out[6] = pl.pair ? 2 : 1; out[7] = 0;
  return DFU_OK;
}
size_t dfu_gemm_workspace(const DfuGemm* desc) {
  dfu::Plan pl;
  if (!desc || dfu::plan_gemm(desc, &pl)) return 0;
  return pl.ws_bytes;
}
}

// diffute_b200 — persistent CTA-pair contraction kernel (tcgen05 cta_group::2) for the shapes with many output tiles:
// batched UNet levels, every VAE conv, the FFN projections.
//
// A cluster of two CTAs (one TPC) owns 256 x block_n output tiles and walks them in a static round-robin schedule:
//   * ONE tcgen05.mma.cta_group::2 per K=16 step covers the whole 256 x block_n tile: each CTA stages its own 128 rows
//     of A and only HALF of the weight tile (block_n / 2 rows), so the L2 -> shared-memory traffic per output element
//     drops by the weight share (the tile-per-CTA kernel at 128 x 160 is bound by exactly that traffic, ~20 TB/s);
//   * two accumulators live in TMEM (2 x block_n columns): the eight epilogue warps of each CTA drain tile i while the
//     tensor core already accumulates tile i + 1 — the epilogue leaves the critical path;
//   * barriers, TMEM allocation, tensor-map prefetch are paid once per CTA, not once per tile.
// Roles per CTA: warp 0 TMA producer (both CTAs; the peer's loads signal the LEADER's full barrier), warp 1 TMEM
// allocator (both) + MMA issuer (leader only; commits are multicast to both CTAs' barriers), warps 2-9 epilogue.
// Same operand model as gemm.cu (matrix or NHWC-image A walked by taps with TMA out-of-bounds zero fill, two operand
// groups, 1 or 3 passes over hi/lo planes) and the same fused epilogues; no split-K (the tile count is the parallelism).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_shared.cuh"

namespace dfu {

constexpr int kG2Threads = 320;

struct Gemm2Params {
  int block_n;       // N of the pair's tile; each CTA stages block_n / 2 weight rows per k-block
  int tiles_m;       // 128-row m-tiles; pair p takes tiles 2p (leader) and 2p + 1 (peer)
  int tiles_n;
  int num_tiles;     // ceil(tiles_m / 2) * tiles_n
  int stages, total_kb, ngroups, npass;
  int kbs;           // k-blocks (K = 64) per pipeline stage: one full-barrier wait and ONE multicast commit per stage
  GroupDev g[2];
  int conv, B, H, W, bw, bh, bn, tiles_x, tiles_y;
  uint32_t a_tx_bytes[2];
  uint32_t b_tx_bytes;  // per CTA
  uint32_t tmem_cols;   // allocated columns (power of two >= 2 * block_n); accumulator s starts at s * tmem_cols / 2
  BatchDev bt;
  int debug;            // experiments only (DFU_G2_DEBUG): bit 0 = every CTA signals its OWN full barrier (results of the
                        // peer half are then unsynchronised: timing only)
  EpiParams e;
};

__global__ void __launch_bounds__(kG2Threads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
             const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
             const __grid_constant__ Gemm2Params p) {
  pdl_trigger();
  DFU_TR_SHARED_DECL();
  DFU_TR_SHARED_BEGIN(TR_GEMM | (2 << 8) | (2 << 16) | (p.block_n << 20));
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];   // leader's are used: both CTAs' loads of a stage land here
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];  // per CTA: "the MMAs that read this slot are done"
  __shared__ __align__(8) uint64_t tmem_full_bar[2];       // per CTA, per accumulator
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];      // leader's are used: 16 epilogue warps of the pair arrive
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t nplane = p.npass == 3 ? 2u : 1u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * 64u;  // block_n / 2 rows of 128 bytes
  const uint32_t kb_bytes = nplane * (kABytes + b_bytes);            // one k-block: [A planes][B planes]
  const uint32_t stage_bytes = static_cast<uint32_t>(p.kbs) * kb_bytes;
  const int nsteps = (p.total_kb + p.kbs - 1) / p.kbs;                // pipeline steps per tile
  // warp index through a shuffle: provably warp-uniform, so that the role loops below are UNIFORM control flow and
  // ptxas keeps the TMA / MMA operands (descriptors, coordinates, barrier addresses) in uniform registers.  With the
  // loops under `if (lane == 0)` every UTCHMMA / UTMALDG was preceded by an ELECT + R2UR.BROADCAST "waterfall"
  // (~250 cycles per MMA, measured: the pair kernel was bound by its single issuing thread, not by the tensor pipe).
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

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
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 16);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before either signals the other
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) DFU_TR_SHARED_MARK(5);

  // tile -> coordinates of THIS CTA's 128 rows
  int a_row0 = 0, b_off = 0, m_end = 0;  // (batched products: A / B row bases and the end of the batch's output rows)
  auto tile_coords = [&](int tile, int& n_tile0, int& m0, int& x0, int& y0, int& img0) {
    const int tn = tile % p.tiles_n;
    const int tm = 2 * (tile / p.tiles_n) + static_cast<int>(rank);
    n_tile0 = tn * p.block_n;
    batch_coords(p.bt, tm, p.e.M, m0, a_row0, b_off, m_end);
    if (tm >= p.tiles_m) m_end = m0;  // the empty peer tile of an odd last pair
    x0 = y0 = img0 = 0;
    if (p.conv) {
      int t = tm;
      x0 = (t % p.tiles_x) * p.bw;
      t /= p.tiles_x;
      y0 = (t % p.tiles_y) * p.bh;
      img0 = (t / p.tiles_y) * p.bn;  // >= B for the peer of an odd last pair: every load is out-of-bounds zero fill
    }
  };

  if (warp == 0) {
    // ===== TMA producer (both CTAs): the whole warp walks the schedule, one elected lane issues ==
    {
      const int kbg0 = p.g[0].kb_per_pass;
      int n_tile0, m0, x0, y0, img0;
      // loads of pipeline step `st` of the current tile (k-blocks st*kbs ...) into ring slot `stage`
      auto issue = [&](int st, int stage, bool load_a, bool load_b) {
        const int kb0 = st * p.kbs;
        const int cnt = min(p.kbs, p.total_kb - kb0);
        const bool localbar = (p.debug & 1) != 0;
        const uint32_t bar = dsmem_addr(smem_u32(&full_bar[stage]), localbar ? rank : 0);  // the leader's barrier
        if (load_b) {
          // the leader arms its barrier with the bytes of ALL tiles of the step, its own and the peer's
          uint32_t bytes = 0;
          for (int j = 0; j < cnt; ++j) bytes += nplane * (p.a_tx_bytes[(kb0 + j) >= kbg0 ? 1 : 0] + p.b_tx_bytes);
          if (localbar) mbar_arrive_expect_tx(&full_bar[stage], bytes);
          else if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * bytes);
        }
        for (int j = 0; j < cnt; ++j) {
          int r = kb0 + j, gi = 0;
          if (r >= kbg0) {
            r -= kbg0;
            gi = 1;
          }
          const GroupDev& G = p.g[gi];
          const int tap = r / G.nchunks;
          const int chunk = r - tap * G.nchunks;
          const CUtensorMap* mA = gi ? &tmA1 : &tmA0;
          const CUtensorMap* mB = gi ? &tmB1 : &tmB0;
          uint8_t* sA = smem + stage * stage_bytes + j * kb_bytes;
          uint8_t* sB = sA + nplane * kABytes;
          if (load_b) {
            for (uint32_t pl = 0; pl < nplane; ++pl)
              tma_load_2d_cg2(sB + pl * b_bytes, mB, bar, (tap * G.nchunks + chunk) * kBlockK,
                              n_tile0 + b_off + static_cast<int>(rank) * (p.block_n >> 1) + static_cast<int>(pl) * G.b_plane);
          }
          if (load_a) {
            for (uint32_t pl = 0; pl < nplane; ++pl) {
              const int a_sel = static_cast<int>(pl) * G.a_plane;
              if (G.a_mode == 0) {
                tma_load_2d_cg2(sA + pl * kABytes, mA, bar, (tap * G.nchunks + chunk) * kBlockK, a_row0 + a_sel);
              } else {
                tma_load_4d_cg2(sA + pl * kABytes, mA, bar, chunk * kBlockK, x0 + G.dx[tap], y0 + G.dy[tap],
                                img0 + G.dn[tap] + a_sel);
              }
            }
          }
        }
      };
      int stage = 0;
      uint32_t phase = 0;
      int npre = 0;
      const bool pre = p.g[0].b_static && (p.ngroups == 1 || p.g[1].b_static);
      if (pair_id < p.num_tiles) {
        tile_coords(pair_id, n_tile0, m0, x0, y0, img0);
        // weights do not depend on the previous kernel: request the first ring-full of weight tiles before waiting
        npre = pre ? min(p.stages, nsteps) : 0;
        if (elect_one())
          for (int i = 0; i < npre; ++i) issue(i, i, false, true);
        __syncwarp();
      }
      pdl_wait();
      if (lane == 0) DFU_TR_SHARED_MARK(6);
      for (int tile = pair_id; tile < p.num_tiles; tile += npairs) {
        tile_coords(tile, n_tile0, m0, x0, y0, img0);
        for (int st = 0; st < nsteps; ++st) {
          if (tile == pair_id && st < npre) {
            if (elect_one()) issue(st, stage, true, false);
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (elect_one()) issue(st, stage, true, true);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only): whole warp in uniform control flow, one elected lane issues =
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16(2 * kBlockM, p.block_n);
      const int nprod = p.npass == 3 ? 3 : 1;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair_id; tile < p.num_tiles; tile += npairs, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1u);  // the pair's epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * (p.tmem_cols >> 1);
        for (int st = 0; st < nsteps; ++st) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0 && it == 0 && st == 0) DFU_TR_SHARED_MARK(7);
          const int cnt = min(p.kbs, p.total_kb - st * p.kbs);
          if (elect_one()) {
            for (int j = 0; j < cnt; ++j) {
              const uint32_t sA = smem_u32(smem + stage * stage_bytes + j * kb_bytes);
              const uint32_t sB = sA + nplane * kABytes;
              for (int ps = 0; ps < nprod; ++ps) {  // hi*hi [, lo*hi, hi*lo] from the same k-block
                const uint64_t adesc = umma_desc_sw128(sA + (ps == 1 ? kABytes : 0u));
                const uint64_t bdesc = umma_desc_sw128(sB + (ps == 2 ? b_bytes : 0u));
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_f16_ss2(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                               (st > 0 || j > 0 || ps > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit2(&empty_bar[stage], 3);  // frees this slot in BOTH CTAs
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma_commit2(&tmem_full_bar[acc], 3);  // accumulator complete, in both CTAs' TMEM
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps (2..9), both CTAs =====================================================
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;  // which half of the 32-column chunks this warp handles
    const int r = q * 32 + lane;
    const EpiParams& e = p.e;
    const bool geglu = e.epi == DFU_EPI_GEGLU;
    const bool has_res = !geglu && e.residual != nullptr;
    const bool f32_out = e.epi == DFU_EPI_F32;
    const bool lo_plane = e.out_planes > 1;
    const bool act = e.act != 0;
    const float alpha = e.alpha;
    const int cq = lane & 7;
    const int cq4 = lane & 3;
    float* stg = reinterpret_cast<float*>(smem + static_cast<size_t>(p.stages) * stage_bytes) + (warp - 2) * kStageFloats;
    const uint32_t tmem_empty_addr[2] = {dsmem_addr(smem_u32(&tmem_empty_bar[0]), 0),
                                         dsmem_addr(smem_u32(&tmem_empty_bar[1]), 0)};
    bool waited = false;
    int it = 0;
    for (int tile = pair_id; tile < p.num_tiles; tile += npairs, ++it) {
      const int acc = it & 1;
      int n_tile0, m0, x0, y0, img0;
      tile_coords(tile, n_tile0, m0, x0, y0, img0);
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
      // (m, valid) of the rows this lane stores after the transpose through shared memory
      const int vmask = valid ? 1 : 0;
      int mrs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = geglu ? (i & 3) * 8 + (lane >> 2) : i * 4 + (lane >> 3);
        const int mr = __shfl_sync(0xffffffffu, m, row);
        const int vr = __shfl_sync(0xffffffffu, vmask, row);
        mrs[i] = vr ? mr : -1;
      }
      int smp0 = 0;
      bool one_sample = true;
      if (e.rowvec != nullptr) {
        const int smp = valid ? m / e.rows_per_sample : -1;
        smp0 = __reduce_max_sync(0xffffffffu, smp);
        one_sample = __all_sync(0xffffffffu, smp < 0 || smp == smp0);
        if (smp0 < 0) smp0 = 0;
      }
      if (!waited) {  // first touch of activations (residual / rowvec) and of buffers earlier kernels may still read
        pdl_wait();
        waited = true;
      }
      float4 res[8];
      auto fetch_res1 = [&](int c, int i) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res && (c + cq * 4 < p.block_n) && mrs[i] >= 0)
          t = *reinterpret_cast<const float4*>(e.residual + static_cast<size_t>(mrs[i]) * e.ldr + n_tile0 + c + cq * 4);
        return t;
      };
#pragma unroll
      for (int i = 0; i < 8; ++i) res[i] = fetch_res1(cg * 32, i);
      mbar_wait_sleep(&tmem_full_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && it == 0) DFU_TR_SHARED_MARK(8);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc) * (p.tmem_cols >> 1);
      bool released = false;
#pragma unroll 1
      for (int c = cg * 32; c < p.block_n; c += 64) {
        uint32_t raw[32];
        tmem_ld32(taddr + static_cast<uint32_t>(c), raw);
        tmem_ld_wait();
        if (threadIdx.x == 64 && it == 0 && c == 0) DFU_TR_SHARED_MARK(12);
        if (c + 64 >= p.block_n) {  // this warp's last read of the accumulator: hand it back to the MMA issuer
          tc_fence_before();
          if (lane == 0) mbar_arrive_cluster(tmem_empty_addr[acc]);
          released = true;
        }
        const int ncol = p.block_n - c;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(stg + lane * kStageLd + 4 * i) =
              make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]), __uint_as_float(raw[4 * i + 2]),
                          __uint_as_float(raw[4 * i + 3]));
        __syncwarp();
        const int n = n_tile0 + c;
        if (geglu) {
          // value / gate quads 16 columns apart inside the 32-column chunk (weights packed in 16 / 16 row blocks)
          float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
          if (e.bias) {
            ba = __ldg(reinterpret_cast<const float4*>(e.bias + n + cq4 * 4));
            bg = __ldg(reinterpret_cast<const float4*>(e.bias + n + 16 + cq4 * 4));
          }
          const int n_out = ((n + cq4 * 4) >> 5) * 16 + ((n + cq4 * 4) & 15);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + (lane >> 2);
            const float4 a = *reinterpret_cast<const float4*>(stg + row * kStageLd + cq4 * 4);
            const float4 g = *reinterpret_cast<const float4*>(stg + row * kStageLd + 16 + cq4 * 4);
            if (mrs[i] >= 0) {
              float4 o;
              o.x = fmaf(a.x, alpha, ba.x) * gelu_erf_f(fmaf(g.x, alpha, bg.x));
              o.y = fmaf(a.y, alpha, ba.y) * gelu_erf_f(fmaf(g.y, alpha, bg.y));
              o.z = fmaf(a.z, alpha, ba.z) * gelu_erf_f(fmaf(g.z, alpha, bg.z));
              o.w = fmaf(a.w, alpha, ba.w) * gelu_erf_f(fmaf(g.w, alpha, bg.w));
              store_f16x4(e.out_f16 + static_cast<size_t>(mrs[i]) * e.ldh + n_out, o, lo_plane, e.out_plane_stride);
            }
          }
        } else {
          const bool qok = cq * 4 < ncol;
          float4 aq = make_float4(0.f, 0.f, 0.f, 0.f);
          if (qok && e.bias) aq = __ldg(reinterpret_cast<const float4*>(e.bias + n + cq * 4));
          if (qok && e.rowvec && one_sample) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(e.rowvec + static_cast<size_t>(smp0) * e.rowvec_ld + n + cq * 4));
            aq.x += t.x; aq.y += t.y; aq.z += t.z; aq.w += t.w;
          }
          const bool per_row_vec = e.rowvec != nullptr && !one_sample;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3);
            const float4 v = *reinterpret_cast<const float4*>(stg + row * kStageLd + cq * 4);
            if (qok && mrs[i] >= 0) {
              float4 o;
              o.x = fmaf(v.x, alpha, aq.x) + res[i].x;
              o.y = fmaf(v.y, alpha, aq.y) + res[i].y;
              o.z = fmaf(v.z, alpha, aq.z) + res[i].z;
              o.w = fmaf(v.w, alpha, aq.w) + res[i].w;
              if (per_row_vec) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(
                    e.rowvec + static_cast<size_t>(mrs[i] / e.rows_per_sample) * e.rowvec_ld + n + cq * 4));
                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
              }
              if (act) {
                o.x = gelu_erf_f(o.x); o.y = gelu_erf_f(o.y); o.z = gelu_erf_f(o.z); o.w = gelu_erf_f(o.w);
              }
              if (f32_out)
                *reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(mrs[i]) * e.ldo + n + cq * 4) = o;
              else
                store_f16x4(e.out_f16 + static_cast<size_t>(mrs[i]) * e.ldh + n + cq * 4, o, lo_plane, e.out_plane_stride);
            }
            if (c + 64 < p.block_n) res[i] = fetch_res1(c + 64, i);
          }
        }
      }
      if (!released) {  // a warp without a chunk of its own (block_n <= 32) still owes its arrival
        tc_fence_before();
        if (lane == 0) mbar_arrive_cluster(tmem_empty_addr[acc]);
      }
      if (threadIdx.x == 64 && it == 0) DFU_TR_SHARED_MARK(9);
    }
  }

  // nobody leaves (or frees TMEM) while the peer may still signal its barriers or the leader's MMAs write its TMEM
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) DFU_TR_SHARED_END();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
int run_gemm2(const DfuGemm* d, const Plan& pl, cudaStream_t stream) {
  CUtensorMap mA[2], mB[2];
  Gemm2Params p;
  memset(&p, 0, sizeof(p));
  for (int g = 0; g < d->ngroups; ++g) {
    int rc = encode_group(d, d->g[g], pl, pl.block_n / 2, &mA[g], &mB[g], &p.a_tx_bytes[g]);
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
  p.num_tiles = ((pl.tiles_m + 1) / 2) * pl.tiles_n;
  // pl.stages counts k-blocks the ring can hold; a pipeline stage groups `kbs` of them
  static const int kbs_env = getenv("DFU_G2_KBS") ? atoi(getenv("DFU_G2_KBS")) : 0;
  int kbs = kbs_env > 0 ? kbs_env : 4;  // measured: one full-barrier round trip + multicast commit costs ~0.26 us per stage
  if (kbs > pl.stages / 2) kbs = pl.stages / 2 > 0 ? pl.stages / 2 : 1;
  if (kbs > pl.total_kb) kbs = pl.total_kb;
  p.kbs = kbs;
  p.stages = pl.stages / kbs;
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
  p.b_tx_bytes = static_cast<uint32_t>(pl.block_n / 2) * kBlockK * 2;
  uint32_t cols = 32;
  while (cols < 2u * static_cast<uint32_t>(pl.block_n)) cols <<= 1;
  p.tmem_cols = cols;
  fill_epi_params(d, p.e);
  fill_batch_dev(d, p.bt);
  static const int debug = getenv("DFU_G2_DEBUG") ? atoi(getenv("DFU_G2_DEBUG")) : 0;
  p.debug = debug;
  if (first_use_on_device(ONCE_GEMM2_ATTR))
    DFU_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int pairs = sms / 2;
  if (pairs > p.num_tiles) pairs = p.num_tiles;
  DFU_CHECK_CUDA(launch_kc(gemm2_kernel, dim3(2 * pairs), dim3(kG2Threads), pl.smem_bytes, stream, 2, mA[0], mB[0], mA[1], mB[1], p));
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

}  // namespace dfu

DFU_TRACE_SETTER(dfu_trace_set_gemm2)

// diffute_b200 — fused attention core for head dim 64 on tcgen05 (softmax(Q K^T * scale) V), sm_100a.
//
// Replaces diffusers' Attention core (attn1 self-attention over H*W tokens and attn2 cross-attention over the
// 577 glyph tokens) in BasicTransformerBlock — SURVEY.md A.1, reached from app.ipynb:814.
//
// One CTA = 128 query rows of one (sample, head); it streams K/V in blocks of 64 keys:
//   warps 0-7 : two softmax warpgroups that take ALTERNATE key blocks (block g belongs to warpgroup g & 1).  Each has
//               its own S buffer (64 TMEM columns), P tile and accumulator O_wg in TMEM, its own running reference
//               max and row sum: while one warpgroup exponentiates block g, the tensor core already computes
//               S(g+1) for the other — the Q K^T round trip that starved a single shared S buffer (ncu: the softmax
//               warps sat on the S barrier) is off the critical path, and nothing is exchanged per block.  Thread
//               (wg, r) owns query row r (= TMEM lane r) of its blocks; the two partial results of a row are merged
//               once per work item (O = (O_0 w_0 + O_1 w_1) / (l_0 w_0 + l_1 w_1), w = 2^((m - max m) c)).  The
//               reference max is lazy (FlashAttention-4 style): probabilities are computed against the current
//               reference in ONE pass over S while the block max is tracked; only when a row outgrows 2^8 of headroom
//               is O_wg rescaled in TMEM (tcgen05.ld / st) and the block redone.  P is written to shared memory as fp16
//               (hi [, lo]) in the 128-byte-swizzled K-major layout the P V MMA consumes.
//   warp 8    : TMA producer (Q once per work item; K and V in 4-deep rings of 64-key tiles; 4-D maps so rows past
//               the sequence end are zero-filled per sample and per plane).
//   warp 9    : tcgen05.mma issuer: S_wg = Q K^T (K-major B) two blocks ahead, O_wg += P_wg V with V consumed MN-major
//               straight from its natural [key, d] layout (no transpose pass).
// In FP16X2 mode both contractions run as three passes over (hi, lo) operand planes.
//
// Work distribution ("stream-K" over key blocks): a work item is one (sample, head, 128-query tile) with nblk key blocks.
// The items x nblk block-units are laid end to end and cut into G equal contiguous ranges, one per CTA, so a CTA runs
// the tail of one item, possibly whole items, and the head of the next: with G = 2 x SMs every SM holds two equally
// loaded CTAs whatever the (tile, head) count is (160 tiles on 148 SMs at 64x64 latents).  A range that covers a whole
// item writes the result; partial ranges write (unnormalised O, row max, row sum) to the workspace and
// attn_merge_kernel combines the pieces of an item in key order (deterministic).
#include "common.cuh"
#include "kernels.h"

namespace dfu {

constexpr int kAttnThreads = 320;
constexpr int kSoftmaxThreads = 256;
constexpr int kBQ = 128;    // queries per CTA
constexpr int kBKV = 64;    // keys per block
constexpr int kRing = 4;    // K / V ring depth (tiles of 64 keys)
constexpr int kD = 64;
constexpr uint32_t kTile = 128 * 128;  // bytes of one [128 rows x 64 halfs] swizzled tile (Q, P)
constexpr uint32_t kKvTile = 64 * 128;  // bytes of one [64 keys x 64 halfs] swizzled tile (K, V)

struct AttnParams {
  int B, heads, Nq, Nk;
  int planes;       // 1 or 2
  float scale_log2; // softmax scale * log2(e)
  int q_col0, k_col0, v_col0;  // column offsets of head 0 inside the Q / K / V matrices
  __half* out;      // [planes][B*Nq][ldo], head h at columns h*64
  int ldo;
  long long out_plane_stride;
  // stream-K distribution: CTA c owns block-units [c*total/G, (c+1)*total/G) of the items x nblk sequence
  int q_tiles, nblk, items, G;
  long long total;  // items * nblk
  float* ws_o;      // [2*G slots][128][64] fp32; slot = 2*cta + (1 if the piece starts at its item's first block)
  float* ws_ml;     // [2*G slots][128][2] (m in raw score units, l)
};

__host__ __device__ __forceinline__ long long attn_range_begin(const AttnParams& p, int c) {
  return static_cast<long long>(c) * p.total / p.G;
}
__host__ __device__ __forceinline__ int attn_cta_of(const AttnParams& p, long long u) {  // CTA whose range holds unit u
  return static_cast<int>(((u + 1) * p.G + p.total - 1) / p.total - 1);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // not volatile: a pure function the scheduler may reorder
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_ATTN | ((p.G > p.items ? 1 : 0) << 8));
  // No static shared memory: the dynamic window then starts 1024-byte aligned at the CTA's base, and the FP16 mode
  // needs 7 x 16 KiB + 256 B, so two CTAs fit one SM (one's softmax overlaps the other's MMAs).
  extern __shared__ __align__(1024) uint8_t smem[];
  const int planes = p.planes;
  // layout: Q[planes] | K[kRing][planes] | V[kRing][planes] | P[planes][2 warpgroups] | barriers
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + planes * kTile;
  uint8_t* sV = sK + kRing * planes * kKvTile;
  uint8_t* sP = sV + kRing * planes * kKvTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * planes * kTile);
  uint64_t& q_full = bars[0];
  uint64_t& q_free = bars[1];    // all Q K^T MMAs of a segment have read Q: the next segment's Q may be loaded
  uint64_t* k_full = bars + 2;   // [kRing]
  uint64_t* k_empty = bars + 6;
  uint64_t* v_full = bars + 10;
  uint64_t* v_empty = bars + 14;
  uint64_t* s_full = bars + 18;  // [2] per warpgroup: its S buffer holds a new block
  uint64_t* s_free = bars + 20;  // [2] 128 arrivals: the warpgroup has read its S buffer
  uint64_t* p_full = bars + 22;  // [2] 128 arrivals: the warpgroup's P tile is written
  uint64_t* o_full = bars + 24;  // [2] P_wg V of a block has completed
  uint64_t& o_free = bars[26];   // 256 arrivals: both accumulators of a finished segment have been read
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 27);

  // (through a shuffle: provably warp-uniform — the producer / issuer loops are then uniform control flow and ptxas keeps
  // TMA / MMA operands in uniform registers instead of an ELECT + R2UR.BROADCAST waterfall per instruction)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const long long u_begin = attn_range_begin(p, blockIdx.x);
  const long long u_end = attn_range_begin(p, blockIdx.x + 1);
  if (u_begin >= u_end) return;  // (only when G > total; the host never launches that)

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full, 1);
    mbar_init(&q_free, 1);
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], kSoftmaxThreads / 2);
      mbar_init(&p_full[i], kSoftmaxThreads / 2);
      mbar_init(&o_full[i], 1);
    }
    mbar_init(&o_free, kSoftmaxThreads);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(&tmem_base_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_S = tmem_base;        // 2 x 64 columns: S_0, S_1
  const uint32_t tmem_O = tmem_base + 128;  // 2 x 64 columns: O_0, O_1
  DFU_TR_MARK(5);
  pdl_wait();
  DFU_TR_MARK(6);

  // A segment = the part of one item inside this CTA's range: item index, first key block jb0, block count n.
  // All three roles walk the same segment list; `g` counts key blocks over the whole range and keys every barrier
  // parity and ring slot, so the pipeline runs straight through segment boundaries.
  auto seg_of = [&](long long u, int& item, int& jb0, int& n) {
    item = static_cast<int>(u / p.nblk);
    jb0 = static_cast<int>(u - static_cast<long long>(item) * p.nblk);
    const long long rest = u_end - u;
    n = (p.nblk - jb0 < rest) ? p.nblk - jb0 : static_cast<int>(rest);
  };
  auto coords = [&](int item, int& b, int& head, int& q0) {
    const int q_tile = item % p.q_tiles;
    const int bh = item / p.q_tiles;
    head = bh % p.heads;
    b = bh / p.heads;
    q0 = q_tile * kBQ;
  };

  if (warp == 8) {
    // ===== TMA producer: the whole warp walks the schedule, one elected lane issues =============
    {
      int g = 0, seg = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        int item, jb0, n, b, head, q0;
        seg_of(u, item, jb0, n);
        coords(item, b, head, q0);
        if (seg > 0) mbar_wait_ns(&q_free, (seg - 1) & 1, 64, 512);  // previous segment's Q K^T MMAs have all read Q
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full, planes * kTile);
          for (int pl = 0; pl < planes; ++pl)
            tma_load_4d(sQ + pl * kTile, &tmQ, &q_full, p.q_col0 + head * kD, q0, b, pl);
        }
        __syncwarp();
        for (int j = 0; j < n; ++j, ++g) {
          const int slot = g % kRing;
          const uint32_t par = ((g / kRing) & 1) ^ 1u;
          mbar_wait_ns(&k_empty[slot], par, 64, 512);  // (a polling warp takes issue slots from the softmax warps)
          if (elect_one()) {
            mbar_arrive_expect_tx(&k_full[slot], planes * kKvTile);
            for (int pl = 0; pl < planes; ++pl)
              tma_load_4d(sK + (slot * planes + pl) * kKvTile, &tmK, &k_full[slot], p.k_col0 + head * kD, (jb0 + j) * kBKV, b, pl);
          }
          __syncwarp();
          mbar_wait_ns(&v_empty[slot], par, 64, 512);
          if (elect_one()) {
            mbar_arrive_expect_tx(&v_full[slot], planes * kKvTile);
            for (int pl = 0; pl < planes; ++pl)
              tma_load_4d(sV + (slot * planes + pl) * kKvTile, &tmV, &v_full[slot], p.v_col0 + head * kD, (jb0 + j) * kBKV, b, pl);
          }
          __syncwarp();
        }
        u += n;
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer: whole warp in uniform control flow, one elected lane issues ==============
    {
      const uint32_t idesc_qk = umma_idesc_f16(128, kBKV, 0);
      const uint32_t idesc_pv = umma_idesc_f16(128, kD, 1);  // B (= V) is MN-major
      const int npass = planes == 2 ? 3 : 1;
      // block gg belongs to warpgroup h = gg & 1 and is that warpgroup's (gg >> 1)-th block
      auto issue_qk = [&](int gg, bool last_of_segment) {
        const int slot = gg % kRing, h = gg & 1, c = gg >> 1;
        mbar_wait(&k_full[slot], (gg / kRing) & 1);
        if (c > 0) mbar_wait(&s_free[h], (c - 1) & 1);  // the warpgroup has read its previous block out of S_h
        tc_fence_after();
        if (elect_one()) {
          uint32_t acc = 0;
          for (int ps = 0; ps < npass; ++ps) {
            const int qa = (ps == 1) ? 1 : 0, kb = (ps == 2) ? 1 : 0;
            const uint64_t ad = umma_desc_sw128(smem_u32(sQ + qa * kTile));
            const uint64_t bd = umma_desc_sw128(smem_u32(sK + (slot * planes + kb) * kKvTile));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) {
              umma_f16_ss(tmem_S + 64 * h, ad + 2 * k, bd + 2 * k, idesc_qk, acc);
              acc = 1;
            }
          }
          umma_commit(&s_full[h]);
          umma_commit(&k_empty[slot]);
          if (last_of_segment) umma_commit(&q_free);
        }
        __syncwarp();
      };
      int g = 0, seg = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        int item, jb0, n;
        seg_of(u, item, jb0, n);
        mbar_wait(&q_full, seg & 1);
        issue_qk(g, n == 1);            // S two blocks ahead of the P V: both warpgroups have work from the start
        if (n > 1) issue_qk(g + 1, n == 2);
        for (int j = 0; j < n; ++j) {
          const int gg = g + j, slot = gg % kRing, h = gg & 1, c = gg >> 1;
          mbar_wait(&p_full[h], c & 1);  // softmax of block gg done (its s_free arrived at the same time)
          // the warpgroup's NEXT S first — it is waiting for it — then the P V of the block it just finished, whose
          // result is only needed one block later
          if (j + 2 < n) issue_qk(gg + 2, j + 3 == n);
          mbar_wait(&v_full[slot], (gg / kRing) & 1);
          if (j == 0 && seg > 0) mbar_wait(&o_free, (seg - 1) & 1);  // the previous item's accumulators were read
          tc_fence_after();
          if (elect_one()) {
            uint32_t acc = j >= 2 ? 1u : 0u;  // O_h accumulates in TMEM over the warpgroup's blocks of a segment
            for (int ps = 0; ps < npass; ++ps) {
              const int pa = (ps == 1) ? 1 : 0, vb = (ps == 2) ? 1 : 0;
              const uint32_t pbase = smem_u32(sP + (pa * 2 + h) * kTile);
              const uint32_t vbase = smem_u32(sV + (slot * planes + vb) * kKvTile);
#pragma unroll
              for (int k = 0; k < kBKV / 16; ++k) {
                const uint64_t ad = umma_desc_sw128(pbase) + 2 * k;
                const uint64_t bd = umma_desc_sw128(vbase + k * 2048);
                umma_f16_ss(tmem_O + 64 * h, ad, bd, idesc_pv, acc);
                acc = 1;
              }
            }
            umma_commit(&o_full[h]);
            umma_commit(&v_empty[slot]);
          }
          __syncwarp();
        }
        g += n;
        u += n;
      }
    }
  } else {
    // ===== softmax warpgroups (warps 0..7) ======================================================
    const int wg = warp >> 2;          // takes the key blocks gg with (gg & 1) == wg; output columns [32*wg, 32*wg+32)
    const int r = threadIdx.x & 127;   // query row within the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float c2 = p.scale_log2;
    const uint32_t pair_bar = 1u + static_cast<uint32_t>(warp & 3);  // named barrier of warps (w, w + 4)
    const uint32_t prow = static_cast<uint32_t>(r) * 128u;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint8_t* tile_hi = sP + wg * kTile + prow;        // P tile of this warpgroup (hi plane), own row
    uint8_t* tile_lo = sP + (2 + wg) * kTile + prow;
    const uint32_t tmem_Sown = tmem_S + 64 * wg + lane_off;
    const uint32_t tmem_Oown = tmem_O + 64 * wg + lane_off;
    int g = 0;
#ifdef DFU_TRACE
    long long tr_ws = 0, tr_wo = 0, tr_cmp = 0, tr_n = 0, tr_t = 0;  // thread 0: cycles waiting for S / for P V, computing
#endif
    for (long long u = u_begin; u < u_end;) {
      int item, jb0, nblk, b, head, q0;
      seg_of(u, item, jb0, nblk);
      coords(item, b, head, q0);
      float m = -INFINITY, l = 0.f;   // reference max (raw score units) and row sum at that reference, own blocks
      int nown = 0;                   // own blocks of this segment done so far
      int c = 0;
      for (int gg = g + ((g & 1) != wg ? 1 : 0); gg < g + nblk; gg += 2, ++nown) {
        c = gg >> 1;                  // this warpgroup's block counter over the whole range
#ifdef DFU_TRACE
        tr_t = clock64();
#endif
        mbar_wait(&s_full[wg], c & 1);
        tc_fence_after();
#ifdef DFU_TRACE
        { const long long t = clock64(); tr_ws += t - tr_t; tr_t = t; }
#endif
        const int kv_valid = p.Nk - (jb0 + (gg - g)) * kBKV;  // columns >= kv_valid are padding
        const bool full = kv_valid >= kBKV;                   // warp-uniform: interior blocks skip the tail predicates
        // P_wg V of the previous own block must be finished before the P tile is overwritten (and before O_wg is
        // touched).  It was issued right after this block's Q K^T and completes ~450-550 cycles after S arrives
        // (scripts/exp_attn_phases.py).  Computing the first 32 probabilities into registers before this wait was
        // measured: the wait disappears but the block's compute phase grows by more (1830 -> 2480 cycles, spills at 96
        // registers), so it stays here
        if (nown > 0) {
          mbar_wait(&o_full[wg], (c - 1) & 1);
          tc_fence_after();
        }
#ifdef DFU_TRACE
        { const long long t = clock64(); tr_wo += t - tr_t; tr_t = t; }
#endif
        if (nown == 0) {
          // first own block of a segment: the reference is its exact max (one extra read of S per segment)
          float mx = -INFINITY;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t raw[32];
            tmem_ld32(tmem_Sown + cc * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = (full || cc * 32 + i < kv_valid) ? __uint_as_float(raw[i]) : -INFINITY;
              mx = fmaxf(mx, v);
            }
          }
          m = mx;
        }
        // ONE pass over S in the common case: probabilities are computed against the current reference while the
        // block max is tracked; only if some row of the warp outgrew the 2^8 headroom is O_wg rescaled and the block
        // redone against the new reference (S is still there: s_free not yet signalled).
        float rowsum = 0.f;
        for (int attempt = 0;; ++attempt) {
          const float mc = (m == -INFINITY) ? 0.f : m * c2;
          float big = 0.f;  // largest 8-element partial sum: > 2^11 means some probability outgrew the headroom
          rowsum = 0.f;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t raw[32];  // (both halves in flight at once was measured: no gain, and it spills at 96 registers)
            tmem_ld32(tmem_Sown + cc * 32, raw);
            tmem_ld_wait();
            if (!full) {  // ragged last block only (warp-uniform): padding columns become -inf -> probability 0
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (cc * 32 + i >= kv_valid) raw[i] = 0xff800000u;
            }
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {  // 16-byte units of 8 probabilities
              float pv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) pv[i] = fast_exp2(__uint_as_float(raw[uu * 8 + i]) * c2 - mc);
              // the kernel is bound by instruction issue (a single spinning warp costs 12%): no per-element max — the
              // unit sums, needed anyway, tell whether the reference has to move
              const float us = ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
              rowsum += us;
              big = fmaxf(big, us);
              __align__(16) __half2 h[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(pv[2 * i], pv[2 * i + 1]);
              const uint32_t unit = static_cast<uint32_t>(cc * 4 + uu);
              const uint32_t off = (unit ^ sw) << 4;
              *reinterpret_cast<uint4*>(tile_hi + off) = *reinterpret_cast<const uint4*>(h);
              if (planes == 2) {
                __align__(16) __half2 lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 hf = __half22float2(h[i]);
                  lo[i] = __floats2half2_rn(pv[2 * i] - hf.x, pv[2 * i + 1] - hf.y);
                }
                *reinterpret_cast<uint4*>(tile_lo + off) = *reinterpret_cast<const uint4*>(lo);
              }
            }
          }
          if (nown == 0 || attempt == 1) break;
          // every probability below 2^11 (fp16-safe, and the fp32 sums are far from overflow) unless a unit sum says
          // otherwise (NaN-safe: inf - inf cannot occur, -inf inputs give 0)
          const bool move = !(big <= 2048.f);
          if (!__any_sync(0xffffffffu, move)) break;
          // rare path: the exact block max becomes the reference (one extra read of S), O_wg and l are rescaled.
          // tcgen05.ld / st are warp-collective: every lane takes part, rows that stay use alpha = 1
          float mx = -INFINITY;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t raw[32];
            tmem_ld32(tmem_Sown + cc * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = (full || cc * 32 + i < kv_valid) ? __uint_as_float(raw[i]) : -INFINITY;
              mx = fmaxf(mx, v);
            }
          }
          const float alpha = move ? fast_exp2((m - mx) * c2) : 1.f;  // (m == -inf: nothing accumulated yet, 0)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t raw[32];
            tmem_ld32(tmem_Oown + cc * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
            tmem_st32(tmem_Oown + cc * 32, raw);
          }
          tmem_st_wait();
          if (move) {
            l *= alpha;
            m = mx;
          }
        }
        tc_fence_before();
        mbar_arrive(&s_free[wg]);   // S_wg may be overwritten by this warpgroup's next Q K^T
        fence_proxy_async_smem();   // make P visible to the tensor-core (async) proxy
        mbar_arrive(&p_full[wg]);
        l += rowsum;
#ifdef DFU_TRACE
        tr_cmp += clock64() - tr_t;
        ++tr_n;
#endif
      }
      // ---- end of the segment: merge the two partial results of every row ----
      if (nown > 0) {
        mbar_wait(&o_full[wg], c & 1);  // own last P V (the partner waits for its own before the barrier below)
        tc_fence_after();
      }
      if (u + nblk >= u_end) DFU_TR_MARK(8);
      // (m, l) of the partner thread of the row through the first 8 bytes of the (now idle) P tile rows; l = 0 marks a
      // warpgroup that had no block in this segment (its accumulator holds stale data and must not be read into the sum)
      *reinterpret_cast<float2*>(tile_hi) = make_float2(m, l);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // only the two warps that share these 32 rows
      const float2 other = *reinterpret_cast<const float2*>(sP + (wg ^ 1) * kTile + prow);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // both have read before either rewrites its tile
      tc_fence_after();
      const float m0 = wg == 0 ? m : other.x, m1 = wg == 0 ? other.x : m;
      const float l0 = wg == 0 ? l : other.y, l1 = wg == 0 ? other.y : l;
      const float M = fmaxf(m0, m1);   // finite: every row has at least one valid key in the segment
      const float w0 = l0 > 0.f ? fast_exp2((m0 - M) * c2) : 0.f;
      const float w1 = l1 > 0.f ? fast_exp2((m1 - M) * c2) : 0.f;
      const float L = l0 * w0 + l1 * w1;
      float acc[32];
      {
        uint32_t raw[32];
        tmem_ld32(tmem_O + lane_off + wg * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = w0 > 0.f ? __uint_as_float(raw[i]) * w0 : 0.f;
        tmem_ld32(tmem_O + 64 + lane_off + wg * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = w1 > 0.f ? fmaf(__uint_as_float(raw[i]), w1, acc[i]) : acc[i];
      }
      tc_fence_before();
      mbar_arrive(&o_free);  // the next segment's first P V may overwrite the accumulators
      const int q = q0 + r;
      if (nblk < p.nblk) {
        // a piece of an item: unnormalised O and (M, L) to this CTA's slot (1 = the piece starts the item)
        const size_t slot = static_cast<size_t>(blockIdx.x) * 2 + (jb0 == 0 ? 1 : 0);
        float4* po = reinterpret_cast<float4*>(p.ws_o + (slot * 128 + r) * 64 + wg * 32);
#pragma unroll
        for (int uu = 0; uu < 8; ++uu)
          __stcg(po + uu, make_float4(acc[4 * uu], acc[4 * uu + 1], acc[4 * uu + 2], acc[4 * uu + 3]));
        if (wg == 0) __stcg(reinterpret_cast<float2*>(p.ws_ml + (slot * 128 + r) * 2), make_float2(M, L));
      } else if (q < p.Nq) {
        const float inv = 1.0f / L;
        __half* dst = p.out + (static_cast<size_t>(b) * p.Nq + q) * p.ldo + head * kD + wg * 32;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          __align__(16) __half2 h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(acc[uu * 8 + 2 * i] * inv, acc[uu * 8 + 2 * i + 1] * inv);
          *reinterpret_cast<uint4*>(dst + uu * 8) = *reinterpret_cast<const uint4*>(h);
          if (planes == 2) {
            __align__(16) __half2 lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 hf = __half22float2(h[i]);
              lo[i] = __floats2half2_rn(acc[uu * 8 + 2 * i] * inv - hf.x, acc[uu * 8 + 2 * i + 1] * inv - hf.y);
            }
            *reinterpret_cast<uint4*>(dst + p.out_plane_stride + uu * 8) = *reinterpret_cast<const uint4*>(lo);
          }
        }
      }
      g += nblk;
      u += nblk;
    }
#ifdef DFU_TRACE
    if (_tr) {  // (thread 0 = warpgroup 0, row 0) sums over its own blocks
      _tr[12] = tr_ws;
      _tr[13] = tr_wo;
      _tr[14] = tr_cmp;
      _tr[15] = tr_n;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  DFU_TR_END();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// (Letting the CTA that delivers an item's LAST piece merge in place — arrival counter, fence, three barriers, L2
// re-read — was built and measured in the captured step: 71 vs 51 + 9 us for level-0 self-attention.  It lost: every CTA
// pays the fence + atomic round trip twice, in the middle of its range.)
// Combine the pieces of every item that was cut across CTAs: O = sum_s O_s 2^((m_s - M) c2) / sum_s l_s 2^((m_s - M) c2),
// pieces in key order (deterministic).  One 4-column quad per thread; every load of a thread (<= 8 pieces x ((m, l) +
// O quad)) is issued before the first use, so the kernel costs one L2 round trip.  Items that one CTA covered
// completely were written by attn_fwd_kernel and are skipped.
constexpr int kMaxKvSplits = 8;
__global__ void __launch_bounds__(256) attn_merge_kernel(AttnParams p) {
  pdl_trigger();
  DFU_TR_BEGIN(TR_ATTN_MERGE);
  pdl_wait();
  DFU_TR_MARK(6);
  const int gidx = blockIdx.x * 256 + threadIdx.x;  // (item, row, quad)
  const int item = gidx >> 11;
  const int r = (gidx >> 4) & 127, cq = gidx & 15;
  const int q_tile = item % p.q_tiles;
  const int bh = item / p.q_tiles;
  const int head = bh % p.heads, b = bh / p.heads;
  const int q = q_tile * kBQ + r;
  const long long u0 = static_cast<long long>(item) * p.nblk;
  const int c_first = attn_cta_of(p, u0), c_last = attn_cta_of(p, u0 + p.nblk - 1);
  const int pieces = c_last - c_first + 1;
  if (q < p.Nq && pieces > 1) {
    float2 ml[kMaxKvSplits];
    float4 t[kMaxKvSplits];
#pragma unroll
    for (int s = 0; s < kMaxKvSplits; ++s)
      if (s < pieces) {
        const int c = c_first + s;
        const size_t slot = static_cast<size_t>(c) * 2 + (attn_range_begin(p, c) <= u0 ? 1 : 0);
        ml[s] = __ldcg(reinterpret_cast<const float2*>(p.ws_ml + (slot * 128 + r) * 2));
        t[s] = __ldcg(reinterpret_cast<const float4*>(p.ws_o + (slot * 128 + r) * 64 + cq * 4));
      }
    float M = -INFINITY;
#pragma unroll
    for (int s = 0; s < kMaxKvSplits; ++s)
      if (s < pieces) M = fmaxf(M, ml[s].x);
    float L = 0.f;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < kMaxKvSplits; ++s)
      if (s < pieces) {
        const float w = fast_exp2((ml[s].x - M) * p.scale_log2);
        L += ml[s].y * w;
        o.x += t[s].x * w; o.y += t[s].y * w; o.z += t[s].z * w; o.w += t[s].w * w;
      }
    const float inv = 1.0f / L;
    o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
    __half* dst = p.out + (static_cast<size_t>(b) * p.Nq + q) * p.ldo + head * kD + cq * 4;
    __align__(8) __half2 h[2];
    h[0] = __floats2half2_rn(o.x, o.y);
    h[1] = __floats2half2_rn(o.z, o.w);
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(h);
    if (p.planes == 2) {
      const float2 a = __half22float2(h[0]), c = __half22float2(h[1]);
      __align__(8) __half2 lo[2];
      lo[0] = __floats2half2_rn(o.x - a.x, o.y - a.y);
      lo[1] = __floats2half2_rn(o.z - c.x, o.w - c.y);
      *reinterpret_cast<uint2*>(dst + p.out_plane_stride) = *reinterpret_cast<const uint2*>(lo);
    }
  }
  DFU_TR_END();
}

static int make_seq_map(CUtensorMap* m, const void* base, int ld, int N, int B, int planes, long long plane_stride,
                        int box_rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(N), static_cast<uint64_t>(B),
                      static_cast<uint64_t>(planes)};
  uint64_t str[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(N) * ld * 2,
                     static_cast<uint64_t>(plane_stride) * 2};
  uint32_t box[4] = {kD, static_cast<uint32_t>(box_rows), 1, 1};
  return make_tmap_f16(m, base, 4, dims, str, box);
}

}  // namespace dfu

DFU_TRACE_SETTER(dfu_trace_set_attn)

using namespace dfu;

// Number of CTAs G for the stream-K distribution.  kv_splits > 0: items * kv_splits (an explicit even cut, 1 = no
// cut); kv_splits <= 0: automatic.
static int attn_grid(int B, int heads, int Nq, int Nk, int kv_splits) {
  const int q_tiles = (Nq + kBQ - 1) / kBQ;
  const int nblk = (Nk + kBKV - 1) / kBKV;
  const long long items = static_cast<long long>(q_tiles) * heads * B;
  const long long total = items * nblk;
  long long G;
  if (kv_splits > 0) {
    G = items * (kv_splits < nblk ? kv_splits : nblk);
  } else {
    // Measured in the captured UNet step (scripts/trace_step.py): two CTAs per SM interleave better than one, so the
    // goal is two equally loaded CTAs per SM with at least four 64-key blocks each; many-item problems (>= 4 waves)
    // balance by themselves and are not cut.
    const int sms = num_sms() > 0 ? num_sms() : 148;
    if (items >= 4LL * sms || Nk < 1024) {
      // short key sequences (the 577 glyph tokens): the merge launch (~6-8 us) costs more than the imbalance it
      // removes (measured: 18.6 us uncut vs 18.5 + 8.0 us cut at 160 tiles x 577 keys)
      G = items;
    } else {
      G = 2LL * sms;
      if (G > total / 4) G = total / 4;
      if (G < items) G = items;
    }
  }
  if (G > items * 6) G = items * 6;  // <= 7 pieces per item (merge kernel holds 8)
  if (G > total) G = total;
  if (G < 1) G = 1;
  return static_cast<int>(G);
}

// Host-side view of the stream-K distribution (no GPU needed): out[0] = G (CTAs), out[1] = work items, out[2] = 64-key
// blocks per item, out[3] = largest number of pieces any item is cut into, out[4] = 1 if every CTA range is non-empty
// and attn_cta_of inverts attn_range_begin on every range boundary.
extern "C" int dfu_attention_plan(int B, int heads, int Nq, int Nk, int kv_splits, int32_t* out) {
  if (!out || B <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return DFU_ERR_INVALID;
  AttnParams p;
  p.q_tiles = (Nq + kBQ - 1) / kBQ;
  p.nblk = (Nk + kBKV - 1) / kBKV;
  p.items = p.q_tiles * heads * B;
  p.total = static_cast<long long>(p.items) * p.nblk;
  p.G = attn_grid(B, heads, Nq, Nk, kv_splits);
  int max_pieces = 0, ok = 1;
  for (int c = 0; c < p.G; ++c) {
    const long long b0 = attn_range_begin(p, c), b1 = attn_range_begin(p, c + 1);
    if (b0 >= b1 || attn_cta_of(p, b0) != c || attn_cta_of(p, b1 - 1) != c) ok = 0;
  }
  for (int i = 0; i < p.items; ++i) {
    const long long u0 = static_cast<long long>(i) * p.nblk;
    const int pieces = attn_cta_of(p, u0 + p.nblk - 1) - attn_cta_of(p, u0) + 1;
    if (pieces > max_pieces) max_pieces = pieces;
  }
  out[0] = p.G; out[1] = p.items; out[2] = p.nblk; out[3] = max_pieces; out[4] = ok;
  return DFU_OK;
}

extern "C" size_t dfu_attention_workspace(int B, int heads, int Nq, int Nk, int kv_splits) {
  const int q_tiles = (Nq + kBQ - 1) / kBQ;
  const int G = attn_grid(B, heads, Nq, Nk, kv_splits);
  if (G <= q_tiles * heads * B) return 0;  // every item inside one CTA's range: no partial pieces
  return static_cast<size_t>(G) * 2 * (128 * 64 + 128 * 2) * sizeof(float);
}

extern "C" int dfu_attention(const void* q, int ldq, int q_col0, int64_t q_plane_stride, const void* k, int ldk,
                             int k_col0, const void* v, int ldv, int v_col0, int64_t kv_plane_stride, int B, int heads,
                             int Nq, int Nk, int planes, float scale, void* out, int ldo, int64_t out_plane_stride,
                             int kv_splits, void* workspace, size_t workspace_bytes, void* stream_) {
  DFU_REQUIRE(planes == 1 || planes == 2, "attention: planes=%d", planes);
  DFU_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nk > 0, "attention: empty problem");
  DFU_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention: leading dims must be x8");
  DFU_REQUIRE(q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0, "attention: column offsets must be x8");
  CUtensorMap mQ, mK, mV;
  int rc;
  if ((rc = make_seq_map(&mQ, q, ldq, Nq, B, planes, q_plane_stride, kBQ))) return rc;
  if ((rc = make_seq_map(&mK, k, ldk, Nk, B, planes, kv_plane_stride, kBKV))) return rc;
  if ((rc = make_seq_map(&mV, v, ldv, Nk, B, planes, kv_plane_stride, kBKV))) return rc;
  AttnParams p;
  p.B = B; p.heads = heads; p.Nq = Nq; p.Nk = Nk; p.planes = planes;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.out = static_cast<__half*>(out);
  p.ldo = ldo;
  p.out_plane_stride = out_plane_stride;
  p.q_tiles = (Nq + kBQ - 1) / kBQ;
  p.nblk = (Nk + kBKV - 1) / kBKV;
  p.items = p.q_tiles * heads * B;
  p.total = static_cast<long long>(p.items) * p.nblk;
  p.G = attn_grid(B, heads, Nq, Nk, kv_splits);
  p.ws_o = nullptr;
  p.ws_ml = nullptr;
  const bool pieces = p.G > p.items;
  if (pieces) {
    const size_t slots = static_cast<size_t>(p.G) * 2;
    const size_t need = slots * (128 * 64 + 128 * 2) * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("attention: cut key ranges need %zu workspace bytes, got %zu", need, workspace_bytes);
      return DFU_ERR_WORKSPACE;
    }
    p.ws_o = static_cast<float*>(workspace);
    p.ws_ml = p.ws_o + slots * 128 * 64;
  }
  const size_t smem = static_cast<size_t>(planes) * (kTile * 3 + 2 * kRing * kKvTile) + 256;
  if (first_use_on_device(ONCE_ATTN_ATTR)) {
    DFU_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  DFU_CHECK_CUDA(launch_k(attn_fwd_kernel, dim3(p.G), dim3(kAttnThreads), smem, static_cast<cudaStream_t>(stream_), mQ, mK, mV, p));
  if (pieces)
    DFU_CHECK_CUDA(launch_k(attn_merge_kernel, dim3(p.items * 8), dim3(256), 0, static_cast<cudaStream_t>(stream_), p));
  return DFU_OK;
}

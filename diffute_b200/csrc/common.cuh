// diffute_b200 — shared device helpers: mbarrier / TMA / tcgen05 PTX wrappers for sm_100a.
// Hand-written inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dfu {

// ----------------------------------------------------------------------------------------------
// error codes of the C-ABI (include/diffute_b200.h)
// ----------------------------------------------------------------------------------------------
enum : int {
  DFU_OK = 0,
  DFU_ERR_INVALID = -1,     // bad argument / unsupported shape
  DFU_ERR_CUDA = -2,        // a CUDA runtime call failed (see dfu_last_error)
  DFU_ERR_DRIVER = -3,      // driver entry point (cuTensorMapEncodeTiled) unavailable
  DFU_ERR_WORKSPACE = -4,   // caller-provided workspace too small
};

constexpr int kWarp = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (process dies with an error) instead of hanging the GPU box.
// The slow path backs off with nanosleep between probes and counts iterations instead of reading the clock: ncu showed
// the plain probe loop (try_wait returns after ~50 cycles; + clock read, compare, branch) taking 40% of ALL issued
// instructions of the attention kernel — issue slots stolen from the warps doing the exponentials.
__device__ __forceinline__ void mbar_wait_ns(uint64_t* bar, uint32_t parity, unsigned ns, unsigned max_ns = 0) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned spins = 0;
  for (;;) {
    __nanosleep(ns);
    if (mbar_try_wait(bar, parity)) return;
    if (ns < max_ns) ns *= 2;  // exponential back-off for waits that are not latency-critical (TMA producers)
    if (++spins > (1u << 26)) {  // >= 2 s
      printf("dfu: mbarrier timeout block(%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_ns(bar, parity, 32); }
// latency-critical single-thread waits (an MMA issuer between two dependent MMAs): probe back to back
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("dfu: mbarrier timeout block(%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
// same with a suspend-time hint (ns) on the probe itself
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Long waits (epilogue warps waiting for the whole main loop): coarser back-off
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) { mbar_wait_ns(bar, parity, 64); }

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch: every kernel lets its successor start launching at once (its CTAs only fill
// SM slots the running grid no longer needs) and waits for its predecessor right before touching global memory,
// so launch latency and the barrier/TMEM prologue overlap the previous kernel's tail.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// In-kernel timeline tracing (only in the -DDFU_TRACE build, libdiffute_b200_trace.so; the product library compiles
// these macros to nothing).  One 16 x u64 record per CTA: [0] %gridid, [1] tag | bid << 32, [2] smid | nctas << 32,
// [3] globaltimer at entry, [4] clock64 at entry, [5..9] clock64 phase marks, [10] clock64 at exit, [11] globaltimer
// at exit, [12..15] extra clock64 marks.  g_tr[0] = next record index (atomic), g_tr[1] = capacity, records start at g_tr + 8.
// ----------------------------------------------------------------------------------------------
#ifdef DFU_TRACE
static __device__ unsigned long long* g_tr = nullptr;
__device__ __forceinline__ unsigned long long trace_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long* trace_begin(unsigned int tag) {
  unsigned long long* base = g_tr;
  if (!base) return nullptr;
  const unsigned long long idx = atomicAdd(base, 1ull);
  if (idx >= base[1]) return nullptr;
  unsigned long long* r = base + 8 + idx * 16;
  unsigned long long gid;
  unsigned int smid;
  asm volatile("mov.u64 %0, %%gridid;" : "=l"(gid));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  r[0] = gid;
  r[1] = tag | (static_cast<unsigned long long>(bid) << 32);
  r[2] = smid | (static_cast<unsigned long long>(gridDim.x * gridDim.y * gridDim.z) << 32);
  r[3] = trace_gtime();
  r[4] = clock64();
#pragma unroll
  for (int i = 5; i < 16; ++i) r[i] = 0;
  return r;
}
__device__ __forceinline__ void trace_mark(unsigned long long* r, int slot) {
  if (r) r[slot] = clock64();
}
__device__ __forceinline__ void trace_end(unsigned long long* r) {
  if (r) {
    r[10] = clock64();
    r[11] = trace_gtime();
  }
}
// per-thread flavour (simple kernels: thread 0 owns the record in a register)
#define DFU_TR_BEGIN(tag) unsigned long long* _tr = (threadIdx.x == 0 && threadIdx.y == 0) ? dfu::trace_begin(tag) : nullptr
#define DFU_TR_MARK(slot) dfu::trace_mark(_tr, slot)
#define DFU_TR_END() dfu::trace_end(_tr)
// CTA-shared flavour (warp-specialised kernels: the record pointer lives in shared memory, any thread may mark)
#define DFU_TR_SHARED_DECL() __shared__ unsigned long long* _trs
#define DFU_TR_SHARED_BEGIN(tag) do { if (threadIdx.x == 0) _trs = dfu::trace_begin(tag); } while (0)
#define DFU_TR_SHARED_MARK(slot) dfu::trace_mark(_trs, slot)
#define DFU_TR_SHARED_END() dfu::trace_end(_trs)
#define DFU_TRACE_SETTER(name) \
  extern "C" int name(void* p) { return cudaMemcpyToSymbol(dfu::g_tr, &p, sizeof(p)) == cudaSuccess ? 0 : -2; }
#else
#define DFU_TR_BEGIN(tag) do {} while (0)
#define DFU_TR_MARK(slot) do {} while (0)
#define DFU_TR_END() do {} while (0)
#define DFU_TR_SHARED_DECL() do {} while (0)
#define DFU_TR_SHARED_BEGIN(tag) do {} while (0)
#define DFU_TR_SHARED_MARK(slot) do {} while (0)
#define DFU_TR_SHARED_END() do {} while (0)
#define DFU_TRACE_SETTER(name) extern "C" int name(void*) { return -1; }
#endif
// kernel tags of the trace records
enum : unsigned int {
  TR_GEMM = 1, TR_SPLITK_REDUCE, TR_ATTN, TR_ATTN_MERGE, TR_GN_STATS, TR_GN_FINALIZE, TR_GN_APPLY, TR_GN_FUSED,
  TR_LAYERNORM, TR_CAST, TR_TEMB, TR_GEMV, TR_CONV_IN, TR_CONV_OUT, TR_MISC
};

// ----------------------------------------------------------------------------------------------
// thread-block clusters: rank, cluster-wide barrier, distributed shared memory loads
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive() {  // split barrier: arrive now ...
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {  // ... wait later
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ double ld_dsmem_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  // volatile (ordered against the volatile cluster barriers) but no memory clobber: independent global loads and
  // stores around it may be scheduled freely, and consecutive remote loads pipeline
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}

// ----------------------------------------------------------------------------------------------
// proxies / fences
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {  // generic-proxy smem writes -> async proxy (UMMA/TMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, mbarrier completion
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM alloc, MMA (kind::f16, cta_group::1), commit, ld
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_128B, tile rows of 128 bytes (64 x 16-bit), 8-row groups 1024 B apart.
// Bit layout (sm_100 "version 1"): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
// [61,64) layout type (2 = SWIZZLE_128B).  Valid for K-major operands and for MN-major operands whose
// MN extent is one 64-element swizzle row (LBO unused in both cases).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: fp16 (or bf16) operands, fp32 accumulate, M x N tile.
// [4,6) c_format=1 (F32), [7,10) a_format, [10,13) b_format (0=F16, 1=BF16), [15] a_major, [16] b_major
// (0 = K-major, 1 = MN-major), [17,23) N>>3, [24,29) M>>4.
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N, int b_mn_major = 0, int bf16 = 0) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (bf16 ? 1u : 0u) << 7;
  d |= (bf16 ? 1u : 0u) << 10;
  d |= (b_mn_major ? 1u : 0u) << 16;
  d |= (static_cast<uint32_t>(N) >> 3) << 17;
  d |= (static_cast<uint32_t>(M) >> 4) << 24;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// cta_group::2 — a CTA pair (cluster of 2 on one TPC) executes ONE tcgen05.mma of M = 256: each CTA supplies its 128
// rows of A and its half of B (N/2 rows) from its own shared memory and receives its 128 x N slice of D in its own
// TMEM.  Only the leader (cluster rank 0) issues; commits are multicast to the barriers of both CTAs; TMA loads of the
// peer signal the LEADER's full barrier (the .cta_group::2 form of cp.async.bulk.tensor allows a peer-CTA mbarrier).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in every CTA of `cta_mask` when all MMAs issued so far are done
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// TMA loads whose completion is signalled on `bar_cluster_addr`, a shared::cluster address (own CTA or the pair's peer)
__device__ __forceinline__ void tma_load_2d_cg2(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive (release at cluster scope) on an mbarrier given by its shared::cluster address (may be the peer CTA's)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

// TMEM -> registers: the warp reads 32 lanes x 32 consecutive fp32 columns; thread i gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: thread i writes lane (base_lane + i), 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// one fp32 per thread: TMEM[lane(thread)][column] <-> register
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// small math helpers
// ----------------------------------------------------------------------------------------------
// x * sigmoid(x) with the approximate divide (2 ulp, no slow-path call: the IEEE '/' cost ~35 instructions and a
// divergent subroutine per element, a third of the GroupNorm+SiLU instruction stream)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// exact-erf GELU (diffusers GEGLU uses F.gelu default = erf form).  erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7
// absolute, below the fp16 rounding of the result by three orders of magnitude) in ~12 instructions instead of libm
// erff's ~45: the GEGLU epilogue is instruction-issue bound.
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));  // argument in [1, inf): no special cases
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = __expf(-ax * ax);
  const float r = fmaf(-p * t, e, 1.0f);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f)); }

// RN split of an fp32 value into fp16 hi + fp16 lo (x ~= hi + lo to ~22 bits)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dfu

// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (UMMA + TMEM).
// Everything here is device-side plumbing shared by the kernels of libpnpvcve.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pnp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a pipeline bug must surface as a trapped launch, not as a hung GPU.
#ifndef PNP_SPIN_LIMIT
#define PNP_SPIN_LIMIT (1u << 26)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > PNP_SPIN_LIMIT) {
      printf("pnp: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}

// One arrival per warp (barrier count = number of warps): every lane has finished its own work
// (tcgen05.wait::ld + fence) before __syncwarp, lane 0 then speaks for the warp.  256 per-thread
// arrivals per tile were a measurable load on the mbarrier unit.
__device__ __forceinline__ void warp_arrive(uint32_t bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

// Patient variant for the many threads that merely wait for work (epilogue warps, TMA producer):
// one lane polls with a hardware suspend-time hint and the warp re-converges.  Measured on B200
// (profiles/r01_notes.md): with all 256 epilogue lanes spinning on try_wait, an already-complete
// mbarrier wait of the MMA-issuing thread took ~375 cycles instead of ~125 and the N=192 MMAs
// themselves ran ~25 % slower -- mbarrier polling competes for the shared-memory pipe.
// single-thread form (the caller is one elected lane, or wants every lane to poll patiently)
__device__ __forceinline__ void mbar_wait_patient(uint32_t bar, uint32_t parity, int tag = 0, uint32_t hint_ns = 200u) {
  uint32_t spins = 0, ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)      // suspend-time hint in ns
        : "memory");
    if (!ok && ++spins > (PNP_SPIN_LIMIT >> 4)) {
      printf("pnp: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  } while (!ok);
}
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int tag = 0, uint32_t hint_ns = 200u) {
  if ((threadIdx.x & 31) == 0) mbar_wait_patient(bar, parity, tag, hint_ns);
  __syncwarp();
}

// ------------------------------------------------------------------ fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load: coordinates fastest-first (c, x, y, n); out-of-range elements are zero filled.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// same with an L2 eviction policy (createpolicy): used for operands that are dead after this read
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                 int c2, int c3, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}
// 4-D tiled store (smem -> global); out-of-range elements are dropped.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// same with an L2 eviction policy for the written lines
__device__ __forceinline__ void tma_store_4d_hint(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// 1-D bulk copy global -> smem (16-byte multiples), completes on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
      : "memory");
}

// ------------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread l of the warp gets TMEM lane (base_lane + l).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ UMMA (tcgen05.mma)
// Shared-memory matrix descriptor, K-major operand stored as 128-byte rows with the 128B swizzle
// (the layout TMA writes for a box whose inner extent is 64 bf16): 8-row atoms of 1024 bytes,
// SBO = 1024.  `start` may be any 16-byte aligned address inside a 1024-aligned tile: measured on
// B200 (round-1 probe, profiles/r01_probe.log) the hardware applies the swizzle XOR to the
// absolute shared-memory address bits, so a view that starts k pixels (k*128 bytes) into a TMA-written
// tile reads the right data with base_offset = 0; setting base_offset = (start >> 7) & 7 as the PTX
// ISA text suggests for unaligned starts double-counts the phase and returns garbage.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t start, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((start & 0x3FFFFu) >> 4);            // bits [0,14)  start address
  d |= static_cast<uint64_t>(1) << 16;                            // bits [16,30) LBO (unused, =1)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // bits [32,46) SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;                            // bits [46,48) descriptor version
  d |= static_cast<uint64_t>(base_offset & 7u) << 49;             // bits [49,52) base offset
  d |= static_cast<uint64_t>(2) << 61;                            // bits [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major, M=128.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4)            // c_format = F32
         | (1u << 7)          // a_format = BF16
         | (1u << 10)         // b_format = BF16
         | ((n >> 3) << 17)   // N >> 3
         | ((m >> 4) << 24);  // M >> 4
}
// True for exactly one lane of a converged warp.  Guarding uniform-datapath instructions (UTCHMMA,
// UTMALDG, UTMASTG) with elect.sync instead of `lane == 0` lets ptxas emit them once instead of
// wrapping each in an elect-and-branch loop over the possibly-active lanes.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// High word shared by every K-major SWIZZLE_128B descriptor here: SBO=1024, version 1, swizzle 128B.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
// K-major operand stored as 64-byte rows with the 64B swizzle (what TMA writes for a box whose inner extent is 32 bf16):
// 8-row atoms of 512 bytes, SBO = 512, layout type SWIZZLE_64B.  K = 32 per row: the two K=16 slices are 32 bytes apart.
constexpr uint32_t kDescHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);
// Low word: start address (>>4) and LBO=1.  Advancing by n bytes == adding n/16 to the low word.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t start) {
  return ((start & 0x3FFFFu) >> 4) | (1u << 16);
}
// D[tmem] (+)= A[smem] * B[smem]^T from descriptor words; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// ------------------------------------------------------------------ CTA pairs (tcgen05 cta_group::2)
// M = 256 across the two CTAs of a cluster: each CTA supplies its own 128 rows of A and HALF of the rows of B from its
// own shared memory (same offsets in both CTAs), the accumulator rows of a CTA's pixels land in its own TMEM.  Issued
// by one thread of the leader CTA (cluster rank 0).  kColl selects the A-operand collector: 0 none, 1 fill (keep A
// after this MMA), 2 use (re-use the kept A and keep it), 3 lastuse.  Measured (tools/umma_collect_bench.cu,
// profiles/r02_umma_collect.log): three N=64 MMAs chained fill/use/lastuse cost 110 cycles against 98 for one N=192
// MMA and 132-146 without the collector -- the 4 KB A tile is fetched from shared memory once.
template <int kColl>
__device__ __forceinline__ void umma2_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
#define PNP_UMMA2(str)                                                                                          \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"   \
               "setp.ne.b32 p, %6, 0;\n\t" str " [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),                       \
               "r"(a_lo), "r"(kDescHiSw128), "r"(b_lo), "r"(kDescHiSw128), "r"(idesc), "r"(accumulate)           \
               : "memory")
  if constexpr (kColl == 0) PNP_UMMA2("tcgen05.mma.cta_group::2.kind::f16");
  if constexpr (kColl == 1) PNP_UMMA2("tcgen05.mma.cta_group::2.kind::f16.collector::a::fill");
  if constexpr (kColl == 2) PNP_UMMA2("tcgen05.mma.cta_group::2.kind::f16.collector::a::use");
  if constexpr (kColl == 3) PNP_UMMA2("tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse");
#undef PNP_UMMA2
}
// the same collector variants for single-CTA MMAs
template <int kColl>
__device__ __forceinline__ void umma1_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
#define PNP_UMMA1(str)                                                                                          \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"   \
               "setp.ne.b32 p, %6, 0;\n\t" str " [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),                       \
               "r"(a_lo), "r"(kDescHiSw128), "r"(b_lo), "r"(kDescHiSw128), "r"(idesc), "r"(accumulate)           \
               : "memory")
  if constexpr (kColl == 0) PNP_UMMA1("tcgen05.mma.cta_group::1.kind::f16");
  if constexpr (kColl == 1) PNP_UMMA1("tcgen05.mma.cta_group::1.kind::f16.collector::a::fill");
  if constexpr (kColl == 2) PNP_UMMA1("tcgen05.mma.cta_group::1.kind::f16.collector::a::use");
  if constexpr (kColl == 3) PNP_UMMA1("tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse");
#undef PNP_UMMA1
}
// commit of a CTA pair's MMAs: the mbarrier at this shared-memory offset arrives in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {  // warp 0 of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {   // warp 0 of BOTH CTAs
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 4-D tiled load into this CTA's shared memory whose completion is signalled on an mbarrier of the pair's leader
// (`cluster_bar`: shared::cluster address, mapa_shared(bar, 0))
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Arrivals / counter bumps that cross to the other CTA of the cluster.  NOT `.release.cluster`: that form compiles to
// MEMBAR.ALL.GPU + ERRBAR in front of the arrive (and `.acquire.cluster` waits to a CCTL.IVALL behind it), which made the
// first pair kernel 75 % slower than the single-CTA one.  What these arrivals hand over are TMEM reads already completed
// by tcgen05.wait::ld (+ tcgen05.fence::before_thread_sync), as in CUTLASS's 2-SM kernels (ClusterBarrier::arrive).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void red_add_cluster(uint32_t cluster_addr) {
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], 1;" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ------------------------------------------------------------------ flag hand-off in shared memory
// A "scout" thread does the (slow, ~250 cycle) mbarrier waits and publishes progress counters; the
// MMA-issuing thread only needs these ~30-cycle acquire loads on its critical path.
__device__ __forceinline__ void st_release_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// counter variant: every lane has finished its reads, lane 0 adds 1 for the warp.
// RELAXED on purpose: what is being handed over are TMEM reads that tcgen05.wait::ld has already
// completed, and a release here would also wait for this thread's outstanding global loads (the
// partition-map prefetch, ~1000 cycles from HBM) -- measured as a 400-1500 cycle stall per row.
__device__ __forceinline__ void warp_flag_add(uint32_t addr) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0)
    asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
// same reasoning for "this TMEM accumulator has been read" arrivals
__device__ __forceinline__ void warp_arrive_relaxed(uint32_t bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0)
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void spin_until_ge(uint32_t addr, uint32_t target, int tag) {
  uint32_t spins = 0;
  while (ld_acquire_shared(addr) < target) {
    if (++spins > PNP_SPIN_LIMIT) {
      printf("pnp: flag wait timed out (tag %d, block %d, target %u)\n", tag, (int)blockIdx.x, target);
      __trap();
    }
  }
}

// same, returning the value that satisfied the wait (callers cache it and skip later polls)
__device__ __forceinline__ uint32_t spin_until_ge_v(uint32_t addr, uint32_t target, int tag) {
  uint32_t spins = 0, v;
  while ((v = ld_acquire_shared(addr)) < target) {
    if (++spins > PNP_SPIN_LIMIT) {
      printf("pnp: flag wait timed out (tag %d, block %d, target %u)\n", tag, (int)blockIdx.x, target);
      __trap();
    }
  }
  return v;
}

// ------------------------------------------------------------------ thread-block clusters / DSMEM
// Cluster / DSMEM helpers (the CTA-pair conv form; first used by round 1's fused residual-block kernel).  Measured on B200 (round-1 microbenchmark,
// profiles/r01_dsmem_bench.log): the SM-to-SM path moves 21.3 B/cycle with 512 contiguous bytes per
// warp store or with cp.async.bulk, but only 10.7 B/cycle when each lane writes 16 B at a 128 B stride.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa_shared)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar),
               "r"(bytes)
               : "memory");
}
// wait on a LOCAL mbarrier whose arrivals come from another CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag = 0) {
  uint32_t spins = 0, ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > PNP_SPIN_LIMIT) {
      printf("pnp: cluster mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  } while (!ok);
}
// bulk copy own shared memory -> shared memory of another CTA; completes (complete_tx) on an mbarrier
// of the destination CTA
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t cluster_dst, uint32_t src, uint32_t bytes,
                                                     uint32_t cluster_bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(cluster_dst), "r"(src), "r"(bytes), "r"(cluster_bar)
      : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ------------------------------------------------------------------ programmatic dependent launch
// launch_dependents: the next kernel in the stream (launched with the programmatic-serialization
// attribute) may start its prologue; wait: block until the previous kernel has completed and its
// memory is visible.  Everything that touches activations comes after griddep_wait().
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------ register re-allocation between warpgroups
// (whole warpgroups of 4 consecutive warps must execute the same instruction)
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------ misc
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// max(x, 0) fused into the conversion
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace pnp

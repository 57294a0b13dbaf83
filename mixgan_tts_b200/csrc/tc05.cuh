// Thin inline-PTX wrappers for the sm_100a features the bf16 path uses: mbarrier, bulk async
// copies (TMA engine, 1-D form), tcgen05 tensor-core MMA with TMEM accumulators, and the
// shared-memory matrix descriptors.  Everything here is device-only and header-only.
//
// Shared-memory operand convention used throughout this repo (K-major, SWIZZLE_NONE):
//   element (row r, k) of an operand tile lives at
//        base + (k / 8) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 8) * 2      [bf16]
//   i.e. 8x8 "core matrices" of 128 contiguous bytes; LBO = byte stride between core matrices
//   along K, SBO = byte stride between 8-row groups.  With SBO = 128 the rows of one 8-channel
//   chunk are a dense [rows][16 B] array, so a +16 B start address shifts the tile by one row —
//   that is how the k=3 convolution taps read the same activation tile three times.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: returns false if the phase did not complete within `max_cycles` SM cycles.
// Every wait in this repo is bounded so that a protocol bug ends the kernel instead of hanging
// the GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, long long max_cycles) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > max_cycles) return false;
  }
  return true;
}

// Hot-path wait: one try_wait on the fast path and nothing that depends on its outcome afterwards
// (so callers keep warp-uniform state in uniform registers).  If the phase does not complete within
// `max_cycles` the watchdog word is set and the kernel traps — a protocol bug fails loudly instead
// of hanging the GPU.  `bar_addr` is a shared::cta address.
__device__ __forceinline__ bool mbar_try_wait_addr(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity, long long max_cycles, int* status,
                                            int code) {
  const long long t0 = clock64();
  while (!mbar_try_wait_addr(bar_addr, parity)) {
    if (clock64() - t0 > max_cycles) {
      if (status) atomicOr(status, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_trap(uint32_t bar_addr, uint32_t parity, long long max_cycles, int* status,
                                               int code) {
  if (!mbar_try_wait_addr(bar_addr, parity)) mbar_wait_slow(bar_addr, parity, max_cycles, status, code);
}

// address-based variants (shared::cta u32 addresses) for the hot loops
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_addr(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_addr(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar_addr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(bar_addr)
               : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- bulk async copy (TMA, 1-D)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Multicast form: the bytes land at the same CTA-relative offset in every CTA of `cta_mask` and each destination
// CTA's mbarrier (same CTA-relative offset) receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_multicast_addr(uint32_t smem_dst, const void* gmem_src, uint32_t bytes,
                                                        uint32_t bar_addr, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar_addr), "h"(cta_mask)
      : "memory");
}
// cta_group::1 commit whose arrive is multicast to the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_mc_addr(uint32_t bar_addr, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar_addr), "h"(cta_mask)
      : "memory");
}
// Tensor-map TMA, 2-D tile: one instruction moves a whole box (measured on B200: a 2 KB cp.async.bulk costs ~60-110 cycles
// of TMA-engine time whatever its size, so many small 1-D copies fill shared memory at ~20-30 B/clk/SM while one 16-64 KB box
// reaches 75-125 B/clk/SM; profiles/r01/bulk_copy_rate_probe.txt).  `tmap` = generic address of a __grid_constant__ CUtensorMap.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar_addr)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* tmap, int c0, int c1, int c2, uint32_t bar_addr) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar_addr)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar_addr,
                                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar_addr), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {  // one full warp
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of 2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 in, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Warp-collective TMEM -> registers: lane i of the warp reads TMEM lane (base_lane + i),
// 32 consecutive fp32 columns.  `taddr` = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
// Warp-collective registers -> TMEM, the mirror image of tmem_ld32 (32 consecutive fp32 columns of the warp's 32 lanes)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// K-major, no swizzle (layout_type 0), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4)                       // D format: F32
         | (1u << 7) | (1u << 10)        // A, B format: BF16
         | (0u << 15) | (0u << 16)       // A, B K-major
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// kind::f16 with both operands fp16 (format 0) or bf16 (format 1): same instruction, same rate; fp16 carries an 11-bit
// significand (TF32's), bf16 an 8-bit one.
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, bool fp16) {
  return (1u << 4) | ((fp16 ? 0u : 1u) << 7) | ((fp16 ? 0u : 1u) << 10)
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// kind::tf32: 4-byte operands (the tensor core reads sign, exponent and the top 10 mantissa bits), fp32 accumulate, K = 8 per
// instruction = 32 bytes per operand row: the shared-memory descriptors are the kind::f16 ones byte for byte (a core matrix is
// 8 rows x 16 bytes = 8 x 4 tf32).  Half the kind::f16 rate.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace tc

// =====================================================================================================
// 2-CTA (cta_group::2) variants: a CTA pair of one cluster shares every tcgen05.mma (M = 256 = 2 x 128
// rows, each CTA's accumulator rows in its own TMEM, each CTA's shared memory supplying its own A rows
// and HALF of the B rows).  Only the leader CTA (cluster rank 0) issues MMAs and commits.
// =====================================================================================================
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// arrive on a barrier that may live in another CTA of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// Remote arrive with the default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id):
// a pure signal.  A cluster-scope release makes the issuing warp wait for a cluster-wide memory fence (measured:
// ~1000 cycles per arrive on B200), which is only needed when the arrive must publish REMOTE shared-memory stores.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// Remote (DSMEM) store that reports its bytes to an mbarrier in the DESTINATION CTA (complete_tx): the consumer
// waits on that barrier instead of the producer paying a cluster-scope fence.  Both addresses are shared::cluster.
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t cluster_bar_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar_addr) : "memory");
}
// wait on a local barrier whose arrivals may come from the peer CTA (acquire at cluster scope)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster_trap(uint32_t bar_addr, uint32_t parity, long long max_cycles,
                                                       int* status, int code) {
  if (mbar_try_wait_cluster(bar_addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar_addr, parity)) {
    if (clock64() - t0 > max_cycles) {
      if (status) atomicOr(status, code);
      __threadfence_system();
      __trap();
    }
  }
}

template <int kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* slot_in_smem) {   // same warp index in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2cta_mc(uint32_t bar_addr) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar_addr), "h"((uint16_t)3)
      : "memory");
}

}  // namespace tc

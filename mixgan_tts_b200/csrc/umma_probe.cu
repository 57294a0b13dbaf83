// tcgen05 descriptor probe: one CTA, raw operand images copied into shared memory, `ksteps`
// tcgen05.mma (M=128, N=n, K=16, bf16 -> fp32 in TMEM), result read back with tcgen05.ld.
// tests/test_umma_probe.py uses it to pin the shared-memory descriptor conventions of tc05.cuh
// (LBO/SBO meaning, +16 B row shift of the start address, K advance) against a host matmul.
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/mixgan_b200_probe.h"
#include "tc05.cuh"

namespace mgb {
namespace {

constexpr long long kProbeTimeout = 200000000LL;  // ~0.1 s of SM cycles

struct ProbeArgs {
  const uint8_t* a_img; const uint8_t* b_img;
  int a_bytes, b_bytes, b_off;
  int a_start, a_lbo, a_sbo, a_kadv;
  int b_start, b_lbo, b_sbo, b_kadv;
  int n, ksteps, use_bulk;
  float* d_out; int* status;
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int status = 0;

  if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
  if (tid == 32) {
    tc::mbar_init(&bar_load, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_barrier_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (p.use_bulk & 1) {
    if (tid == 0) {
      tc::mbar_arrive_expect_tx(&bar_load, (uint32_t)(p.a_bytes + p.b_bytes));
      tc::bulk_g2s(smem, p.a_img, (uint32_t)p.a_bytes, &bar_load);
      tc::bulk_g2s(smem + p.b_off, p.b_img, (uint32_t)p.b_bytes, &bar_load);
    }
    if (!tc::mbar_wait(&bar_load, 0, kProbeTimeout)) status = 1;
  } else {
    const uint4* ga = reinterpret_cast<const uint4*>(p.a_img);
    const uint4* gb = reinterpret_cast<const uint4*>(p.b_img);
    uint4* sa = reinterpret_cast<uint4*>(smem);
    uint4* sb = reinterpret_cast<uint4*>(smem + p.b_off);
    for (int i = tid; i < p.a_bytes / 16; i += blockDim.x) sa[i] = ga[i];
    for (int i = tid; i < p.b_bytes / 16; i += blockDim.x) sb[i] = gb[i];
    tc::fence_proxy_async_smem();
  }
  __syncthreads();

  if (warp == 0 && status == 0) {
    tc::tc_fence_after();
    if (tc::elect_one()) {
      // use_bulk bit 1: A operand is MN-major (M contiguous), bit 2: B operand is MN-major (N contiguous)
      const uint32_t idesc = tc::make_idesc_bf16(128, p.n) | (((uint32_t)p.use_bulk >> 1 & 1u) << 15) |
                             (((uint32_t)p.use_bulk >> 2 & 1u) << 16);
      const uint32_t a0 = tc::smem_u32(smem) + p.a_start;
      const uint32_t b0 = tc::smem_u32(smem + p.b_off) + p.b_start;
      for (int k = 0; k < p.ksteps; ++k) {
        const uint64_t ad = tc::make_smem_desc(a0 + k * p.a_kadv, p.a_lbo, p.a_sbo);
        const uint64_t bd = tc::make_smem_desc(b0 + k * p.b_kadv, p.b_lbo, p.b_sbo);
        tc::umma_bf16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
      }
      tc::umma_commit(&bar_mma);
    }
    __syncwarp();
  }
  if (!tc::mbar_wait(&bar_mma, 0, kProbeTimeout)) status |= 2;
  tc::tc_fence_after();

  if (status == 0) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.n; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < p.n) p.d_out[(size_t)row * p.n + c0 + j] = __uint_as_float(r[j]);
    }
  }
  if (status) atomicOr(p.status, status);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}


// ---- 2-CTA probe: cluster of two CTAs, M = 256 (128 rows per CTA), N = n, B rows split in halves ----
struct Probe2Args {
  const uint8_t* a_img;   // [2][a_bytes]: per-CTA A image (128 rows)
  const uint8_t* b_img;   // [2][b_bytes]: per-CTA B image (n/2 rows)
  int a_bytes, b_bytes, b_off;
  int a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv;
  int n, ksteps;
  float* d_out;           // [256][n]
  int* status;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_probe2_kernel(const Probe2Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_load, bar_peer, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = tc::cluster_ctarank();

  if (warp == 0) tc::tmem_alloc_2cta<256>(&tmem_slot);
  if (tid == 32) {
    tc::mbar_init(&bar_load, 1);
    tc::mbar_init(&bar_peer, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_barrier_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    tc::mbar_arrive_expect_tx(&bar_load, (uint32_t)(p.a_bytes + p.b_bytes));
    tc::bulk_g2s(smem, p.a_img + (size_t)rank * p.a_bytes, (uint32_t)p.a_bytes, &bar_load);
    tc::bulk_g2s(smem + p.b_off, p.b_img + (size_t)rank * p.b_bytes, (uint32_t)p.b_bytes, &bar_load);
  }
  // every CTA waits for its own operands; CTA 1 then tells the leader
  tc::mbar_wait_trap(tc::smem_u32(&bar_load), 0, kProbeTimeout, p.status, 1);
  __syncthreads();
  if (rank == 1 && tid == 0) tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bar_peer), 0));
  if (rank == 0 && warp == 0) {
    tc::mbar_wait_cluster_trap(tc::smem_u32(&bar_peer), 0, kProbeTimeout, p.status, 2);
    tc::tc_fence_after();
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(256, p.n);
      const uint32_t a0 = tc::smem_u32(smem), b0 = tc::smem_u32(smem + p.b_off);
      for (int k = 0; k < p.ksteps; ++k)
        tc::umma_bf16_2cta(tmem, tc::make_smem_desc(a0 + k * p.a_kadv, p.a_lbo, p.a_sbo),
                           tc::make_smem_desc(b0 + k * p.b_kadv, p.b_lbo, p.b_sbo), idesc, k > 0 ? 1u : 0u);
      tc::umma_commit_2cta_mc(tc::smem_u32(&bar_mma));
    }
    __syncwarp();
  }
  tc::mbar_wait_trap(tc::smem_u32(&bar_mma), 0, kProbeTimeout, p.status, 4);
  tc::tc_fence_after();
  const int row = (int)rank * 128 + warp * 32 + lane;
  for (int c0 = 0; c0 < p.n; c0 += 32) {
    uint32_t r[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < p.n) p.d_out[(size_t)row * p.n + c0 + j] = __uint_as_float(r[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  if (warp == 0) tc::tmem_dealloc_2cta<256>(tmem);
}


// ---- issue-rate probe: `reps` back-to-back tcgen05.mma on resident (zero) operands, cycles from the first
// issue to the completion of the final commit.  cta2 != 0 runs a CTA pair with cta_group::2 (M = 256).
struct RateArgs {
  int a_lbo, a_sbo, b_lbo, b_sbo, a_kadv, b_kadv, b_off, n, ksteps, reps, nacc;
  long long* cycles;   // [gridDim.x]
  int* status;
  int fill;            // 0: zero operands; 1: pseudo-random values in (-1, 1) (operand toggling as in a real GEMM)
  int fp16;            // operand format: 0 bf16, 1 fp16
};

template <bool CTA2>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(const RateArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = CTA2 ? tc::cluster_ctarank() : 0u;
  for (int i = tid; i < (p.b_off * 2) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (p.fill) {
    for (int i = tid; i < p.b_off; i += 128) {     // b_off * 2 bytes = b_off 16-bit operands
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const float v = (float)(int)(h & 0xFFFFu) * (1.f / 32768.f) - 1.f;
      if (p.fp16) reinterpret_cast<__half*>(smem)[i] = __float2half_rn(v);
      else reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16_rn(v);
    }
  }
  if (warp == 0) { if (CTA2) tc::tmem_alloc_2cta<512>(&tmem_slot); else tc::tmem_alloc<512>(&tmem_slot); }
  if (tid == 32) { tc::mbar_init(&bar_mma, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  if (CTA2) tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long cyc = 0;
  if (rank == 0 && warp == 0) {
    const uint32_t idesc = tc::make_idesc_16(CTA2 ? 256 : 128, p.n, p.fp16 != 0);
    const uint32_t a0 = tc::smem_u32(smem), b0 = tc::smem_u32(smem + p.b_off);
    const uint64_t ad = tc::make_smem_desc(a0, p.a_lbo, p.a_sbo), bd = tc::make_smem_desc(b0, p.b_lbo, p.b_sbo);
    const long long t0 = clock64();
    if (tc::elect_one()) {
      for (int r = 0; r < p.reps; ++r) {
        const uint32_t d = tmem + (uint32_t)((r % p.nacc) * p.n);
        for (int k = 0; k < p.ksteps; ++k) {
          if (CTA2) tc::umma_bf16_2cta(d, ad + (uint64_t)(((k & 3) * p.a_kadv) >> 4), bd + (uint64_t)(((k & 3) * p.b_kadv) >> 4), idesc, k > 0);
          else tc::umma_bf16(d, ad + (uint64_t)(((k & 3) * p.a_kadv) >> 4), bd + (uint64_t)(((k & 3) * p.b_kadv) >> 4), idesc, k > 0);
        }
      }
      if (CTA2) tc::umma_commit_2cta_mc(tc::smem_u32(&bar_mma)); else tc::umma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait_trap(tc::smem_u32(&bar_mma), 0, 2000000000LL, p.status, 2);
    cyc = clock64() - t0;
    if ((tid & 31) == 0) p.cycles[blockIdx.x] = cyc;
  } else if (CTA2 && warp == 0) {
    tc::mbar_wait_trap(tc::smem_u32(&bar_mma), 0, 2000000000LL, p.status, 4);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CTA2) tc::cluster_sync_all();
  if (warp == 0) { if (CTA2) tc::tmem_dealloc_2cta<512>(tmem); else tc::tmem_dealloc<512>(tmem); }
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" int mgb_probe_umma(const void* a_img, int a_bytes, const void* b_img, int b_bytes, int a_start,
                              int a_lbo, int a_sbo, int a_kadv, int b_start, int b_lbo, int b_sbo, int b_kadv,
                              int n, int ksteps, int use_bulk_copy, float* d_out, int* status_out, void* stream) {
  MGB_REQUIRE(a_img && b_img && d_out && status_out, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(a_bytes > 0 && b_bytes > 0 && a_bytes % 16 == 0 && b_bytes % 16 == 0, MGB_E_ARG,
              "operand images must be positive multiples of 16 bytes");
  MGB_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && ksteps >= 1 && ksteps <= 64, MGB_E_ARG, "bad n/ksteps");
  MGB_REQUIRE(((a_start | a_lbo | a_sbo | a_kadv | b_start | b_lbo | b_sbo | b_kadv) & 15) == 0, MGB_E_ARG,
              "descriptor byte fields must be multiples of 16");
  if (int rc = check_arch()) return rc;
  ProbeArgs p{};
  p.a_img = static_cast<const uint8_t*>(a_img); p.b_img = static_cast<const uint8_t*>(b_img);
  p.a_bytes = a_bytes; p.b_bytes = b_bytes; p.b_off = (int)align_up((size_t)a_bytes, 1024);
  p.a_start = a_start; p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.a_kadv = a_kadv;
  p.b_start = b_start; p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_kadv = b_kadv;
  p.n = n; p.ksteps = ksteps; p.use_bulk = use_bulk_copy; p.d_out = d_out; p.status = status_out;
  const size_t smem = (size_t)p.b_off + align_up((size_t)b_bytes, 1024);
  MGB_REQUIRE(smem <= 200 * 1024, MGB_E_ARG, "operand images too large for shared memory");
  MGB_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

extern "C" int mgb_probe_umma_2cta(const void* a_img, int a_bytes, const void* b_img, int b_bytes, int a_lbo, int a_sbo,
                                   int a_kadv, int b_lbo, int b_sbo, int b_kadv, int n, int ksteps, float* d_out,
                                   int* status_out, void* stream) {
  MGB_REQUIRE(a_img && b_img && d_out && status_out, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(a_bytes > 0 && b_bytes > 0 && a_bytes % 16 == 0 && b_bytes % 16 == 0, MGB_E_ARG,
              "operand images must be positive multiples of 16 bytes");
  MGB_REQUIRE(n >= 32 && n <= 256 && n % 32 == 0 && ksteps >= 1 && ksteps <= 64, MGB_E_ARG, "bad n/ksteps");
  MGB_REQUIRE(((a_lbo | a_sbo | a_kadv | b_lbo | b_sbo | b_kadv) & 15) == 0, MGB_E_ARG,
              "descriptor byte fields must be multiples of 16");
  if (int rc = check_arch()) return rc;
  Probe2Args p{};
  p.a_img = static_cast<const uint8_t*>(a_img); p.b_img = static_cast<const uint8_t*>(b_img);
  p.a_bytes = a_bytes; p.b_bytes = b_bytes; p.b_off = (int)align_up((size_t)a_bytes, 1024);
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.a_kadv = a_kadv; p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_kadv = b_kadv;
  p.n = n; p.ksteps = ksteps; p.d_out = d_out; p.status = status_out;
  const size_t smem = (size_t)p.b_off + align_up((size_t)b_bytes, 1024);
  MGB_REQUIRE(smem <= 200 * 1024, MGB_E_ARG, "operand images too large for shared memory");
  MGB_CUDA_CHECK(cudaFuncSetAttribute(umma_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe2_kernel<<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

// Issue-rate probe (scripts/umma_rate.py): runs `grid` CTAs (or CTA pairs when cta2) each issuing
// reps*ksteps tcgen05.mma on zeroed operands; cycles_out[i] = SM cycles of leader CTA i from first issue to
// commit completion.  Operand regions: A at 0, B at b_off (each b_off bytes, <= 100 KB).
extern "C" int mgb_probe_umma_rate_data(int cta2, int grid, int n, int ksteps, int reps, int nacc, int a_lbo, int a_sbo,
                                        int a_kadv, int b_lbo, int b_sbo, int b_kadv, int b_off, int fill, int fp16,
                                        long long* cycles_out, int* status_out, void* stream);
extern "C" int mgb_probe_umma_rate(int cta2, int grid, int n, int ksteps, int reps, int nacc, int a_lbo, int a_sbo,
                                   int a_kadv, int b_lbo, int b_sbo, int b_kadv, int b_off, long long* cycles_out,
                                   int* status_out, void* stream) {
  return mgb_probe_umma_rate_data(cta2, grid, n, ksteps, reps, nacc, a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv, b_off, 0, 0,
                                  cycles_out, status_out, stream);
}
// The same probe with a choice of operand contents (fill: 0 zeros, 1 pseudo-random values) and operand format (fp16: 0 bf16,
// 1 fp16): tensor-pipe throughput on B200 depends on operand toggling once the board is power-managed.
extern "C" int mgb_probe_umma_rate_data(int cta2, int grid, int n, int ksteps, int reps, int nacc, int a_lbo, int a_sbo,
                                        int a_kadv, int b_lbo, int b_sbo, int b_kadv, int b_off, int fill, int fp16,
                                        long long* cycles_out, int* status_out, void* stream) {
  MGB_REQUIRE(cycles_out && status_out, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && ksteps >= 1 && reps >= 1 && nacc >= 1 && nacc * n <= 512, MGB_E_ARG,
              "bad n/ksteps/reps/nacc");
  MGB_REQUIRE(b_off > 0 && b_off % 1024 == 0 && b_off <= 100 * 1024, MGB_E_ARG, "bad b_off");
  if (int rc = check_arch()) return rc;
  RateArgs p{};
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.a_kadv = a_kadv; p.b_kadv = b_kadv;
  p.b_off = b_off; p.n = n; p.ksteps = ksteps; p.reps = reps; p.nacc = nacc; p.cycles = cycles_out; p.status = status_out;
  p.fill = fill; p.fp16 = fp16;
  const size_t smem = (size_t)2 * b_off;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cta2 ? 2 * grid : grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cta2 ? 2 : 1; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  if (cta2) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(umma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MGB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, umma_rate_kernel<true>, p));
  } else {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(umma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MGB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, umma_rate_kernel<false>, p));
  }
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}


// ---- bulk-copy (TMA engine, 1-D) ingest-rate probe: every CTA streams `total_bytes` from an L2-resident buffer into a
// 2-slot shared-memory ring with copies of `copy_bytes` each (`copies_per_slot` per barrier phase) and reports cycles.
namespace mgb {
namespace {
struct BulkRateArgs { const uint8_t* src; long long src_bytes; int copy_bytes, copies_per_slot, slots, iters; long long* cycles; int* status; };
__global__ void __launch_bounds__(64, 1) bulk_rate_kernel(const BulkRateArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[8];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < p.slots; ++i) tc::mbar_init(&bar[i], 1); tc::fence_barrier_init(); }
  __syncthreads();
  if (tid == 0) {
    const int slot_bytes = p.copy_bytes * p.copies_per_slot;
    const long long span = p.src_bytes - slot_bytes;
    long long off = ((long long)blockIdx.x * 7919 * slot_bytes) % span;
    off &= ~15LL;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters + p.slots; ++it) {
      const int slot = it % p.slots;
      if (it >= p.slots) {                              // wait for the copies issued `slots` iterations ago
        if (!tc::mbar_wait(&bar[slot], ((it / p.slots) - 1) & 1, 400000000LL)) { atomicOr(p.status, 1); break; }
      }
      if (it < p.iters) {
        tc::mbar_arrive_expect_tx(&bar[slot], (uint32_t)slot_bytes);
        for (int c = 0; c < p.copies_per_slot; ++c)
          tc::bulk_g2s(smem + (size_t)slot * slot_bytes + (size_t)c * p.copy_bytes, p.src + off + (size_t)c * p.copy_bytes,
                       (uint32_t)p.copy_bytes, &bar[slot]);
        off += slot_bytes;
        if (off > span) off = 0;
      }
    }
    p.cycles[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace
}  // namespace mgb

extern "C" int mgb_probe_bulk_rate(const void* src, long long src_bytes, int grid, int copy_bytes, int copies_per_slot,
                                   int slots, int iters, long long* cycles_out, int* status_out, void* stream) {
  using namespace mgb;
  MGB_REQUIRE(src && cycles_out && status_out && grid > 0 && copy_bytes % 16 == 0 && copy_bytes > 0 && copies_per_slot > 0 &&
              slots >= 1 && slots <= 8 && iters > 0, MGB_E_ARG, "bad argument");
  const size_t smem = (size_t)copy_bytes * copies_per_slot * slots;
  MGB_REQUIRE(smem <= 200 * 1024 && (long long)copy_bytes * copies_per_slot * 2 < src_bytes, MGB_E_ARG, "ring too large");
  if (int rc = check_arch()) return rc;
  BulkRateArgs a{static_cast<const uint8_t*>(src), src_bytes, copy_bytes, copies_per_slot, slots, iters, cycles_out, status_out};
  MGB_CUDA_CHECK(cudaFuncSetAttribute(bulk_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bulk_rate_kernel<<<grid, 64, smem, static_cast<cudaStream_t>(stream)>>>(a);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

// bf16 tensor-core path of the Denoiser (MGB_PREC_BF16): a group of residual blocks is chained in
// ONE kernel per 128-frame tile.  Weights stream from L2 through a ring of 16 KB shared-memory
// slots filled by the TMA engine (1-D bulk copies of pre-packed operand images); every convolution
// is a set of tcgen05.mma (M=128 frames, N=128 channels, K=16) accumulating in TMEM; the gate, the
// residual update, the conditioner add and (in the tail) the skip/out projections and the posterior
// update are epilogues that read TMEM with tcgen05.ld.  Activations never leave the SM inside a
// group: the conv input and the gate output live in shared memory as bf16 MMA operands, the fp32
// residual stream lives in the epilogue threads' registers, the skip sum accumulates in TMEM.
//
// Reference behaviour restated: ResidualBlock.forward model/blocks.py:1157-1176, Denoiser.forward
// model/modules.py:420-446, q_posterior_sample model/diffusion.py:104-119.
//
// Algebra.  With r = x + d_l (the residual, blocks.py:1166) and u_l = r + c_l (+ s_l) the conv
// input (c_l = Wc_l cond + bc_l, s_l the speaker term), the block is
//     g      = sigmoid(conv3(u_l)[:C]) * tanh(conv3(u_l)[C:])
//     x'     = (Wo_x g + bo_x + r) / sqrt(2),      skip += Wo_s g + bo_s
// and the next conv input follows from the previous one without materialising x or c:
//     u_{l+1} = ( u_l + [g | cond] [Wo_x ; sqrt(2) Wc_{l+1} - Wc_l]^T ) / sqrt(2) + k_l
//     k_l     = (bo_x,l - bc_l - s_l)/sqrt(2) + d_{l+1} + bc_{l+1} + s_{l+1}      (per utterance)
// so one K=512 GEMM per block produces the residual update and the conditioner projection at once.
//
// Tiling.  A kernel launch runs layers [lb, le) for every tile; a tile is 128 consecutive frames of
// one utterance of which the middle 128 - 2*(le-lb) are exact after le-lb k=3 convolutions (halo
// recompute).  Between groups u (fp32) and the partial skip sum are spilled to HBM.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..11 = epilogue (thread = one frame row x half of a 128-column chunk).
#include <cstdlib>

#include "common.cuh"
#include "small_ops.cuh"
#include "tc05.cuh"

namespace mgb {
namespace {

using namespace smallops;

constexpr int C = 256;
constexpr int SLOT_BYTES = 16384;      // one operand image: [8 k-chunks][128 rows][8 bf16]
constexpr int NSLOTS = 6;
constexpr int A_ROWS = 130;            // 128 tile rows + one zero/halo row each side
constexpr uint32_t A_LBO = A_ROWS * 16;
constexpr uint32_t G_LBO = 128 * 16;
constexpr uint32_t W_LBO = 128 * 16;
constexpr uint32_t SBO = 128;
constexpr int SMEM_A = 32 * A_LBO;
constexpr int SMEM_G = 32 * G_LBO;
constexpr int SMEM_SLOTS = NSLOTS * SLOT_BYTES;
constexpr int SMEM_BARS = 256;
constexpr int SMEM_TOTAL = SMEM_A + SMEM_G + SMEM_SLOTS + SMEM_BARS;
constexpr int NTHREADS = 384;
constexpr int MAX_GROUP_LAYERS = 24;
constexpr int COND_PAD_LO = 32;        // zero rows in front of each utterance in the cond image
constexpr long long WAIT_CYCLES = 400000000LL;   // ~0.2 s: a protocol bug ends the kernel, never hangs it

// weight-image slot indices (see pack_images_kernel)
constexpr int W_IN = 0, W_P0 = 4, W_SKIPP = 12, W_OUT = 20, W_LAYER0 = 24, W_PER_LAYER = 72;

constexpr float RSQRT2 = 0.70710678118654752440f;

// barrier indices
enum { B_FULL = 0, B_EMPTY = NSLOTS, B_TFULL = 2 * NSLOTS, B_TEMPTY = B_TFULL + 2, B_AREADY = B_TEMPTY + 2,
       B_GREADY = B_AREADY + 1, B_SKIPDONE = B_GREADY + 4, B_COUNT = B_SKIPDONE + 1 };

struct FusedParams {
  const uint8_t* wimg;          // weight slot images
  const __nv_bfloat16* condT;   // [B][32][Tp][8] bf16, rows shifted by COND_PAD_LO, zero outside [0,T)
  int Tp;
  const float* x_t;             // [B][M][T]
  const float* noise;           // [B][M][T] or null
  float* x_prev;                // [B][M][T] or null
  float* x0_out;                // [B][M][T] or null
  const float* sched;           // [3][K] or null
  const int64_t* t;             // [B]
  int K, clip, n_mel;
  const float* ktab;            // [B][L][C]
  const float* k00;             // [B][C]
  const float* conv_bias;       // [L][4][128]  (per chunk: 64 gate then 64 filter)
  const float* bsum_skip;       // [C]
  const float* b_in;            // [C]
  const float* b_skip;          // [C]
  const float* b_out;           // [128]
  const float* U_in;            // [B*T][C] fp32 u spilled by the previous group
  float* U_out;                 // [B*T][C] fp32 u for the next group (ping-pong: neighbours read U_in meanwhile)
  float* S;                     // [B*T][C] fp32 partial skip sum between groups
  int B, T, L, lb, le, V, halo, tiles_per_utt;
  int* status;
  long long* prof;              // debug (MGB_PROFILE): per-tile cycle counters, 16 per tile
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct Ring {
  int slot = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++slot == NSLOTS) { slot = 0; phase ^= 1; }
  }
};

template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1) fused_group_kernel(const FusedParams p) {
  const long long t_start = PROF ? clock64() : 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sG = smem + SMEM_A;
  uint8_t* sSlots = sG + SMEM_G;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSlots + SMEM_SLOTS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int b = tile / p.tiles_per_utt;
  const int f0 = (tile - b * p.tiles_per_utt) * p.V - p.halo;   // frame of tile row 0
  const bool first_group = p.lb == 0, last_group = p.le == p.L;

  // ---- setup -------------------------------------------------------------------------------
  {  // zero the operand tiles (halo rows of A stay zero for the whole kernel)
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < (SMEM_A + SMEM_G) / 16; i += NTHREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int i = 0; i < NSLOTS; ++i) { tc::mbar_init(&bars[B_FULL + i], 1); tc::mbar_init(&bars[B_EMPTY + i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars[B_TFULL + i], 1); tc::mbar_init(&bars[B_TEMPTY + i], 8); }
    tc::mbar_init(&bars[B_AREADY], 8);
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[B_GREADY + i], 8);
    tc::mbar_init(&bars[B_SKIPDONE], 1);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t TM_SKIP = tmem, TM_TEMP0 = tmem + 256, TM_TEMP1 = tmem + 384;

  long long t_tfull_out = 0;
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // Both roles below run with WARP-UNIFORM control flow (all 32 lanes wait on the barriers and walk
    // the schedule); only the instructions that must come from one thread (bulk copies, tcgen05.mma,
    // tcgen05.commit, expect_tx) are predicated with elect_one().  Waits never feed state back into the
    // loops (a timeout traps), so ring positions and descriptors stay in uniform registers.
    const uint32_t bar0 = tc::smem_u32(bars);
    const uint32_t slots0 = tc::smem_u32(sSlots);
    if (warp == 0) {
      // =========================== TMA PRODUCER ===========================
      uint32_t slot = 0, phase = 0;
      const uint8_t* condb = reinterpret_cast<const uint8_t*>(p.condT) +
                             ((size_t)b * 32 * p.Tp + (size_t)(f0 + COND_PAD_LO)) * 16;
      auto advance = [&]() {
        const bool wrap = slot == NSLOTS - 1;
        slot = wrap ? 0u : slot + 1u;
        phase ^= wrap ? 1u : 0u;
      };
      long long t_empty = 0;
      auto wait_empty = [&]() {
        const long long t0 = PROF ? clock64() : 0;
        tc::mbar_wait_trap(bar0 + (B_EMPTY + slot) * 8, phase ^ 1, WAIT_CYCLES, p.status, 1);
        if (PROF) t_empty += clock64() - t0;
      };
      auto load_w = [&](int widx, uint32_t bytes) {
        wait_empty();
        if (tc::elect_one()) {
          const uint32_t fb = bar0 + (B_FULL + slot) * 8;
          tc::mbar_arrive_expect_tx_addr(fb, bytes);
          tc::bulk_g2s_addr(slots0 + slot * SLOT_BYTES, p.wimg + (size_t)widx * SLOT_BYTES, bytes, fb);
        }
        __syncwarp();
        advance();
      };
      auto load_cond = [&](int kb) {   // cond channels [64kb, 64kb+64) of the tile's 128 frames
        wait_empty();
        if (tc::elect_one()) {
          const uint32_t fb = bar0 + (B_FULL + slot) * 8;
          tc::mbar_arrive_expect_tx_addr(fb, SLOT_BYTES);
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8)
            tc::bulk_g2s_addr(slots0 + slot * SLOT_BYTES + k8 * 2048, condb + (size_t)(kb * 8 + k8) * p.Tp * 16, 2048, fb);
        }
        __syncwarp();
        advance();
      };
      if (first_group) {
        for (int c = 0; c < 2; ++c) { load_w(W_IN + 2 * c, SLOT_BYTES); load_w(W_IN + 2 * c + 1, 4096); }
        for (int c = 0; c < 2; ++c)
          for (int j = 0; j < 4; ++j) { load_cond(j); load_w(W_P0 + 4 * c + j, SLOT_BYTES); }
      }
      for (int l = p.lb; l < p.le; ++l) {
        const int base = W_LAYER0 + l * W_PER_LAYER;
        for (int i = 0; i < 48; ++i) load_w(base + i, SLOT_BYTES);
        if (l < p.L - 1) {
          for (int c = 0; c < 2; ++c) {
            for (int j = 0; j < 4; ++j) { load_cond(j); load_w(base + 48 + c * 8 + j, SLOT_BYTES); }
            for (int j = 0; j < 4; ++j) load_w(base + 48 + c * 8 + 4 + j, SLOT_BYTES);
          }
        }
        for (int i = 0; i < 8; ++i) load_w(base + 64 + i, SLOT_BYTES);
      }
      if (last_group) {
        for (int i = 0; i < 8; ++i) load_w(W_SKIPP + i, SLOT_BYTES);
        for (int i = 0; i < 4; ++i) load_w(W_OUT + i, SLOT_BYTES);
      }
      if (PROF && lane == 0) { p.prof[tile * 16 + 12] = t_empty; p.prof[tile * 16 + 13] = clock64() - t_start; }
    } else if (warp == 1) {
      // =========================== MMA ISSUER ===========================
      uint32_t slot = 0, phase = 0;
      const uint32_t idesc = tc::make_idesc_bf16(128, 128);
      // descriptor templates; a byte offset is added to the 14-bit start-address field (>> 4)
      const uint64_t dA = tc::make_smem_desc(tc::smem_u32(sA), A_LBO, SBO);
      const uint64_t dG = tc::make_smem_desc(tc::smem_u32(sG), G_LBO, SBO);
      const uint64_t dS = tc::make_smem_desc(slots0, W_LBO, SBO);
      uint32_t n_use = 0;            // temp-buffer uses so far; they strictly alternate 0,1,0,1,...
      uint32_t n_aready = 0, n_gready = 0;

      auto advance = [&]() {
        const bool wrap = slot == NSLOTS - 1;
        slot = wrap ? 0u : slot + 1u;
        phase ^= wrap ? 1u : 0u;
      };
      long long t_full = 0, t_temp = 0, t_ar = 0, t_gr = 0;
      auto wait_full = [&]() {
        const long long t0 = PROF ? clock64() : 0;
        tc::mbar_wait_trap(bar0 + (B_FULL + slot) * 8, phase, WAIT_CYCLES, p.status, 2);
        if (PROF) t_full += clock64() - t0;
        tc::tc_fence_after();
      };
      // one weight slot = NK k-steps of K=16 with the A operand from a resident tile (descriptor a0)
      auto mma_w4 = [&](uint64_t a0, uint32_t a_kstep16, uint32_t d_tmem, uint32_t acc_first) {
        wait_full();
        if (tc::elect_one()) {
          const uint64_t b0 = dS + (uint64_t)(slot * (SLOT_BYTES >> 4));
          tc::umma_bf16(d_tmem, a0, b0, idesc, acc_first);
          tc::umma_bf16(d_tmem, a0 + a_kstep16, b0 + (2 * W_LBO >> 4), idesc, 1u);
          tc::umma_bf16(d_tmem, a0 + 2 * a_kstep16, b0 + 2 * (2 * W_LBO >> 4), idesc, 1u);
          tc::umma_bf16(d_tmem, a0 + 3 * a_kstep16, b0 + 3 * (2 * W_LBO >> 4), idesc, 1u);
          tc::umma_commit_addr(bar0 + (B_EMPTY + slot) * 8);
        }
        __syncwarp();
        advance();
      };
      auto mma_w1 = [&](uint64_t a0, uint32_t d_tmem, uint32_t acc_first) {   // a K=16 slot (input projection tail)
        wait_full();
        if (tc::elect_one()) {
          tc::umma_bf16(d_tmem, a0, dS + (uint64_t)(slot * (SLOT_BYTES >> 4)), idesc, acc_first);
          tc::umma_commit_addr(bar0 + (B_EMPTY + slot) * 8);
        }
        __syncwarp();
        advance();
      };
      // a cond slot (A operand) followed by its weight slot
      auto mma_cond = [&](uint32_t d_tmem, uint32_t acc_first) {
        wait_full();
        const uint32_t sa = slot;
        advance();
        wait_full();
        if (tc::elect_one()) {
          const uint64_t a0 = dS + (uint64_t)(sa * (SLOT_BYTES >> 4));
          const uint64_t b0 = dS + (uint64_t)(slot * (SLOT_BYTES >> 4));
          tc::umma_bf16(d_tmem, a0, b0, idesc, acc_first);
          tc::umma_bf16(d_tmem, a0 + (2 * W_LBO >> 4), b0 + (2 * W_LBO >> 4), idesc, 1u);
          tc::umma_bf16(d_tmem, a0 + 2 * (2 * W_LBO >> 4), b0 + 2 * (2 * W_LBO >> 4), idesc, 1u);
          tc::umma_bf16(d_tmem, a0 + 3 * (2 * W_LBO >> 4), b0 + 3 * (2 * W_LBO >> 4), idesc, 1u);
          tc::umma_commit_addr(bar0 + (B_EMPTY + sa) * 8);
          tc::umma_commit_addr(bar0 + (B_EMPTY + slot) * 8);
        }
        __syncwarp();
        advance();
      };
      auto temp_acquire = [&](uint32_t tb) {   // wait until the epilogue has drained the previous use of buffer tb
        const long long t0 = PROF ? clock64() : 0;
        tc::mbar_wait_trap(bar0 + (B_TEMPTY + tb) * 8, ((n_use >> 1) + 1) & 1, WAIT_CYCLES, p.status, 2);
        if (PROF) t_temp += clock64() - t0;
        tc::tc_fence_after();
      };
      auto temp_publish = [&](uint32_t tb) {
        if (tc::elect_one()) tc::umma_commit_addr(bar0 + (B_TFULL + tb) * 8);
        __syncwarp();
        ++n_use;
      };
      auto wait_bar = [&](uint32_t bar, uint32_t n) {
        const long long t0 = PROF ? clock64() : 0;
        tc::mbar_wait_trap(bar0 + bar * 8, n & 1, WAIT_CYCLES, p.status, 2);
        if (PROF) { if (bar == B_AREADY) t_ar += clock64() - t0; else t_gr += clock64() - t0; }
        tc::tc_fence_after();
      };
      auto tm_t = [&](uint32_t tb) { return TM_TEMP0 + tb * 128u; };
      constexpr uint32_t A_K16 = (2 * A_LBO) >> 4, G_K16 = (2 * G_LBO) >> 4;   // descriptor step per K=16

      if (first_group) {
        wait_bar(B_AREADY, n_aready++);                 // x_t tile as bf16, channels 0..79, rows 1..128
        for (uint32_t c = 0; c < 2; ++c) {              // input projection, K = 80
          temp_acquire(c);
          mma_w4(dA + (16 >> 4), A_K16, tm_t(c), 0u);
          mma_w1(dA + ((16 + 8 * A_LBO) >> 4), tm_t(c), 1u);
          temp_publish(c);
        }
        for (uint32_t c = 0; c < 2; ++c) {              // conditioner projection of layer 0
          temp_acquire(c);
          for (int j = 0; j < 4; ++j) mma_cond(tm_t(c), j ? 1u : 0u);
          temp_publish(c);
        }
      }
      for (int l = p.lb; l < p.le; ++l) {
        wait_bar(B_AREADY, n_aready++);                 // conv input u_l in sA
#pragma unroll 1
        for (uint32_t i = 0; i < 4; ++i) {              // k=3 conv, chunk i = 64 gate + 64 filter channels
          const uint32_t tb = i & 1;
          temp_acquire(tb);
          uint32_t acc = 0;
#pragma unroll 1
          for (uint32_t tap = 0; tap < 3; ++tap) {      // tap = +16 B row shift of the start address
            uint64_t a = dA + tap;
#pragma unroll 1
            for (uint32_t kb = 0; kb < 4; ++kb) {       // 64-channel blocks
              mma_w4(a, A_K16, tm_t(tb), acc);
              acc = 1;
              a += (8 * A_LBO) >> 4;
            }
          }
          temp_publish(tb);
        }
        if (l < p.L - 1) {
#pragma unroll 1
          for (uint32_t c = 0; c < 2; ++c) {            // residual-out + conditioner delta, 128 channels each
            temp_acquire(c);
#pragma unroll 1
            for (uint32_t j = 0; j < 4; ++j) mma_cond(tm_t(c), j ? 1u : 0u);
            uint64_t a = dG;
#pragma unroll 1
            for (uint32_t j = 0; j < 4; ++j) {
              if (c == 0) wait_bar(B_GREADY + j, n_gready);
              mma_w4(a, G_K16, tm_t(c), 1u);
              a += (8 * G_LBO) >> 4;
            }
            temp_publish(c);
          }
        } else {
          for (int j = 0; j < 4; ++j) wait_bar(B_GREADY + j, n_gready);
        }
        ++n_gready;
#pragma unroll 1
        for (uint32_t c = 0; c < 2; ++c) {              // skip projection accumulates across layers
          uint64_t a = dG;
#pragma unroll 1
          for (uint32_t j = 0; j < 4; ++j) {
            mma_w4(a, G_K16, TM_SKIP + c * 128, (j == 0 && l == p.lb) ? 0u : 1u);
            a += (8 * G_LBO) >> 4;
          }
        }
      }
      if (tc::elect_one()) tc::umma_commit_addr(bar0 + B_SKIPDONE * 8);
      __syncwarp();
      if (last_group) {
        wait_bar(B_AREADY, n_aready++);                 // skip sum / sqrt(L) as bf16 in sA rows 1..128
        for (uint32_t c = 0; c < 2; ++c) {
          temp_acquire(c);
          for (int j = 0; j < 4; ++j) mma_w4(dA + ((16 + j * 8 * A_LBO) >> 4), A_K16, tm_t(c), j ? 1u : 0u);
          temp_publish(c);
        }
        wait_bar(B_GREADY + 0, n_gready++);             // relu(skip projection) as bf16 in sG
        temp_acquire(0);
        for (int j = 0; j < 4; ++j) mma_w4(dG + ((j * 8 * G_LBO) >> 4), G_K16, tm_t(0), j ? 1u : 0u);
        temp_publish(0);
      }
      if (PROF && lane == 0) {
        long long* q = p.prof + tile * 16;
        q[0] = t_full; q[1] = t_temp; q[2] = t_ar; q[3] = t_gr; q[4] = clock64() - t_start;
      }
    }
  } else {
    // =========================== EPILOGUE (warps 4..11) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int ew = warp - 4;
    const int q = ew & 3;            // TMEM lane quadrant == warp % 4
    const int h = ew >> 2;           // which half of a 128-column chunk
    const int r = q * 32 + lane;     // tile row
    const int f = f0 + r;            // frame
    const bool in_seq = f >= 0 && f < p.T;
    const bool is_out = in_seq && r >= p.halo && r < 128 - p.halo;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    auto tm_t = [&](int tb) { return TM_TEMP0 + lane_off + (uint32_t)tb * 128u; };
    const uint32_t aA = tc::smem_u32(sA), aG = tc::smem_u32(sG);
    const size_t row_g = (size_t)b * p.T + (in_seq ? f : 0);
    uint32_t n_use = 0;              // temp-buffer uses so far (alternate 0,1,0,1,...)
    float u[128];                    // fp32 residual stream: channels 128c + 64h + j at index 64c + j

    long long t_tfull = 0;
    auto temp_wait = [&](int tb) {
      const long long t0 = PROF ? clock64() : 0;
      tc::mbar_wait_trap(tc::smem_u32(&bars[B_TFULL + tb]), (n_use >> 1) & 1, WAIT_CYCLES, p.status, 4);
      if (PROF) t_tfull += clock64() - t0;
      ++n_use;
      tc::tc_fence_after();
    };
    auto temp_release = [&](int tb) {
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[B_TEMPTY + tb]);
    };
    auto publish = [&](int bar) {     // smem operand tile written by this warp is ready for the MMA
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[bar]);
    };
    // write u[64c .. 64c+64) as bf16 into the conv-input tile (zero outside the utterance)
    auto write_A = [&](int c, const float* v) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = in_seq ? pack_bf16(v[jj * 8 + 2 * e], v[jj * 8 + 2 * e + 1]) : 0u;
        st_shared_v4(aA + (uint32_t)(16 * c + 8 * h + jj) * A_LBO + (uint32_t)(r + 1) * 16, w[0], w[1], w[2], w[3]);
      }
    };

    // ---- group start: produce u_lb ----
    if (first_group) {
      // x_t tile -> bf16 A operand (channels 0..79): this thread converts bins [40h, 40h+40)
#pragma unroll
      for (int jj = 0; jj < 5; ++jj) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int n = 40 * h + jj * 8 + e;
          v[e] = (in_seq && n < p.n_mel) ? p.x_t[((size_t)b * p.n_mel + n) * p.T + f] : 0.f;
        }
        st_shared_v4(aA + (uint32_t)(5 * h + jj) * A_LBO + (uint32_t)(r + 1) * 16, pack_bf16(v[0], v[1]),
                     pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
      publish(B_AREADY);
#pragma unroll
      for (int c = 0; c < 2; ++c) {     // u = relu(W_in x + b_in)
        temp_wait(c);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t a[32];
          tc::tmem_ld32(tm_t(c) + 64 * h + 32 * hh, a);
          tc::tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(p.b_in + 128 * c + 64 * h + 32 * hh);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bv = __ldg(bp + j4);
            u[64 * c + 32 * hh + 4 * j4 + 0] = fmaxf(__uint_as_float(a[4 * j4 + 0]) + bv.x, 0.f);
            u[64 * c + 32 * hh + 4 * j4 + 1] = fmaxf(__uint_as_float(a[4 * j4 + 1]) + bv.y, 0.f);
            u[64 * c + 32 * hh + 4 * j4 + 2] = fmaxf(__uint_as_float(a[4 * j4 + 2]) + bv.z, 0.f);
            u[64 * c + 32 * hh + 4 * j4 + 3] = fmaxf(__uint_as_float(a[4 * j4 + 3]) + bv.w, 0.f);
          }
        }
        temp_release(c);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {     // u += Wc_0 cond + (d_0 + bc_0 + s_0)
        temp_wait(c);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t a[32];
          tc::tmem_ld32(tm_t(c) + 64 * h + 32 * hh, a);
          tc::tmem_ld_wait();
          const float4* kp = reinterpret_cast<const float4*>(p.k00 + (size_t)b * C + 128 * c + 64 * h + 32 * hh);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 kv = __ldg(kp + j4);
            u[64 * c + 32 * hh + 4 * j4 + 0] += __uint_as_float(a[4 * j4 + 0]) + kv.x;
            u[64 * c + 32 * hh + 4 * j4 + 1] += __uint_as_float(a[4 * j4 + 1]) + kv.y;
            u[64 * c + 32 * hh + 4 * j4 + 2] += __uint_as_float(a[4 * j4 + 2]) + kv.z;
            u[64 * c + 32 * hh + 4 * j4 + 3] += __uint_as_float(a[4 * j4 + 3]) + kv.w;
          }
        }
        temp_release(c);
        write_A(c, &u[64 * c]);
      }
      publish(B_AREADY);
    } else {
      // reload u_lb spilled by the previous group
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4* up = reinterpret_cast<const float4*>(p.U_in + row_g * C + 128 * c + 64 * h);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 v = in_seq ? __ldg(up + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
          u[64 * c + 4 * j4 + 0] = v.x; u[64 * c + 4 * j4 + 1] = v.y;
          u[64 * c + 4 * j4 + 2] = v.z; u[64 * c + 4 * j4 + 3] = v.w;
        }
        write_A(c, &u[64 * c]);
      }
      publish(B_AREADY);
    }

    // ---- residual blocks ----
    for (int l = p.lb; l < p.le; ++l) {
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {     // conv chunk i: gate columns [32h,32h+32), filter columns 64+[32h,32h+32)
        const int tb = i & 1;
        temp_wait(tb);
        uint32_t ga[32], fa[32];
        tc::tmem_ld32(tm_t(tb) + 32 * h, ga);
        tc::tmem_ld32(tm_t(tb) + 64 + 32 * h, fa);
        tc::tmem_ld_wait();
        temp_release(tb);
        const float4* bg = reinterpret_cast<const float4*>(p.conv_bias + ((size_t)l * 4 + i) * 128 + 32 * h);
        const float4* bf = reinterpret_cast<const float4*>(p.conv_bias + ((size_t)l * 4 + i) * 128 + 64 + 32 * h);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float g[8];
#pragma unroll
          for (int e4 = 0; e4 < 2; ++e4) {
            const float4 bgv = __ldg(bg + jj * 2 + e4), bfv = __ldg(bf + jj * 2 + e4);
            const float bgs[4] = {bgv.x, bgv.y, bgv.z, bgv.w}, bfs[4] = {bfv.x, bfv.y, bfv.z, bfv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = jj * 8 + e4 * 4 + e;
              const float a = __uint_as_float(ga[j]) + bgs[e];
              const float fl = __uint_as_float(fa[j]) + bfs[e];
              const float sg = fmaf(tanh_approx(0.5f * a), 0.5f, 0.5f);   // sigmoid(a)
              g[e4 * 4 + e] = sg * tanh_approx(fl);
            }
          }
          st_shared_v4(aG + (uint32_t)(8 * i + 4 * h + jj) * G_LBO + (uint32_t)r * 16, pack_bf16(g[0], g[1]),
                       pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
        }
        publish(B_GREADY + i);
      }
      if (l < p.L - 1) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {   // u <- (u + acc)/sqrt(2) + k_l ; write the next conv input
          temp_wait(c);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t a[32];
            tc::tmem_ld32(tm_t(c) + 64 * h + 32 * hh, a);
            tc::tmem_ld_wait();
            if (hh == 1) temp_release(c);
            const float4* kp =
                reinterpret_cast<const float4*>(p.ktab + ((size_t)b * p.L + l) * C + 128 * c + 64 * h + 32 * hh);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 kv = __ldg(kp + j4);
              float* uu = &u[64 * c + 32 * hh + 4 * j4];
              uu[0] = fmaf(uu[0] + __uint_as_float(a[4 * j4 + 0]), RSQRT2, kv.x);
              uu[1] = fmaf(uu[1] + __uint_as_float(a[4 * j4 + 1]), RSQRT2, kv.y);
              uu[2] = fmaf(uu[2] + __uint_as_float(a[4 * j4 + 2]), RSQRT2, kv.z);
              uu[3] = fmaf(uu[3] + __uint_as_float(a[4 * j4 + 3]), RSQRT2, kv.w);
            }
          }
          write_A(c, &u[64 * c]);
        }
        publish(B_AREADY);
      }
    }

    // ---- group end ----
    tc::mbar_wait_trap(tc::smem_u32(&bars[B_SKIPDONE]), 0, WAIT_CYCLES, p.status, 4);
    tc::tc_fence_after();
    if (!last_group) {
      if (is_out) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float4* up = reinterpret_cast<float4*>(p.U_out + row_g * C + 128 * c + 64 * h);
#pragma unroll
          for (int j4 = 0; j4 < 16; ++j4)
            up[j4] = make_float4(u[64 * c + 4 * j4], u[64 * c + 4 * j4 + 1], u[64 * c + 4 * j4 + 2], u[64 * c + 4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {  // partial skip sum, channels [128h + 32cc, +32)
        uint32_t a[32];
        tc::tmem_ld32(TM_SKIP + lane_off + 128 * h + 32 * cc, a);
        tc::tmem_ld_wait();
        if (is_out) {
          float4* sp = reinterpret_cast<float4*>(p.S + row_g * C + 128 * h + 32 * cc);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 v = first_group ? make_float4(0.f, 0.f, 0.f, 0.f) : sp[j4];
            v.x += __uint_as_float(a[4 * j4 + 0]); v.y += __uint_as_float(a[4 * j4 + 1]);
            v.z += __uint_as_float(a[4 * j4 + 2]); v.w += __uint_as_float(a[4 * j4 + 3]);
            sp[j4] = v;
          }
        }
      }
    } else {
      // ---- tail: skip/sqrt(L) -> skip_projection -> ReLU -> output_projection -> posterior ----
      const float inv_sqrt_l = 1.0f / sqrtf((float)p.L);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t a[32];
        tc::tmem_ld32(TM_SKIP + lane_off + 128 * h + 32 * cc, a);
        tc::tmem_ld_wait();
        const float4* sp = reinterpret_cast<const float4*>(p.S + row_g * C + 128 * h + 32 * cc);
        const float4* bp = reinterpret_cast<const float4*>(p.bsum_skip + 128 * h + 32 * cc);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float v[8];
#pragma unroll
          for (int e4 = 0; e4 < 2; ++e4) {
            const float4 bv = __ldg(bp + jj * 2 + e4);
            const float4 sv = (!first_group && in_seq) ? sp[jj * 2 + e4] : make_float4(0.f, 0.f, 0.f, 0.f);
            v[e4 * 4 + 0] = (__uint_as_float(a[jj * 8 + e4 * 4 + 0]) + sv.x + bv.x) * inv_sqrt_l;
            v[e4 * 4 + 1] = (__uint_as_float(a[jj * 8 + e4 * 4 + 1]) + sv.y + bv.y) * inv_sqrt_l;
            v[e4 * 4 + 2] = (__uint_as_float(a[jj * 8 + e4 * 4 + 2]) + sv.z + bv.z) * inv_sqrt_l;
            v[e4 * 4 + 3] = (__uint_as_float(a[jj * 8 + e4 * 4 + 3]) + sv.w + bv.w) * inv_sqrt_l;
          }
          st_shared_v4(aA + (uint32_t)(16 * h + 4 * cc + jj) * A_LBO + (uint32_t)(r + 1) * 16, pack_bf16(v[0], v[1]),
                       pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
      }
      publish(B_AREADY);
#pragma unroll
      for (int c = 0; c < 2; ++c) {     // relu(skip_projection): channels 128c + 64h + [0,64) -> sG
        temp_wait(c);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t a[32];
          tc::tmem_ld32(tm_t(c) + 64 * h + 32 * hh, a);
          tc::tmem_ld_wait();
          if (hh == 1) temp_release(c);
          const float4* bp = reinterpret_cast<const float4*>(p.b_skip + 128 * c + 64 * h + 32 * hh);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float4 b0 = __ldg(bp + jj * 2), b1 = __ldg(bp + jj * 2 + 1);
            const float v0 = fmaxf(__uint_as_float(a[jj * 8 + 0]) + b0.x, 0.f), v1 = fmaxf(__uint_as_float(a[jj * 8 + 1]) + b0.y, 0.f);
            const float v2 = fmaxf(__uint_as_float(a[jj * 8 + 2]) + b0.z, 0.f), v3 = fmaxf(__uint_as_float(a[jj * 8 + 3]) + b0.w, 0.f);
            const float v4 = fmaxf(__uint_as_float(a[jj * 8 + 4]) + b1.x, 0.f), v5 = fmaxf(__uint_as_float(a[jj * 8 + 5]) + b1.y, 0.f);
            const float v6 = fmaxf(__uint_as_float(a[jj * 8 + 6]) + b1.z, 0.f), v7 = fmaxf(__uint_as_float(a[jj * 8 + 7]) + b1.w, 0.f);
            st_shared_v4(aG + (uint32_t)(16 * c + 8 * h + 4 * hh + jj) * G_LBO + (uint32_t)r * 16, pack_bf16(v0, v1),
                         pack_bf16(v2, v3), pack_bf16(v4, v5), pack_bf16(v6, v7));
          }
        }
      }
      publish(B_GREADY + 0);
      // output projection: mel bins 64h + [0, 64) (only bins < n_mel exist), then the posterior update
      temp_wait(0);
      float c1 = 0.f, c2 = 0.f, sg = 0.f;
      if (p.sched) {
        const int tb = (int)p.t[b];
        c1 = p.sched[tb]; c2 = p.sched[p.K + tb]; sg = p.sched[2 * p.K + tb];
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int n0 = 64 * h + 32 * hh;
        if (n0 < p.n_mel) {              // warp-uniform
          uint32_t a[32];
          tc::tmem_ld32(tm_t(0) + n0, a);
          tc::tmem_ld_wait();
          if (is_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + j;
              if (n < p.n_mel) {
                float x0 = __uint_as_float(a[j]) + __ldg(p.b_out + n);
                if (p.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
                const size_t o = ((size_t)b * p.n_mel + n) * p.T + f;
                if (p.x0_out) p.x0_out[o] = x0;
                if (p.sched) {
                  const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, p.x_t[o]));
                  p.x_prev[o] = __fadd_rn(mean, __fmul_rn(sg, p.noise[o]));
                }
              }
            }
          }
        }
      }
      temp_release(0);
    }
    if (PROF) t_tfull_out = t_tfull;
  }

  if (PROF && warp == 4 && lane == 0) { p.prof[tile * 16 + 8] = t_tfull_out; p.prof[tile * 16 + 9] = clock64() - t_start; }
  // ---- teardown ----
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc<512>(tmem);
}

// ---- cond [B][T][H] fp32 -> [B][32][Tp][8] bf16 with COND_PAD_LO leading zero rows ---------------
__global__ void cond_pack_kernel(const float* __restrict__ cond, __nv_bfloat16* __restrict__ out, int T, int Tp) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);   // 8 rows per block, 32 chunks per row
  const int c8 = threadIdx.x & 31;
  if (row >= Tp) return;
  const int f = row - COND_PAD_LO;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (f >= 0 && f < T) {
    const float4* src = reinterpret_cast<const float4*>(cond + ((size_t)b * T + f) * C + c8 * 8);
    const float4 a = __ldg(src), c = __ldg(src + 1);
    v = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
  }
  reinterpret_cast<uint4*>(out)[((size_t)b * 32 + c8) * Tp + row] = v;
}

// ---- per-utterance constants of the u recurrence ----------------------------------------------------
// ktab[b][l] = (bo_x,l - ctab[b][l]) / sqrt(2) + dtab[b][l+1] + ctab[b][l+1]  (l < L-1), k00[b] = dtab[b][0] + ctab[b][0]
__global__ void ktab_kernel(const float* __restrict__ dtab, const float* __restrict__ ctab,
                            const float* __restrict__ bo_x, float* __restrict__ ktab, float* __restrict__ k00, int L) {
  const int b = blockIdx.y, l = blockIdx.x, c = threadIdx.x;
  const size_t i = ((size_t)b * L + l) * C + c;
  float v = 0.f;
  if (l < L - 1) v = (bo_x[(size_t)l * C + c] - ctab[i]) * RSQRT2 + dtab[i + C] + ctab[i + C];
  ktab[i] = v;
  if (l == 0) k00[(size_t)b * C + c] = dtab[i] + ctab[i];
}

// ---- weight images --------------------------------------------------------------------------------
struct SmallOff {   // fp32 section of the packed buffer (float offsets)
  size_t mlp0_wt, mlp2_wt, dproj_wt, sproj_wt, cproj_b, conv_bias, bo_x, bsum_skip, b_in, b_skip, b_out, total;
};
SmallOff small_layout(const mgb_model_dims& d) {
  const size_t L = d.layers, H = d.d_encoder;
  SmallOff o{};
  size_t p = 0;
  o.mlp0_wt = p; p += (size_t)C * 4 * C;
  o.mlp2_wt = p; p += (size_t)4 * C * C;
  o.dproj_wt = p; p += L * C * C;
  o.sproj_wt = p; if (d.multi_speaker) p += L * H * C;
  o.cproj_b = p; p += L * C;
  o.conv_bias = p; p += L * 512;
  o.bo_x = p; p += L * C;
  o.bsum_skip = p; p += C;
  o.b_in = p; p += C;
  o.b_skip = p; p += C;
  o.b_out = p; p += 128;
  o.total = align_up(p, 256);
  return o;
}
inline int num_wslots(const mgb_model_dims& d) { return W_LAYER0 + d.layers * W_PER_LAYER; }

__global__ void pack_images_kernel(const float* __restrict__ flat, const FlatOffsets f, const int L, const int n_mel,
                                   __nv_bfloat16* __restrict__ img) {
  const int slot = blockIdx.x;
  const float* conv_w = nullptr; const float* oproj_w = nullptr; const float* cproj_w = nullptr;
  const float* cproj_next = nullptr;
  int kind, c = 0, j = 0, ci = 0;
  if (slot < W_P0) { kind = 0; c = slot >> 1; j = slot & 1; }
  else if (slot < W_SKIPP) { kind = 1; c = (slot - W_P0) >> 2; j = (slot - W_P0) & 3; cproj_w = flat + f.layer0 + f.rel.cproj_w; }
  else if (slot < W_OUT) { kind = 2; c = (slot - W_SKIPP) >> 2; j = (slot - W_SKIPP) & 3; }
  else if (slot < W_LAYER0) { kind = 3; j = slot - W_OUT; }
  else {
    const int l = (slot - W_LAYER0) / W_PER_LAYER, rr = (slot - W_LAYER0) % W_PER_LAYER;
    const float* fl = flat + f.layer0 + (size_t)l * f.layer_stride;
    conv_w = fl + f.rel.conv_w; oproj_w = fl + f.rel.oproj_w; cproj_w = fl + f.rel.cproj_w;
    cproj_next = (l + 1 < L) ? fl + f.layer_stride + f.rel.cproj_w : nullptr;
    if (rr < 48) { kind = 4; ci = rr / 12; j = rr % 12; }
    else if (rr < 64) { const int qq = rr - 48; c = qq >> 3; j = qq & 7; kind = j < 4 ? 5 : 6; j &= 3; }
    else { kind = 7; c = (rr - 64) >> 2; j = (rr - 64) & 3; }
  }
  for (int unit = threadIdx.x; unit < 1024; unit += blockDim.x) {
    const int k8 = unit >> 7, n = unit & 127;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k8 * 8 + e;
      float x = 0.f;
      switch (kind) {
        case 0: { const int kk = j * 64 + k; if (kk < n_mel) x = flat[f.in_w + (size_t)(128 * c + n) * n_mel + kk]; break; }
        case 1: x = cproj_w[(size_t)(128 * c + n) * C + 64 * j + k]; break;
        case 2: x = flat[f.skip_w + (size_t)(128 * c + n) * C + 64 * j + k]; break;
        case 3: if (n < n_mel) x = flat[f.out_w + (size_t)n * C + 64 * j + k]; break;
        case 4: {
          const int oc = n < 64 ? 64 * ci + n : C + 64 * ci + (n - 64);
          const int tap = j >> 2, cin = (j & 3) * 64 + k;
          x = conv_w[((size_t)oc * C + cin) * 3 + tap];
          break;
        }
        case 5:
          if (cproj_next) {
            const size_t o = (size_t)(128 * c + n) * C + 64 * j + k;
            x = 1.41421356237309504880f * cproj_next[o] - cproj_w[o];
          }
          break;
        case 6: x = oproj_w[(size_t)(128 * c + n) * C + 64 * j + k]; break;
        default: x = oproj_w[(size_t)(C + 128 * c + n) * C + 64 * j + k]; break;
      }
      v[e] = x;
    }
    reinterpret_cast<uint4*>(img)[(size_t)slot * 1024 + unit] =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// small fp32 vectors in kernel order
__global__ void pack_small_kernel(const float* __restrict__ flat, const FlatOffsets f, const int L, const int n_mel,
                                  float* __restrict__ conv_bias, float* __restrict__ bo_x, float* __restrict__ bsum,
                                  float* __restrict__ b_in, float* __restrict__ b_skip, float* __restrict__ b_out) {
  const int c = threadIdx.x;   // 256 threads
  float s = 0.f;
  for (int l = 0; l < L; ++l) {
    const float* fl = flat + f.layer0 + (size_t)l * f.layer_stride;
    for (int n = c; n < 512; n += 256) {   // chunk i = n/128: 64 gate then 64 filter channels
      const int i = n >> 7, pos = n & 127;
      const int oc = pos < 64 ? 64 * i + pos : C + 64 * i + (pos - 64);
      conv_bias[(size_t)l * 512 + n] = fl[f.rel.conv_b + oc];
    }
    bo_x[(size_t)l * C + c] = fl[f.rel.oproj_b + c];
    s += fl[f.rel.oproj_b + C + c];
  }
  bsum[c] = s;
  b_in[c] = flat[f.in_b + c];
  b_skip[c] = flat[f.skip_b + c];
  if (c < 128) b_out[c] = c < n_mel ? flat[f.out_b + c] : 0.f;
}

struct WorkBf16 {
  size_t condT, d, h, dtab, ctab, ktab, k00, U, U2, S, status, total;
  int Tp;
};
WorkBf16 work_layout(const mgb_model_dims& d, int B, int T) {
  WorkBf16 w{};
  w.Tp = (int)align_up((size_t)T + COND_PAD_LO + 128 + MAX_GROUP_LAYERS, 8);
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t r = p; p += align_up(bytes, 256); return r; };
  w.condT = take((size_t)B * 32 * w.Tp * 16);
  w.d = take((size_t)B * C * 4);
  w.h = take((size_t)B * 4 * C * 4);
  w.dtab = take((size_t)B * d.layers * C * 4);
  w.ctab = take((size_t)B * d.layers * C * 4);
  w.ktab = take((size_t)B * d.layers * C * 4);
  w.k00 = take((size_t)B * C * 4);
  w.U = take((size_t)B * T * C * 4);
  w.U2 = take((size_t)B * T * C * 4);
  w.S = take((size_t)B * T * C * 4);
  w.status = take(256);
  w.total = p;
  return w;
}

int group_layers() {
  static int g = [] {
    const char* e = getenv("MGB_GROUP_LAYERS");
    int v = e ? atoi(e) : 10;
    return v < 1 ? 1 : (v > MAX_GROUP_LAYERS ? MAX_GROUP_LAYERS : v);
  }();
  return g;
}

}  // namespace

size_t bf16_packed_bytes(const mgb_model_dims& d) {
  return small_layout(d).total * sizeof(float) + (size_t)num_wslots(d) * SLOT_BYTES;
}
size_t bf16_workspace_bytes(const mgb_model_dims& d, int B, int T, int) { return work_layout(d, B, T).total; }
size_t bf16_status_offset(const mgb_model_dims& d, int B, int T) { return work_layout(d, B, T).status; }

int bf16_pack(const mgb_model_dims& d, const float* flat, void* packed, cudaStream_t s) {
  const FlatOffsets f = flat_offsets(d);
  const SmallOff o = small_layout(d);
  float* P = static_cast<float*>(packed);
  const int H = d.d_encoder, L = d.layers;
  launch_pack(flat + f.mlp0_w, P + o.mlp0_wt, 4 * C, C, 1, 4 * C, 0, 0, s);
  launch_pack(flat + f.mlp2_w, P + o.mlp2_wt, C, 4 * C, 1, C, 0, 0, s);
  for (int l = 0; l < L; ++l) {
    const float* fl = flat + f.layer0 + (size_t)l * f.layer_stride;
    launch_pack(fl + f.rel.dproj_w, P + o.dproj_wt + (size_t)l * C * C, C, C, 1, C, 0, 0, s);
    if (d.multi_speaker) launch_pack(fl + f.rel.sproj_w, P + o.sproj_wt + (size_t)l * H * C, C, H, 1, C, 0, 0, s);
    pack_bias_kernel<<<2, 128, 0, s>>>(fl + f.rel.cproj_b, P + o.cproj_b + (size_t)l * C, C, C, 0, 0);
  }
  pack_small_kernel<<<1, 256, 0, s>>>(flat, f, L, d.n_mel, P + o.conv_bias, P + o.bo_x, P + o.bsum_skip, P + o.b_in,
                                      P + o.b_skip, P + o.b_out);
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(P + o.total);
  pack_images_kernel<<<num_wslots(d), 256, 0, s>>>(flat, f, L, d.n_mel, img);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int bf16_denoiser(const mgb_model_dims& d, const void* packed, const float* x, const int64_t* t, const float* cond,
                  const float* spk, const float* noise, const float* sched, int K, int clip, float* x_prev,
                  float* out_x0, int B, int T, void* ws, bool cond_ready, cudaStream_t s) {
  MGB_REQUIRE(d.n_mel == 80, MGB_E_UNSUPPORTED, "the bf16 path is built for n_mel == 80 (got %d)", d.n_mel);
  const SmallOff o = small_layout(d);
  const WorkBf16 w = work_layout(d, B, T);
  const float* P = static_cast<const float*>(packed);
  uint8_t* W = static_cast<uint8_t*>(ws);
  const int L = d.layers, H = d.d_encoder;
  float* dvec = reinterpret_cast<float*>(W + w.d);
  float* dtab = reinterpret_cast<float*>(W + w.dtab);
  float* ctab = reinterpret_cast<float*>(W + w.ctab);
  float* ktab = reinterpret_cast<float*>(W + w.ktab);
  float* k00 = reinterpret_cast<float*>(W + w.k00);
  __nv_bfloat16* condT = reinterpret_cast<__nv_bfloat16*>(W + w.condT);
  int* status = reinterpret_cast<int*>(W + w.status);

  static bool attr_set = false;
  if (!attr_set) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(fused_group_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set = true;
  }
  if (!cond_ready) {
    dim3 grid((w.Tp + 7) / 8, B);
    cond_pack_kernel<<<grid, 256, 0, s>>>(cond, condT, T, w.Tp);
    MGB_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));
    note_launch();
  }
  launch_step_mlp(t, P + o.mlp0_wt, P + o.mlp2_wt, reinterpret_cast<float*>(W + w.h), dvec, B, C, s);
  {
    dim3 grid(L, (B + TAB_UB - 1) / TAB_UB);
    proj_table_kernel<<<grid, 256, (size_t)TAB_UB * C * sizeof(float), s>>>(dvec, C, P + o.dproj_wt, (size_t)C * C,
                                                                            nullptr, 0, dtab, B, L, C);
    proj_table_kernel<<<grid, 256, (size_t)TAB_UB * H * sizeof(float), s>>>(
        d.multi_speaker ? spk : nullptr, H, P + o.sproj_wt, (size_t)H * C, P + o.cproj_b, (size_t)C, ctab, B, L, C);
    dim3 kgrid(L, B);
    ktab_kernel<<<kgrid, 256, 0, s>>>(dtab, ctab, P + o.bo_x, ktab, k00, L);
    note_launch(5);   // step MLP (2), two projection tables, ktab
  }
  FusedParams p{};
  p.wimg = reinterpret_cast<const uint8_t*>(P + o.total);
  p.condT = condT; p.Tp = w.Tp;
  p.x_t = x; p.noise = noise; p.x_prev = x_prev; p.x0_out = out_x0; p.sched = sched; p.t = t;
  p.K = K; p.clip = clip; p.n_mel = d.n_mel;
  p.ktab = ktab; p.k00 = k00; p.conv_bias = P + o.conv_bias; p.bsum_skip = P + o.bsum_skip;
  p.b_in = P + o.b_in; p.b_skip = P + o.b_skip; p.b_out = P + o.b_out;
  float* Ubuf[2] = {reinterpret_cast<float*>(W + w.U), reinterpret_cast<float*>(W + w.U2)};
  p.S = reinterpret_cast<float*>(W + w.S);
  p.B = B; p.T = T; p.L = L; p.status = status;
  const int gl = group_layers();
  const int ngroups = (L + gl - 1) / gl;
  for (int g = 0; g < ngroups; ++g) {
    p.lb = g * L / ngroups;
    p.le = (g + 1) * L / ngroups;
    p.halo = p.le - p.lb;
    p.V = 128 - 2 * p.halo;
    p.tiles_per_utt = (T + p.V - 1) / p.V;
    p.U_in = Ubuf[g & 1];
    p.U_out = Ubuf[(g + 1) & 1];
    prof_begin(s);
    static const bool do_prof = getenv("MGB_PROFILE") != nullptr;
    if (do_prof) {
      const int ntile = B * p.tiles_per_utt;
      long long* dprof = nullptr;
      cudaMalloc(&dprof, (size_t)ntile * 16 * sizeof(long long));
      cudaMemset(dprof, 0, (size_t)ntile * 16 * sizeof(long long));
      p.prof = dprof;
      cudaFuncSetAttribute(fused_group_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
      fused_group_kernel<true><<<ntile, NTHREADS, SMEM_TOTAL, s>>>(p);
      cudaStreamSynchronize(s);
      long long* h = (long long*)malloc((size_t)ntile * 16 * sizeof(long long));
      cudaMemcpy(h, dprof, (size_t)ntile * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
      double a[16] = {0};
      for (int i = 0; i < ntile; ++i) for (int k = 0; k < 16; ++k) a[k] += (double)h[i * 16 + k] / ntile;
      fprintf(stderr, "[mgb profile] layers [%d,%d) tiles %d | MMA warp: total %.0f wait_full %.0f wait_temp %.0f wait_aready %.0f "
              "wait_gready %.0f | epilogue w4: total %.0f wait_tfull %.0f | producer: total %.0f wait_empty %.0f (cycles, mean per tile)\n",
              p.lb, p.le, ntile, a[4], a[0], a[1], a[2], a[3], a[9], a[8], a[13], a[12]);
      free(h); cudaFree(dprof); p.prof = nullptr;
    } else {
      fused_group_kernel<false><<<B * p.tiles_per_utt, NTHREADS, SMEM_TOTAL, s>>>(p);
    }
    prof_end(s);
    note_launch();
  }
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

}  // namespace mgb

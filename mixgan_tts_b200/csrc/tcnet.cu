// Tensor-core Conv1d / Linear engine + self-attention kernel (tcnet.cuh header comment has the layouts).
#include "tcnet.cuh"

#include "tc05.cuh"
#include "tmap.cuh"

namespace mgb {
namespace tcnet {
namespace {

constexpr long long kTimeout = 400000000LL;   // ~0.2 s of SM cycles: a protocol bug traps instead of hanging

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// fp32 stream [C/4][Rp][4]: a thread owns one row, so a warp's 16-byte access to chunk q is 512 contiguous bytes
__device__ __forceinline__ void load_s32(const float* base, size_t Rp, size_t row, int ch0, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(base + ((size_t)(ch0 / 4 + q) * Rp + row) * 4);
    v[q * 4] = a.x; v[q * 4 + 1] = a.y; v[q * 4 + 2] = a.z; v[q * 4 + 3] = a.w;
  }
}
__device__ __forceinline__ void store_s32(float* base, size_t Rp, size_t row, int ch0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(base + ((size_t)(ch0 / 4 + q) * Rp + row) * 4) =
        make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
}
// fp16 image [C/8][Rp][8]
__device__ __forceinline__ void store_img32(__half* img, size_t Rp, size_t row, int ch0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u = make_uint4(pack_h2(v[q * 8], v[q * 8 + 1]), pack_h2(v[q * 8 + 2], v[q * 8 + 3]),
                         pack_h2(v[q * 8 + 4], v[q * 8 + 5]), pack_h2(v[q * 8 + 6], v[q * 8 + 7]));
    *reinterpret_cast<uint4*>(img + ((size_t)(ch0 / 8 + q) * Rp + row) * 8) = u;
  }
}

// low-order halves: lo = fp16(v - float(fp16(v)))
__device__ __forceinline__ void store_img32_lo(__half* img, size_t Rp, size_t row, int ch0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = v[q * 8 + e] - __half2float(__float2half_rn(v[q * 8 + e]));
    uint4 u = make_uint4(pack_h2(w[0], w[1]), pack_h2(w[2], w[3]), pack_h2(w[4], w[5]), pack_h2(w[6], w[7]));
    *reinterpret_cast<uint4*>(img + ((size_t)(ch0 / 8 + q) * Rp + row) * 8) = u;
  }
}

struct KArgs {
  const __half* Wpk; const float* bias;
  int taps, dil, kspt;
  int B, T, Tg;
  const int* lens; int len_mul;
  int Cout, Cout32;            // channels of the output tensor, rounded up to 32
  int act; float scale;
  const float* res1; const float* res2;
  const float* ln_g; const float* ln_b;
  float* stream_out; __half* img_out; float img_slope;
  __half* img_lo_out; int split;
  float* user_out; int user_ld;
  int up; long long Rp_out;
  int ntiles, ntn;             // work list: row tiles x column tiles
  int rows, halo8, resident;   // rows of an A halo tile (128 + 2 * halo8, halo8 = the padding rounded up to 8 rows; 128 = plain
                               // 2-D box); 1 = the whole weight matrix stays in shared memory
  int* status;
  long long* trace;            // debug library only (MGB_TC_TRACE=1): SM-cycle stamps of CTA 0's first 64 work items
};
#ifdef MGB_DEBUG_BUILD
#define MGB_TC_STAMP(slot) do { if (p.trace && blockIdx.x == 0 && j < 32) p.trace[j * 16 + (slot)] = clock64(); } while (0)
#else
#define MGB_TC_STAMP(slot) do { } while (0)
#endif

// Shared memory of the persistent kernel: an A ring of halo tiles ((128 + 2*halo) rows x KC channels: ALL taps of a
// convolution read the same tile through descriptors shifted by tap*dilation rows, so an input row crosses the TMA engine
// once per k-step instead of once per tap) and a W region that either holds the layer's whole weight matrix (small layers:
// loaded once per CTA) or a ring of per-(k-step, tap) tiles.
constexpr int ROWS_MAX = 192;          // 128 + 2 * 32 halo rows
template <int NT, int KC>
struct Smem {
  // wide tiles stream 32 KB weight tiles (each used once per work item): give them the deeper ring; narrow tiles keep
  // whole weight matrices resident and want more A stages in flight
  static constexpr int W_REGION = (NT == 256 ? 128 : 96) * 1024;
  static constexpr int A_REGION = (NT == 256 ? 72 : 96) * 1024;
  static constexpr int A_STAGE = ROWS_MAX * KC * 2;                 // 24 KB (KC = 64) / 12 KB (KC = 32)
  static constexpr int A_STAGES = A_REGION / A_STAGE;               // 3-4 / 6-8
  static constexpr int W_STAGE = NT * KC * 2;
  static constexpr int W_RAW = W_REGION / W_STAGE;
  static constexpr int W_STAGES = W_RAW > 8 ? 8 : W_RAW;            // 4 (NT = 256, KC = 64) ... 8
  static constexpr int TOTAL = W_REGION + A_STAGES * A_STAGE + 1024;
};

// tcgen05.ld without the wait: the epilogue keeps the NEXT 32 columns in flight while it works on the current ones
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[32]) { tc::tmem_ld32(taddr, r); }

// One tile row per thread.  Uniform options (activation, residuals, slope, outputs) are tested OUTSIDE the element
// loops; the accumulator columns are software-pipelined (the tcgen05.ld of column group cg + 1 is in flight while group
// cg is processed).
// 32 consecutive floats of a shared-memory vector, the same address in every lane (broadcast reads)
__device__ __forceinline__ void lds_f32x32(const float* p, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(p + q * 4);
    v[q * 4] = a.x; v[q * 4 + 1] = a.y; v[q * 4 + 2] = a.z; v[q * 4 + 3] = a.w;
  }
}

// s_bias: the layer's whole (padded) bias vector, s_ln: LayerNorm weight | bias, staged in shared memory once per CTA —
// global (even L1-resident) loads of these per-column constants stalled every column group on the long scoreboard.
template <int NT>
__device__ __forceinline__ void epilogue(const KArgs& p, const float* s_bias, const float* s_ln, uint32_t trow, int tile,
                                         int ntile, int i, long long* tr) {
  const unsigned row = (unsigned)tile * TILE + i;
  const unsigned b = row / (unsigned)p.Tg, t = row - b * (unsigned)p.Tg;
  const bool inrange = b < (unsigned)p.B && t < (unsigned)p.T;
  bool valid = inrange;
  if (valid && p.lens) valid = (int)t < p.lens[b] * p.len_mul;
  const size_t orow = p.up > 1 ? (size_t)row * p.up + ntile : (size_t)row;
  const int nbase = p.up > 1 ? 0 : ntile * NT;
  const float* bias = s_bias + ntile * NT;
  const size_t Rp = (size_t)p.Rp_out;
  float* uo = nullptr;
  if (p.user_out && inrange) uo = p.user_out + ((size_t)(b * (unsigned)p.T + t) * p.up + (p.up > 1 ? ntile : 0)) * p.user_ld;
  constexpr int NCG = NT / 32;
  // 2 epilogue groups x 168 registers (NT = 256): room for a second accumulator buffer, so the tcgen05.ld of column group
  // cg + 1 is in flight while group cg is processed; the narrower tiles run 4 groups and hide the latency across groups.
  constexpr bool PIPE = NT == 256;
  uint32_t ra[32], rb[PIPE ? 32 : 1];

  if (p.ln_g != nullptr) {
    if constexpr (NT == 256) {
      // LayerNorm(acc + bias + residual) over the 256 channels of this thread's row (nn.LayerNorm: biased variance, eps
      // 1e-5).  Pass A accumulates sum(x - s) and sum((x - s)^2) around a per-row shift s (the first element), which is
      // as accurate as the two-pass form when |mean| >> std; pass B normalises and stores.
      // Pass A also writes x = acc + bias + residual back over the accumulator (tcgen05.st), so pass B reads TMEM only; the
      // residual of column group cg + 1 is fetched (HBM / L2 latency) while group cg is processed.
      const bool has_res = valid && p.res1 != nullptr;
      float s1 = 0.f, s2 = 0.f, shift = 0.f;
      float rra[32], rrb[32];
      auto pass_a = [&](int cg, uint32_t (&cur)[32], uint32_t (&nxt)[32], float (&rcur)[32], float (&rnxt)[32]) {
        tc::tmem_ld_wait();
        if (cg + 1 < NCG) {
          tmem_ld_issue(trow + (cg + 1) * 32, nxt);
          if (has_res) load_s32(p.res1, Rp, orow, (cg + 1) * 32, rnxt);
        }
        {
          float bv[32];
          lds_f32x32(bias + cg * 32, bv);
#pragma unroll
          for (int j = 0; j < 32; ++j) bv[j] += __uint_as_float(cur[j]);
          if (has_res) {
#pragma unroll
            for (int j = 0; j < 32; ++j) bv[j] += rcur[j];
          }
          if (cg == 0) shift = bv[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float d = bv[j] - shift; s1 += d; s2 = fmaf(d, d, s2); cur[j] = __float_as_uint(bv[j]); }
        }
        tc::tmem_st32(trow + cg * 32, cur);
      };
      if (has_res) load_s32(p.res1, Rp, orow, 0, rra);
      tmem_ld_issue(trow, ra);
#pragma unroll 1
      for (int cg = 0; cg < NCG; cg += 2) { pass_a(cg, ra, rb, rra, rrb); pass_a(cg + 1, rb, ra, rrb, rra); }
      tc::tmem_st_wait();
      const float dm = s1 * (1.f / 256.f);                      // mean - shift
      const float mean = shift + dm;
      const float var = fmaxf(s2 * (1.f / 256.f) - dm * dm, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      auto pass_b = [&](int cg, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
        tc::tmem_ld_wait();
        if (cg + 1 < NCG) tmem_ld_issue(trow + (cg + 1) * 32, nxt);
        float v[32];
        {
          float g[32];
          lds_f32x32(s_ln + cg * 32, g);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(cur[j]) - mean) * rstd * g[j];
        }
        {
          float be[32];
          lds_f32x32(s_ln + 256 + cg * 32, be);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? v[j] + be[j] : 0.f;
        }
        if (p.stream_out) store_s32(p.stream_out, Rp, orow, cg * 32, v);
        if (p.img_out) store_img32(p.img_out, Rp, orow, cg * 32, v);
        if (uo) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(uo + cg * 32 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
      };
      tmem_ld_issue(trow, ra);
#pragma unroll 1
      for (int cg = 0; cg < NCG; cg += 2) { pass_b(cg, ra, rb); pass_b(cg + 1, rb, ra); }
    }
    return;
  }

  const int ncg = min(NCG, (p.Cout32 - nbase + 31) / 32);         // column groups that exist in the output tensor
  const int act = p.act;
  const float scale = p.scale, slope = p.img_slope;
  const bool res1 = valid && p.res1 != nullptr, res2 = valid && p.res2 != nullptr;
  auto body = [&](int cg, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
    const int n0 = nbase + cg * 32;
    if (!PIPE) tmem_ld_issue(trow + cg * 32, cur);
    tc::tmem_ld_wait();
#ifdef MGB_DEBUG_BUILD
    if (tr && cg < 2) tr[8 + cg * 3] = clock64();
#endif
    if (PIPE && cg + 1 < ncg) tmem_ld_issue(trow + (cg + 1) * 32, nxt);
    float v[32];
    {
      float bv[32];
      lds_f32x32(bias + cg * 32, bv);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]) + bv[j];
    }
    if (act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == ACT_TANH) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
    }
    if (res1) {
      float r1[32];
      load_s32(p.res1, Rp, orow, n0, r1);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r1[j];
    }
    if (res2) {
      float r2[32];
      load_s32(p.res2, Rp, orow, n0, r2);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r2[j];
    }
    if (scale != 1.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= scale;
    }
    if (!valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
#ifdef MGB_DEBUG_BUILD
    if (tr && cg < 2) tr[9 + cg * 3] = clock64() + (long long)(v[0] != 12345.f ? 0 : 1);
#endif
    if (p.stream_out) store_s32(p.stream_out, Rp, orow, n0, v);
    if (uo) {
      if (n0 + 32 <= p.Cout && (p.user_ld & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(uo + n0 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else {
#pragma unroll                                                   // static indices: a rolled loop would put v[] in local memory
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.Cout) uo[n0 + j] = v[j];
      }
    }
    if (p.img_out) {
      if (slope != 1.f) {                                          // the consumer's leaky_relu, applied once here
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + slope * fminf(v[j], 0.f);
      }
      store_img32(p.img_out, Rp, orow, n0, v);
      if (p.img_lo_out) store_img32_lo(p.img_lo_out, Rp, orow, n0, v);
    }
#ifdef MGB_DEBUG_BUILD
    if (tr && cg < 2) tr[10 + cg * 3] = clock64();
#endif
  };
  if constexpr (PIPE) {
    tmem_ld_issue(trow, ra);
#pragma unroll 1
    for (int cg = 0; cg < ncg; cg += 2) {                          // ncg is even for every 256-wide tile (Cout32 % 64 == 0 there)
      body(cg, ra, rb);
      if (cg + 1 < ncg) body(cg + 1, rb, ra);
    }
  } else {
#pragma unroll 1
    for (int cg = 0; cg < ncg; ++cg) body(cg, ra, ra);
  }
}

// Persistent kernel: one CTA per SM walks the (row tile, column tile) work list with a stride of gridDim.x.  The TMA
// warp and the MMA warp run ahead over work items through one shared-memory ring; the accumulator is double-buffered in
// TMEM (2 x NT columns) and the two epilogue warpgroups take alternate work items, so a tile's epilogue (the HBM side:
// residual reads, stream / image stores) overlaps the next tiles' loads and MMAs.  Column tiles of one row tile are
// consecutive work items (the A box is re-read from L2 while it is hot).
// Waits of the roles that are NOT on the critical path (epilogue groups waiting for their accumulator, the TMA warp waiting
// for a free stage) back off with nanosleep: a dozen warps polling mbarriers at full speed took 3/4 of the issue slots of
// the sub-partition the single MMA-issuing thread lives on (measured: ~290 cycles per tcgen05.mma issued, 160 with fewer pollers).
__device__ __forceinline__ void wait_backoff(uint32_t bar_addr, uint32_t parity, int* status, int code) {
  if (tc::mbar_try_wait_addr(bar_addr, parity)) return;
  const long long t0 = clock64();
  while (!tc::mbar_try_wait_addr(bar_addr, parity)) {
    __nanosleep(64);
    if (clock64() - t0 > kTimeout) {
      if (status) atomicOr(status, code);
      __threadfence_system();
      __trap();
    }
  }
}

constexpr int MAX_BIAS = 2048;     // widest padded GEMM N: a transposed convolution with stride 8 into 256 channels

template <int NT>
struct Groups { static constexpr int NBUF = NT <= 128 ? 4 : 2; static constexpr int THREADS = 64 + NBUF * 128; };

template <int NT, int KC>
__global__ void __launch_bounds__(Groups<NT>::THREADS, 1) tcconv_kernel(const KArgs p, const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmA2) {
  using S = Smem<NT, KC>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NBUF = Groups<NT>::NBUF;
  __shared__ __align__(8) uint64_t a_full[S::A_STAGES], a_empty[S::A_STAGES], w_full[S::W_STAGES], w_empty[S::W_STAGES],
      w_res, acc_full[NBUF], acc_empty[NBUF];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[MAX_BIAS], s_ln[512];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // provably warp-uniform: role code stays on the uniform datapath

  if (warp == 1) tc::tmem_alloc<NBUF * NT>(&tmem_slot);
  if (tid == 0) {
    for (int i = 0; i < S::A_STAGES; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < S::W_STAGES; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    tc::mbar_init(&w_res, 1);
    for (int i = 0; i < NBUF; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 4); }
    tc::fence_barrier_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_trigger();
  pdl_wait();
  for (int k = tid; k < p.ntn * NT; k += blockDim.x) s_bias[k] = p.bias[k];
  if (p.ln_g) for (int k = tid; k < 256; k += blockDim.x) { s_ln[k] = p.ln_g[k]; s_ln[256 + k] = p.ln_b[k]; }
  __syncthreads();

  const int nparts = p.split ? 2 : 1;                    // weight tiles per (k-step, tap): hi [, lo]
  const int nsteps = p.taps * p.kspt * nparts;           // per work item; order: k-step outer, tap, part inner
  const int nwork = p.ntiles * p.ntn;
  const uint32_t w_base = tc::smem_u32(smem);
  const uint32_t a_base = w_base + S::W_REGION;
  const int halo = (p.taps >> 1) * p.dil;
  const int row_shift = p.halo8 - halo;                 // the first tap's row inside the tile
  const uint32_t a_bytes = (uint32_t)p.rows * KC * 2;

  // The TMA and MMA roles run with WARP-UNIFORM control flow (all 32 lanes wait on the barriers and walk the work list);
  // only the instructions that must come from one thread are predicated with elect_one().  Ring positions and operand
  // descriptors then live in uniform registers and a tcgen05.mma costs two 64-bit adds instead of a chain of
  // vector-to-uniform register moves (measured: ~290 cycles per MMA issued from an `if (lane == 0)` loop).
  if (warp == 0) {
    if (p.resident && tc::elect_one()) {                 // ntn == 1: the whole matrix, once
      const uint32_t total = (uint32_t)nsteps * S::W_STAGE;
      tc::mbar_arrive_expect_tx_addr(tc::smem_u32(&w_res), total);
      for (uint32_t off = 0; off < total; off += 16384) {
        const uint32_t n = total - off < 16384 ? total - off : 16384;
        tc::bulk_g2s_addr(w_base + off, reinterpret_cast<const uint8_t*>(p.Wpk) + off, n, tc::smem_u32(&w_res));
      }
    }
    int ia = 0, iw = 0, j = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++j) {
      const int tile = w / p.ntn, ntile = w - tile * p.ntn;
      if (lane == 0) MGB_TC_STAMP(0);
      const __half* bsrc = p.Wpk + (size_t)ntile * nsteps * (size_t)(NT * KC);
      for (int kc = 0; kc < p.kspt; ++kc) {
        for (int part = 0; part < nparts; ++part, ++ia) {       // the hi tile [and the lo tile] of this k-step
          const int st = ia % S::A_STAGES, ph = (ia / S::A_STAGES) & 1;
          wait_backoff(tc::smem_u32(&a_empty[st]), ph ^ 1, p.status, 1);
          if (tc::elect_one()) {
            const uint32_t fb = tc::smem_u32(&a_full[st]);
            const void* tm = part ? &tmA2 : &tmA;
            tc::mbar_arrive_expect_tx_addr(fb, a_bytes);
            // (128 + 2*halo8) rows x KC channels in ONE box, zero-filled outside the row axis
            if (p.halo8 == 0) tc::tma_load_2d(a_base + st * S::A_STAGE, tm, 2 * tile * TILE, kc * (KC / 8), fb);
            else tc::tma_load_3d(a_base + st * S::A_STAGE, tm, 0, (tile * TILE - p.halo8) >> 3, kc * (KC / 8), fb);
          }
        }
        if (!p.resident) {
          for (int t2 = 0; t2 < p.taps * nparts; ++t2, ++iw) {
            const int ws = iw % S::W_STAGES, wph = (iw / S::W_STAGES) & 1;
            wait_backoff(tc::smem_u32(&w_empty[ws]), wph ^ 1, p.status, 1);
            if (tc::elect_one()) {
              const uint32_t wb = tc::smem_u32(&w_full[ws]);
              tc::mbar_arrive_expect_tx_addr(wb, S::W_STAGE);
              tc::bulk_g2s_addr(w_base + ws * S::W_STAGE, bsrc + (size_t)(kc * p.taps * nparts + t2) * (NT * KC), S::W_STAGE, wb);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = tc::make_idesc_16(128, NT, true);
    const uint32_t lbo_a = (uint32_t)p.rows * 16;        // byte stride between 8-channel chunks of the A tile
    const uint64_t a_desc0 = tc::make_smem_desc(a_base, lbo_a, 128);
    const uint64_t b_desc0 = tc::make_smem_desc(w_base, NT * 16, 128);
    const uint32_t a_kstep = (2 * lbo_a) >> 4;           // descriptor address units (16 B): one K = 16 step = 2 chunks
    constexpr uint32_t b_kstep = (2 * NT * 16) >> 4;
    if (p.resident) { tc::mbar_wait_trap(tc::smem_u32(&w_res), 0, kTimeout, p.status, 2); tc::tc_fence_after(); }
    int ia = 0, iw = 0, j = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++j) {
      const int buf = j % NBUF;
      if (lane == 0) MGB_TC_STAMP(1);
      tc::mbar_wait_trap(tc::smem_u32(&acc_empty[buf]), ((j / NBUF) & 1) ^ 1, kTimeout, p.status, 2);
      tc::tc_fence_after();
      if (lane == 0) MGB_TC_STAMP(2);
      const uint32_t d_tmem = tmem + buf * NT;
      for (int kc = 0; kc < p.kspt; ++kc) {
        const int st = ia % S::A_STAGES, ph = (ia / S::A_STAGES) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&a_full[st]), ph, kTimeout, p.status, 2);
        int st2 = st;
        if (p.split) {
          st2 = (ia + 1) % S::A_STAGES;
          tc::mbar_wait_trap(tc::smem_u32(&a_full[st2]), ((ia + 1) / S::A_STAGES) & 1, kTimeout, p.status, 2);
        }
        ia += nparts;
        tc::tc_fence_after();
        const uint64_t a_st = a_desc0 + (uint64_t)((st * S::A_STAGE) >> 4);
        const uint64_t a_st2 = a_desc0 + (uint64_t)((st2 * S::A_STAGE) >> 4);
        for (int tap = 0; tap < p.taps; ++tap) {
          uint64_t b, b2 = 0;
          int ws = 0, ws2 = 0;
          if (p.resident) {
            b = b_desc0 + (uint64_t)(((kc * p.taps + tap) * nparts) * (S::W_STAGE >> 4));
            b2 = b + (S::W_STAGE >> 4);
          } else {
            ws = iw % S::W_STAGES;
            tc::mbar_wait_trap(tc::smem_u32(&w_full[ws]), (iw / S::W_STAGES) & 1, kTimeout, p.status, 2);
            b = b_desc0 + (uint64_t)(ws * (S::W_STAGE >> 4));
            ++iw;
            if (p.split) {
              ws2 = iw % S::W_STAGES;
              tc::mbar_wait_trap(tc::smem_u32(&w_full[ws2]), (iw / S::W_STAGES) & 1, kTimeout, p.status, 2);
              b2 = b_desc0 + (uint64_t)(ws2 * (S::W_STAGE >> 4));
              ++iw;
            }
            tc::tc_fence_after();
          }
          const uint32_t roff = row_shift + tap * p.dil;               // the tap: the same tile, tap*dil rows (16 B each) further
          const uint64_t a = a_st + (uint64_t)roff, a2 = a_st2 + (uint64_t)roff;
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              tc::umma_bf16(d_tmem, a + (uint64_t)(k * a_kstep), b + (uint64_t)(k * b_kstep), idesc, (kc | tap | k) ? 1u : 0u);
            if (p.split) {                                             // + a_hi w_lo + a_lo w_hi
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                tc::umma_bf16(d_tmem, a + (uint64_t)(k * a_kstep), b2 + (uint64_t)(k * b_kstep), idesc, 1u);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                tc::umma_bf16(d_tmem, a2 + (uint64_t)(k * a_kstep), b + (uint64_t)(k * b_kstep), idesc, 1u);
            }
            if (!p.resident) { tc::umma_commit(&w_empty[ws]); if (p.split) tc::umma_commit(&w_empty[ws2]); }
          }
        }
        if (tc::elect_one()) { tc::umma_commit(&a_empty[st]); if (p.split) tc::umma_commit(&a_empty[st2]); }
      }
      if (tc::elect_one()) tc::umma_commit(&acc_full[buf]);
      if (lane == 0) MGB_TC_STAMP(3);
    }
  } else {
    const int g = (warp - 2) >> 2;                         // epilogue group = TMEM buffer = local work index mod NBUF
    const int i = (warp & 3) * 32 + lane;                  // TMEM lane = tile row
    int j = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++j) {
      if (j % NBUF != g) continue;
      const int tile = w / p.ntn, ntile = w - tile * p.ntn;
      if ((tid & 127) == 64) MGB_TC_STAMP(4);
      wait_backoff(tc::smem_u32(&acc_full[g]), (j / NBUF) & 1, p.status, 4);
      tc::tc_fence_after();
      if ((tid & 127) == 64) MGB_TC_STAMP(5);
      long long* tr = nullptr;
#ifdef MGB_DEBUG_BUILD
      if (p.trace && blockIdx.x == 0 && j < 32 && (tid & 127) == 64) tr = p.trace + j * 16;
#endif
      epilogue<NT>(p, s_bias, s_ln, tmem + g * NT + ((uint32_t)((warp & 3) * 32) << 16), tile, ntile, i, tr);
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[g]);
      if ((tid & 127) == 64) MGB_TC_STAMP(6);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<NBUF * NT>(tmem);
}

template <int NT, int KC>
int launch_conv(KArgs a, const __half* in, const __half* in_lo, int in_chunks, int Rp_in, int ntiles, int ntn, cudaStream_t s) {
  using S = Smem<NT, KC>;
  static PerDeviceOnce once;
  static int sms[256];
  if (once.pending()) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(tcconv_kernel<NT, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    int dev = 0, n = 0;
    MGB_CUDA_CHECK(cudaGetDevice(&dev));
    MGB_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sms[dev & 255] = n;
    once.done();
  }
  CUtensorMap m, m2;
  a.halo8 = ((a.taps >> 1) * a.dil + 7) / 8 * 8;
  a.rows = TILE + 2 * a.halo8;
  MGB_REQUIRE(a.rows <= ROWS_MAX, MGB_E_UNSUPPORTED, "convolution padding %d exceeds %d rows", (a.taps >> 1) * a.dil, (ROWS_MAX - TILE) / 2);
  if (a.halo8 == 0) { if (int rc = make_image_map(&m, in, in_chunks, Rp_in, KC / 8)) return rc; }
  else if (int rc = make_image_map3(&m, in, in_chunks, Rp_in, a.rows, KC / 8)) return rc;
  m2 = m;
  if (a.split) {
    MGB_REQUIRE(in_lo != nullptr, MGB_E_ARG, "a split-precision layer needs the low-order image of its input");
    if (a.halo8 == 0) { if (int rc = make_image_map(&m2, in_lo, in_chunks, Rp_in, KC / 8)) return rc; }
    else if (int rc = make_image_map3(&m2, in_lo, in_chunks, Rp_in, a.rows, KC / 8)) return rc;
  }
  a.ntiles = ntiles; a.ntn = ntn;
  a.resident = (ntn == 1 && (long long)a.taps * a.kspt * (a.split ? 2 : 1) * S::W_STAGE <= S::W_REGION) ? 1 : 0;
  const long long nwork = (long long)ntiles * ntn;
  const int nsm = sms[PerDeviceOnce::current()];
  const int grid = (int)(nwork < nsm ? nwork : nsm);
  MGB_CUDA_CHECK(launch_pdl(tcconv_kernel<NT, KC>, dim3(grid), dim3(Groups<NT>::THREADS), S::TOTAL, s, 1, a, m, m2));
  note_launch();
  return MGB_OK;
}

// ---- weight packing: torch [Cout][Cin][k] (conv / linear) or [Cin][Cout][k] (transposed conv) -> streamed fp16 tiles
//      wp[ntile][step = kc*taps + tap][part: hi [, lo]][KC/8][NT][8] -------------------------------------------------------------------
__global__ void pack_w_kernel(const float* __restrict__ w, const float* __restrict__ oscale, __half* __restrict__ wp,
                              int Cin, int Cout, int k, int NT, int KC, int kspt, int taps, int up, int tpad, int nparts,
                              long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx % 8);
  const int nl = (int)((idx / 8) % NT);
  const int c8 = (int)((idx / (8LL * NT)) % (KC / 8));
  const long long per_step = (long long)NT * KC;
  const int nsteps = taps * kspt * nparts;
  const int s2 = (int)((idx / per_step) % nsteps);
  const int nt = (int)(idx / (per_step * nsteps));
  const int part = s2 % nparts, s = s2 / nparts;       // part 0 = fp16(w), part 1 = fp16(w - float(fp16(w)))
  const int kc = s / taps, tap = s - kc * taps;        // streamed order: k-step outer, tap, part inner
  const int ci = kc * KC + c8 * 8 + e;
  float val = 0.f;
  if (up > 1) {
    const int co = nl, j = nt + tpad - up * (tap - 1);       // output phase nt, input offset tap - 1
    if (ci < Cin && co < Cout && j >= 0 && j < k) val = w[((size_t)ci * Cout + co) * k + j];
  } else {
    const int co = nt * NT + nl;
    if (ci < Cin && co < Cout) val = w[((size_t)co * Cin + ci) * k + tap] * (oscale ? oscale[co] : 1.f);
  }
  const __half hi = __float2half_rn(val);
  wp[idx] = part ? __float2half_rn(val - __half2float(hi)) : hi;
}
__global__ void pack_b_kernel(const float* __restrict__ bias, const float* __restrict__ oscale, const float* __restrict__ oshift,
                              float* __restrict__ bp, int Cout, int NT, int ntn, int up) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntn * NT) return;
  const int co = up > 1 ? i % NT : i;
  float v = 0.f;
  if (co < Cout) {
    v = bias ? bias[co] : 0.f;
    if (oscale) v = v * oscale[co] + oshift[co];
  }
  bp[i] = v;
}

// thread = (8-channel chunk, row), rows fastest: a warp writes 512 contiguous bytes of one image chunk and reads one
// full 32-byte sector per thread
__global__ void pack_rows_kernel(const float* __restrict__ user, const float* __restrict__ pos, int C, int B, int T, int Tg,
                                 long long Rp, __half* __restrict__ img, float slope, float* __restrict__ stream,
                                 __half* __restrict__ img_lo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nch = C / 8;
  if (idx >= Rp * nch) return;
  const long long row = idx % Rp;
  const int ch = (int)(idx / Rp);
  const int b = (int)(row / Tg), t = (int)(row - (long long)b * Tg);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (b < B && t < T) {
    const float* src = user + ((size_t)b * T + t) * C + ch * 8;
    const float4 a = *reinterpret_cast<const float4*>(src), c = *reinterpret_cast<const float4*>(src + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    if (pos) {
      const float* ps = pos + (size_t)t * C + ch * 8;
      const float4 pa = *reinterpret_cast<const float4*>(ps), pc = *reinterpret_cast<const float4*>(ps + 4);
      v[0] += pa.x; v[1] += pa.y; v[2] += pa.z; v[3] += pa.w; v[4] += pc.x; v[5] += pc.y; v[6] += pc.z; v[7] += pc.w;
    }
  }
  if (stream) {
    *reinterpret_cast<float4*>(stream + ((size_t)(ch * 2) * Rp + row) * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(stream + ((size_t)(ch * 2 + 1) * Rp + row) * 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (img) {
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = v[e] > 0.f ? v[e] : v[e] * slope;
    *reinterpret_cast<uint4*>(img + ((size_t)ch * Rp + row) * 8) =
        make_uint4(pack_h2(w[0], w[1]), pack_h2(w[2], w[3]), pack_h2(w[4], w[5]), pack_h2(w[6], w[7]));
    if (img_lo) {
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] -= __half2float(__float2half_rn(w[e]));
      *reinterpret_cast<uint4*>(img_lo + ((size_t)ch * Rp + row) * 8) =
          make_uint4(pack_h2(w[0], w[1]), pack_h2(w[2], w[3]), pack_h2(w[4], w[5]), pack_h2(w[6], w[7]));
    }
  }
}

}  // namespace

int pack_rows(const float* user, const float* pos, int C, const Rows& r, __half* img, float img_slope, float* stream,
              cudaStream_t s, __half* img_lo) {
  MGB_REQUIRE(C % 8 == 0, MGB_E_UNSUPPORTED, "pack_rows: channels must be a multiple of 8");
  const long long total = (long long)r.Rp * (C / 8);
  pack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(user, pos, C, r.B, r.T, r.Tg, r.Rp, img, img_slope, stream, img_lo);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int pack_conv(const Layer& l, void* packed, const float* w, const float* bias, const float* oscale, const float* oshift,
              int ntile0, int Cout_part, cudaStream_t s) {
  const int nparts = l.split ? 2 : 1;
  const int nsteps = l.taps * l.kspt * nparts;
  const int ntn = l.up > 1 ? l.ntn : (Cout_part + l.NT - 1) / l.NT;
  const long long total = (long long)ntn * nsteps * l.NT * l.KC;
  __half* wp = static_cast<__half*>(packed) + l.w_off + (size_t)ntile0 * nsteps * l.NT * l.KC;
  pack_w_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, oscale, wp, l.Cin, Cout_part, l.k, l.NT, l.KC, l.kspt,
                                                                 l.taps, l.up, l.tpad, nparts, total);
  float* bp = reinterpret_cast<float*>(packed) + l.b_off + (size_t)ntile0 * l.NT;
  pack_b_kernel<<<(ntn * l.NT + 255) / 256, 256, 0, s>>>(bias, oscale, oshift, bp, Cout_part, l.NT, ntn, l.up);
  note_launch(2);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int run_conv(const Layer& l, const void* packed, const Rows& rin, const ConvIO& io, int* status, cudaStream_t s) {
  KArgs a{};
  a.Wpk = static_cast<const __half*>(packed) + l.w_off;
  a.bias = reinterpret_cast<const float*>(packed) + l.b_off;
  a.taps = l.taps; a.dil = l.dil; a.kspt = l.kspt;
  a.B = rin.B; a.T = rin.T; a.Tg = rin.Tg;
  a.lens = io.lens; a.len_mul = io.len_mul;
  a.Cout = l.Cout; a.Cout32 = (l.Cout + 31) / 32 * 32;
  a.act = io.act; a.scale = io.scale;
  a.res1 = io.res1; a.res2 = io.res2; a.ln_g = io.ln_g; a.ln_b = io.ln_b;
  a.stream_out = io.stream_out; a.img_out = io.img_out; a.img_slope = io.img_slope;
  a.img_lo_out = io.img_lo_out; a.split = l.split;
  a.user_out = io.user_out; a.user_ld = io.user_ld;
  a.up = l.up; a.Rp_out = (long long)rin.Rp * l.up;
  a.status = status;
  a.trace = nullptr;
#ifdef MGB_DEBUG_BUILD
  {   // the callers' status region is 8 KB: word 0 = watchdog, bytes [1024, 1024 + 64*8*8) = trace of the LAST launch
    // MGB_TC_TRACE=<n>: trace the n-th run_conv call of the process (0-based); MGB_TC_TRACE=all: every call (the last one stays)
    static const int which = [] { const char* e = getenv("MGB_TC_TRACE"); return !e ? -2 : (*e == 'a' ? -1 : atoi(e)); }();
    static int counter = 0;
    if (which != -2 && status && (which == -1 || which == counter)) a.trace = reinterpret_cast<long long*>(reinterpret_cast<char*>(status) + 1024);
    ++counter;
  }
#endif
  MGB_REQUIRE(io.ln_g == nullptr || (l.NT == 256 && l.Cout == 256 && l.up == 1), MGB_E_UNSUPPORTED,
              "the fused LayerNorm epilogue needs a 256-channel output");
  MGB_REQUIRE(io.in_chunks * 8 >= l.Cin, MGB_E_ARG, "input image narrower than the layer's input channels");
  MGB_REQUIRE(l.ntn * l.NT <= MAX_BIAS, MGB_E_UNSUPPORTED, "layer wider than %d output columns", MAX_BIAS);
#define MGB_TC_CASE(NT_, KC_)                                                                       \
  if (l.NT == NT_ && l.KC == KC_) return launch_conv<NT_, KC_>(a, io.in, io.in_lo, io.in_chunks, rin.Rp, rin.ntiles, l.ntn, s);
  MGB_TC_CASE(256, 64) MGB_TC_CASE(256, 32) MGB_TC_CASE(128, 64) MGB_TC_CASE(128, 32)
  MGB_TC_CASE(64, 64) MGB_TC_CASE(64, 32) MGB_TC_CASE(32, 64) MGB_TC_CASE(32, 32)
#undef MGB_TC_CASE
  MGB_REQUIRE(false, MGB_E_UNSUPPORTED, "no tensor-core tile for NT=%d KC=%d", l.NT, l.KC);
}

}  // namespace tcnet
}  // namespace mgb

// Generic fp32 Conv1d on frames-major activations, forward and backward: the building block of the JCU discriminator
// (reference: model/mixgantts.py:186-288 — ConvNorm model/blocks.py:326-371 with stride 1 or 2, LinearNorm :278-291,
// F.leaky_relu(., 0.2), Mish :894-896, DiffusionEmbedding :899-913) as train.py:126-184 drives it: 4 forwards and
// 2 backwards per step.  The discriminator is 0.65 MFLOP per mel frame (2.7 % of a Denoiser call), and its gradient has
// to match the reference's fp32 autograd.  Two implementations of the same implicit GEMMs live here:
//   conv_gemm_tc_kernel    forward and data gradient on the tensor cores at fp32 accuracy (3 x TF32: hi / lo operand split,
//                          three tcgen05.mma.kind::tf32 per k8-step, fp32 accumulation in TMEM) for channel counts that are
//                          multiples of 32 — every layer of the discriminator but the two logit convolutions
//   conv_gemm_f32_kernel, conv_wgrad_f32_kernel   exact-fp32 CUDA-core kernels: every other shape, and the weight gradient
// History of conv_gemm_tc_kernel on the 128 -> 512 (k = 5, stride 2) layer at 16 utterances x T = 800, B200, ncu with warm
// L2.  First versions: forward 62 us / data gradient 51 us, and nothing moved them by more than 10 % — fence.proxy.async
// skipped, mbarrier.test_wait spinning instead of try_wait, weight ring 3 -> 4 slots, one row per thread instead of eight
// lanes per row, B staged by the threads instead of pre-split tiles by cp.async.bulk, one __syncthreads per k-block instead
// of the mbarrier ring; only staging zeros instead of loading A did (30 / 49 us).  The cause was one line: the forward's
// "+ rowbias" was added inside the load function, right behind the load — also as a no-op select when there is no rowbias —
// so the loaded registers were consumed at once and the two-k-block prefetch never existed in the forward kernels.  With the
// rowbias values in their own prefetched registers and the add at stage-write time: forward 28 us, data gradient 50 us (the
// stride-2 data gradient stages, and multiplies, the zero rows of the taps that do not hit an input row).
//
// Layout: x [B][Tin][Cin], y [B][Tout][Cout] fp32, one row per frame (channels contiguous), Tout = (Tin - 1) / stride + 1
// for the reference's padding (k - 1) / 2.  A kernel tap is a row offset, so no im2col buffer exists anywhere:
//   forward   y[b][to][co]  = act(bias[co] + sum_j sum_ci xin[b][to*s + j - pad][ci] * w[co][ci][j])
//             xin = x (+ rowbias[b][ci] on the rows that exist: the discriminator's "x + diffusion_step (+ speaker)",
//             mixgantts.py:275-276, fused into the operand gather; padding rows stay zero as in the reference)
//   dgrad     dx[b][ti][ci] = sum_j sum_co dz[b][(ti + pad - j) / s][co] * w[co][ci][j]      (rows with an exact quotient)
//   wgrad     dw[co][ci][j] = sum_{b,to} xin[b][to*s + j - pad][ci] * dz[b][to][co]          (frames split over CTAs, partial
//             sums reduced in a fixed order: deterministic, no atomics)
// dz = dy * act'(.) is formed by a small elementwise kernel first; bias and rowbias gradients are column sums.
#include <algorithm>

#include "common.cuh"
#include "tc05.cuh"

namespace mgb {
namespace {

constexpr int TM = 128, TN = 64, TK = 16, NTHR = 256;      // CTA tile; every thread owns an 8 x 4 block of it
constexpr int LDA = TM + 4, LDB = TN + 4;

enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_MISH = 2, ACT_RELU = 3 };

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_MISH) {
    const float sp = v > 20.f ? v : log1pf(expf(v));   // F.softplus default threshold
    return v * tanhf(sp);
  }
  return v;
}

// Epilogue form: leaky_relu / relu / identity are one select with a slope (0.2 / 0 / 1); Mish is a call, not inlined — the
// per-element act_fwd() switch inlined 128 times per thread made the tensor-core forward kernel 23 000 SASS lines (395 MUFU
// sites, 360 KB of code: instruction-cache misses in every epilogue) and the forward 15 us slower than the data gradient of
// the same shape.
__device__ __noinline__ float mish_call(float v) {
  const float sp = v > 20.f ? v : log1pf(expf(v));
  return v * tanhf(sp);
}
__device__ __forceinline__ float act_slope(int act) { return act == ACT_LEAKY ? 0.2f : (act == ACT_RELU ? 0.f : 1.f); }
__device__ __forceinline__ float act_apply(float v, float slope, bool mish) {
  return mish ? mish_call(v) : (v > 0.f ? v : slope * v);
}

struct ConvShape { int B, Tin, Tout, Cin, Cout, k, stride, pad; };

// ---- weights: torch [Cout][Cin][k] -> wp [k][Cin][Cout] (forward / wgrad B operand) and wq [k][Cout][Cin] (dgrad) ----
__global__ void pack_w_kernel(const float* __restrict__ w, float* __restrict__ wp, float* __restrict__ wq, int Cin, int Cout, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * k) return;
  const int j = i % k, ci = (i / k) % Cin, co = i / (k * Cin);
  const float v = w[i];
  if (wp) wp[((size_t)j * Cin + ci) * Cout + co] = v;
  if (wq) wq[((size_t)j * Cout + co) * Cin + ci] = v;
}

// One 16-deep step of the register-tiled product: thread (tx = tid & 15, ty = tid >> 4) owns rows {4 ty .. 4 ty + 3} and
// {64 + 4 ty ..} x columns {4 tx .. 4 tx + 3} of the 128 x 64 CTA tile: 3 shared-memory float4 loads per 32 FFMA, 8 warps
// per CTA.  (History, ncu on the 128 -> 512 k = 5 stride-2 data gradient: 4 x 4 blocks — 2 loads per 16 FFMA — were bound
// by shared-memory bandwidth; 8 x 8 blocks with 4 warps per CTA left one warp per scheduler, which stalled 2.8 cycles per
// issued instruction on shared-memory and FFMA latencies.)
__device__ __forceinline__ void tile_fma(const float (*__restrict__ As)[LDA], const float (*__restrict__ Bs)[LDB], int tx, int ty,
                                         float (&acc)[8][4]) {
#pragma unroll
  for (int q = 0; q < TK; ++q) {
    const float4 a0 = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
    const float4 a1 = *reinterpret_cast<const float4*>(&As[q][64 + ty * 4]);
    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// ---- forward (DGRAD = false) and data gradient (DGRAD = true): out[M][N] = A[M][K] * Bm[K][N] ----------------------
//   forward: M = B*Tout rows, K = k*Cin, N = Cout, A(m, j*Cin + ci) = xin[b][to*s + j - pad][ci], Bm = wp
//   dgrad  : M = B*Tin  rows, K = k*Cout, N = Cin, A(m, j*Cout + co) = dz[b][(ti + pad - j)/s][co], Bm = wq
// Double-buffered shared-memory tiles: the operands of k-step i + 1 travel from global memory into registers while
// k-step i is multiplied.  gridDim.z > 1 = split K: CTA z accumulates k-steps [z * steps_per_split, ...) and writes raw
// partial sums to `part[z][M][N]`; splitk_epilogue_kernel adds them in a fixed order and applies bias / activation (the
// discriminator's deep layers are 25-100 tiles of K = 2560: without the split two thirds of the SMs idle).
template <bool DGRAD>
__global__ void __launch_bounds__(NTHR, 2) conv_gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             const float* __restrict__ bias, const float* __restrict__ rowbias,
                                                             float* __restrict__ out, float* __restrict__ pre,
                                                             float* __restrict__ part, ConvShape s, int act, int steps_per_split) {
  __shared__ __align__(16) float As[2][TK][LDA];
  __shared__ __align__(16) float Bs[2][TK][LDB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int rows_per_b = DGRAD ? s.Tin : s.Tout;       // rows of `out` per utterance
  const int src_per_b = DGRAD ? s.Tout : s.Tin;        // rows of `A` per utterance
  const int Ca = DGRAD ? s.Cout : s.Cin;               // channels of A
  const int N = DGRAD ? s.Cin : s.Cout;
  const int M = s.B * rows_per_b, K = s.k * Ca;
  const int nks_all = (K + TK - 1) / TK;
  const int ks0 = blockIdx.z * steps_per_split;
  const int nks = min(steps_per_split, nks_all - ks0);
  const bool vecA = (Ca & 3) == 0, vecB = (N & 3) == 0;
  // A loader: thread -> rows a_r + 64 i (i < 2), 4 consecutive k at a_k4; B loader: k row b_k, 4 n at b_n4
  const int a_r = tid >> 2, a_k4 = (tid & 3) * 4;
  const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;
  int a_b[2], a_t[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + a_r + 64 * i;
    a_b[i] = m < M ? m / rows_per_b : -1;
    a_t[i] = m < M ? m - a_b[i] * rows_per_b : 0;
  }
  // (tap j, channel c) of this thread's first k of the current step, advanced by TK per step
  int kj = (ks0 * TK + a_k4) / Ca, kc = (ks0 * TK + a_k4) - kj * Ca;

  auto src_row = [&](int t, int j, int& src) -> bool {
    if (DGRAD) { const int q = t + s.pad - j; src = q / s.stride; return q >= 0 && q % s.stride == 0 && src < s.Tout; }
    src = t * s.stride + j - s.pad;
    return src >= 0 && src < s.Tin;
  };
  const bool has_rb = !DGRAD && rowbias != nullptr;
  float4 pa[2], prb[2], pb;        // prb: the rowbias values, added when the tile is stored (not behind the load: see the header)
  auto load_g = [&](int ks) {
    const int kk = ks * TK + a_k4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      prb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_b[i] >= 0 && kk < K) {
        if (vecA) {
          int src;
          if (src_row(a_t[i], kj, src)) {
            v = *reinterpret_cast<const float4*>(A + ((size_t)a_b[i] * src_per_b + src) * Ca + kc);
            if (has_rb) prb[i] = *reinterpret_cast<const float4*>(rowbias + (size_t)a_b[i] * Ca + kc);
          }
        } else {
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ke = kk + e;
            if (ke >= K) break;
            const int j = ke / Ca, c = ke - j * Ca;
            int src;
            if (src_row(a_t[i], j, src))
              e4[e] = A[((size_t)a_b[i] * src_per_b + src) * Ca + c] + ((!DGRAD && rowbias) ? rowbias[(size_t)a_b[i] * Ca + c] : 0.f);
          }
          v = make_float4(e4[0], e4[1], e4[2], e4[3]);
        }
      }
      pa[i] = v;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kr = ks * TK + b_k, n = n0 + b_n4;
      if (kr < K) {
        if (vecB && n + 3 < N) v = *reinterpret_cast<const float4*>(Bm + (size_t)kr * N + n);
        else {
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < N) e4[e] = Bm[(size_t)kr * N + n + e];
          v = make_float4(e4[0], e4[1], e4[2], e4[3]);
        }
      }
      pb = v;
    }
    kc += TK;
    while (kc >= Ca) { kc -= Ca; ++kj; }
  };
  auto store_s = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = pa[i];
      if (has_rb) { v.x += prb[i].x; v.y += prb[i].y; v.z += prb[i].z; v.w += prb[i].w; }
      As[buf][a_k4 + 0][a_r + 64 * i] = v.x; As[buf][a_k4 + 1][a_r + 64 * i] = v.y;
      As[buf][a_k4 + 2][a_r + 64 * i] = v.z; As[buf][a_k4 + 3][a_r + 64 * i] = v.w;
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n4]) = pb;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (nks > 0) {
    load_g(ks0);
    store_s(0);
    __syncthreads();
    for (int it = 0; it < nks; ++it) {
      const int buf = it & 1;
      if (it + 1 < nks) load_g(ks0 + it + 1);
      tile_fma(As[buf], Bs[buf], tx, ty, acc);
      if (it + 1 < nks) store_s(buf ^ 1);
      __syncthreads();
    }
  }
  const bool split = gridDim.z > 1;
  float* dst = split ? part + (size_t)blockIdx.z * M * N : out;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    {
      const int n = n0 + tx * 4;
      float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
      if (!split && !DGRAD) {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (bias && n + e < N) v[e] += bias[n + e];
        if (pre) {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < N) pre[(size_t)m * N + n + e] = v[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = act_apply(v[e], act_slope(act), act == ACT_MISH);
      }
      if (vecB && n + 3 < N) *reinterpret_cast<float4*>(dst + (size_t)m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (n + e < N) dst[(size_t)m * N + n + e] = v[e];
      }
    }
  }
}

// out = act(bias + sum_z part[z]) in a fixed order (the epilogue of a split-K conv_gemm_f32_kernel)
__global__ void splitk_epilogue_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ out,
                                       float* __restrict__ pre, size_t MN, int N, int nsplit, int act) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float v = 0.f;
  for (int z = 0; z < nsplit; ++z) v += part[(size_t)z * MN + i];
  if (bias) v += bias[i % N];
  if (pre) pre[i] = v;
  out[i] = act_fwd(v, act);
}

// ---- the same GEMM on the tensor cores at fp32 accuracy: 3 x TF32 ------------------------------------------------------
// out[128 x 64 tile] = sum over 32-deep k-blocks of A B with every fp32 operand split as x = hi + lo, hi = x with the low 13
// mantissa bits cleared (exactly what kind::tf32 reads), lo = x - hi (exact in fp32): three tcgen05.mma.kind::tf32 per
// k8-step, lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM; the dropped terms are below 2^-21 of a product, so the result is
// fp32-grade (the parity tests hold it to the same 1e-5 as the CUDA-core kernels).
//   A operand: gathered by the CTA's own threads, not by TMA — it is the implicit-convolution gather (tap shift, stride,
//     padding rows, "+ rowbias") of conv_gemm_f32_kernel, 32 channels (128 contiguous bytes) of a row per k-block, split and
//     stored as K-major core matrices (8 rows x 16 bytes; per 4-channel group the 128 rows are one dense [128][16 B] array,
//     SBO = 128 B, LBO = 2 KB + 16 B).
//   B operand: the weights, split and tiled ONCE per call by pack_w_tc_kernel into exactly the shared-memory image of a
//     (64-column tile, k-block) pair — hi then lo, 16 KB — so one cp.async.bulk per k-block brings it (4-slot ring, issued
//     two k-blocks ahead, completion on the slot's mbarrier).  (Staging B with the threads as well took 480
//     instructions per warp and k-block at 7.5 stall cycles each: ncu, profiles/r02.)
// Warp-specialised, no CTA-wide barrier in the loop: 128 staging threads (their global loads run two k-blocks ahead in
// registers) publish an A stage through an mbarrier; one thread of a fifth warp issues the B copies and the MMAs as soon as
// both operands of a k-block have landed, and tcgen05.commit on the stage's mbarrier hands it back.  (The first version —
// one __syncthreads per k-block, MMAs issued by staging thread 0 — spent 4 400 cycles per k-block on that serial chain.)  Needs Ca % 32 == 0 (a k-block inside one tap) and N % 4 == 0; other
// shapes take the CUDA-core kernel.  Split K as there (blockIdx.z), same epilogue kernel.
constexpr int TC_BM = 128, TC_BK = 32, TC_THREADS = 160;    // 4 staging / epilogue warps + 1 warp issuing MMAs and B copies
constexpr int TC_A_LBO = TC_BM * 16 + 16;                   // + 16 B, so the 8 channel groups of a row hit 8 bank groups
constexpr int TC_A_BYTES = 8 * TC_A_LBO;                    // one of hi / lo: 16.1 KB
constexpr int TC_A_STAGE = 2 * TC_A_BYTES;                  // hi + lo
constexpr int TC_NA = 2, TC_NB = 4;                         // B copies run two k-blocks ahead: a 32 KB bulk copy from L2 takes
                                                            // ~2 000 cycles; one k-block ahead it paced the whole loop (1.1 us per k-block)
// column tile BN = 128 (64 for narrow outputs): a tcgen05.mma costs the issuing thread ~300 cycles whatever its size (twelve
// N = 64 MMAs per k-block were 3 800 cycles, the whole loop time), so the tile is as wide as shared memory allows
template <int BN> struct TcCfg {
  static constexpr int B_LBO = BN * 16, B_BYTES = 8 * B_LBO, B_STAGE = 2 * B_BYTES;      // BN = 128: 32 KB per slot
  static constexpr int SMEM = TC_NA * TC_A_STAGE + TC_NB * B_STAGE + 128;                // 194 KB (BN = 128), 130 KB (64)
};
constexpr long long TC_TIMEOUT = 400000000LL;

__device__ __forceinline__ float cut_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void split_tf32(const float4 v, float4& hi, float4& lo) {
  hi = make_float4(cut_tf32(v.x), cut_tf32(v.y), cut_tf32(v.z), cut_tf32(v.w));
  lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);     // the tensor core reads its top 19 bits
}

// B[kk][n] (forward: kk = j*Cin + ci, n = co; data gradient: kk = j*Cout + co, n = ci) = w[co][ci][j], split and written as
// tiles [n / BN][kk / 32] of {hi, lo} x [kc = (kk % 32) / 4][n % BN][kk % 4]; columns beyond N are zero
__global__ void pack_w_tc_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cin, int Cout, int k, int dgrad, int BN) {
  const int Ca = dgrad ? Cout : Cin, N = dgrad ? Cin : Cout;
  const int K = k * Ca, Npad = (N + BN - 1) / BN * BN, nkb = K / TC_BK;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * Npad) return;
  // the fastest thread index walks w with the smaller stride: ci (stride k floats) — that is kk forward, n in the data gradient
  const int n = dgrad ? i % Npad : i / K, kk = dgrad ? i / Npad : i % K;
  const int j = kk / Ca, c = kk - j * Ca;
  float v = 0.f;
  if (n < N) v = dgrad ? w[((size_t)c * Cin + n) * k + j] : w[((size_t)n * Cin + c) * k + j];
  const float hi = cut_tf32(v), lo = v - hi;
  const int nt = n / BN, kb = kk / TC_BK, kc = (kk % TC_BK) / 4, e = kk & 3, nn = n % BN;
  const int half = 8 * BN * 4;                                 // floats of one of hi / lo
  float* t = dst + ((size_t)nt * nkb + kb) * (2 * half);
  t[kc * (BN * 4) + nn * 4 + e] = hi;
  t[half + kc * (BN * 4) + nn * 4 + e] = lo;
}

template <bool DGRAD, int BN>
__global__ void __launch_bounds__(TC_THREADS) conv_gemm_tc_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                                  const float* __restrict__ bias, const float* __restrict__ rowbias,
                                                                  float* __restrict__ out, float* __restrict__ pre,
                                                                  float* __restrict__ part, ConvShape s, int act, int blocks_per_split) {
  using Cfg = TcCfg<BN>;
  constexpr int TC_BN = BN, TC_B_LBO = Cfg::B_LBO, TC_B_BYTES = Cfg::B_BYTES, TC_B_STAGE = Cfg::B_STAGE;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* smA = tc_smem;
  uint8_t* smB = tc_smem + TC_NA * TC_A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + TC_NA * TC_A_STAGE + TC_NB * TC_B_STAGE);
  uint64_t* bar_free = bars;               // [2] the tensor core has read A stage i (and the B slot used with it)
  uint64_t* bar_afull = bars + 2;          // [2] all 128 staging threads have written A stage i
  uint64_t* bar_bfull = bars + 4;          // [4] B slot landed
  uint64_t* bar_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BN;
  const int rows_per_b = DGRAD ? s.Tin : s.Tout, src_per_b = DGRAD ? s.Tout : s.Tin;
  const int Ca = DGRAD ? s.Cout : s.Cin, N = DGRAD ? s.Cin : s.Cout;
  const int M = s.B * rows_per_b, K = s.k * Ca;
  const int nkb_all = K / TC_BK;
  const int kb0 = blockIdx.z * blocks_per_split;
  const int nkb = min(blocks_per_split, nkb_all - kb0);

  if (tid == 0) {
    tc::mbar_init(&bar_free[0], 1); tc::mbar_init(&bar_free[1], 1);
    tc::mbar_init(&bar_afull[0], 128); tc::mbar_init(&bar_afull[1], 128);
    for (int i = 0; i < TC_NB; ++i) tc::mbar_init(&bar_bfull[i], 1);
    tc::mbar_init(bar_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 4) tc::tmem_alloc<TC_BN>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ---- one elected thread of this warp: B copies one k-block ahead, MMAs as soon as both operands of a k-block are in
    // shared memory.  The whole warp walks the loop (uniform control flow, elect-predicated issue): the descriptors stay
    // in uniform registers.
    if (nkb > 0) {
      const float* btile = Bt + ((size_t)blockIdx.y * nkb_all + kb0) * (TC_B_STAGE / 4);
      constexpr uint32_t IDESC = tc::make_idesc_tf32(TC_BM, TC_BN);
      const uint32_t smA_u = tc::smem_u32(smA), smB_u = tc::smem_u32(smB);
      if (tc::elect_one()) {
        for (int i = 0; i < 2 && i < nkb; ++i) {
          tc::mbar_arrive_expect_tx(&bar_bfull[i], TC_B_STAGE);
          tc::bulk_g2s(smB + i * TC_B_STAGE, btile + (size_t)i * (TC_B_STAGE / 4), TC_B_STAGE, &bar_bfull[i]);
        }
      }
      for (int it = 0; it < nkb; ++it) {
        const int stage = it & 1, slot = it % TC_NB;
        if (it + 2 < nkb) {
          // B slot (it + 2) % 4 was read by the MMAs of k-block it - 2
          if (it >= 2 && !tc::mbar_wait(&bar_free[stage], ((it >> 1) + 1) & 1, TC_TIMEOUT)) __trap();
          const int ns = (it + 2) % TC_NB;
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(&bar_bfull[ns], TC_B_STAGE);
            tc::bulk_g2s(smB + ns * TC_B_STAGE, btile + (size_t)(it + 2) * (TC_B_STAGE / 4), TC_B_STAGE, &bar_bfull[ns]);
          }
        }
        if (!tc::mbar_wait(&bar_afull[stage], (it >> 1) & 1, TC_TIMEOUT)) __trap();
        if (!tc::mbar_wait(&bar_bfull[slot], (it / TC_NB) & 1, TC_TIMEOUT)) __trap();
        tc::tc_fence_after();
        const uint32_t abase = smA_u + stage * TC_A_STAGE, bbase = smB_u + slot * TC_B_STAGE;
        // descriptors of k8-step 0; a step further is two core-matrix columns = 2 LBO bytes (>> 4 in the address field)
        const uint64_t a_hi = tc::make_smem_desc(abase, TC_A_LBO, 128), a_lo = tc::make_smem_desc(abase + TC_A_BYTES, TC_A_LBO, 128);
        const uint64_t b_hi = tc::make_smem_desc(bbase, TC_B_LBO, 128), b_lo = tc::make_smem_desc(bbase + TC_B_BYTES, TC_B_LBO, 128);
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            const uint64_t da = (uint64_t)(ks * 2 * TC_A_LBO >> 4), db = (uint64_t)(ks * 2 * TC_B_LBO >> 4);
            tc::umma_tf32(tmem, a_lo + da, b_hi + db, IDESC, (it > 0 || ks > 0) ? 1u : 0u);      // small terms first
            tc::umma_tf32(tmem, a_hi + da, b_lo + db, IDESC, 1u);
            tc::umma_tf32(tmem, a_hi + da, b_hi + db, IDESC, 1u);
          }
          tc::umma_commit(&bar_free[stage]);
          if (it + 1 == nkb) tc::umma_commit(bar_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- 128 staging threads.  Eight lanes read the 128 contiguous bytes (32 channels) of one A row, so a warp-wide 16-byte
    // load touches 4 lines (one row per thread touched 32, and the loop waited on the L1 request queue: ncu long-scoreboard
    // stalls at the first use of the loaded registers, even two k-blocks ahead).  Thread -> rows 32 warp + 4 i + lane / 8
    // (i < 8), 4-channel group lane % 8; per row only (base offset, first source row) are kept, a k-block costs an add and
    // two compares per row.  The global loads run two k-blocks ahead in registers.
    const int lane = tid & 31, a_kc = lane & 7, a_r0 = warp * 32 + (lane >> 3);
    long long a_base[8];          // element offset of (utterance, source row 0) + this thread's channel group; < 0: no such row
    int a_q[8];                   // forward: t * stride - pad (source row = a_q + j); data gradient: t + pad (q = a_q - j)
    int a_rb[8];                  // forward: offset of this thread's channel group in the utterance's rowbias row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mi = m0 + a_r0 + 4 * i;
      if (mi < M) {
        const int b = mi / rows_per_b, t = mi - b * rows_per_b;
        a_base[i] = (long long)b * src_per_b * Ca + a_kc * 4;
        a_q[i] = DGRAD ? t + s.pad : t * s.stride - s.pad;
        a_rb[i] = b * Ca + a_kc * 4;
      } else { a_base[i] = -1; a_q[i] = 0; a_rb[i] = 0; }
    }
    const bool has_rb = !DGRAD && rowbias != nullptr;
    // k-block kb of the whole K axis -> registers; the rowbias values travel in their own registers and are added when the
    // stage is written (adding them here made every load's use immediate: no prefetch for the layer that has a rowbias)
    auto load_g = [&](int kb, float4 (&ra)[8], float4 (&rb)[8]) {
      const int kk = kb * TC_BK;
      const int j = kk / Ca, c0 = kk - j * Ca;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int src;
        bool ok = a_base[i] >= 0;
        if (DGRAD) {
          const int q = a_q[i] - j;
          if (s.stride == 1) src = q;
          else if (s.stride == 2) { src = q >> 1; ok = ok && (q & 1) == 0; }
          else { src = q / s.stride; ok = ok && q % s.stride == 0; }
          ok = ok && q >= 0 && src < s.Tout;
        } else {
          src = a_q[i] + j;
          ok = ok && src >= 0 && src < s.Tin;
        }
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          v = *reinterpret_cast<const float4*>(A + a_base[i] + (long long)src * Ca + c0);
          if (has_rb) r = *reinterpret_cast<const float4*>(rowbias + a_rb[i] + c0);
        }
        ra[i] = v;
        if (!DGRAD) rb[i] = r;
      }
    };
    auto stage_step = [&](int it, float4 (&ra)[8], float4 (&rb)[8]) {   // publish k-block `it` from ra, then refill ra with it + 2
      const int stage = it & 1;
      if (it >= 2 && !tc::mbar_wait(&bar_free[stage], ((it >> 1) + 1) & 1, TC_TIMEOUT)) __trap();
      uint8_t* st = smA + stage * TC_A_STAGE + a_kc * TC_A_LBO + a_r0 * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 hi, lo, v = ra[i];
        if (has_rb) { v.x += rb[i].x; v.y += rb[i].y; v.z += rb[i].z; v.w += rb[i].w; }
        split_tf32(v, hi, lo);
        *reinterpret_cast<float4*>(st + i * 64) = hi;
        *reinterpret_cast<float4*>(st + TC_A_BYTES + i * 64) = lo;
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&bar_afull[stage]);
      if (it + 2 < nkb) load_g(kb0 + it + 2, ra, rb);
    };
    float4 r0[8], r1[8], q0[8], q1[8];
    if (nkb > 0) load_g(kb0, r0, q0);
    if (nkb > 1) load_g(kb0 + 1, r1, q1);
    for (int it = 0; it < nkb; it += 2) {
      stage_step(it, r0, q0);
      if (it + 1 < nkb) stage_step(it + 1, r1, q1);
    }
    const int m = m0 + tid;                               // epilogue: thread = output row = TMEM lane
    // epilogue
    const bool split = gridDim.z > 1;
    const float slope = act_slope(act);
    const bool mish = act == ACT_MISH;
    float* dst = split ? part + (size_t)blockIdx.z * M * N : out;
    if (nkb > 0 && !tc::mbar_wait(bar_done, 0, TC_TIMEOUT)) __trap();
    tc::tc_fence_after();
#pragma unroll
    for (int h = 0; h < TC_BN / 32; ++h) {
      uint32_t r[32];
      if (nkb > 0) {
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + h * 32, r);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (m < M) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const int n = n0 + h * 32 + c4 * 4;
          if (n >= N) break;                                   // N % 4 == 0
          float v[4] = {__uint_as_float(r[c4 * 4]), __uint_as_float(r[c4 * 4 + 1]), __uint_as_float(r[c4 * 4 + 2]),
                        __uint_as_float(r[c4 * 4 + 3])};
          if (!split && !DGRAD) {
            if (bias) {
              const float4 bv = *reinterpret_cast<const float4*>(bias + n);
              v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
            }
            if (pre) *reinterpret_cast<float4*>(pre + (size_t)m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = act_apply(v[e], slope, mish);
          }
          *reinterpret_cast<float4*>(dst + (size_t)m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<TC_BN>(tmem);
}

// ---- dz = dy * act'(.)  (leaky / relu from the sign of y, mish from the stored pre-activation) ----
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ pre,
                               float* __restrict__ dz, size_t n, int act) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = dy[i];
  if (act == ACT_LEAKY) g = y[i] > 0.f ? g : 0.2f * g;
  else if (act == ACT_RELU) g = y[i] > 0.f ? g : 0.f;
  else if (act == ACT_MISH) {
    const float x = pre[i];
    const float sp = x > 20.f ? x : log1pf(expf(x));
    const float th = tanhf(sp);
    const float sg = 1.f / (1.f + expf(-x));                    // d softplus / dx (1 above the threshold, where sg == 1 in fp32)
    g *= th + x * (1.f - th * th) * sg;
  }
  dz[i] = g;
}

// ---- weight gradient partials: part[split][K][Cout] = sum over the split's rows of xin(row, kk) * dz(row, co) ----
// The same 128 x 64 register-tiled product with the frame axis as the reduction: As[q][kk] = xin(row q, kk) and
// Bs[q][co] = dz(row q, co) are both contiguous along their tile axis, so the tiles are stored as they are loaded.
__global__ void __launch_bounds__(NTHR, 2) conv_wgrad_f32_kernel(const float* __restrict__ x, const float* __restrict__ rowbias,
                                                              const float* __restrict__ dz, float* __restrict__ part, ConvShape s,
                                                              int rows_per_split) {
  __shared__ __align__(16) float As[2][TK][LDA];
  __shared__ __align__(16) float Bs[2][TK][LDB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * TM, n0 = blockIdx.y * TN, split = blockIdx.z;
  const int K = s.k * s.Cin, M = s.B * s.Tout;
  const int mb = split * rows_per_split, me = min(M, mb + rows_per_split);
  const bool vecA = (s.Cin & 3) == 0, vecB = (s.Cout & 3) == 0;
  // A loader: rows a_q + 8 i (i < 2), 4 consecutive kk at a_k4 (fixed for the whole kernel, so is its tap / channel)
  const int a_q = tid >> 5, a_k4 = (tid & 31) * 4;
  const int b_q = tid >> 4, b_n4 = (tid & 15) * 4;
  const int kk = k0 + a_k4;
  const int kj = kk / s.Cin, kc = kk - kj * s.Cin;
  // the rowbias values travel in their own registers (prb) and are added when the tile is stored: an add right behind the
  // load — even predicated off — waits for the load and cancels the prefetch (see the header)
  float4 pa[2], prb[2], pb;
  auto load_g = [&](int mq) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      prb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int m = mq + a_q + 8 * i;
      if (m < me && kk < K) {
        const int b = m / s.Tout, to = m - b * s.Tout;
        if (vecA) {
          const int ti = to * s.stride + kj - s.pad;
          if (ti >= 0 && ti < s.Tin) {
            v = *reinterpret_cast<const float4*>(x + ((size_t)b * s.Tin + ti) * s.Cin + kc);
            if (rowbias) prb[i] = *reinterpret_cast<const float4*>(rowbias + (size_t)b * s.Cin + kc);
          }
        } else {
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ke = kk + e;
            if (ke >= K) break;
            const int j = ke / s.Cin, c = ke - j * s.Cin;
            const int ti = to * s.stride + j - s.pad;
            if (ti >= 0 && ti < s.Tin) e4[e] = x[((size_t)b * s.Tin + ti) * s.Cin + c] + (rowbias ? rowbias[(size_t)b * s.Cin + c] : 0.f);
          }
          v = make_float4(e4[0], e4[1], e4[2], e4[3]);
        }
      }
      pa[i] = v;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int m = mq + b_q, n = n0 + b_n4;
      if (m < me) {
        if (vecB && n + 3 < s.Cout) v = *reinterpret_cast<const float4*>(dz + (size_t)m * s.Cout + n);
        else {
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < s.Cout) e4[e] = dz[(size_t)m * s.Cout + n + e];
          v = make_float4(e4[0], e4[1], e4[2], e4[3]);
        }
      }
      pb = v;
    }
  };
  auto store_s = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = pa[i];
      if (rowbias) { v.x += prb[i].x; v.y += prb[i].y; v.z += prb[i].z; v.w += prb[i].w; }
      *reinterpret_cast<float4*>(&As[buf][a_q + 8 * i][a_k4]) = v;
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_q][b_n4]) = pb;
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nch = (me - mb + TK - 1) / TK;
  if (nch > 0) {
    load_g(mb);
    store_s(0);
    __syncthreads();
    for (int it = 0; it < nch; ++it) {
      const int buf = it & 1;
      if (it + 1 < nch) load_g(mb + (it + 1) * TK);
      tile_fma(As[buf], Bs[buf], tx, ty, acc);
      if (it + 1 < nch) store_s(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kr = k0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (kr >= K) continue;
    {
      const int n = n0 + tx * 4;
      float* d = part + ((size_t)split * K + kr) * s.Cout + n;
      if (vecB && n + 3 < s.Cout) *reinterpret_cast<float4*>(d) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (n + e < s.Cout) d[e] = acc[i][e];
      }
    }
  }
}

// dw[co][ci][j] (torch layout) = sum over splits (fixed order) of part[split][j*Cin + ci][co].  Thread = one (kk, co) element
// of the partial layout: the nsplit reads of a warp are 128 contiguous bytes each, only the single write is scattered
// (indexing by the torch layout instead made every one of the nsplit reads a 4-byte access Cin * Cout floats apart).
__global__ void wgrad_reduce_f32_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cin, int Cout, int k, int nsplit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = Cout * Cin * k;
  if (i >= n) return;
  const int co = i % Cout, kk = i / Cout;
  const int j = kk / Cin, ci = kk - j * Cin;
  float a = 0.f;
  for (int s = 0; s < nsplit; ++s) a += part[(size_t)s * n + i];
  dw[((size_t)co * Cin + ci) * k + j] = a;
}

// out[g][c] = sum over the group's rows of a[row][c]   (bias gradient: one group; rowbias gradient: one group per utterance).
// Two deterministic stages: 256-thread blocks (32 columns x 8 row lanes, 128-byte coalesced rows) sum one row split each
// into part[split][g][c] in a fixed order, then one thread per output adds the splits in order.  (A single 128-thread
// block walking all 6400 rows took 100-390 us per call and a quarter of the discriminator's time.)
constexpr int CS_MAX_SPLITS = 64;
__global__ void __launch_bounds__(256) colsum_part_kernel(const float* __restrict__ a, float* __restrict__ part, int rows_per_group,
                                                          int C, int rows_per_split) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, g = blockIdx.y, sp = blockIdx.z;
  const int r0 = sp * rows_per_split, r1 = min(rows_per_group, r0 + rows_per_split);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    const float* p = a + (size_t)g * rows_per_group * C + c;
    int r = r0 + ry;
    for (; r + 24 < r1; r += 32) {
      s0 += p[(size_t)r * C]; s1 += p[(size_t)(r + 8) * C]; s2 += p[(size_t)(r + 16) * C]; s3 += p[(size_t)(r + 24) * C];
    }
    for (; r < r1; r += 8) s0 += p[(size_t)r * C];
  }
  red[ry][cx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][cx];
    part[((size_t)sp * gridDim.y + g) * C + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int GC, int nsplit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= GC) return;
  float t = 0.f;
  for (int sp = 0; sp < nsplit; ++sp) t += part[(size_t)sp * GC + i];
  out[i] = t;
}
inline int colsum_splits(int rows_per_group) {
  int sp = (rows_per_group + 127) / 128;
  return sp < 1 ? 1 : (sp > CS_MAX_SPLITS ? CS_MAX_SPLITS : sp);
}
inline void launch_colsum(const float* a, float* out, float* part, int groups, int rows_per_group, int C, cudaStream_t st) {
  const int sp = colsum_splits(rows_per_group);
  const int rps = ((rows_per_group + sp - 1) / sp + 7) / 8 * 8;
  colsum_part_kernel<<<dim3((C + 31) / 32, groups, sp), 256, 0, st>>>(a, part, rows_per_group, C, rps);
  colsum_final_kernel<<<(groups * C + 255) / 256, 256, 0, st>>>(part, out, groups * C, sp);
}

// emb[b] = [sin(t f_i), cos(t f_i)], f_i = exp(-i ln(1e4) / (dim/2 - 1))     (blocks.py:906-913)
__global__ void step_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ emb, int dim) {
  const int b = blockIdx.x, halfd = dim / 2;
  const float tv = (float)t[b];
  const float scale = (float)(9.210340371976184 / (double)(halfd - 1));
  for (int i = threadIdx.x; i < halfd; i += blockDim.x) {
    const float a = tv * expf((float)i * -scale);
    emb[(size_t)b * dim + i] = sinf(a);
    emb[(size_t)b * dim + halfd + i] = cosf(a);
  }
}

bool shape_ok(int B, int Tin, int Cin, int Cout, int k, int stride) {
  return B > 0 && Tin > 0 && Cin > 0 && Cout > 0 && k > 0 && (k & 1) == 1 && stride >= 1 && stride <= 8;
}
ConvShape make_shape(int B, int Tin, int Cin, int Cout, int k, int stride) {
  ConvShape s{};
  s.B = B; s.Tin = Tin; s.Cin = Cin; s.Cout = Cout; s.k = k; s.stride = stride; s.pad = (k - 1) / 2;
  s.Tout = (Tin + 2 * s.pad - k) / stride + 1;
  return s;
}
int wgrad_splits(const ConvShape& s) {
  const int M = s.B * s.Tout;
  const int tiles = ((s.k * s.Cin + TM - 1) / TM) * ((s.Cout + TN - 1) / TN);
  int want = 296 / tiles;                                // at most two CTAs per SM in total: one wave
  const int max_by_rows = (M + 127) / 128;               // at least 128 rows per split
  if (want > max_by_rows) want = max_by_rows;
  return want < 1 ? 1 : (want > 64 ? 64 : want);
}
// split K of the forward / data-gradient GEMM: up to two CTAs per SM in total, at least 8 k-steps (128 deep) per split
int gemm_splits(int M, int N, int K) {
  const int tiles = ((M + TM - 1) / TM) * ((N + TN - 1) / TN), nks = (K + TK - 1) / TK;
  int want = 296 / tiles;                                // never more CTAs than fit at once (2 per SM): a 300-CTA grid
                                                         // ran 4 of them alone, after the rest, for twice the time
  if (want > nks / 8) want = nks / 8;
  return want < 1 ? 1 : (want > 16 ? 16 : want);
}
// tensor-core plan: 128 x BN tiles (BN = 128, or 64 for outputs of at most 64 columns), 32-deep k-blocks; split K until
// about one CTA per SM exists, at least 4 k-blocks per split
bool tc_eligible(int N, int Ca) { return Ca % TC_BK == 0 && N % 4 == 0 && N >= 16; }
int tc_bn(int N) { return N > 64 ? 128 : 64; }
int tc_splits(int M, int N, int K) {
  const int bn = tc_bn(N);
  const int tiles = ((M + TC_BM - 1) / TC_BM) * ((N + bn - 1) / bn), nkb = K / TC_BK;
  int want = 148 / tiles;
  if (want > nkb / 4) want = nkb / 4;
  return want < 1 ? 1 : (want > 16 ? 16 : want);
}
struct Work { size_t wp, wq, dz, part, cs, cs2, gpart, total; };
Work work_layout(const ConvShape& s) {
  Work w{};
  size_t p = 0;
  auto take = [&](size_t n) { size_t r = p; p += align_up(n * sizeof(float), 256); return r; };
  const size_t wn = (size_t)s.k * s.Cin * s.Cout;
  size_t wbuf = wn;                                // one packing buffer: [K][N] (CUDA-core kernels) or the hi / lo tiles
  auto padn = [](int n) { const int bn = tc_bn(n); return (size_t)((n + bn - 1) / bn * bn); };
  if (tc_eligible(s.Cout, s.Cin)) wbuf = std::max(wbuf, 2 * (size_t)s.k * s.Cin * padn(s.Cout));
  if (tc_eligible(s.Cin, s.Cout)) wbuf = std::max(wbuf, 2 * (size_t)s.k * s.Cout * padn(s.Cin));
  w.wp = take(wbuf); w.wq = w.wp;
  w.dz = take((size_t)s.B * s.Tout * s.Cout);
  w.part = take((size_t)wgrad_splits(s) * wn);
  w.cs = take((size_t)CS_MAX_SPLITS * (size_t)((size_t)s.B * s.Cin > (size_t)s.Cout ? (size_t)s.B * s.Cin : (size_t)s.Cout));
  w.cs2 = take((size_t)CS_MAX_SPLITS * (size_t)s.Cout);          // the bias column sum runs beside the rowbias one (side stream)
  int sf = gemm_splits(s.B * s.Tout, s.Cout, s.k * s.Cin), sd = gemm_splits(s.B * s.Tin, s.Cin, s.k * s.Cout);
  if (tc_eligible(s.Cout, s.Cin)) sf = tc_splits(s.B * s.Tout, s.Cout, s.k * s.Cin);
  if (tc_eligible(s.Cin, s.Cout)) sd = tc_splits(s.B * s.Tin, s.Cin, s.k * s.Cout);
  const size_t gf = sf > 1 ? (size_t)sf * s.B * s.Tout * s.Cout : 0, gd = sd > 1 ? (size_t)sd * s.B * s.Tin * s.Cin : 0;
  w.gpart = take(gf > gd ? gf : gd);
  w.total = p;
  return w;
}

// forward (DGRAD = false) or data gradient; returns the number of launches
// forward (DGRAD = false) or data gradient; `w` = torch-layout weights [Cout][Cin][k], `wbuf` = the packing buffer of the
// workspace; returns the number of launches
template <bool DGRAD>
int launch_conv_gemm(const float* A, const float* w, float* wbuf, const float* bias, const float* rowbias, float* out, float* pre,
                     float* gpart, const ConvShape& s, int act, cudaStream_t st) {
  const int M = DGRAD ? s.B * s.Tin : s.B * s.Tout, N = DGRAD ? s.Cin : s.Cout, Ca = DGRAD ? s.Cout : s.Cin, K = s.k * Ca;
  const size_t MN = (size_t)M * N;
  if (tc_eligible(N, Ca)) {
    static PerDeviceOnce once;
    if (once.pending()) {
      cudaFuncSetAttribute(conv_gemm_tc_kernel<DGRAD, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::SMEM);
      cudaFuncSetAttribute(conv_gemm_tc_kernel<DGRAD, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM);
      once.done();
    }
    const int bn = tc_bn(N), Npad = (N + bn - 1) / bn * bn;
    pack_w_tc_kernel<<<(unsigned)(((size_t)K * Npad + 255) / 256), 256, 0, st>>>(w, wbuf, s.Cin, s.Cout, s.k, DGRAD ? 1 : 0, bn);
    const int nkb = K / TC_BK;
    int splits = tc_splits(M, N, K);
    const int bps = (nkb + splits - 1) / splits;
    splits = (nkb + bps - 1) / bps;
    dim3 grid((M + TC_BM - 1) / TC_BM, Npad / bn, splits);
    if (bn == 64)
      conv_gemm_tc_kernel<DGRAD, 64><<<grid, TC_THREADS, TcCfg<64>::SMEM, st>>>(A, wbuf, bias, rowbias, out, pre, gpart, s, act, bps);
    else
      conv_gemm_tc_kernel<DGRAD, 128><<<grid, TC_THREADS, TcCfg<128>::SMEM, st>>>(A, wbuf, bias, rowbias, out, pre, gpart, s, act, bps);
    if (splits == 1) return 2;
    splitk_epilogue_kernel<<<(unsigned)((MN + 255) / 256), 256, 0, st>>>(gpart, DGRAD ? nullptr : bias, out, DGRAD ? nullptr : pre, MN, N,
                                                                          splits, DGRAD ? ACT_NONE : act);
    return 3;
  }
  const int wn = s.Cout * s.Cin * s.k;
  pack_w_kernel<<<(wn + 255) / 256, 256, 0, st>>>(w, DGRAD ? nullptr : wbuf, DGRAD ? wbuf : nullptr, s.Cin, s.Cout, s.k);
  const int nks = (K + TK - 1) / TK;
  int splits = gemm_splits(M, N, K);
  const int sps = (nks + splits - 1) / splits;
  splits = (nks + sps - 1) / sps;                              // no empty split
  dim3 grid((M + TM - 1) / TM, (N + TN - 1) / TN, splits);
  conv_gemm_f32_kernel<DGRAD><<<grid, NTHR, 0, st>>>(A, wbuf, bias, rowbias, out, pre, gpart, s, act, sps);
  if (splits == 1) return 2;
  splitk_epilogue_kernel<<<(unsigned)((MN + 255) / 256), 256, 0, st>>>(gpart, DGRAD ? nullptr : bias, out, DGRAD ? nullptr : pre, MN, N,
                                                                        splits, DGRAD ? ACT_NONE : act);
  return 3;
}

// The backward of a layer is two independent halves once dz exists: the data gradient (tensor cores, what the previous layer's
// backward waits for) and the weight / bias gradients (CUDA cores).  With both requested, the second half runs on a
// library-owned side stream (one per device, created at first use), forked and joined by events inside the call — the caller's
// stream sees finished gradients when the call's work completes, and inside a CUDA-graph capture the fork/join become edges.
struct ConvSide { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
ConvSide* conv_side() {
  static ConvSide ctx[256];
  static PerDeviceOnce once;
  const int dev = PerDeviceOnce::current();
  if (once.pending()) {
    ConvSide& c = ctx[dev];
    if (cudaStreamCreateWithFlags(&c.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    once.done();
  }
  return &ctx[dev];
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_conv1d_out_len(int Tin, int k, int stride) { return (Tin + 2 * ((k - 1) / 2) - k) / stride + 1; }

size_t mgb_conv1d_workspace_bytes(int B, int Tin, int Cin, int Cout, int k, int stride) {
  if (!shape_ok(B, Tin, Cin, Cout, k, stride)) return 0;
  return work_layout(make_shape(B, Tin, Cin, Cout, k, stride)).total;
}

int mgb_conv1d_forward(const float* x, const float* w, const float* bias, const float* rowbias, float* y, float* pre, int B,
                       int Tin, int Cin, int Cout, int k, int stride, int act, void* workspace, size_t workspace_bytes,
                       void* stream) {
  MGB_REQUIRE(x && w && y && workspace, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(shape_ok(B, Tin, Cin, Cout, k, stride), MGB_E_ARG, "bad conv shape (odd kernel, stride 1..8)");
  MGB_REQUIRE(act >= ACT_NONE && act <= ACT_RELU, MGB_E_ARG, "unknown activation %d", act);
  MGB_REQUIRE(act != ACT_MISH || pre, MGB_E_ARG, "the Mish backward needs the pre-activation: pass `pre`");
  if (int rc = check_arch()) return rc;
  const ConvShape s = make_shape(B, Tin, Cin, Cout, k, stride);
  const Work wl = work_layout(s);
  MGB_REQUIRE(workspace_bytes >= wl.total, MGB_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, wl.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* W = static_cast<uint8_t*>(workspace);
  float* wp = reinterpret_cast<float*>(W + wl.wp);
  note_launch(launch_conv_gemm<false>(x, w, wp, bias, rowbias, y, pre, reinterpret_cast<float*>(W + wl.gpart), s, act, st));
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_conv1d_backward(const float* x, const float* w, const float* rowbias, const float* y, const float* pre,
                        const float* grad_y, float* grad_x, float* grad_w, float* grad_bias, float* grad_rowbias, int B,
                        int Tin, int Cin, int Cout, int k, int stride, int act, void* workspace, size_t workspace_bytes,
                        void* stream) {
  MGB_REQUIRE(x && w && grad_y && workspace, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(shape_ok(B, Tin, Cin, Cout, k, stride), MGB_E_ARG, "bad conv shape (odd kernel, stride 1..8)");
  MGB_REQUIRE(act >= ACT_NONE && act <= ACT_RELU, MGB_E_ARG, "unknown activation %d", act);
  MGB_REQUIRE((act != ACT_LEAKY && act != ACT_RELU) || y, MGB_E_ARG, "this activation's backward needs the output y");
  MGB_REQUIRE(act != ACT_MISH || pre, MGB_E_ARG, "the Mish backward needs the pre-activation");
  MGB_REQUIRE(!grad_rowbias || grad_x, MGB_E_ARG, "grad_rowbias is the column sum of grad_x: pass grad_x too");
  if (int rc = check_arch()) return rc;
  const ConvShape s = make_shape(B, Tin, Cin, Cout, k, stride);
  const Work wl = work_layout(s);
  MGB_REQUIRE(workspace_bytes >= wl.total, MGB_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, wl.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* W = static_cast<uint8_t*>(workspace);
  float* wq = reinterpret_cast<float*>(W + wl.wq);
  float* dzb = reinterpret_cast<float*>(W + wl.dz);
  float* part = reinterpret_cast<float*>(W + wl.part);
  const size_t ny = (size_t)B * s.Tout * Cout;
  const float* dz = grad_y;
  int launches = 0;
  if (act != ACT_NONE) {
    act_bwd_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, st>>>(grad_y, y, pre, dzb, ny, act);
    dz = dzb;
    ++launches;
  }
  const int wn = Cout * Cin * k;
  ConvSide* side = (grad_x && (grad_w || grad_bias)) ? conv_side() : nullptr;
  cudaStream_t ss = side ? side->s : st;                    // stream of the weight / bias gradient half
  if (side) {
    MGB_CUDA_CHECK(cudaEventRecord(side->fork, st));
    MGB_CUDA_CHECK(cudaStreamWaitEvent(ss, side->fork, 0));
  }
  if (grad_w) {
    const int splits = wgrad_splits(s);
    const int M = B * s.Tout;
    const int rows_per_split = ((M + splits - 1) / splits + TK - 1) / TK * TK;
    dim3 grid((k * Cin + TM - 1) / TM, (Cout + TN - 1) / TN, splits);
    conv_wgrad_f32_kernel<<<grid, NTHR, 0, ss>>>(x, rowbias, dz, part, s, rows_per_split);
    wgrad_reduce_f32_kernel<<<(wn + 255) / 256, 256, 0, ss>>>(part, grad_w, Cin, Cout, k, splits);
    launches += 2;
  }
  if (grad_bias) {
    launch_colsum(dz, grad_bias, reinterpret_cast<float*>(W + wl.cs2), 1, B * s.Tout, Cout, ss);
    launches += 2;
  }
  if (grad_x) {
    launches += launch_conv_gemm<true>(dz, w, wq, nullptr, nullptr, grad_x, nullptr, reinterpret_cast<float*>(W + wl.gpart), s, ACT_NONE, st);
    if (grad_rowbias) {
      // d rowbias[b][ci] = sum over the rows of utterance b of grad_x: exact, because every existing input row carries
      // the bias once and grad_x is the gradient with respect to (x + rowbias)
      launch_colsum(grad_x, grad_rowbias, reinterpret_cast<float*>(W + wl.cs), B, Tin, Cin, st);
      launches += 2;
    }
  }
  if (side) {
    MGB_CUDA_CHECK(cudaEventRecord(side->join, ss));
    MGB_CUDA_CHECK(cudaStreamWaitEvent(st, side->join, 0));
  }
  note_launch(launches);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_step_embedding(const int64_t* t, float* emb, int B, int dim, void* stream) {
  MGB_REQUIRE(t && emb && B > 0 && dim >= 4 && dim % 2 == 0, MGB_E_ARG, "bad argument");
  if (int rc = check_arch()) return rc;
  step_embedding_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, emb, dim);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

}  // extern "C"

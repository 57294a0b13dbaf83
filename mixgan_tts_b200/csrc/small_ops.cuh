// Small per-utterance / per-weight kernels shared by the fp32 and bf16 paths: weight transposes,
// the diffusion-step embedding + MLP (blocks.py:906-913, modules.py:433-434) and the per-layer
// projection tables (blocks.py:1159,1161-1164).
#pragma once

#include "common.cuh"

namespace mgb {
namespace smallops {

// ---- weight packing: dst[k][n'] = src[orig(n')][k] -------------------------------------------
// src is [n_src][Kin][taps] (state_dict layout); dst is [taps*Kin][ldd], k = tap*Kin + ci.
// perm: 128-wide column tiles hold 64 first-half then 64 second-half output channels.
// blockIdx.z walks `gridDim.z` matrices of the same shape (one per residual block), src_zs / dst_zs floats apart.
static __global__ void pack_wt_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_src,
                               int Kin, int taps, int ldd, int perm, int half_n, size_t src_zs, size_t dst_zs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (n >= ldd) return;
  src += blockIdx.z * src_zs;
  dst += blockIdx.z * dst_zs;
  int orig = n;
  if (perm) {
    const int tile = n >> 7, pos = n & 127;
    orig = pos < 64 ? tile * 64 + pos : half_n + tile * 64 + (pos - 64);
  }
  const int tap = k / Kin, ci = k - tap * Kin;
  float v = 0.f;
  if (orig < n_src) v = src[((size_t)orig * Kin + ci) * taps + tap];
  dst[(size_t)k * ldd + n] = v;
}

static __global__ void pack_bias_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_src,
                                 int n_dst, int perm, int half_n, size_t src_zs = 0, size_t dst_zs = 0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_dst) return;
  src += blockIdx.y * src_zs;
  dst += blockIdx.y * dst_zs;
  int orig = n;
  if (perm) {
    const int tile = n >> 7, pos = n & 127;
    orig = pos < 64 ? tile * 64 + pos : half_n + tile * 64 + (pos - 64);
  }
  dst[n] = orig < n_src ? src[orig] : 0.f;
}

// ---- small per-utterance ops --------------------------------------------------------------------
// d[b] = W2 * mish(W0 * [sin(t f), cos(t f)])   (blocks.py:906-913, modules.py:433-434)
static __global__ void __launch_bounds__(256) step_mlp_kernel(const int64_t* __restrict__ t,
                                                       const float* __restrict__ w0t,   // [C][4C]
                                                       const float* __restrict__ w2t,   // [4C][C]
                                                       float* __restrict__ d, int C) {
  extern __shared__ float sm[];
  float* emb = sm;          // [C]
  float* h = sm + C;        // [4C]
  const int b = blockIdx.x, tid = threadIdx.x;
  const int halfd = C / 2;
  const float tv = (float)t[b];
  const float scale = (float)(9.210340371976184 / (double)(halfd - 1));  // ln(10000)/(half-1)
  for (int i = tid; i < halfd; i += blockDim.x) {
    const float f = expf((float)i * -scale);
    const float a = tv * f;
    emb[i] = sinf(a);
    emb[halfd + i] = cosf(a);
  }
  __syncthreads();
  for (int j = tid; j < 4 * C; j += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < C; ++k) s = fmaf(w0t[(size_t)k * 4 * C + j], emb[k], s);
    const float sp = s > 20.f ? s : log1pf(expf(s));  // F.softplus default threshold
    h[j] = s * tanhf(sp);
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < 4 * C; ++k) s = fmaf(w2t[(size_t)k * C + c], h[k], s);
    d[(size_t)b * C + c] = s;
  }
}

// Batched version of the same MLP: one launch per layer, weights read once per 8 rows.
//   out[b][n] = f( sum_k in[b][k] * wt[k][n] ),  EMB: in[b] = sinusoidal embedding of t[b],  MISH: f = mish
// Block = 32 output columns x 8 K-slices (256 threads): the K loop is split eight ways so that the few rows of a
// sampling call (one per diffusion step) do not serialise a 1024-long chain of L2-latency weight loads.
constexpr int MLP_UB = 8, MLP_KS = 8;
template <bool EMB, bool MISH>
static __global__ void __launch_bounds__(256) step_mlp_layer_kernel(const int64_t* __restrict__ t,
                                                                    const float* __restrict__ in,
                                                                    const float* __restrict__ wt,
                                                                    float* __restrict__ out, int B, int K, int N) {
  extern __shared__ float xs[];   // [MLP_UB][K] inputs, then [MLP_KS][MLP_UB][32] partial sums
  float* part = xs + MLP_UB * K;
  const int col = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int b0 = blockIdx.y * MLP_UB, n = blockIdx.x * 32 + col;
  const int halfd = K / 2;
  const float scale = (float)(9.210340371976184 / (double)(halfd - 1));  // ln(10000)/(half-1)
  for (int i = threadIdx.x; i < MLP_UB * K; i += 256) {
    const int u = i / K, k = i - u * K, b = b0 + u;
    float v = 0.f;
    if (b < B) {
      if (EMB) {
        const int kk = k < halfd ? k : k - halfd;
        const float a = (float)t[b] * expf((float)kk * -scale);
        v = k < halfd ? sinf(a) : cosf(a);
      } else {
        v = in[(size_t)b * K + k];
      }
    }
    xs[i] = v;
  }
  __syncthreads();
  float acc[MLP_UB];
#pragma unroll
  for (int u = 0; u < MLP_UB; ++u) acc[u] = 0.f;
  const int kper = K / MLP_KS;             // K % (8 * 16) == 0 for K = 256, 1024
  const int kbeg = ks * kper;
  if (n < N) {
    for (int k0 = kbeg; k0 < kbeg + kper; k0 += 16) {     // 16 independent weight loads in flight
      float w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = __ldg(wt + (size_t)(k0 + i) * N + n);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int u = 0; u < MLP_UB; ++u) acc[u] = fmaf(w[i], xs[u * K + k0 + i], acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < MLP_UB; ++u) part[(ks * MLP_UB + u) * 32 + col] = acc[u];
  __syncthreads();
  {                                        // thread (u = ks, col) sums the 8 K-slices in a fixed order
    const int u = ks;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < MLP_KS; ++q) s += part[(q * MLP_UB + u) * 32 + col];
    if (n < N && b0 + u < B) {
      if (MISH) {
        const float sp = s > 20.f ? s : log1pf(expf(s));  // F.softplus default threshold
        s = s * tanhf(sp);
      }
      out[(size_t)(b0 + u) * N + n] = s;
    }
  }
}

// d[B][C] = W2 * mish(W0 * emb(t)); `hbuf` is [B][4C] scratch.  Two launches.
static inline void launch_step_mlp(const int64_t* t, const float* w0t, const float* w2t, float* hbuf, float* d,
                                   int B, int C, cudaStream_t s) {
  dim3 g1(4 * C / 32, (B + MLP_UB - 1) / MLP_UB), g2(C / 32, (B + MLP_UB - 1) / MLP_UB);
  const size_t sm1 = (size_t)(MLP_UB * C + MLP_KS * MLP_UB * 32) * sizeof(float);
  const size_t sm2 = (size_t)(MLP_UB * 4 * C + MLP_KS * MLP_UB * 32) * sizeof(float);
  step_mlp_layer_kernel<true, true><<<g1, 256, sm1, s>>>(t, nullptr, w0t, hbuf, B, C, 4 * C);
  step_mlp_layer_kernel<false, false><<<g2, 256, sm2, s>>>(nullptr, hbuf, w2t, d, B, 4 * C, C);
}

// tab[b][l][c] = sum_k Wt_l[k][c] * v[b][k] (+ bias_l[c]);  v == nullptr -> bias only.
constexpr int TAB_UB = 8;
static __global__ void __launch_bounds__(256) proj_table_kernel(const float* __restrict__ v, int Kin,
                                                         const float* __restrict__ wt0, size_t w_stride,
                                                         const float* __restrict__ bias0, size_t b_stride,
                                                         float* __restrict__ tab, int B, int L, int C) {
  extern __shared__ float vs[];  // [TAB_UB][Kin]
  const int l = blockIdx.x, b0 = blockIdx.y * TAB_UB, c = threadIdx.x;
  const int nb = min(TAB_UB, B - b0);
  float acc[TAB_UB];
#pragma unroll
  for (int u = 0; u < TAB_UB; ++u) acc[u] = 0.f;
  if (v) {
    for (int i = threadIdx.x; i < TAB_UB * Kin; i += blockDim.x) {
      const int u = i / Kin, k = i - u * Kin;
      vs[i] = u < nb ? v[(size_t)(b0 + u) * Kin + k] : 0.f;
    }
    __syncthreads();
    const float* wt = wt0 + (size_t)l * w_stride;
    if (c < C) {
      for (int k0 = 0; k0 < Kin; k0 += 16) {   // Kin % 16 == 0
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = __ldg(wt + (size_t)(k0 + i) * C + c);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int u = 0; u < TAB_UB; ++u) acc[u] = fmaf(w[i], vs[u * Kin + k0 + i], acc[u]);
        }
      }
    }
  }
  if (c < C) {
    const float bv = bias0 ? bias0[(size_t)l * b_stride + c] : 0.f;
    for (int u = 0; u < nb; ++u) tab[((size_t)(b0 + u) * L + l) * C + c] = acc[u] + bv;
  }
}


static inline void launch_pack(const float* src, float* dst, int n_src, int Kin, int taps, int ldd, int perm,
                 int half_n, cudaStream_t s, int nmat = 1, size_t src_zs = 0, size_t dst_zs = 0) {
  dim3 grid((ldd + 127) / 128, taps * Kin, nmat);
  pack_wt_kernel<<<grid, 128, 0, s>>>(src, dst, n_src, Kin, taps, ldd, perm, half_n, src_zs, dst_zs);
}


}  // namespace smallops
}  // namespace mgb

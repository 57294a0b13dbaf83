// Small kernels shared by the fp32 and bf16 training paths (train_fp32.cu, train_bf16.cu): the fixed-order reduction of
// split weight-gradient partials, the per-layer bias / projection terms, and the step-MLP backward.  All fp32.
#pragma once

#include "common.cuh"

namespace mgb {
namespace trainsmall {

// dst[(m*Kin + ci)*taps + tap] = sum_s part[s][m][tap*Kin + ci]   (state_dict layout [out][in][k])
static __global__ void wgrad_reduce_kernel(const float* __restrict__ part, int S, int Mo, int N, int Kin, int taps,
                                    float* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)Mo * N;
  if (i >= tot) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += part[(size_t)k * tot + i];
  const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
  const int tap = n / Kin, ci = n - tap * Kin;
  dst[((size_t)m * Kin + ci) * taps + tap] = s;
}

// Same reduction for partials stored in the chunked layout part[s][n/4][m][4] (the tcgen05 weight-gradient kernel owns
// one output row m per thread, so a warp's 16-byte store to chunk n/4 is 512 contiguous bytes).
static __global__ void wgrad_reduce_chunked_kernel(const float* __restrict__ part, int S, int Mo, int N, int Kin, int taps,
                                                   float* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)Mo * N;
  if (j >= tot) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += part[(size_t)k * tot + j];
  const int nq = (int)(j / ((size_t)Mo * 4));
  const int rem = (int)(j - (size_t)nq * Mo * 4);
  const int m = rem >> 2, n = nq * 4 + (rem & 3);
  const int tap = n / Kin, ci = n - tap * Kin;
  dst[((size_t)m * Kin + ci) * taps + tap] = s;
}

// ------------------------------------------------------------------------------------------------
// per-layer small terms (B rows each): bias gradients, diffusion/speaker projection gradients, and the
// accumulation of d loss / d dvec and d loss / d spk over layers.
struct LayerSmallArgs {
  const float* usumE;   // [B][C]   per-utterance column sums of e_l (nullptr for the top block: zero)
  const float* usumZ;   // [B][2C]  of dZ
  const float* usumY;   // [B][C]   of dY
  const float* usumS;   // [B][C]   of dS
  const float* dvec;    // [B][C]
  const float* spk;     // [B][H] or nullptr
  const float* Wd;      // [C][C]   diffusion_projection.linear.weight  (raw [out][in])
  const float* Ws;      // [C][H]   speaker_projection.linear.weight or nullptr
  float *g_conv_b, *g_oproj_b, *g_cproj_b, *g_dproj_w, *g_sproj_w;
  float* dd_l;          // [B][C]  this layer's slot of dd_all [L][B][C]
  float* ds_l;          // [B][C]  this layer's slot of ds_all [L][B][C] (or nullptr)
  int B, C, H;
};
// grid: C blocks (row co of dWd / dWs) + B blocks (ddvec / dspk rows) + 1 block (biases); 256 threads, C == H == 256
static __global__ void __launch_bounds__(256) layer_small_kernel(const LayerSmallArgs p) {
  pdl_trigger();
  pdl_wait();
  const int C = p.C, tid = threadIdx.x;
  const int blk = blockIdx.x;
  if (blk < C) {
    // dWd[co][ci] = sum_b dd[b][co] * dvec[b][ci],  dd = usumE + usumY ; dWs[co][h] = sum_b usumY[b][co] * spk[b][h]
    const int co = blk;
    float gd = 0.f, gs = 0.f;
    for (int b = 0; b < p.B; ++b) {
      const float y = p.usumY[(size_t)b * C + co];
      const float e = p.usumE ? p.usumE[(size_t)b * C + co] : 0.f;
      gd = fmaf(e + y, p.dvec[(size_t)b * C + tid], gd);
      if (p.Ws) gs = fmaf(y, p.spk[(size_t)b * p.H + tid], gs);
    }
    p.g_dproj_w[(size_t)co * C + tid] = gd;
    if (p.Ws) p.g_sproj_w[(size_t)co * p.H + tid] = gs;
  } else if (blk < C + p.B) {
    // keep dd_l = usumE + usumY and ds_l = usumY per layer: d loss / d dvec and d loss / d spk are contracted with all
    // layers' projection weights in one launch at the head (layers_rowgemm_kernel) instead of a serial chain per layer
    const int b = blk - C;
    const float y = p.usumY[(size_t)b * C + tid];
    p.dd_l[(size_t)b * C + tid] = (p.usumE ? p.usumE[(size_t)b * C + tid] : 0.f) + y;
    if (p.Ws) p.ds_l[(size_t)b * C + tid] = y;
  } else {
    // conv bias [2C], output-projection bias [x-half from e_l | skip half from dS], conditioner bias [C]
    for (int c = tid; c < 2 * C; c += blockDim.x) {
      float z = 0.f;
      for (int b = 0; b < p.B; ++b) z += p.usumZ[(size_t)b * 2 * C + c];
      p.g_conv_b[c] = z;
    }
    for (int c = tid; c < C; c += blockDim.x) {
      float e = 0.f, sk = 0.f, y = 0.f;
      for (int b = 0; b < p.B; ++b) {
        if (p.usumE) e += p.usumE[(size_t)b * C + c];
        sk += p.usumS[(size_t)b * C + c];
        y += p.usumY[(size_t)b * C + c];
      }
      p.g_oproj_b[c] = e;
      p.g_oproj_b[C + c] = sk;
      p.g_cproj_b[c] = y;
    }
  }
}

// part[l][b][n] = sum_k in[l][b][k] * W_l[k][n]  (K = N = 256; W_l = raw [out=k][in=n] projection weight of layer l).
// grid (N/32, ceil(B/8), L), 256 threads = 32 columns x 8 K-slices; 16 weight loads in flight per thread.
static __global__ void __launch_bounds__(256) layers_rowgemm_kernel(const float* __restrict__ in, const float* __restrict__ w0,
                                                                    size_t w_lstride, float* __restrict__ part, int B) {
  constexpr int K = 256, N = 256, UB = 8, KS = 8;
  __shared__ float xs[UB * K];
  __shared__ float ps[KS * UB * 32];
  const int l = blockIdx.z, b0 = blockIdx.y * UB, col = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + col;
  const float* inl = in + (size_t)l * B * K;
  for (int i = threadIdx.x; i < UB * K; i += 256) {
    const int u = i / K, k = i - u * K;
    xs[i] = b0 + u < B ? inl[(size_t)(b0 + u) * K + k] : 0.f;
  }
  __syncthreads();
  const float* w = w0 + (size_t)l * w_lstride;
  float acc[UB];
#pragma unroll
  for (int u = 0; u < UB; ++u) acc[u] = 0.f;
  const int kbeg = ks * (K / KS);
  for (int k0 = kbeg; k0 < kbeg + K / KS; k0 += 16) {
    float wv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) wv[i] = __ldg(w + (size_t)(k0 + i) * N + n);
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
      for (int u = 0; u < UB; ++u) acc[u] = fmaf(wv[i], xs[u * K + k0 + i], acc[u]);
  }
#pragma unroll
  for (int u = 0; u < UB; ++u) ps[(ks * UB + u) * 32 + col] = acc[u];
  __syncthreads();
  const int u = ks;
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < KS; ++q) sum += ps[(q * UB + u) * 32 + col];
  if (b0 + u < B) part[((size_t)l * B + b0 + u) * N + n] = sum;
}
// out[i] = sum_l part[l][i]   (fixed order)
static __global__ void sum_layers_kernel(const float* __restrict__ part, int L, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int l = 0; l < L; ++l) s += part[(size_t)l * n + i];
  out[i] = s;
}
// ddvec[B][C] = sum_l dd_all[l] Wd_l ; dspk[B][H] = sum_l ds_all[l] Ws_l (when spk_w0 != nullptr)
static inline void launch_dvec_contraction(const float* dd_all, const float* ds_all, const float* dproj_w0, const float* sproj_w0,
                                           size_t w_lstride, float* part, float* ddvec, float* dspk, int B, int L,
                                           cudaStream_t s) {
  const dim3 grid(256 / 32, (B + 7) / 8, L);
  layers_rowgemm_kernel<<<grid, 256, 0, s>>>(dd_all, dproj_w0, w_lstride, part, B);
  sum_layers_kernel<<<(B * 256 + 255) / 256, 256, 0, s>>>(part, L, B * 256, ddvec);
  note_launch(2);
  if (sproj_w0) {
    layers_rowgemm_kernel<<<grid, 256, 0, s>>>(ds_all, sproj_w0, w_lstride, part, B);
    sum_layers_kernel<<<(B * 256 + 255) / 256, 256, 0, s>>>(part, L, B * 256, dspk);
    note_launch(2);
  }
}

// out[c] = sum_b usum[b][c]
static __global__ void bias_from_usum_kernel(const float* __restrict__ usum, int B, int n, int ld, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += usum[(size_t)b * ld + c];
  out[c] = s;
}

// ---- step MLP backward (modules.py:433-434, blocks.py:894-913); all operands are B rows ----------
// phase 0: dW2[c][j] = sum_b ddvec[b][c] * h[b][j]                    grid C blocks x 256 threads (j strided)
// phase 1: dpre[b][j] = (sum_c ddvec[b][c] * W2[c][j]) * mish'(pre[b][j]),  pre = W0 emb(t_b)      grid (4C/256, B)
// phase 2: dW0[j][k] = sum_b dpre[b][j] * emb[b][k]                   grid 4C blocks x C threads
static __device__ __forceinline__ float emb_value(float tv, int k, int C) {
  const int halfd = C / 2;
  const float scale = (float)(9.210340371976184 / (double)(halfd - 1));
  const int kk = k < halfd ? k : k - halfd;
  const float a = tv * expf((float)kk * -scale);
  return k < halfd ? sinf(a) : cosf(a);
}
static __global__ void __launch_bounds__(256) mlp_bwd_w2_kernel(const float* __restrict__ ddvec, const float* __restrict__ h,
                                                         float* __restrict__ gW2, int B, int C) {
  const int c = blockIdx.x;
  for (int j = threadIdx.x; j < 4 * C; j += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(ddvec[(size_t)b * C + c], h[(size_t)b * 4 * C + j], s);
    gW2[(size_t)c * 4 * C + j] = s;
  }
}
static __global__ void __launch_bounds__(256) mlp_bwd_pre_kernel(const int64_t* __restrict__ t, const float* __restrict__ ddvec,
                                                          const float* __restrict__ W0, const float* __restrict__ W2,
                                                          float* __restrict__ dpre, int C) {
  // grid (4C / 256, B).  pre[j] = W0[j][:] . emb: one WARP per output j (lanes stride the contiguous row: coalesced; a
  // thread-per-row walk over the [4C][C] matrix was 90 us of serialised L2 latency), results parked in shared memory;
  // then dh[j] = sum_c ddvec[c] W2[c][j] with one thread per j (coalesced over j).
  __shared__ float emb[256];
  __shared__ float pre_s[256];
  const int b = blockIdx.y, j0 = blockIdx.x * 256, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float tv = (float)t[b];
  for (int k = threadIdx.x; k < C; k += blockDim.x) emb[k] = emb_value(tv, k, C);
  __syncthreads();
  for (int jj = warp; jj < 256; jj += 8) {
    const float* row = W0 + (size_t)(j0 + jj) * C;
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) acc = fmaf(row[k], emb[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) pre_s[jj] = acc;
  }
  __syncthreads();
  const int j = j0 + threadIdx.x;
  const float pre = pre_s[threadIdx.x];
  float dh = 0.f;
#pragma unroll 8
  for (int c = 0; c < C; ++c) dh = fmaf(ddvec[(size_t)b * C + c], W2[(size_t)c * 4 * C + j], dh);
  // mish(x) = x tanh(softplus(x));  mish'(x) = tanh(sp) + x (1 - tanh(sp)^2) sigmoid(x)
  const float sp = pre > 20.f ? pre : log1pf(expf(pre));
  const float th = tanhf(sp);
  const float sg = 1.0f / (1.0f + expf(-pre));
  dpre[(size_t)b * 4 * C + j] = dh * (th + pre * (1.0f - th * th) * sg);
}
static __global__ void __launch_bounds__(256) mlp_bwd_w0_kernel(const int64_t* __restrict__ t, const float* __restrict__ dpre,
                                                         float* __restrict__ gW0, int B, int C) {
  const int j = blockIdx.x, k = threadIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s = fmaf(dpre[(size_t)b * 4 * C + j], emb_value((float)t[b], k, C), s);
  gW0[(size_t)j * C + k] = s;
}


}  // namespace trainsmall
}  // namespace mgb

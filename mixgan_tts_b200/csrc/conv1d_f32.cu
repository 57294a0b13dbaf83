// Generic fp32 Conv1d on frames-major activations, forward and backward: the building block of the JCU discriminator
// (reference: model/mixgantts.py:186-288 — ConvNorm model/blocks.py:326-371 with stride 1 or 2, LinearNorm :278-291,
// F.leaky_relu(., 0.2), Mish :894-896, DiffusionEmbedding :899-913) as train.py:126-184 drives it: 4 forwards and
// 2 backwards per step.  The discriminator is 0.65 MFLOP per mel frame (2.7 % of a Denoiser call), and its gradient has
// to match the reference's fp32 autograd, so these kernels are exact-fp32 CUDA-core implicit GEMMs, not tensor-core tiles.
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
#include "common.cuh"

namespace mgb {
namespace {

constexpr int TM = 64, TN = 64, TK = 16, NTHR = 256;

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

struct ConvShape { int B, Tin, Tout, Cin, Cout, k, stride, pad; };

// ---- weights: torch [Cout][Cin][k] -> wp [k][Cin][Cout] (forward / wgrad B operand) and wq [k][Cout][Cin] (dgrad) ----
__global__ void pack_w_kernel(const float* __restrict__ w, float* __restrict__ wp, float* __restrict__ wq, int Cin, int Cout, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * k) return;
  const int j = i % k, ci = (i / k) % Cin, co = i / (k * Cin);
  const float v = w[i];
  wp[((size_t)j * Cin + ci) * Cout + co] = v;
  if (wq) wq[((size_t)j * Cout + co) * Cin + ci] = v;
}

// ---- forward (DGRAD = false) and data gradient (DGRAD = true): out[M][N] = A[M][K] * Bm[K][N] ----------------------
//   forward: M = B*Tout rows, K = k*Cin, N = Cout, A(m, j*Cin + ci) = xin[b][to*s + j - pad][ci], Bm = wp
//   dgrad  : M = B*Tin  rows, K = k*Cout, N = Cin, A(m, j*Cout + co) = dz[b][(ti + pad - j)/s][co], Bm = wq
// The operand tiles of k-step i + 1 are fetched into registers while k-step i is multiplied out of shared memory: with
// 50-200 CTAs on 148 SMs these GEMMs are latency-bound, and an un-prefetched loop paid one full global-memory latency
// (~1 us) per 16-deep k-step (measured 227 us for the 512 -> 128, k = 5 layer; same accumulation order, same bits).
template <bool DGRAD>
__global__ void __launch_bounds__(NTHR) conv_gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             const float* __restrict__ bias, const float* __restrict__ rowbias,
                                                             float* __restrict__ out, float* __restrict__ pre, ConvShape s, int act) {
  __shared__ float As[TK][TM + 4], Bs[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int rows_per_b = DGRAD ? s.Tin : s.Tout;       // rows of `out` per utterance
  const int src_per_b = DGRAD ? s.Tout : s.Tin;        // rows of `A` per utterance
  const int Ca = DGRAD ? s.Cout : s.Cin;               // channels of A
  const int N = DGRAD ? s.Cin : s.Cout;
  const int M = s.B * rows_per_b, K = s.k * Ca;
  // A-tile loader: thread -> (row a_r, 4 consecutive k)
  const int a_r = tid >> 2, a_k4 = (tid & 3) * 4;
  const int a_m = m0 + a_r;
  const int a_b = a_m < M ? a_m / rows_per_b : 0, a_t = a_m - a_b * rows_per_b;
  // B-tile loader: thread -> (k row b_k, 4 consecutive n)
  const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;
  const bool vecA = (Ca & 3) == 0, vecB = (N & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  auto src_row = [&](int j, int& src) -> bool {
    if (DGRAD) { const int q = a_t + s.pad - j; src = q / s.stride; return q >= 0 && q % s.stride == 0 && q / s.stride < s.Tout; }
    src = a_t * s.stride + j - s.pad;
    return src >= 0 && src < s.Tin;
  };
  auto load_a = [&](int k0, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    const int kk = k0 + a_k4;
    if (a_m >= M || kk >= K) return;
    if (vecA) {
      const int j = kk / Ca, c = kk - j * Ca;
      int src;
      if (src_row(j, src)) {
        const float4 x4 = *reinterpret_cast<const float4*>(A + ((size_t)a_b * src_per_b + src) * Ca + c);
        v[0] = x4.x; v[1] = x4.y; v[2] = x4.z; v[3] = x4.w;
        if (!DGRAD && rowbias) {
          const float4 r4 = *reinterpret_cast<const float4*>(rowbias + (size_t)a_b * Ca + c);
          v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ke = kk + e;
        if (ke >= K) break;
        const int j = ke / Ca, c = ke - j * Ca;
        int src;
        if (src_row(j, src)) v[e] = A[((size_t)a_b * src_per_b + src) * Ca + c] + ((!DGRAD && rowbias) ? rowbias[(size_t)a_b * Ca + c] : 0.f);
      }
    }
  };
  auto load_b = [&](int k0, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    const int kk = k0 + b_k, n = n0 + b_n4;
    if (kk >= K) return;
    if (vecB && n + 3 < N) {
      const float4 w4 = *reinterpret_cast<const float4*>(Bm + (size_t)kk * N + n);
      v[0] = w4.x; v[1] = w4.y; v[2] = w4.z; v[3] = w4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (n + e < N) v[e] = Bm[(size_t)kk * N + n + e];
    }
  };

  // PF k-steps of operand tiles are in flight in registers (static indices: the k loop is unrolled by PF)
  constexpr int PF = 4;
  float pa[PF][4], pb[PF][4];
#pragma unroll
  for (int u = 0; u < PF; ++u) { load_a(u * TK, pa[u]); load_b(u * TK, pb[u]); }
  for (int kb = 0; kb < K; kb += PF * TK) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int k0 = kb + u * TK;
      if (k0 >= K) break;                                             // uniform over the block
#pragma unroll
      for (int e = 0; e < 4; ++e) As[a_k4 + e][a_r] = pa[u][e];
      *reinterpret_cast<float4*>(&Bs[b_k][b_n4]) = make_float4(pb[u][0], pb[u][1], pb[u][2], pb[u][3]);
      __syncthreads();
      load_a(k0 + PF * TK, pa[u]);                                    // zero beyond K; in flight during the next PF multiplies
      load_b(k0 + PF * TK, pb[u]);
#pragma unroll
      for (int kq = 0; kq < TK; ++kq) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kq][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kq][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (!DGRAD) {
        if (bias) v += bias[n];
        if (pre) pre[(size_t)m * N + n] = v;
        v = act_fwd(v, act);
      }
      out[(size_t)m * N + n] = v;
    }
  }
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
__global__ void __launch_bounds__(NTHR) conv_wgrad_f32_kernel(const float* __restrict__ x, const float* __restrict__ rowbias,
                                                              const float* __restrict__ dz, float* __restrict__ part, ConvShape s,
                                                              int rows_per_split) {
  __shared__ float As[TK][TM + 4], Bs[TK][TN + 4];     // As[m][kk], Bs[m][co]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * TM, n0 = blockIdx.y * TN, split = blockIdx.z;
  const int K = s.k * s.Cin, M = s.B * s.Tout;
  const int mb = split * rows_per_split, me = min(M, mb + rows_per_split);
  const int l_m = tid >> 4, l_c4 = (tid & 15) * 4;    // loader: (row within the 16-row chunk, 4 consecutive columns)
  const bool vecA = (s.Cin & 3) == 0, vecB = (s.Cout & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  auto load_ab = [&](int m0, float (&va)[4], float (&vb)[4]) {
    va[0] = va[1] = va[2] = va[3] = 0.f;
    vb[0] = vb[1] = vb[2] = vb[3] = 0.f;
    const int m = m0 + l_m;
    if (m >= me) return;
    const int b = m / s.Tout, to = m - b * s.Tout;
    const int kk = k0 + l_c4;
    if (vecA) {                  // K = k * Cin is a multiple of 4 and a float4 never straddles two taps
      if (kk < K) {
        const int j = kk / s.Cin, c = kk - j * s.Cin;
        const int ti = to * s.stride + j - s.pad;
        if (ti >= 0 && ti < s.Tin) {
          const float4 x4 = *reinterpret_cast<const float4*>(x + ((size_t)b * s.Tin + ti) * s.Cin + c);
          va[0] = x4.x; va[1] = x4.y; va[2] = x4.z; va[3] = x4.w;
          if (rowbias) {
            const float4 r4 = *reinterpret_cast<const float4*>(rowbias + (size_t)b * s.Cin + c);
            va[0] += r4.x; va[1] += r4.y; va[2] += r4.z; va[3] += r4.w;
          }
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ke = kk + e;
        if (ke < K) {
          const int j = ke / s.Cin, c = ke - j * s.Cin;
          const int ti = to * s.stride + j - s.pad;
          if (ti >= 0 && ti < s.Tin)
            va[e] = x[((size_t)b * s.Tin + ti) * s.Cin + c] + (rowbias ? rowbias[(size_t)b * s.Cin + c] : 0.f);
        }
      }
    }
    const int n = n0 + l_c4;
    if (vecB && n + 3 < s.Cout) {
      const float4 d4 = *reinterpret_cast<const float4*>(dz + (size_t)m * s.Cout + n);
      vb[0] = d4.x; vb[1] = d4.y; vb[2] = d4.z; vb[3] = d4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (n + e < s.Cout) vb[e] = dz[(size_t)m * s.Cout + n + e];
    }
  };
  constexpr int PF = 4;                                  // row chunks in flight in registers
  float va[PF][4], vb[PF][4];
#pragma unroll
  for (int u = 0; u < PF; ++u) load_ab(mb + u * TK, va[u], vb[u]);
  for (int mq = mb; mq < me; mq += PF * TK) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int m0 = mq + u * TK;
      if (m0 >= me) break;                               // uniform over the block
      *reinterpret_cast<float4*>(&As[l_m][l_c4]) = make_float4(va[u][0], va[u][1], va[u][2], va[u][3]);
      *reinterpret_cast<float4*>(&Bs[l_m][l_c4]) = make_float4(vb[u][0], vb[u][1], vb[u][2], vb[u][3]);
      __syncthreads();
      load_ab(m0 + PF * TK, va[u], vb[u]);               // zero beyond the split
#pragma unroll
      for (int q = 0; q < TK; ++q) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kk = k0 + ty * 4 + i;
    if (kk >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < s.Cout) part[((size_t)split * K + kk) * s.Cout + n] = acc[i][j];
    }
  }
}

// dw[co][ci][j] (torch layout) = sum over splits (fixed order) of part[split][j*Cin + ci][co]
__global__ void wgrad_reduce_f32_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cin, int Cout, int k, int nsplit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Cin * k) return;
  const int j = i % k, ci = (i / k) % Cin, co = i / (k * Cin);
  const size_t K = (size_t)k * Cin, kk = (size_t)j * Cin + ci;
  float a = 0.f;
  for (int s = 0; s < nsplit; ++s) a += part[((size_t)s * K + kk) * Cout + co];
  dw[i] = a;
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
  int want = (296 + tiles - 1) / tiles;                  // about two CTAs per SM in total
  const int max_by_rows = (M + 127) / 128;               // at least 128 rows per split
  if (want > max_by_rows) want = max_by_rows;
  return want < 1 ? 1 : (want > 64 ? 64 : want);
}
struct Work { size_t wp, wq, dz, part, cs, total; };
Work work_layout(const ConvShape& s) {
  Work w{};
  size_t p = 0;
  auto take = [&](size_t n) { size_t r = p; p += align_up(n * sizeof(float), 256); return r; };
  const size_t wn = (size_t)s.k * s.Cin * s.Cout;
  w.wp = take(wn); w.wq = take(wn);
  w.dz = take((size_t)s.B * s.Tout * s.Cout);
  w.part = take((size_t)wgrad_splits(s) * wn);
  w.cs = take((size_t)CS_MAX_SPLITS * (size_t)((size_t)s.B * s.Cin > (size_t)s.Cout ? (size_t)s.B * s.Cin : (size_t)s.Cout));
  w.total = p;
  return w;
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
  const int wn = Cout * Cin * k;
  pack_w_kernel<<<(wn + 255) / 256, 256, 0, st>>>(w, wp, nullptr, Cin, Cout, k);
  dim3 grid((B * s.Tout + TM - 1) / TM, (Cout + TN - 1) / TN);
  conv_gemm_f32_kernel<false><<<grid, NTHR, 0, st>>>(x, wp, bias, rowbias, y, pre, s, act);
  note_launch(2);
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
  float* wp = reinterpret_cast<float*>(W + wl.wp);
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
  if (grad_x) {
    pack_w_kernel<<<(wn + 255) / 256, 256, 0, st>>>(w, wp, wq, Cin, Cout, k);
    dim3 grid((B * Tin + TM - 1) / TM, (Cin + TN - 1) / TN);
    conv_gemm_f32_kernel<true><<<grid, NTHR, 0, st>>>(dz, wq, nullptr, nullptr, grad_x, nullptr, s, ACT_NONE);
    launches += 2;
    if (grad_rowbias) {
      // d rowbias[b][ci] = sum over the rows of utterance b of grad_x: exact, because every existing input row carries
      // the bias once and grad_x is the gradient with respect to (x + rowbias)
      launch_colsum(grad_x, grad_rowbias, reinterpret_cast<float*>(W + wl.cs), B, Tin, Cin, st);
      launches += 2;
    }
  }
  if (grad_w) {
    const int splits = wgrad_splits(s);
    const int M = B * s.Tout;
    const int rows_per_split = ((M + splits - 1) / splits + TK - 1) / TK * TK;
    dim3 grid((k * Cin + TM - 1) / TM, (Cout + TN - 1) / TN, splits);
    conv_wgrad_f32_kernel<<<grid, NTHR, 0, st>>>(x, rowbias, dz, part, s, rows_per_split);
    wgrad_reduce_f32_kernel<<<(wn + 255) / 256, 256, 0, st>>>(part, grad_w, Cin, Cout, k, splits);
    launches += 2;
  }
  if (grad_bias) {
    launch_colsum(dz, grad_bias, reinterpret_cast<float*>(W + wl.cs), 1, B * s.Tout, Cout, st);
    launches += 2;
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

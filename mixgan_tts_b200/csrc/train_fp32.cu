// fp32 backward of the Denoiser (training config, BASELINE configs[4]; reference: autograd through
// Denoiser.forward model/modules.py:420-446 and ResidualBlock.forward model/blocks.py:1157-1176).
//
// The forward (fp32_path.cu with a stash) keeps, per frame, the conv input y_l, the conv pre-activation z_l and
// the gate output g_l of every block.  The backward walks the blocks in reverse.  With
//     e_l   = d loss / d (x-half of the block's output projection) = d loss / d x_{l+1} / sqrt(2)
//     dS    = d loss / d skip_l  (the same matrix for every block: the skip sum is a plain sum)
// one block is
//     dG  = [e_l | dS] Wo_l                 data-grad GEMM, epilogue = gate backward -> dZ
//     dWo = [e_l | dS]^T g_l                weight-grad GEMM (reduction over frames)
//     dY  = conv3^T(dZ)                     data-grad GEMM (three shifted taps), epilogue: dx_l = e_l + dY,
//                                           e_{l-1} = dx_l / sqrt(2)   (layer 0: ReLU mask of the input projection)
//     dW3 = dZ^T shift(y_l), dWc = dY^T cond, dCond += dY Wc_l
//     bias / per-utterance terms from per-utterance column sums of e_l, dZ, dY.
// Weight-grad GEMMs split the frame axis over CTAs and write partial sums that a second kernel adds in a fixed
// order (deterministic, no atomics) straight into the flat gradient in state_dict layout.
//
// Segments (tail, layers L-1..0, head) can be run separately so that the host can start the NCCL all-reduce of
// a finished gradient bucket while the next segment computes.
#include "common.cuh"
#include "gemm_fp32.cuh"
#include "train_small.cuh"

namespace mgb {

namespace {

using namespace trainsmall;
using gemm32::BK; using gemm32::BM; using gemm32::BMP; using gemm32::BN; using gemm32::NT;

// ------------------------------------------------------------------------------------------------
// data-grad GEMMs (frames on M) with backward epilogues
enum BEpi { B_GATE = 0, B_DX = 1, B_ACC = 2, B_RELUMASK = 3, B_SCALE = 4, B_DXT = 5 };

struct BArgs {
  gemm32::FrameGemm g;
  float* out;          // GATE: dZ [rows][2C]; DX: e (in/out) [rows][C]; ACC/RELUMASK/SCALE: [rows][ldo]; DXT: [B][n_mel][T]
  float* out2;         // DX: dY [rows][C]
  const float* aux;    // GATE: z [rows][2C]; DX: X0 [rows][C] (layer 0, ReLU mask) ; RELUMASK: activation [rows][ldo]
  int ldo, C;
  int first;           // DX: e is implicitly zero on input (top block); ACC: overwrite instead of accumulate
  int relu_mask;       // DX: layer 0 -> write (x0 > 0 ? dx : 0) instead of dx / sqrt(2)
  float scale;         // SCALE
  int n_mel;           // DXT
};

template <int EPI>
__global__ void __launch_bounds__(NT) bwd_frame_gemm_kernel(const BArgs p) {
  __shared__ __align__(16) float As[2][BK][BMP];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[8][8];
  gemm32::frame_gemm_mainloop(p.g, m0, n0, acc, As, Bs);

  const int C = p.C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + gemm32::acc_row(ty, i);
    if (m >= p.g.rows) continue;
    if constexpr (EPI == B_DXT) {
      const int b = m / p.g.T, t = m - b * p.g.T;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = n0 + gemm32::acc_col(tx, j);
        if (n < p.n_mel) p.out[((size_t)b * p.n_mel + n) * p.g.T + t] = acc[i][j];
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + h * 64 + tx * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][h * 4 + j];
        if constexpr (EPI == B_GATE) {
          // g = sigmoid(za) * tanh(zb)  (blocks.py:1170-1171)
          const float4 za = *reinterpret_cast<const float4*>(p.aux + (size_t)m * 2 * C + n);
          const float4 zb = *reinterpret_cast<const float4*>(p.aux + (size_t)m * 2 * C + C + n);
          const float a[4] = {za.x, za.y, za.z, za.w}, f[4] = {zb.x, zb.y, zb.z, zb.w};
          float da[4], df[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float sg = 1.0f / (1.0f + expf(-a[j]));
            const float th = tanhf(f[j]);
            da[j] = v[j] * th * (sg * (1.0f - sg));
            df[j] = v[j] * sg * (1.0f - th * th);
          }
          *reinterpret_cast<float4*>(p.out + (size_t)m * 2 * C + n) = make_float4(da[0], da[1], da[2], da[3]);
          *reinterpret_cast<float4*>(p.out + (size_t)m * 2 * C + C + n) = make_float4(df[0], df[1], df[2], df[3]);
        } else if constexpr (EPI == B_DX) {
          float4* ep = reinterpret_cast<float4*>(p.out + (size_t)m * C + n);
          *reinterpret_cast<float4*>(p.out2 + (size_t)m * C + n) = make_float4(v[0], v[1], v[2], v[3]);
          float4 e = p.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *ep;
          float dx[4] = {e.x + v[0], e.y + v[1], e.z + v[2], e.w + v[3]};
          if (p.relu_mask) {
            const float4 x0 = *reinterpret_cast<const float4*>(p.aux + (size_t)m * C + n);
            const float xm[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) dx[j] = xm[j] > 0.f ? dx[j] : 0.f;
          } else {
            const float SQRT2 = 1.41421356237309504880f;
#pragma unroll
            for (int j = 0; j < 4; ++j) dx[j] = __fdiv_rn(dx[j], SQRT2);
          }
          *ep = make_float4(dx[0], dx[1], dx[2], dx[3]);
        } else {
          float4* op = reinterpret_cast<float4*>(p.out + (size_t)m * p.ldo + n);
          if constexpr (EPI == B_ACC) {
            if (!p.first) {
              const float4 o = *op;
              v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
            }
          } else if constexpr (EPI == B_RELUMASK) {
            const float4 a = *reinterpret_cast<const float4*>(p.aux + (size_t)m * p.ldo + n);
            v[0] = a.x > 0.f ? v[0] : 0.f; v[1] = a.y > 0.f ? v[1] : 0.f;
            v[2] = a.z > 0.f ? v[2] : 0.f; v[3] = a.w > 0.f ? v[3] : 0.f;
          } else {  // B_SCALE
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] *= p.scale;
          }
          *op = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
}

template <int EPI>
void launch_bgemm(const BArgs& a, int N, cudaStream_t s) {
  dim3 grid((a.g.rows + BM - 1) / BM, (N + BN - 1) / BN);
  bwd_frame_gemm_kernel<EPI><<<grid, NT, 0, s>>>(a);
  note_launch();
}

// ------------------------------------------------------------------------------------------------
// weight-grad GEMM: part[s][m][n] = sum over the split's frames f of P[f][m] * Q[f + sh(n)][ci(n)],
// n = tap*Kin + ci, sh = tap - taps/2, rows outside the utterance contribute zero.
struct WArgs {
  const float* P;    // [rows][ldp]   columns m < msplit
  const float* P2;   // [rows][ldp2]  columns m >= msplit (or nullptr)
  const float* Q;    // [rows][ldq]
  float* part;       // [S][Mo][N]
  int ldp, ldp2, msplit, Mo, ldq, Kin, taps, N, rows, T, chunk;
};

__global__ void __launch_bounds__(NT) wgrad_kernel(const WArgs p) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, sp = blockIdx.z;
  const int f_begin = sp * p.chunk;
  const int f_end = min(p.rows, f_begin + p.chunk);

  // loader roles: 8 frames x 128 columns per operand and k-step = one float4 per thread
  const int l_k = tid >> 5, l_c = (tid & 31) * 4;
  // P operand column -> source matrix
  const int pm = m0 + l_c;
  const float* psrc = nullptr;
  int pld = 0, pcol = 0;
  if (pm < p.Mo) {
    if (p.P2 && pm >= p.msplit) { psrc = p.P2; pld = p.ldp2; pcol = pm - p.msplit; }
    else { psrc = p.P; pld = p.ldp; pcol = pm; }
  }
  // Q operand column -> (tap, ci)
  const int qn = n0 + l_c;
  const int tap = qn / p.Kin, ci = qn - tap * p.Kin;
  const int sh = tap - (p.taps >> 1);
  const bool q_ok = qn < p.N;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  auto gload = [&](int f0) {
    const int f = f0 + l_k;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (f < f_end) {
      if (psrc) ra = *reinterpret_cast<const float4*>(psrc + (size_t)f * pld + pcol);
      const int t = f % p.T + sh;
      if (q_ok && t >= 0 && t < p.T) rb = *reinterpret_cast<const float4*>(p.Q + (size_t)(f + sh) * p.ldq + ci);
    }
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][l_k][l_c]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][l_k][l_c]) = rb;
  };

  const int nk = (f_end - f_begin + BK - 1) / BK;
  if (nk > 0) {
    gload(f_begin);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) gload(f_begin + (kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(cur ^ 1);
      __syncthreads();
    }
  }
  float* part = p.part + (size_t)sp * p.Mo * p.N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + gemm32::acc_row(ty, i);
    if (m >= p.Mo) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n < p.N)   // N % 4 == 0
        *reinterpret_cast<float4*>(part + (size_t)m * p.N + n) =
            make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
    }
  }
}

struct WgradPlan { int chunk, S; };
WgradPlan plan_wgrad(int rows, int Mo, int N) {
  const int tiles = ((Mo + BM - 1) / BM) * ((N + BN - 1) / BN);
  int S = (148 * 4 + tiles - 1) / tiles;          // about four CTAs per SM in total
  const int maxS = (rows + 63) / 64;
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  int chunk = (rows + S - 1) / S;
  chunk = (chunk + BK - 1) / BK * BK;
  S = (rows + chunk - 1) / chunk;
  return {chunk, S};
}
size_t wgrad_part_floats(int rows, int Mo, int N) {
  return (size_t)plan_wgrad(rows, Mo, N).S * Mo * N;
}

void launch_wgrad(const float* P, int ldp, const float* P2, int ldp2, int msplit, int Mo, const float* Q, int ldq,
                  int Kin, int taps, int rows, int T, float* part, float* dst, cudaStream_t s) {
  WArgs a{};
  a.P = P; a.ldp = ldp; a.P2 = P2; a.ldp2 = ldp2; a.msplit = msplit; a.Mo = Mo; a.Q = Q; a.ldq = ldq;
  a.Kin = Kin; a.taps = taps; a.N = Kin * taps; a.rows = rows; a.T = T; a.part = part;
  const WgradPlan pl = plan_wgrad(rows, Mo, a.N);
  a.chunk = pl.chunk;
  dim3 grid((Mo + BM - 1) / BM, (a.N + BN - 1) / BN, pl.S);
  wgrad_kernel<<<grid, NT, 0, s>>>(a);
  const size_t tot = (size_t)Mo * a.N;
  wgrad_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(part, pl.S, Mo, a.N, Kin, taps, dst);
  note_launch(2);
}

// ------------------------------------------------------------------------------------------------
// per-utterance column sums: out[b][c] = sum_t A[b*T + t][c]; up to three matrices in one launch
struct ColsumArgs {
  const float* A[3];
  float* out[3];
  int lda[3], ncols[3], ldo[3];
  int T;
};
__global__ void __launch_bounds__(256) utt_colsum_kernel(const ColsumArgs p) {
  __shared__ float red[8][33];
  int cg = blockIdx.x, which = 0;
  while (which < 2 && cg >= p.ncols[which] / 32) { cg -= p.ncols[which] / 32; ++which; }
  const int b = blockIdx.y, lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = cg * 32 + lane;
  const float* A = p.A[which] + (size_t)b * p.T * p.lda[which] + c;
  float s = 0.f;
  for (int t = rg; t < p.T; t += 8) s += A[(size_t)t * p.lda[which]];
  red[rg][lane] = s;
  __syncthreads();
  if (rg == 0) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][lane];
    p.out[which][(size_t)b * p.ldo[which] + c] = v;
  }
}
void launch_colsum(const ColsumArgs& a, int n_mats, int B, cudaStream_t s) {
  int groups = 0;
  for (int i = 0; i < n_mats; ++i) groups += a.ncols[i] / 32;
  utt_colsum_kernel<<<dim3(groups, B), 256, 0, s>>>(a);
  note_launch();
}

// [B][M][T] -> [B*T][M]  (gradient of the output w.r.t. the frames-major layout)
__global__ void bmt_to_btm_kernel2(const float* __restrict__ src, float* __restrict__ dst, int M, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int m = m0 + r, t = t0 + threadIdx.x;
    tile[r][threadIdx.x] = (m < M && t < T) ? src[((size_t)b * M + m) * T + t] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r, m = m0 + threadIdx.x;
    if (t < T && m < M) dst[((size_t)b * T + t) * M + m] = tile[threadIdx.x][r];
  }
}

// w3t[(tap'*2C + co)*C + ci] = W3[co][ci][2 - tap']   (transposed convolution as a 3-tap frame GEMM)
__global__ void pack_w3t_kernel(const float* __restrict__ w3, float* __restrict__ w3t, int C) {
  const int ci = threadIdx.x, co = blockIdx.x, tp = blockIdx.y;
  w3t[((size_t)tp * 2 * C + co) * C + ci] = w3[((size_t)co * C + ci) * 3 + (2 - tp)];
}
// dst[c][128] = src[c][0..M) zero-padded
__global__ void pad_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, int M) {
  const int c = blockIdx.x, n = threadIdx.x;
  dst[(size_t)c * 128 + n] = n < M ? src[(size_t)c * M + n] : 0.f;
}

struct TrainWork {
  size_t fwd;                                   // the forward's scratch (WorkF32) lives at the front
  size_t doutf, dPre, dS, E, dZ, dY, w3t, winpad, part, usumE, usumZ, usumY, usumS, usumT, ddvec, dspk, dpre, dd_all, ds_all, lpart, total;
};
TrainWork train_work_layout(const mgb_model_dims& d, int B, int T) {
  const size_t C = d.channels, H = d.d_encoder, M = d.n_mel, F = (size_t)B * T;
  TrainWork w{};
  size_t p = fp32_workspace_bytes(d, B, T) / sizeof(float);
  p = align_up(p, 64);
  auto take = [&](size_t n) { size_t r = p; p += align_up(n, 64); return r; };
  w.fwd = 0;
  w.doutf = take(F * M);
  w.dPre = take(F * C);
  w.dS = take(F * C);
  w.E = take(F * C);
  w.dZ = take(F * 2 * C);
  w.dY = take(F * C);
  w.w3t = take(3 * 2 * C * C);
  w.winpad = take(C * 128);
  size_t part = wgrad_part_floats((int)F, 2 * C, 3 * C);
  auto mx = [&](size_t v) { if (v > part) part = v; };
  mx(wgrad_part_floats((int)F, 2 * C, C));
  mx(wgrad_part_floats((int)F, C, H));
  mx(wgrad_part_floats((int)F, C, C));
  mx(wgrad_part_floats((int)F, M, C));
  mx(wgrad_part_floats((int)F, C, M));
  w.part = take(part);
  w.usumE = take((size_t)B * C);
  w.usumZ = take((size_t)B * 2 * C);
  w.usumY = take((size_t)B * C);
  w.usumS = take((size_t)B * C);
  w.usumT = take((size_t)B * C);
  w.ddvec = take((size_t)B * C);
  w.dspk = take((size_t)B * H);
  w.dpre = take((size_t)B * 4 * C);
  w.dd_all = take((size_t)d.layers * B * C);
  w.ds_all = take((size_t)d.layers * B * C);
  w.lpart = take((size_t)d.layers * B * C);
  w.total = p;
  return w;
}

}  // namespace

size_t train_workspace_bytes(const mgb_model_dims& d, int B, int T) {
  return train_work_layout(d, B, T).total * sizeof(float);
}

int fp32_train_backward(const mgb_model_dims& d, const float* flat, const float* saved, const int64_t* t,
                        const float* cond, const float* spk, const float* grad_out, float* grad_flat, float* grad_cond,
                        float* grad_spk, float* grad_x, int B, int T, int seg_begin, int seg_end, void* ws,
                        cudaStream_t s) {
  const FlatOffsets f = flat_offsets(d);
  const TrainSaved sv = train_saved_layout(d, B, T);
  const TrainWork w = train_work_layout(d, B, T);
  float* W = static_cast<float*>(ws);
  const int C = d.channels, H = d.d_encoder, M = d.n_mel, L = d.layers;
  const int rows = B * T;
  const float inv_sqrtL = 1.0f / sqrtf((float)L);

  auto frame = [&](const float* A, int lda, const float* Wt, int ldw, int Kin, int taps) {
    gemm32::FrameGemm g{};
    g.A = A; g.lda = lda; g.Wt = Wt; g.ldw = ldw; g.rows = rows; g.T = T; g.Kin = Kin; g.taps = taps;
    return g;
  };

  for (int seg = seg_begin; seg < seg_end; ++seg) {
    if (seg == 0) {
      // ---------------- tail: output projection, ReLU, skip projection, 1/sqrt(L)   (modules.py:441-444)
      {
        dim3 grid((T + 31) / 32, (M + 31) / 32, B), block(32, 8);
        bmt_to_btm_kernel2<<<grid, block, 0, s>>>(grad_out, W + w.doutf, M, T);
        note_launch();
      }
      launch_wgrad(W + w.doutf, M, nullptr, 0, 0, M, saved + sv.P, C, C, 1, rows, T, W + w.part, grad_flat + f.out_w, s);
      {
        ColsumArgs c{};
        c.A[0] = W + w.doutf; c.lda[0] = M; c.ncols[0] = M / 32 * 32; c.out[0] = W + w.usumT; c.ldo[0] = C; c.T = T;
        // n_mel = 80 is not a multiple of 32: sum 64 columns here and the last 16 through a second, overlapping group
        launch_colsum(c, 1, B, s);
        if (M % 32) {
          ColsumArgs c2{};
          c2.A[0] = W + w.doutf + (M - 32); c2.lda[0] = M; c2.ncols[0] = 32; c2.out[0] = W + w.usumT + (M - 32);
          c2.ldo[0] = C; c2.T = T;
          launch_colsum(c2, 1, B, s);
        }
        bias_from_usum_kernel<<<1, 128, 0, s>>>(W + w.usumT, B, M, C, grad_flat + f.out_b);
        note_launch();
      }
      {
        BArgs a{};
        a.g = frame(W + w.doutf, M, flat + f.out_w, C, M, 1);       // dP = dout * Wout  (raw [n_mel][C])
        a.out = W + w.dPre; a.ldo = C; a.C = C; a.aux = saved + sv.P;
        launch_bgemm<B_RELUMASK>(a, C, s);
      }
      launch_wgrad(W + w.dPre, C, nullptr, 0, 0, C, saved + sv.Sn, C, C, 1, rows, T, W + w.part, grad_flat + f.skip_w, s);
      {
        ColsumArgs c{};
        c.A[0] = W + w.dPre; c.lda[0] = C; c.ncols[0] = C; c.out[0] = W + w.usumT; c.ldo[0] = C; c.T = T;
        launch_colsum(c, 1, B, s);
        bias_from_usum_kernel<<<(C + 127) / 128, 128, 0, s>>>(W + w.usumT, B, C, C, grad_flat + f.skip_b);
        note_launch();
      }
      {
        BArgs a{};
        a.g = frame(W + w.dPre, C, flat + f.skip_w, C, C, 1);       // dS = (dPre * Wsk) / sqrt(L)
        a.out = W + w.dS; a.ldo = C; a.C = C; a.scale = inv_sqrtL;
        launch_bgemm<B_SCALE>(a, C, s);
      }
      {
        ColsumArgs c{};
        c.A[0] = W + w.dS; c.lda[0] = C; c.ncols[0] = C; c.out[0] = W + w.usumS; c.ldo[0] = C; c.T = T;
        launch_colsum(c, 1, B, s);
      }
    } else if (seg <= L) {
      // ---------------- residual block l   (blocks.py:1157-1176)
      const int l = L - seg;
      const bool top = (l == L - 1);
      const float* fl = flat + f.layer0 + (size_t)l * f.layer_stride;
      float* gl = grad_flat + f.layer0 + (size_t)l * f.layer_stride;
      const float* sl = saved + sv.layer0 + (size_t)l * sv.layer_stride;
      // 1. dG = [e | dS] * Wo (raw [2C][C]); gate backward -> dZ.  The top block's x output is unused: e = 0.
      {
        BArgs a{};
        if (top) {
          a.g = frame(W + w.dS, C, fl + f.rel.oproj_w + (size_t)C * C, C, C, 1);
        } else {
          a.g = frame(W + w.E, C, fl + f.rel.oproj_w, C, 2 * C, 1);
          a.g.A2 = W + w.dS; a.g.lda2 = C; a.g.ksplit = C;
        }
        a.out = W + w.dZ; a.aux = sl + sv.rZ; a.C = C;
        launch_bgemm<B_GATE>(a, C, s);
      }
      // 2. dWo = [e | dS]^T g
      if (top) {
        MGB_CUDA_CHECK(cudaMemsetAsync(gl + f.rel.oproj_w, 0, sizeof(float) * (size_t)C * C, s));
        launch_wgrad(W + w.dS, C, nullptr, 0, 0, C, sl + sv.rG, C, C, 1, rows, T, W + w.part,
                     gl + f.rel.oproj_w + (size_t)C * C, s);
      } else {
        launch_wgrad(W + w.E, C, W + w.dS, C, C, 2 * C, sl + sv.rG, C, C, 1, rows, T, W + w.part, gl + f.rel.oproj_w, s);
      }
      // 3. per-utterance column sums of e_l and dZ
      {
        ColsumArgs c{};
        c.T = T;
        c.A[0] = W + w.dZ; c.lda[0] = 2 * C; c.ncols[0] = 2 * C; c.out[0] = W + w.usumZ; c.ldo[0] = 2 * C;
        int n = 1;
        if (!top) {
          c.A[1] = W + w.E; c.lda[1] = C; c.ncols[1] = C; c.out[1] = W + w.usumE; c.ldo[1] = C;
          n = 2;
        }
        launch_colsum(c, n, B, s);
      }
      // 4. dY = conv3^T(dZ); dx_l = e_l + dY; e_{l-1} = dx_l / sqrt(2)  (layer 0: ReLU mask of the input projection)
      {
        pack_w3t_kernel<<<dim3(2 * C, 3), C, 0, s>>>(fl + f.rel.conv_w, W + w.w3t, C);
        note_launch();
        BArgs a{};
        a.g = frame(W + w.dZ, 2 * C, W + w.w3t, C, 2 * C, 3);
        a.out = W + w.E; a.out2 = W + w.dY; a.C = C; a.first = top ? 1 : 0;
        a.relu_mask = (l == 0) ? 1 : 0; a.aux = saved + sv.X0;
        launch_bgemm<B_DX>(a, C, s);
      }
      // 5. dW3 = dZ^T shift(y_l)
      launch_wgrad(W + w.dZ, 2 * C, nullptr, 0, 0, 2 * C, sl + sv.rY, C, C, 3, rows, T, W + w.part, gl + f.rel.conv_w, s);
      // 6. dWc = dY^T cond;  dCond (+)= dY * Wc (raw [C][H])
      launch_wgrad(W + w.dY, C, nullptr, 0, 0, C, cond, H, H, 1, rows, T, W + w.part, gl + f.rel.cproj_w, s);
      if (grad_cond) {
        BArgs a{};
        a.g = frame(W + w.dY, C, fl + f.rel.cproj_w, H, C, 1);
        a.out = grad_cond; a.ldo = H; a.C = C; a.first = top ? 1 : 0;
        launch_bgemm<B_ACC>(a, H, s);
      }
      // 7. per-utterance sums of dY, then the small per-layer terms
      {
        ColsumArgs c{};
        c.T = T;
        c.A[0] = W + w.dY; c.lda[0] = C; c.ncols[0] = C; c.out[0] = W + w.usumY; c.ldo[0] = C;
        launch_colsum(c, 1, B, s);
        LayerSmallArgs q{};
        q.usumE = top ? nullptr : W + w.usumE; q.usumZ = W + w.usumZ; q.usumY = W + w.usumY; q.usumS = W + w.usumS;
        q.dvec = saved + sv.dvec; q.spk = d.multi_speaker ? spk : nullptr;
        q.Wd = fl + f.rel.dproj_w; q.Ws = d.multi_speaker ? fl + f.rel.sproj_w : nullptr;
        q.g_conv_b = gl + f.rel.conv_b; q.g_oproj_b = gl + f.rel.oproj_b; q.g_cproj_b = gl + f.rel.cproj_b;
        q.g_dproj_w = gl + f.rel.dproj_w; q.g_sproj_w = d.multi_speaker ? gl + f.rel.sproj_w : nullptr;
        q.dd_l = W + w.dd_all + (size_t)l * B * C; q.ds_l = d.multi_speaker ? W + w.ds_all + (size_t)l * B * C : nullptr;
        q.B = B; q.C = C; q.H = H;
        layer_small_kernel<<<C + B + 1, 256, 0, s>>>(q);
        note_launch();
      }
    } else if (seg == L + 1) {
      // ---------------- head: input projection (E now holds relu-masked d loss / d (Win x + b)) and the step MLP
      launch_wgrad(W + w.E, C, nullptr, 0, 0, C, saved + sv.xt, M, M, 1, rows, T, W + w.part, grad_flat + f.in_w, s);
      {
        ColsumArgs c{};
        c.A[0] = W + w.E; c.lda[0] = C; c.ncols[0] = C; c.out[0] = W + w.usumT; c.ldo[0] = C; c.T = T;
        launch_colsum(c, 1, B, s);
        bias_from_usum_kernel<<<(C + 127) / 128, 128, 0, s>>>(W + w.usumT, B, C, C, grad_flat + f.in_b);
        note_launch();
      }
      if (grad_x) {
        pad_cols_kernel<<<C, 128, 0, s>>>(flat + f.in_w, W + w.winpad, M);
        BArgs a{};
        a.g = frame(W + w.E, C, W + w.winpad, 128, C, 1);
        a.out = grad_x; a.C = C; a.n_mel = M;
        launch_bgemm<B_DXT>(a, 128, s);
        note_launch();
      }
      launch_dvec_contraction(W + w.dd_all, W + w.ds_all, flat + f.layer0 + f.rel.dproj_w,
                              d.multi_speaker ? flat + f.layer0 + f.rel.sproj_w : nullptr, f.layer_stride, W + w.lpart,
                              W + w.ddvec, W + w.dspk, B, L, s);
      mlp_bwd_w2_kernel<<<C, 256, 0, s>>>(W + w.ddvec, saved + sv.h, grad_flat + f.mlp2_w, B, C);
      mlp_bwd_pre_kernel<<<dim3(4 * C / 256, B), 256, 0, s>>>(t, W + w.ddvec, flat + f.mlp0_w, flat + f.mlp2_w,
                                                             W + w.dpre, C);
      mlp_bwd_w0_kernel<<<4 * C, C, 0, s>>>(t, W + w.dpre, grad_flat + f.mlp0_w, B, C);
      note_launch(3);
      if (grad_spk && d.multi_speaker)
        MGB_CUDA_CHECK(cudaMemcpyAsync(grad_spk, W + w.dspk, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, s));
    }
  }
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

}  // namespace mgb

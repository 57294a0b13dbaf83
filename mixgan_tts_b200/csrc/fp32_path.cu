// fp32 parity path of the Denoiser: every convolution is an implicit GEMM on the CUDA cores
// (fp32 operands, fp32 accumulation) with the surrounding elementwise work fused into the GEMM
// epilogues.  This is the MGB_PREC_FP32 arithmetic; the tcgen05 bf16 path lives in fused_bf16.cu.
//
// Reference being restated (behaviour only): Denoiser.forward model/modules.py:420-446,
// ResidualBlock.forward model/blocks.py:1157-1176, DiffusionEmbedding model/blocks.py:906-913,
// Mish model/blocks.py:894-896, q_posterior_sample model/diffusion.py:104-119.
//
// Activation layout in HBM: frames-major [B*T][C] fp32 (one row per mel frame), so a k=3
// convolution is three row-shifted GEMMs over the same matrix with zero rows at utterance edges.
#include "common.cuh"
#include "gemm_fp32.cuh"
#include "small_ops.cuh"

namespace mgb {

namespace {

using namespace smallops;

using gemm32::BK; using gemm32::BM; using gemm32::BMP; using gemm32::BN; using gemm32::NT;

enum Epi { EPI_RELU = 0, EPI_COND = 1, EPI_GATE = 2, EPI_OUT = 3, EPI_FINAL = 4 };

struct GemmArgs {
  const float* A;    // [rows][lda]
  const float* Wt;   // [taps*Kin][ldw]
  const float* bias; // [N] (packed column order)
  int lda, ldw, rows, T, Kin, taps;
  // epilogue operands
  float* out;          // RELU/COND/GATE: [rows][C]; OUT: X in/out; FINAL: x_prev or x0 [B][M][T]
  float* out2;         // OUT: skip accumulator; FINAL: optional x0 copy
  const float* x;      // COND: X [rows][C]; OUT: residual-stream input when it is not `out` (training: X0 is kept)
  float* zsave;        // GATE: optional [rows][2C] copy of the pre-activation (gate | filter halves), for the backward
  const float* dtab;   // COND/OUT: per-utterance step bias for this layer, row stride tab_stride
  const float* ctab;   // COND: per-utterance conditioner-side bias (bc + speaker), same stride
  int tab_stride;
  int C;               // channels (ld of out)
  int first, last;     // OUT: first/last layer
  float inv_div;       // OUT last: sqrt(L) divisor
  // FINAL
  const float* x_t; const float* noise; const float* sched; const int64_t* t;
  int K, clip, n_mel;
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

template <int EPI>
__global__ void __launch_bounds__(NT) conv_gemm_kernel(const GemmArgs p) {
  __shared__ __align__(16) float As[2][BK][BMP];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[8][8];
  {
    gemm32::FrameGemm g{p.A, nullptr, p.Wt, p.lda, 0, 0, p.ldw, p.rows, p.T, p.Kin, p.taps};
    gemm32::frame_gemm_mainloop(g, m0, n0, acc, As, Bs);
  }

  // ---------------------------------------------------------------- epilogues
  const int C = p.C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.rows) continue;
    const int b = m / p.T;
    if constexpr (EPI == EPI_RELU || EPI == EPI_COND) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + h * 64 + tx * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][h * 4 + j];
        if constexpr (EPI == EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j] + p.bias[n + j], 0.f);
        } else {
          const float4 xv = *reinterpret_cast<const float4*>(p.x + (size_t)m * C + n);
          const float4 dv = *reinterpret_cast<const float4*>(p.dtab + (size_t)b * p.tab_stride + n);
          const float4 cv = *reinterpret_cast<const float4*>(p.ctab + (size_t)b * p.tab_stride + n);
          // (x + d) + (Wc*cond + bc) [+ s]: same association as blocks.py:1166-1168
          v[0] = (xv.x + dv.x) + (v[0] + cv.x);
          v[1] = (xv.y + dv.y) + (v[1] + cv.y);
          v[2] = (xv.z + dv.z) + (v[2] + cv.z);
          v[3] = (xv.w + dv.w) + (v[3] + cv.w);
        }
        *reinterpret_cast<float4*>(p.out + (size_t)m * C + n) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else if constexpr (EPI == EPI_GATE || EPI == EPI_OUT) {
      // packed column tile: [0,64) first-half channels (gate / x), [64,128) second half (filter / skip)
      const int ch = (n0 >> 1) + tx * 4;
      const int nb = n0 + tx * 4;
      float lo[4], hi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lo[j] = acc[i][j] + p.bias[nb + j];
        hi[j] = acc[i][4 + j] + p.bias[nb + 64 + j];
      }
      if constexpr (EPI == EPI_GATE) {
        float g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = sigmoidf_(lo[j]) * tanhf(hi[j]);
        *reinterpret_cast<float4*>(p.out + (size_t)m * C + ch) = make_float4(g[0], g[1], g[2], g[3]);
        if (p.zsave) {
          *reinterpret_cast<float4*>(p.zsave + (size_t)m * 2 * C + ch) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<float4*>(p.zsave + (size_t)m * 2 * C + C + ch) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        }
      } else {
        float4* xp = reinterpret_cast<float4*>(p.out + (size_t)m * C + ch);
        float4* sp = reinterpret_cast<float4*>(p.out2 + (size_t)m * C + ch);
        const float4 xv = p.x ? *reinterpret_cast<const float4*>(p.x + (size_t)m * C + ch) : *xp;
        const float4 dv = *reinterpret_cast<const float4*>(p.dtab + (size_t)b * p.tab_stride + ch);
        const float SQRT2 = 1.41421356237309504880f;
        float4 xn;
        xn.x = __fdiv_rn(lo[0] + (xv.x + dv.x), SQRT2);
        xn.y = __fdiv_rn(lo[1] + (xv.y + dv.y), SQRT2);
        xn.z = __fdiv_rn(lo[2] + (xv.z + dv.z), SQRT2);
        xn.w = __fdiv_rn(lo[3] + (xv.w + dv.w), SQRT2);
        *xp = xn;
        float4 sv = p.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *sp;
        sv.x += hi[0]; sv.y += hi[1]; sv.z += hi[2]; sv.w += hi[3];
        if (p.last) {
          sv.x = __fdiv_rn(sv.x, p.inv_div); sv.y = __fdiv_rn(sv.y, p.inv_div);
          sv.z = __fdiv_rn(sv.z, p.inv_div); sv.w = __fdiv_rn(sv.w, p.inv_div);
        }
        *sp = sv;
      }
    } else {  // EPI_FINAL
      const int t = m - b * p.T;
      float c1 = 0.f, c2 = 0.f, sg = 0.f;
      if (p.sched) {
        const int tb = min(max((int)p.t[b], 0), p.K - 1);   // never read outside the schedule (the host validates t)
        c1 = p.sched[tb]; c2 = p.sched[p.K + tb]; sg = p.sched[2 * p.K + tb];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        if (n >= p.n_mel) continue;
        float x0 = acc[i][j] + p.bias[n];
        if (p.clip && x0 == x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);   // NaN propagates, as torch.clamp does
        const size_t o = ((size_t)b * p.n_mel + n) * p.T + t;
        if (p.out2) p.out2[o] = x0;
        if (p.sched) {
          const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, p.x_t[o]));
          p.out[o] = __fadd_rn(mean, __fmul_rn(sg, p.noise[o]));
        } else if (!p.out2 || p.out != p.out2) {
          p.out[o] = x0;
        }
      }
    }
  }
}

// [B][M][T] -> [B*T][M]
__global__ void bmt_to_btm_kernel(const float* __restrict__ src, float* __restrict__ dst, int M, int T) {
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

struct PackedF32 {
  size_t in_wt, in_b, mlp0_wt, mlp2_wt, layer0, layer_stride;
  size_t r_cproj_wt, r_conv_wt, r_conv_b, r_oproj_wt, r_oproj_b, r_cproj_b, r_dproj_wt, r_sproj_wt;
  size_t skip_wt, skip_b, out_wt, out_b, total;
};

PackedF32 packed_layout(const mgb_model_dims& d) {
  const size_t C = d.channels, H = d.d_encoder, M = d.n_mel;
  PackedF32 o{};
  size_t p = 0;
  o.in_wt = p; p += M * C;
  o.in_b = p; p += C;
  o.mlp0_wt = p; p += C * 4 * C;
  o.mlp2_wt = p; p += 4 * C * C;
  o.layer0 = p;
  size_t q = 0;
  o.r_cproj_wt = q; q += H * C;
  o.r_conv_wt = q; q += 3 * C * 2 * C;
  o.r_conv_b = q; q += 2 * C;
  o.r_oproj_wt = q; q += C * 2 * C;
  o.r_oproj_b = q; q += 2 * C;
  o.r_cproj_b = q; q += C;
  o.r_dproj_wt = q; q += C * C;
  o.r_sproj_wt = q; if (d.multi_speaker) q += H * C;
  o.layer_stride = q;
  p += q * d.layers;
  o.skip_wt = p; p += C * C;
  o.skip_b = p; p += C;
  o.out_wt = p; p += C * 128;
  o.out_b = p; p += 128;
  o.total = p;
  return o;
}

struct WorkF32 {
  size_t X, Y, G, S, xt, d, h, dtab, ctab, total;
};
WorkF32 work_layout(const mgb_model_dims& d, int B, int T) {
  const size_t C = d.channels, BT = (size_t)B * T;
  WorkF32 w{};
  size_t p = 0;
  auto take = [&](size_t n) { size_t r = p; p += align_up(n, 64); return r; };
  w.X = take(BT * C); w.Y = take(BT * C); w.G = take(BT * C); w.S = take(BT * C);
  w.xt = take(BT * d.n_mel);
  w.d = take((size_t)B * C);
  w.h = take((size_t)B * 4 * C);
  w.dtab = take((size_t)B * d.layers * C);
  w.ctab = take((size_t)B * d.layers * C);
  w.total = p;
  return w;
}

template <int EPI>
void launch_gemm(const GemmArgs& a, int N, cudaStream_t s) {
  dim3 grid((a.rows + BM - 1) / BM, N / BN);
  if (EPI == EPI_GATE) prof_begin(s);   // the dominant kernel of this path (k=3 conv + gate)
  conv_gemm_kernel<EPI><<<grid, NT, 0, s>>>(a);
  if (EPI == EPI_GATE) prof_end(s);
  note_launch();
}

}  // namespace

size_t fp32_packed_bytes(const mgb_model_dims& d) { return packed_layout(d).total * sizeof(float); }
size_t fp32_workspace_bytes(const mgb_model_dims& d, int B, int T) {
  return work_layout(d, B, T).total * sizeof(float);
}

int fp32_pack(const mgb_model_dims& d, const float* flat, void* packed, cudaStream_t s) {
  const FlatOffsets f = flat_offsets(d);
  const PackedF32 o = packed_layout(d);
  float* P = static_cast<float*>(packed);
  const int C = d.channels, H = d.d_encoder, M = d.n_mel;
  launch_pack(flat + f.in_w, P + o.in_wt, C, M, 1, C, 0, 0, s);
  pack_bias_kernel<<<(C + 127) / 128, 128, 0, s>>>(flat + f.in_b, P + o.in_b, C, C, 0, 0);
  launch_pack(flat + f.mlp0_w, P + o.mlp0_wt, 4 * C, C, 1, 4 * C, 0, 0, s);
  launch_pack(flat + f.mlp2_w, P + o.mlp2_wt, C, 4 * C, 1, C, 0, 0, s);
  for (int l = 0; l < d.layers; ++l) {
    const float* fl = flat + f.layer0 + (size_t)l * f.layer_stride;
    float* pl = P + o.layer0 + (size_t)l * o.layer_stride;
    launch_pack(fl + f.rel.cproj_w, pl + o.r_cproj_wt, C, H, 1, C, 0, 0, s);
    launch_pack(fl + f.rel.conv_w, pl + o.r_conv_wt, 2 * C, C, 3, 2 * C, 1, C, s);
    pack_bias_kernel<<<(2 * C + 127) / 128, 128, 0, s>>>(fl + f.rel.conv_b, pl + o.r_conv_b, 2 * C, 2 * C, 1, C);
    launch_pack(fl + f.rel.oproj_w, pl + o.r_oproj_wt, 2 * C, C, 1, 2 * C, 1, C, s);
    pack_bias_kernel<<<(2 * C + 127) / 128, 128, 0, s>>>(fl + f.rel.oproj_b, pl + o.r_oproj_b, 2 * C, 2 * C, 1, C);
    pack_bias_kernel<<<(C + 127) / 128, 128, 0, s>>>(fl + f.rel.cproj_b, pl + o.r_cproj_b, C, C, 0, 0);
    launch_pack(fl + f.rel.dproj_w, pl + o.r_dproj_wt, C, C, 1, C, 0, 0, s);
    if (d.multi_speaker) launch_pack(fl + f.rel.sproj_w, pl + o.r_sproj_wt, C, H, 1, C, 0, 0, s);
  }
  launch_pack(flat + f.skip_w, P + o.skip_wt, C, C, 1, C, 0, 0, s);
  pack_bias_kernel<<<(C + 127) / 128, 128, 0, s>>>(flat + f.skip_b, P + o.skip_b, C, C, 0, 0);
  launch_pack(flat + f.out_w, P + o.out_wt, M, C, 1, 128, 0, 0, s);
  pack_bias_kernel<<<1, 128, 0, s>>>(flat + f.out_b, P + o.out_b, M, 128, 0, 0);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

// Only what fp32_step_tables reads (step MLP, diffusion / speaker projections, conditioner bias), same offsets as the full
// pack, ONE launch: the bf16 training forward re-packs these every step (the parameters change every step).
struct TablesPack {
  size_t src[4], dst[4];   // [0] mlp0 [4C][C] -> [C][4C], [1] mlp2 [C][4C] -> [4C][C], [2] dproj [C][C] (per layer), [3] sproj
  size_t src_cb, dst_cb, src_lstride, dst_lstride;
  int C, H, L, multi;
};
__global__ void __launch_bounds__(256) pack_tables_kernel(const float* __restrict__ flat, float* __restrict__ P, const TablesPack a) {
  // blockIdx.y: 0 = mlp0, 1 = mlp2, 2 + l = layer l (dproj, sproj, cproj_b)
  const int which = blockIdx.y, C = a.C, H = a.H;
  auto transpose = [&](const float* src, float* dst, int n_out, int n_in) {   // dst[k][n] = src[n][k]
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out * n_in; i += gridDim.x * blockDim.x) {
      const int k = i / n_out, n = i - k * n_out;
      dst[i] = src[(size_t)n * n_in + k];
    }
  };
  if (which == 0) transpose(flat + a.src[0], P + a.dst[0], 4 * C, C);
  else if (which == 1) transpose(flat + a.src[1], P + a.dst[1], C, 4 * C);
  else {
    const int l = which - 2;
    const float* fl = flat + (size_t)l * a.src_lstride;
    float* pl = P + (size_t)l * a.dst_lstride;
    transpose(fl + a.src[2], pl + a.dst[2], C, C);
    if (a.multi) transpose(fl + a.src[3], pl + a.dst[3], C, H);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C; i += gridDim.x * blockDim.x) pl[a.dst_cb + i] = fl[a.src_cb + i];
  }
}
int fp32_pack_tables(const mgb_model_dims& d, const float* flat, void* packed, cudaStream_t s) {
  const FlatOffsets f = flat_offsets(d);
  const PackedF32 o = packed_layout(d);
  TablesPack a{};
  a.C = d.channels; a.H = d.d_encoder; a.L = d.layers; a.multi = d.multi_speaker;
  a.src[0] = f.mlp0_w; a.dst[0] = o.mlp0_wt;
  a.src[1] = f.mlp2_w; a.dst[1] = o.mlp2_wt;
  a.src[2] = f.layer0 + f.rel.dproj_w; a.dst[2] = o.layer0 + o.r_dproj_wt;
  a.src[3] = f.layer0 + f.rel.sproj_w; a.dst[3] = o.layer0 + o.r_sproj_wt;
  a.src_cb = f.layer0 + f.rel.cproj_b; a.dst_cb = o.layer0 + o.r_cproj_b;
  a.src_lstride = f.layer_stride; a.dst_lstride = o.layer_stride;
  pack_tables_kernel<<<dim3(32, 2 + d.layers), 256, 0, s>>>(flat, static_cast<float*>(packed), a);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

// d[B][C] = step MLP(t); dtab[b][l] = Wd_l d[b]; ctab[b][l] = bc_l (+ Ws_l spk[b])   (blocks.py:1159-1164)
int fp32_step_tables(const mgb_model_dims& d, const void* packed, const int64_t* t, const float* spk, int B, float* d_buf,
                     float* h_buf, float* dtab, float* ctab, cudaStream_t s) {
  const PackedF32 o = packed_layout(d);
  const float* P = static_cast<const float*>(packed);
  const int C = d.channels, H = d.d_encoder, L = d.layers;
  launch_step_mlp(t, P + o.mlp0_wt, P + o.mlp2_wt, h_buf, d_buf, B, C, s);
  dim3 grid(L, (B + TAB_UB - 1) / TAB_UB);
  proj_table_kernel<<<grid, 256, (size_t)TAB_UB * C * sizeof(float), s>>>(
      d_buf, C, P + o.layer0 + o.r_dproj_wt, o.layer_stride, nullptr, 0, dtab, B, L, C);
  proj_table_kernel<<<grid, 256, (size_t)TAB_UB * H * sizeof(float), s>>>(
      d.multi_speaker ? spk : nullptr, H, P + o.layer0 + o.r_sproj_wt, o.layer_stride,
      P + o.layer0 + o.r_cproj_b, o.layer_stride, ctab, B, L, C);
  note_launch(4);   // step MLP (2), two projection tables
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int fp32_denoiser(const mgb_model_dims& d, const void* packed, const float* x, const int64_t* t,
                  const float* cond, const float* spk, const float* noise, const float* sched, int K,
                  int clip, float* x_prev, float* out_x0, int B, int T, void* ws, cudaStream_t s, float* saved) {
  const PackedF32 o = packed_layout(d);
  const WorkF32 w = work_layout(d, B, T);
  const float* P = static_cast<const float*>(packed);
  float* W = static_cast<float*>(ws);
  const int C = d.channels, H = d.d_encoder, M = d.n_mel, L = d.layers;
  const int rows = B * T;
  // training: activations go straight into the caller's stash instead of the reused scratch buffers
  const TrainSaved sv = train_saved_layout(d, B, T);
  float* const xt_buf = saved ? saved + sv.xt : W + w.xt;
  float* const x0_buf = saved ? saved + sv.X0 : W + w.X;
  float* const s_buf = saved ? saved + sv.Sn : W + w.S;
  float* const p_buf = saved ? saved + sv.P : W + w.Y;
  float* const d_buf = saved ? saved + sv.dvec : W + w.d;
  float* const h_buf = saved ? saved + sv.h : W + w.h;

  // per-utterance step embedding, MLP and the per-layer projection tables
  fp32_step_tables(d, packed, t, spk, B, d_buf, h_buf, W + w.dtab, W + w.ctab, s);
  {
    dim3 grid((T + 31) / 32, (M + 31) / 32, B), block(32, 8);
    bmt_to_btm_kernel<<<grid, block, 0, s>>>(x, xt_buf, M, T);
    note_launch();
  }
  GemmArgs a{};
  a.rows = rows; a.T = T; a.C = C; a.tab_stride = L * C;
  // input projection + ReLU (the second F.relu at modules.py:431 is idempotent)
  a.A = xt_buf; a.lda = M; a.Wt = P + o.in_wt; a.ldw = C; a.bias = P + o.in_b; a.Kin = M; a.taps = 1;
  a.out = x0_buf;
  launch_gemm<EPI_RELU>(a, C, s);
  for (int l = 0; l < L; ++l) {
    const float* pl = P + o.layer0 + (size_t)l * o.layer_stride;
    float* const sl = saved ? saved + sv.layer0 + (size_t)l * sv.layer_stride : nullptr;
    const float* x_in = l == 0 ? x0_buf : W + w.X;
    float* const y_buf = sl ? sl + sv.rY : W + w.Y;
    float* const g_buf = sl ? sl + sv.rG : W + w.G;
    GemmArgs c = a;
    c.A = cond; c.lda = H; c.Wt = pl + o.r_cproj_wt; c.ldw = C; c.Kin = H; c.taps = 1; c.bias = nullptr;
    c.x = x_in; c.dtab = W + w.dtab + (size_t)l * C; c.ctab = W + w.ctab + (size_t)l * C;
    c.out = y_buf;
    launch_gemm<EPI_COND>(c, C, s);
    GemmArgs g = a;
    g.A = y_buf; g.lda = C; g.Wt = pl + o.r_conv_wt; g.ldw = 2 * C; g.Kin = C; g.taps = 3;
    g.bias = pl + o.r_conv_b; g.out = g_buf; g.zsave = sl ? sl + sv.rZ : nullptr;
    launch_gemm<EPI_GATE>(g, 2 * C, s);
    GemmArgs q = a;
    q.A = g_buf; q.lda = C; q.Wt = pl + o.r_oproj_wt; q.ldw = 2 * C; q.Kin = C; q.taps = 1;
    q.bias = pl + o.r_oproj_b; q.out = W + w.X; q.out2 = s_buf; q.dtab = W + w.dtab + (size_t)l * C;
    q.x = (x_in == W + w.X) ? nullptr : x_in;   // layer 0 of a training forward reads X0 and leaves it intact
    q.first = (l == 0); q.last = (l == L - 1); q.inv_div = sqrtf((float)L);
    launch_gemm<EPI_OUT>(q, 2 * C, s);
  }
  GemmArgs k = a;
  k.A = s_buf; k.lda = C; k.Wt = P + o.skip_wt; k.ldw = C; k.Kin = C; k.taps = 1; k.bias = P + o.skip_b;
  k.out = p_buf;
  launch_gemm<EPI_RELU>(k, C, s);
  GemmArgs f = a;
  f.A = p_buf; f.lda = C; f.Wt = P + o.out_wt; f.ldw = 128; f.Kin = C; f.taps = 1; f.bias = P + o.out_b;
  f.n_mel = M; f.clip = clip; f.K = K; f.t = t; f.sched = sched; f.x_t = x; f.noise = noise;
  f.out = sched ? x_prev : out_x0; f.out2 = sched ? out_x0 : nullptr;
  launch_gemm<EPI_FINAL>(f, 128, s);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int fp32_train_forward(const mgb_model_dims& d, const void* packed, const float* x, const int64_t* t, const float* cond,
                       const float* spk, float* out, float* saved, int B, int T, void* ws, cudaStream_t s) {
  return fp32_denoiser(d, packed, x, t, cond, spk, nullptr, nullptr, 0, 0, nullptr, out, B, T, ws, s, saved);
}

}  // namespace mgb

// Aux decoder of the shallow-diffusion configs (second half): the FastSpeech2 decoder that
// produces the coarse mel the reverse diffusion starts from.  Reference (model/mixgantts.py:139-143):
//     coarse = mel_linear(decoder(output, mel_masks));  coarse = postnet(coarse) + coarse
// with  Decoder            transformer/Models.py:103-171   (+ position_enc, 6 FFTBlocks, eval branch)
//       FFTBlock           transformer/Layers.py:11-31     (self-attention, masked_fill, conv FFN, masked_fill)
//       MultiHeadAttention transformer/SubLayers.py:8-59   (2 heads x 128, key mask, fc, residual, LayerNorm)
//       ScaledDotProduct   transformer/Modules.py:6-24     (softmax(q k^T / sqrt(d_k)) v)
//       PositionwiseFFN    transformer/SubLayers.py:62-97  (Conv1d k=9 -> ReLU -> Conv1d k=1, residual, LayerNorm)
//       PostNet            transformer/Layers.py:67-137    (5 x Conv1d k=5 + BatchNorm1d (eval) + tanh)
// Inference only (dropout = identity, BatchNorm running statistics folded into the packed weights).
//
// Every matrix product runs on tcgen05 (fp16 operands, fp32 accumulation): the linear / convolution layers through the
// engine of tcnet.cuh with bias, ReLU / tanh, residual, LayerNorm and the padding mask fused into the GEMM epilogues;
// the attention through attn_kernel below.  5 launches per FFT block, 38 per call.
#include "tcnet.cuh"

#include "tc05.cuh"
#include "tmap.cuh"

#include <vector>

namespace mgb {
namespace {

using namespace tcnet;

constexpr long long kTimeout = 400000000LL;
constexpr int GAP = 4;            // zero rows between utterances: (9 - 1) / 2, the widest padding on this path
constexpr int DK = 128;           // head width the attention kernel is written for

// =====================================================================================================
// Self-attention on the tensor cores, two passes over the keys of one utterance per 128-query tile:
//   pass 1   S = Q K^T per 128-key block (TMEM, double-buffered) -> row max (registers)
//   pass 2   S again -> P = exp(S/sqrt(d) - max) <= 1, unnormalised, as fp16 into shared memory (K-major A operand), row
//            sums in fp32 -> O += P V with V read as an MN-major B operand straight from its image box (no transpose)
// so O never needs rescaling, every exponential is taken once, and O leaves TMEM once, scaled by 1 / sum.  Keys at or beyond the utterance's length are masked (-inf, as
// SubLayers.py:47 / Modules.py:19-20) and key blocks beyond it are never loaded; query tiles beyond it are skipped
// (FFTBlock zero-fills those rows, Layers.py:27 — the next GEMM's epilogue does that here).
// 320 threads: TMA warp, MMA warp, 8 softmax warps (thread = query row = TMEM lane x one half of a key block's columns).
// =====================================================================================================
constexpr int ATT_TILE_BYTES = 128 * DK * 2;     // 32 KB: 128 rows x 128 channels fp16
constexpr int ATT_RING = 4;
constexpr int ATT_SMEM = (1 + ATT_RING + 2) * ATT_TILE_BYTES + 1024;

struct AttArgs {
  const int* lens; int B, T, Tg; long long Rp;
  __half* out;              // image [d_model/8][Rp][8]
  int q_chunk0, k_chunk0, v_chunk0;   // first 8-channel chunk of Q / K / V inside the fused QKV image (head 0)
  float c2;                 // log2(e) / sqrt(d_k)
  int* status;
};

__global__ void __launch_bounds__(320, 1) attn_kernel(const AttArgs p, const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_q, ring_full[ATT_RING], ring_empty[ATT_RING], s_full[2], s_free[2], p_full[2],
      p_free[2], o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_stat[2][128];          // row max, then row sum, of the two column halves

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  int len = p.T;
  if (p.lens) { len = p.lens[b]; len = len < 1 ? 1 : (len > p.T ? p.T : len); }   // lens is an input of the call, not of the predecessor
  if (qt * 128 >= len) return;                       // whole tile is padding: its rows are zero-filled downstream
  const int nblk = (len + 127) >> 7;

  if (warp == 1) tc::tmem_alloc<512>(&tmem_slot);
  if (tid == 0) {
    tc::mbar_init(&bar_q, 1);
    for (int i = 0; i < ATT_RING; ++i) { tc::mbar_init(&ring_full[i], 1); tc::mbar_init(&ring_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_free[i], 8);
      tc::mbar_init(&p_full[i], 8); tc::mbar_init(&p_free[i], 1);
    }
    tc::mbar_init(&o_full, 1);
    tc::fence_barrier_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_trigger();
  pdl_wait();

  const uint32_t q_base = tc::smem_u32(smem);
  const uint32_t ring_base = q_base + ATT_TILE_BYTES;
  const uint32_t p_base = ring_base + ATT_RING * ATT_TILE_BYTES;
  const int row_u = b * p.Tg;                        // first row of this utterance

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx_addr(tc::smem_u32(&bar_q), ATT_TILE_BYTES);
      tc::tma_load_2d(q_base, &tm, 2 * (row_u + qt * 128), p.q_chunk0 + h * (DK / 8), tc::smem_u32(&bar_q));
      int idx = 0;
      auto load = [&](int chunk0, int j) {
        const int slot = idx % ATT_RING, ph = (idx / ATT_RING) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&ring_empty[slot]), ph ^ 1, kTimeout, p.status, 1);
        const uint32_t fb = tc::smem_u32(&ring_full[slot]);
        tc::mbar_arrive_expect_tx_addr(fb, ATT_TILE_BYTES);
        tc::tma_load_2d(ring_base + slot * ATT_TILE_BYTES, &tm, 2 * (row_u + j * 128), chunk0 + h * (DK / 8), fb);
        ++idx;
      };
      for (int j = 0; j < nblk; ++j) load(p.k_chunk0, j);
      for (int j = 0; j < nblk; ++j) {               // the order the MMA warp consumes them in
        if (j == 0) load(p.k_chunk0, 0);
        if (j + 1 < nblk) load(p.k_chunk0, j + 1);
        load(p.v_chunk0, j);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = tc::make_idesc_16(128, 128, true);
      const uint32_t idesc_o = idesc_s | (1u << 16);          // B (= V) is MN-major: its K axis is the key-row axis
      tc::mbar_wait_trap(tc::smem_u32(&bar_q), 0, kTimeout, p.status, 2);
      int idx = 0, c = 0;
      auto issue_s = [&]() {
        const int slot = idx % ATT_RING, ph = (idx / ATT_RING) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&ring_full[slot]), ph, kTimeout, p.status, 2);
        if (c >= 2) tc::mbar_wait_trap(tc::smem_u32(&s_free[c & 1]), ((c >> 1) - 1) & 1, kTimeout, p.status, 2);
        tc::tc_fence_after();
        const uint32_t kb = ring_base + slot * ATT_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < DK / 16; ++kk) {
          const uint64_t ad = tc::make_smem_desc(q_base + kk * 2 * 2048, 2048, 128);
          const uint64_t bd = tc::make_smem_desc(kb + kk * 2 * 2048, 2048, 128);
          tc::umma_bf16(tmem + (c & 1) * 128, ad, bd, idesc_s, kk ? 1u : 0u);
        }
        tc::umma_commit(&ring_empty[slot]);
        tc::umma_commit(&s_full[c & 1]);
        ++c; ++idx;
      };
      for (int j = 0; j < nblk; ++j) issue_s();
      issue_s();
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s();
        const int slot = idx % ATT_RING, ph = (idx / ATT_RING) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&ring_full[slot]), ph, kTimeout, p.status, 2);
        tc::mbar_wait_trap(tc::smem_u32(&p_full[j & 1]), (j >> 1) & 1, kTimeout, p.status, 2);
        tc::tc_fence_after();
        const uint32_t vb = ring_base + slot * ATT_TILE_BYTES, pb = p_base + (j & 1) * ATT_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 128 / 16; ++kk) {      // 16 keys per MMA
          const uint64_t ad = tc::make_smem_desc(pb + kk * 2 * 2048, 2048, 128);
          const uint64_t bd = tc::make_smem_desc(vb + kk * 256, 128, 2048);
          tc::umma_bf16(tmem + 256, ad, bd, idesc_o, (j | kk) ? 1u : 0u);
        }
        tc::umma_commit(&ring_empty[slot]);
        tc::umma_commit(&p_free[j & 1]);
        ++idx;
      }
      tc::umma_commit(&o_full);
    }
  } else {
    // 8 softmax warps: warp w owns TMEM lanes 32*(w%4).. (= query rows) and the column half (w-2)/4 of every 128-key block
    const int i = (warp & 3) * 32 + lane;
    const int half = (warp - 2) >> 2;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float c2 = p.c2;
    auto ex2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
    float m = -INFINITY;
    // pass 1: row max only (the exponentials are taken once, in pass 2, relative to the final max)
    for (int c = 0; c < nblk; ++c) {
      tc::mbar_wait_trap(tc::smem_u32(&s_full[c & 1]), (c >> 1) & 1, kTimeout, p.status, 4);
      tc::tc_fence_after();
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        uint32_t r[32];
        tc::tmem_ld32(trow + (c & 1) * 128 + half * 64 + cg * 32, r);
        tc::tmem_ld_wait();
        const int key0 = c * 128 + half * 64 + cg * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, key0 + j < len ? __uint_as_float(r[j]) : -INFINITY);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_free[c & 1]);
    }
    s_stat[half][i] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");               // the 8 softmax warps only
    m = fmaxf(s_stat[0][i], s_stat[1][i]);                       // finite: key 0 of every utterance is valid
    const float mc = m * c2;
    float l = 0.f;
    // pass 2: unnormalised probabilities exp(s/sqrt(d) - max) <= 1 -> fp16 in shared memory (A operand of P V), row sums in fp32
    for (int j = 0; j < nblk; ++j) {
      const int c = nblk + j;
      tc::mbar_wait_trap(tc::smem_u32(&s_full[c & 1]), (c >> 1) & 1, kTimeout, p.status, 4);
      if (j >= 2) tc::mbar_wait_trap(tc::smem_u32(&p_free[j & 1]), ((j >> 1) - 1) & 1, kTimeout, p.status, 4);
      tc::tc_fence_after();
      uint8_t* pb = smem + ATT_TILE_BYTES * (1 + ATT_RING + (j & 1));
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        uint32_t r[32];
        tc::tmem_ld32(trow + (c & 1) * 128 + half * 64 + cg * 32, r);
        tc::tmem_ld_wait();
        const int key0 = j * 128 + half * 64 + cg * 32;
        float pr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          pr[e] = key0 + e < len ? ex2(fmaf(__uint_as_float(r[e]), c2, -mc)) : 0.f;
          l += pr[e];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __half2 h0 = __floats2half2_rn(pr[q * 8], pr[q * 8 + 1]), h1 = __floats2half2_rn(pr[q * 8 + 2], pr[q * 8 + 3]);
          __half2 h2 = __floats2half2_rn(pr[q * 8 + 4], pr[q * 8 + 5]), h3 = __floats2half2_rn(pr[q * 8 + 6], pr[q * 8 + 7]);
          *reinterpret_cast<uint4*>(pb + (half * 8 + cg * 4 + q) * 2048 + i * 16) =
              make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                         *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
      }
      tc::fence_proxy_async_smem();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(&p_full[j & 1]); tc::mbar_arrive(&s_free[c & 1]); }
    }
    s_stat[half][i] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv_l = 1.f / (s_stat[0][i] + s_stat[1][i]);
    // O / l -> fp16 image, channels h*128 + half*64 ...
    tc::mbar_wait_trap(tc::smem_u32(&o_full), 0, kTimeout, p.status, 4);
    tc::tc_fence_after();
    const int tq = qt * 128 + i;
    const size_t orow = (size_t)row_u + tq;
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
      uint32_t r[32];
      tc::tmem_ld32(trow + 256 + half * 64 + cg * 32, r);
      tc::tmem_ld_wait();
      if (tq < p.T) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __half2 h0 = __floats2half2_rn(__uint_as_float(r[q * 8]) * inv_l, __uint_as_float(r[q * 8 + 1]) * inv_l);
          __half2 h1 = __floats2half2_rn(__uint_as_float(r[q * 8 + 2]) * inv_l, __uint_as_float(r[q * 8 + 3]) * inv_l);
          __half2 h2 = __floats2half2_rn(__uint_as_float(r[q * 8 + 4]) * inv_l, __uint_as_float(r[q * 8 + 5]) * inv_l);
          __half2 h3 = __floats2half2_rn(__uint_as_float(r[q * 8 + 6]) * inv_l, __uint_as_float(r[q * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(p.out + ((size_t)(h * (DK / 8) + half * 8 + cg * 4 + q) * p.Rp + orow) * 8) =
              make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                         *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

int run_attention(const __half* qkv, int qkv_chunks, const Rows& r, int n_head, const int* lens, __half* out, int* status,
                  cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.pending()) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    once.done();
  }
  CUtensorMap m;
  if (int rc = make_image_map(&m, qkv, qkv_chunks, r.Rp, DK / 8)) return rc;
  AttArgs a{};
  a.lens = lens; a.B = r.B; a.T = r.T; a.Tg = r.Tg; a.Rp = r.Rp; a.out = out;
  a.q_chunk0 = 0; a.k_chunk0 = n_head * (DK / 8); a.v_chunk0 = 2 * n_head * (DK / 8);
  a.c2 = 1.4426950408889634f / sqrtf((float)DK);
  a.status = status;
  MGB_CUDA_CHECK(launch_pdl(attn_kernel, dim3((r.T + 127) / 128, n_head, r.B), dim3(320), ATT_SMEM, s, 1, a, m));
  note_launch();
  return MGB_OK;
}

// =====================================================================================================
// Plan: layers, flat-parameter offsets (include/mixgan_b200.h order), packed-buffer offsets, workspace
// =====================================================================================================
struct FftLayer {
  Layer qkv, fc, w1, w2;
  size_t f_wq, f_bq, f_wk, f_bk, f_wv, f_bv, f_ln1g, f_ln1b, f_wfc, f_bfc, f_w1, f_b1, f_w2, f_b2, f_ln2g, f_ln2b;   // flat
  size_t p_ln1g, p_ln1b, p_ln2g, p_ln2b;          // packed fp32 (float index)
};
struct PostLayer { Layer conv; size_t f_w, f_b, f_g, f_be, f_mean, f_var; size_t p_scale, p_shift; };
struct AuxPlan {
  std::vector<FftLayer> fft;
  Layer mel; size_t f_melw, f_melb;
  std::vector<PostLayer> post;
  size_t flat_total, packed_bytes;
};

bool aux_dims_ok(const mgb_auxdec_dims* d) {
  return d && d->d_model == 256 && d->n_head == 2 && d->n_mel > 0 && d->n_mel <= 128 && d->n_mel % 8 == 0 &&
         d->d_inner > 0 && d->d_inner % 64 == 0 && d->ffn_kernel % 2 == 1 && d->ffn_kernel <= 2 * GAP + 1 &&
         d->layers >= 1 && d->layers <= 32 && d->postnet_layers >= 2 && d->postnet_layers <= 16 &&
         d->postnet_dim > 0 && d->postnet_dim % 64 == 0 && d->postnet_kernel % 2 == 1 && d->postnet_kernel <= 2 * GAP + 1;
}

AuxPlan make_plan(const mgb_auxdec_dims& d) {
  AuxPlan pl;
  const size_t D = d.d_model, H = d.d_inner, M = d.n_mel, P = d.postnet_dim;
  size_t f = 0, pb = 0;      // flat floats, packed bytes
  auto takef = [&](size_t n) { size_t r = f; f += n; return r; };
  auto place = [&](Layer& l) {
    pb = align_up(pb, 128); l.w_off = pb / 2; pb += l.w_halves() * 2;
    pb = align_up(pb, 16); l.b_off = pb / 4; pb += l.b_floats() * 4;
  };
  auto takep = [&](size_t n) { pb = align_up(pb, 16); size_t r = pb / 4; pb += n * 4; return r; };
  for (int i = 0; i < d.layers; ++i) {
    FftLayer L{};
    L.qkv = plan_layer((int)D, 3 * (int)D, 1, 1); L.fc = plan_layer((int)D, (int)D, 1, 1);
    L.w1 = plan_layer((int)D, (int)H, d.ffn_kernel, 1); L.w2 = plan_layer((int)H, (int)D, 1, 1);
    L.f_wq = takef(D * D); L.f_bq = takef(D); L.f_wk = takef(D * D); L.f_bk = takef(D); L.f_wv = takef(D * D); L.f_bv = takef(D);
    L.f_ln1g = takef(D); L.f_ln1b = takef(D); L.f_wfc = takef(D * D); L.f_bfc = takef(D);
    L.f_w1 = takef(H * D * d.ffn_kernel); L.f_b1 = takef(H); L.f_w2 = takef(D * H); L.f_b2 = takef(D);
    L.f_ln2g = takef(D); L.f_ln2b = takef(D);
    place(L.qkv); place(L.fc); place(L.w1); place(L.w2);
    L.p_ln1g = takep(D); L.p_ln1b = takep(D); L.p_ln2g = takep(D); L.p_ln2b = takep(D);
    pl.fft.push_back(L);
  }
  pl.mel = plan_layer((int)D, (int)M, 1, 1);
  pl.f_melw = takef(M * D); pl.f_melb = takef(M);
  place(pl.mel);
  for (int i = 0; i < d.postnet_layers; ++i) {
    PostLayer L{};
    const int cin = i == 0 ? (int)M : (int)P, cout = i == d.postnet_layers - 1 ? (int)M : (int)P;
    L.conv = plan_layer(cin, cout, d.postnet_kernel, 1);
    L.f_w = takef((size_t)cout * cin * d.postnet_kernel); L.f_b = takef(cout);
    L.f_g = takef(cout); L.f_be = takef(cout); L.f_mean = takef(cout); L.f_var = takef(cout);
    place(L.conv);
    L.p_scale = takep(cout); L.p_shift = takep(cout);
    pl.post.push_back(L);
  }
  pl.flat_total = f;
  pl.packed_bytes = align_up(pb, 256);
  return pl;
}

// BatchNorm1d (eval) after a convolution: y = (conv + b - mean) * g / sqrt(var + eps) + beta
//   => scale = g / sqrt(var + eps) folded into the fp16 weights, bias' = b * scale + (beta - mean * scale)
__global__ void bn_fold_kernel(const float* g, const float* be, const float* mean, const float* var, float* scale, float* shift,
                               int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sc = g[i] / sqrtf(var[i] + 1e-5f);
  scale[i] = sc;
  shift[i] = be[i] - mean[i] * sc;
}

struct AuxWs {
  size_t xs[2], xi[2], qkv, att, hid, mels, meli, pi[2], status, total;
};
AuxWs aux_ws(const mgb_auxdec_dims& d, const Rows& r) {
  AuxWs w{};
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t o = p; p += align_up(bytes, 1024); return o; };
  const size_t Rp = r.Rp, D = d.d_model, M32 = (d.n_mel + 31) / 32 * 32;
  w.status = take(8192);
  for (int i = 0; i < 2; ++i) { w.xs[i] = take(Rp * D * 4); w.xi[i] = take(Rp * D * 2); }
  w.qkv = take(Rp * 3 * D * 2);
  w.att = take(Rp * D * 2);
  w.hid = take(Rp * (size_t)d.d_inner * 2);
  w.mels = take(Rp * M32 * 4);
  w.meli = take(Rp * M32 * 2);
  for (int i = 0; i < 2; ++i) w.pi[i] = take(Rp * (size_t)d.postnet_dim * 2);
  w.total = p;
  return w;
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" {

size_t mgb_auxdec_flat_count(const mgb_auxdec_dims* dims) { return aux_dims_ok(dims) ? make_plan(*dims).flat_total : 0; }
size_t mgb_auxdec_packed_bytes(const mgb_auxdec_dims* dims) { return aux_dims_ok(dims) ? make_plan(*dims).packed_bytes : 0; }
size_t mgb_auxdec_workspace_bytes(const mgb_auxdec_dims* dims, int B, int T) {
  if (!aux_dims_ok(dims) || B <= 0 || T <= 0) return 0;
  return aux_ws(*dims, make_rows(B, T, GAP)).total;
}

int mgb_auxdec_pack(const mgb_auxdec_dims* dims, const float* flat, void* packed, size_t packed_bytes, void* stream) {
  MGB_REQUIRE(aux_dims_ok(dims), MGB_E_UNSUPPORTED,
              "aux decoder dims unsupported (d_model 256, 2 heads, kernels <= 9, widths multiples of 64)");
  MGB_REQUIRE(flat && packed, MGB_E_ARG, "NULL pointer argument");
  const AuxPlan pl = make_plan(*dims);
  MGB_REQUIRE(packed_bytes >= pl.packed_bytes, MGB_E_WORKSPACE, "packed buffer too small");
  if (int rc = check_arch()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* pf = static_cast<float*>(packed);
  const int D = dims->d_model;
  for (const FftLayer& L : pl.fft) {
    if (int rc = pack_conv(L.qkv, packed, flat + L.f_wq, flat + L.f_bq, nullptr, nullptr, 0, D, s)) return rc;
    if (int rc = pack_conv(L.qkv, packed, flat + L.f_wk, flat + L.f_bk, nullptr, nullptr, 1, D, s)) return rc;
    if (int rc = pack_conv(L.qkv, packed, flat + L.f_wv, flat + L.f_bv, nullptr, nullptr, 2, D, s)) return rc;
    if (int rc = pack_conv(L.fc, packed, flat + L.f_wfc, flat + L.f_bfc, nullptr, nullptr, 0, D, s)) return rc;
    if (int rc = pack_conv(L.w1, packed, flat + L.f_w1, flat + L.f_b1, nullptr, nullptr, 0, dims->d_inner, s)) return rc;
    if (int rc = pack_conv(L.w2, packed, flat + L.f_w2, flat + L.f_b2, nullptr, nullptr, 0, D, s)) return rc;
    MGB_CUDA_CHECK(cudaMemcpyAsync(pf + L.p_ln1g, flat + L.f_ln1g, D * 4, cudaMemcpyDeviceToDevice, s));
    MGB_CUDA_CHECK(cudaMemcpyAsync(pf + L.p_ln1b, flat + L.f_ln1b, D * 4, cudaMemcpyDeviceToDevice, s));
    MGB_CUDA_CHECK(cudaMemcpyAsync(pf + L.p_ln2g, flat + L.f_ln2g, D * 4, cudaMemcpyDeviceToDevice, s));
    MGB_CUDA_CHECK(cudaMemcpyAsync(pf + L.p_ln2b, flat + L.f_ln2b, D * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (int rc = pack_conv(pl.mel, packed, flat + pl.f_melw, flat + pl.f_melb, nullptr, nullptr, 0, dims->n_mel, s)) return rc;
  for (const PostLayer& L : pl.post) {
    const int n = L.conv.Cout;
    bn_fold_kernel<<<(n + 255) / 256, 256, 0, s>>>(flat + L.f_g, flat + L.f_be, flat + L.f_mean, flat + L.f_var,
                                                   pf + L.p_scale, pf + L.p_shift, n);
    note_launch();
    if (int rc = pack_conv(L.conv, packed, flat + L.f_w, flat + L.f_b, pf + L.p_scale, pf + L.p_shift, 0, n, s)) return rc;
  }
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_auxdec_forward(const mgb_auxdec_dims* dims, const void* packed, const float* x, const float* pos, const int32_t* lens,
                       float* coarse, float* dec_out, float* mel_before, int B, int T, void* workspace, size_t workspace_bytes,
                       void* stream) {
  MGB_REQUIRE(aux_dims_ok(dims), MGB_E_UNSUPPORTED,
              "aux decoder dims unsupported (d_model 256, 2 heads, kernels <= 9, widths multiples of 64)");
  MGB_REQUIRE(packed && x && pos && coarse && workspace, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0 && (long long)B * (T + GAP) < (1LL << 30), MGB_E_ARG, "bad shape");
  const Rows r = make_rows(B, T, GAP);
  const AuxWs w = aux_ws(*dims, r);
  MGB_REQUIRE(workspace_bytes >= w.total, MGB_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total);
  if (int rc = check_arch()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const AuxPlan pl = make_plan(*dims);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const float* pf = static_cast<const float*>(packed);
  int* status = reinterpret_cast<int*>(ws + w.status);
  float* xs[2] = {reinterpret_cast<float*>(ws + w.xs[0]), reinterpret_cast<float*>(ws + w.xs[1])};
  __half* xi[2] = {reinterpret_cast<__half*>(ws + w.xi[0]), reinterpret_cast<__half*>(ws + w.xi[1])};
  __half* qkv = reinterpret_cast<__half*>(ws + w.qkv);
  __half* att = reinterpret_cast<__half*>(ws + w.att);
  __half* hid = reinterpret_cast<__half*>(ws + w.hid);
  float* mels = reinterpret_cast<float*>(ws + w.mels);
  __half* meli = reinterpret_cast<__half*>(ws + w.meli);
  __half* pi[2] = {reinterpret_cast<__half*>(ws + w.pi[0]), reinterpret_cast<__half*>(ws + w.pi[1])};
  const int D = dims->d_model, DC = D / 8;
  const int* lens_i = reinterpret_cast<const int*>(lens);

  MGB_CUDA_CHECK(cudaMemsetAsync(status, 0, 8192, s));
  // dec_output = enc_seq + position_enc[:T]   (Models.py:155-157)
  if (int rc = pack_rows(x, pos, D, r, xi[0], 1.f, xs[0], s)) return rc;
  for (size_t li = 0; li < pl.fft.size(); ++li) {
    const FftLayer& L = pl.fft[li];
    {  // q, k, v projections (SubLayers.py:40-42): one N = 768 GEMM
      ConvIO io = conv_io(xi[0], DC);
      io.img_out = qkv;
      if (int rc = run_conv(L.qkv, packed, r, io, status, s)) return rc;
    }
    if (int rc = run_attention(qkv, 3 * DC, r, dims->n_head, lens_i, att, status, s)) return rc;
    {  // fc, + residual, LayerNorm, masked_fill (SubLayers.py:55-56, Layers.py:27)
      ConvIO io = conv_io(att, DC);
      io.res1 = xs[0]; io.ln_g = pf + L.p_ln1g; io.ln_b = pf + L.p_ln1b; io.lens = lens_i;
      io.stream_out = xs[1]; io.img_out = xi[1];
      if (int rc = run_conv(L.fc, packed, r, io, status, s)) return rc;
    }
    {  // w_1 (k = 9) + ReLU (SubLayers.py:91)
      ConvIO io = conv_io(xi[1], DC);
      io.act = ACT_RELU; io.img_out = hid;
      if (int rc = run_conv(L.w1, packed, r, io, status, s)) return rc;
    }
    {  // w_2, + residual, LayerNorm, masked_fill (SubLayers.py:91-95, Layers.py:30)
      ConvIO io = conv_io(hid, dims->d_inner / 8);
      io.res1 = xs[1]; io.ln_g = pf + L.p_ln2g; io.ln_b = pf + L.p_ln2b; io.lens = lens_i;
      io.stream_out = xs[0]; io.img_out = xi[0];
      if (li + 1 == pl.fft.size() && dec_out) { io.user_out = dec_out; io.user_ld = D; }
      if (int rc = run_conv(L.w2, packed, r, io, status, s)) return rc;
    }
  }
  const int MC = (dims->n_mel + 31) / 32 * 4;       // chunks of the (32-rounded) mel images
  {  // mel_linear (mixgantts.py:140)
    ConvIO io = conv_io(xi[0], DC);
    io.stream_out = mels; io.img_out = meli;
    if (mel_before) { io.user_out = mel_before; io.user_ld = dims->n_mel; }
    if (int rc = run_conv(pl.mel, packed, r, io, status, s)) return rc;
  }
  const __half* cur = meli;
  int cur_chunks = MC;
  for (size_t i = 0; i < pl.post.size(); ++i) {     // PostNet (Layers.py:128-137) + residual (mixgantts.py:141)
    ConvIO io = conv_io(cur, cur_chunks);
    if (i + 1 < pl.post.size()) {
      io.act = ACT_TANH; io.img_out = pi[i & 1];
    } else {
      io.res1 = mels; io.user_out = coarse; io.user_ld = dims->n_mel;
    }
    if (int rc = run_conv(pl.post[i].conv, packed, r, io, status, s)) return rc;
    cur = pi[i & 1]; cur_chunks = dims->postnet_dim / 8;
  }
  return MGB_OK;
}

int mgb_auxdec_debug_status(const mgb_auxdec_dims* dims, int B, int T, const void* workspace, int* host_status) {
  MGB_REQUIRE(aux_dims_ok(dims) && workspace && host_status, MGB_E_ARG, "bad argument");
  const AuxWs w = aux_ws(*dims, make_rows(B, T, GAP));
  MGB_CUDA_CHECK(cudaMemcpy(host_status, static_cast<const uint8_t*>(workspace) + w.status, sizeof(int), cudaMemcpyDeviceToHost));
  return MGB_OK;
}

}  // extern "C"

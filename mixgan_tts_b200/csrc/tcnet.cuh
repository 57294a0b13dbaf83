// Generic tensor-core Conv1d / Linear engine for the stages either side of the diffusion decoder (SURVEY.md 8(f) ranks 2
// and 4): the aux decoder (transformer/Models.py:103-171 Decoder, transformer/Layers.py:11-31 FFTBlock, :67-137 PostNet,
// mel_linear model/mixgantts.py:139-143) and the HiFi-GAN generator (hifigan/models.py:112-173).
//
// Activations are fp16 "images" [C/8][Rp][8] on a row axis that holds every utterance (utterance b owns rows
// b*(T+G) + [0,T); the G rows between utterances are zero = the convolutions' zero padding, and the tensor-map TMA
// zero-fills rows in front of / behind the axis), so a convolution is an implicit GEMM whose taps are the SAME image
// box shifted by (j - k/2) * dilation rows — no im2col, no per-utterance tile quantisation.  Optional fp32 "streams"
// [C/4][Rp][4] carry the residual paths (LayerNorm inputs, HiFi-GAN residual sums) at full precision.
//
//   D[128 rows][NT] = sum over taps, Cin-steps of A[128][KC] * W[NT][KC]^T      tcgen05.mma kind::f16 (fp16 operands,
//   fp32 accumulation in TMEM), A by one tensor-map TMA box per step, W by one bulk copy per step from tiles packed
//   once per load_state_dict in exactly the streamed order.  192 threads: TMA warp, MMA warp, 4 epilogue warps;
//   two CTAs per SM so one tile's epilogue overlaps the other's main loop.
// Fused epilogue: + bias -> relu / tanh -> + up to two fp32 residual streams -> * scale -> [LayerNorm over the 256
// channels of the row] -> row mask -> fp32 stream and/or fp16 image (with the NEXT layer's leaky_relu already applied)
// and/or the caller's [B][T][C] tensor; a ConvTranspose1d (stride s) is the same GEMM with N = s*Cout whose column
// tile p is output phase p, written to row s*r + p.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace mgb {
namespace tcnet {

constexpr int TILE = 128;

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct Rows {            // the row axis of one time resolution
  int B, T, Tg;          // utterances, frames per utterance, T + gap
  int ntiles, Rp;        // 128-row tiles covering B*Tg rows; Rp = ntiles * 128 rows allocated per image chunk
};
inline Rows make_rows(int B, int T, int gap) {
  Rows r{};
  r.B = B; r.T = T; r.Tg = T + gap;
  r.ntiles = (int)(((long long)B * r.Tg + TILE - 1) / TILE);
  r.Rp = r.ntiles * TILE;
  return r;
}
inline Rows upsampled(const Rows& r, int up) {
  Rows o = r;
  o.T = r.T * up; o.Tg = r.Tg * up; o.ntiles = r.ntiles * up; o.Rp = r.Rp * up;
  return o;
}

// One convolution / linear layer as the GEMM sees it
struct Layer {
  int Cin, Cout, k, dil;       // Cout = channels of the OUTPUT tensor (per phase for a transposed convolution)
  int up, tpad;                // transposed convolution: stride and padding (up == 1: ordinary convolution)
  int NT, KC, taps, kspt, ntn; // column tile, channels per k-step, GEMM taps, k-steps per tap, column tiles
  int split;                   // 1 = split precision: operands as fp16 hi + lo pairs, three MMAs per product
                               //     (a_hi w_hi + a_hi w_lo + a_lo w_hi: ~2^-21 relative instead of 2^-11)
  size_t w_off, b_off;         // offsets into the packed buffer: fp16 tiles (in halves), fp32 bias (in floats)
  size_t w_halves() const { return (size_t)ntn * taps * kspt * NT * KC * (split ? 2 : 1); }
  size_t b_floats() const { return (size_t)ntn * NT; }
};
inline Layer plan_layer(int Cin, int Cout, int k, int dil, int up = 1, int tpad = 0, int split = 0) {
  Layer l{};
  l.Cin = Cin; l.Cout = Cout; l.k = k; l.dil = dil; l.up = up; l.tpad = tpad; l.split = split;
  if (up > 1) { l.NT = Cout; l.ntn = up; l.taps = 3; l.dil = 1; }
  else {
    l.NT = Cout > 128 ? 256 : Cout > 64 ? 128 : Cout > 32 ? 64 : 32;
    l.ntn = (Cout + l.NT - 1) / l.NT;
    l.taps = k;
  }
  l.KC = (Cin % 64 == 0) ? 64 : 32;
  l.kspt = (Cin + l.KC - 1) / l.KC;
  return l;
}

struct ConvIO {                      // per-launch operands of run_conv
  const __half* in; int in_chunks;   // input image and its channel chunks (Cin rounded up to 8)
  const __half* in_lo;               // split layers: the low-order image of the input (same geometry)
  __half* img_lo_out;                // low-order companion of img_out (for a split consumer), or NULL
  int act; float scale;
  const float* res1; const float* res2;        // fp32 streams (output row space) added after the activation
  const float* ln_g; const float* ln_b;        // LayerNorm over the row (NT == Cout == 256 only)
  const int* lens; int len_mul;                // rows t >= lens[b] * len_mul are written as zeros (masked_fill)
  float* stream_out;
  __half* img_out; float img_slope;            // image = leaky_relu(v, img_slope); 1.0 = identity
  float* user_out; int user_ld;                // [B][T*up][user_ld]
};
inline ConvIO conv_io(const __half* in, int in_chunks) {
  ConvIO io{};
  io.in = in; io.in_chunks = in_chunks; io.act = ACT_NONE; io.scale = 1.f; io.len_mul = 1; io.img_slope = 1.f;
  return io;
}

// user [B][T][C] fp32 (+ pos [T][C]) -> fp16 image (leaky_relu slope applied; 1 = identity) and/or fp32 stream, every
// one of the Rp rows written (zeros outside the utterances).  C % 8 == 0.
int pack_rows(const float* user, const float* pos, int C, const Rows& r, __half* img, float img_slope, float* stream,
              cudaStream_t s, __half* img_lo = nullptr);
int run_conv(const Layer& l, const void* packed, const Rows& rin, const ConvIO& io, int* status, cudaStream_t s);
int pack_conv(const Layer& l, void* packed, const float* w, const float* bias, const float* oscale, const float* oshift,
              int ntile0, int Cout_part, cudaStream_t s);

}  // namespace tcnet
}  // namespace mgb

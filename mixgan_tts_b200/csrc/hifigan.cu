// HiFi-GAN generator, the vocoder after the diffusion decoder.  Reference:
//   Generator.forward  hifigan/models.py:151-166   conv_pre -> 4 x [leaky_relu(0.1) -> ConvTranspose1d -> mean of 3 ResBlocks]
//                                                  -> leaky_relu(0.01) -> conv_post -> tanh
//   ResBlock.forward   hifigan/models.py:96-103    3 x [x += c2(leaky_relu(c1(leaky_relu(x))))], c1 dilated (1, 3, 5)
//   vocoder_infer      utils/model.py:103-121      wavs = vocoder(mels).squeeze(1)   (mels [B][n_mel][T], weight norm removed)
// Inference only.  Every convolution is a tcgen05 implicit GEMM of tcnet.cuh; a transposed convolution with stride s is
// the GEMM with N = s * Cout whose column tile p is output phase p.  The residual sums stay fp32 (streams), the
// convolution operands are fp16 images that already carry the next layer's leaky_relu.  109 launches per call.
#include "tcnet.cuh"

#include <vector>

namespace mgb {
namespace {

using namespace tcnet;

constexpr int GAP0 = 4;        // zero rows between utterances at the mel rate (conv_pre / conv_post padding is 3)
constexpr int MAX_UP = 8, MAX_RES = 4;

struct ResPlan { Layer c1[3], c2[3]; size_t f_c1w[3], f_c1b[3], f_c2w[3], f_c2b[3]; };
struct StagePlan { Layer up; size_t f_upw, f_upb; int ch; std::vector<ResPlan> res; };
struct VocPlan {
  Layer pre, post; size_t f_prew, f_preb, f_postw, f_postb;
  std::vector<StagePlan> st;
  size_t flat_total, packed_bytes;
  int total_up;
};

bool voc_dims_ok(const mgb_hifigan_dims* d, const char** why) {
  static const char* msg = "";
  if (why) *why = msg;
  auto fail = [&](const char* m) { if (why) *why = m; return false; };
  if (!d) return fail("NULL dims");
  if (d->n_mel <= 0 || d->n_mel % 8 || d->n_mel > 128) return fail("n_mel must be a multiple of 8 up to 128");
  if (d->n_up < 1 || d->n_up > MAX_UP || d->n_res < 1 || d->n_res > MAX_RES) return fail("1-8 upsampling stages, 1-4 resblocks");
  if (d->split_mode < 0 || d->split_mode > 2) return fail("split_mode must be 0, 1 or 2");
  if (d->initial_channel % 64) return fail("initial channel count must be a multiple of 64");
  long long gap = GAP0;
  for (int i = 0; i < d->n_up; ++i) {
    const int u = d->up_rates[i], k = d->up_kernels[i], ch = d->initial_channel >> (i + 1);
    if (ch != 32 && ch != 64 && ch != 128 && ch != 256) return fail("stage channels must be 32, 64, 128 or 256");
    if (u < 2 || k < u || (k - u) % 2) return fail("upsampling needs kernel >= stride >= 2 and an even difference");
    const int pad = (k - u) / 2;
    // input offsets d' used by some phase: j = p + pad - u * d' in [0, k)  ->  must lie in {-1, 0, 1}
    if ((u - 1 + pad) / u > 1 || (k - 1 - pad) / u > 1) return fail("transposed kernel reaches beyond one neighbour frame");
    gap *= u;
    for (int j = 0; j < d->n_res; ++j) {
      if (d->res_kernels[j] % 2 == 0) return fail("resblock kernels must be odd");
      for (int m = 0; m < 3; ++m)
        if ((long long)(d->res_kernels[j] - 1) / 2 * d->res_dilations[j][m] > gap) return fail("resblock padding exceeds the row gap");
    }
  }
  return true;
}

VocPlan make_plan(const mgb_hifigan_dims& d) {
  VocPlan pl;
  size_t f = 0, pb = 0;
  auto takef = [&](size_t n) { size_t r = f; f += n; return r; };
  auto place = [&](Layer& l) {
    pb = align_up(pb, 128); l.w_off = pb / 2; pb += l.w_halves() * 2;
    pb = align_up(pb, 16); l.b_off = pb / 4; pb += l.b_floats() * 4;
  };
  const int C0 = d.initial_channel;
  // split_mode 1: the layers whose inputs have the widest dynamic range (conv_pre on the log-mel, the transposed
  // convolutions on the averaged resblock sums, conv_post) run with hi + lo operand pairs: < 5 % of the FLOPs, ~40 % of the
  // fp16 operand-rounding error.  split_mode 2: every layer (parity mode, 3 MMAs per product).
  const int sp_edge = d.split_mode >= 1 ? 1 : 0, sp_all = d.split_mode >= 2 ? 1 : 0;
  pl.pre = plan_layer(d.n_mel, C0, 7, 1, 1, 0, sp_edge);
  pl.f_prew = takef((size_t)C0 * d.n_mel * 7); pl.f_preb = takef(C0);
  place(pl.pre);
  pl.total_up = 1;
  pl.st.resize(d.n_up);
  for (int i = 0; i < d.n_up; ++i) {             // state_dict order: conv_pre, ups.*, resblocks.*, conv_post
    StagePlan& S = pl.st[i];
    const int cin = C0 >> i, ch = C0 >> (i + 1), u = d.up_rates[i], k = d.up_kernels[i];
    S.ch = ch;
    S.up = plan_layer(cin, ch, k, 1, u, (k - u) / 2, sp_edge);
    S.f_upw = takef((size_t)cin * ch * k); S.f_upb = takef(ch);
    place(S.up);
    pl.total_up *= u;
  }
  for (int i = 0; i < d.n_up; ++i) {
    StagePlan& S = pl.st[i];
    S.res.resize(d.n_res);
    for (int j = 0; j < d.n_res; ++j) {
      ResPlan& R = S.res[j];
      const int k = d.res_kernels[j];
      for (int m = 0; m < 3; ++m) { R.c1[m] = plan_layer(S.ch, S.ch, k, d.res_dilations[j][m], 1, 0, sp_all); R.f_c1w[m] = takef((size_t)S.ch * S.ch * k); R.f_c1b[m] = takef(S.ch); place(R.c1[m]); }
      for (int m = 0; m < 3; ++m) { R.c2[m] = plan_layer(S.ch, S.ch, k, 1, 1, 0, sp_all); R.f_c2w[m] = takef((size_t)S.ch * S.ch * k); R.f_c2b[m] = takef(S.ch); place(R.c2[m]); }
    }
  }
  const int chl = C0 >> d.n_up;
  pl.post = plan_layer(chl, 1, 7, 1, 1, 0, sp_edge);
  pl.f_postw = takef((size_t)chl * 7); pl.f_postb = takef(1);
  place(pl.post);
  pl.flat_total = f;
  pl.packed_bytes = align_up(pb, 256);
  return pl;
}

struct VocWs { size_t status, meli, prei, meli_lo, prei_lo, s32[4], h16[4], lo16[4], total; size_t E; };
VocWs voc_ws(const mgb_hifigan_dims& d, const Rows& r0) {
  VocWs w{};
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t o = p; p += align_up(bytes, 1024); return o; };
  size_t E = 0;
  long long Rp = r0.Rp;
  for (int i = 0; i < d.n_up; ++i) {
    Rp *= d.up_rates[i];
    const size_t e = (size_t)Rp * (size_t)(d.initial_channel >> (i + 1));
    if (e > E) E = e;
  }
  w.E = E;
  w.status = take(8192);
  w.meli = take((size_t)r0.Rp * d.n_mel * 2);
  w.prei = take((size_t)r0.Rp * d.initial_channel * 2);
  for (int i = 0; i < 4; ++i) w.s32[i] = take(E * 4);
  for (int i = 0; i < 4; ++i) w.h16[i] = take(E * 2);
  if (d.split_mode >= 1) {              // low-order companions of the images a split layer reads
    w.meli_lo = take((size_t)r0.Rp * d.n_mel * 2);
    w.prei_lo = take((size_t)r0.Rp * d.initial_channel * 2);
    w.lo16[2] = take(E * 2);            // of I[0], the stage output
    if (d.split_mode >= 2) { w.lo16[0] = take(E * 2); w.lo16[1] = take(E * 2); w.lo16[3] = take(E * 2); }
  }
  w.total = p;
  return w;
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" {

size_t mgb_hifigan_flat_count(const mgb_hifigan_dims* dims) { return voc_dims_ok(dims, nullptr) ? make_plan(*dims).flat_total : 0; }
size_t mgb_hifigan_packed_bytes(const mgb_hifigan_dims* dims) { return voc_dims_ok(dims, nullptr) ? make_plan(*dims).packed_bytes : 0; }
int mgb_hifigan_hop(const mgb_hifigan_dims* dims) { return voc_dims_ok(dims, nullptr) ? make_plan(*dims).total_up : 0; }
size_t mgb_hifigan_workspace_bytes(const mgb_hifigan_dims* dims, int B, int T) {
  if (!voc_dims_ok(dims, nullptr) || B <= 0 || T <= 0) return 0;
  return voc_ws(*dims, make_rows(B, T, GAP0)).total;
}

int mgb_hifigan_pack(const mgb_hifigan_dims* dims, const float* flat, void* packed, size_t packed_bytes, void* stream) {
  const char* why = "";
  MGB_REQUIRE(voc_dims_ok(dims, &why), MGB_E_UNSUPPORTED, "HiFi-GAN dims unsupported: %s", why);
  MGB_REQUIRE(flat && packed, MGB_E_ARG, "NULL pointer argument");
  const VocPlan pl = make_plan(*dims);
  MGB_REQUIRE(packed_bytes >= pl.packed_bytes, MGB_E_WORKSPACE, "packed buffer too small");
  if (int rc = check_arch()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = pack_conv(pl.pre, packed, flat + pl.f_prew, flat + pl.f_preb, nullptr, nullptr, 0, pl.pre.Cout, s)) return rc;
  for (const StagePlan& S : pl.st) {
    if (int rc = pack_conv(S.up, packed, flat + S.f_upw, flat + S.f_upb, nullptr, nullptr, 0, S.ch, s)) return rc;
    for (const ResPlan& R : S.res)
      for (int m = 0; m < 3; ++m) {
        if (int rc = pack_conv(R.c1[m], packed, flat + R.f_c1w[m], flat + R.f_c1b[m], nullptr, nullptr, 0, S.ch, s)) return rc;
        if (int rc = pack_conv(R.c2[m], packed, flat + R.f_c2w[m], flat + R.f_c2b[m], nullptr, nullptr, 0, S.ch, s)) return rc;
      }
  }
  return pack_conv(pl.post, packed, flat + pl.f_postw, flat + pl.f_postb, nullptr, nullptr, 0, 1, s);
}

/* mel [B][T][n_mel] (frames-major, the layout the diffusion decoder returns) -> wav [B][T * hop] */
int mgb_hifigan_forward(const mgb_hifigan_dims* dims, const void* packed, const float* mel, float* wav, int B, int T,
                        void* workspace, size_t workspace_bytes, void* stream) {
  const char* why = "";
  MGB_REQUIRE(voc_dims_ok(dims, &why), MGB_E_UNSUPPORTED, "HiFi-GAN dims unsupported: %s", why);
  MGB_REQUIRE(packed && mel && wav && workspace, MGB_E_ARG, "NULL pointer argument");
  const VocPlan pl = make_plan(*dims);
  MGB_REQUIRE(B > 0 && T > 0 && (long long)B * (T + GAP0) * pl.total_up < (1LL << 30), MGB_E_ARG,
              "bad shape (B * (T + 4) * hop must stay below 2^30 rows)");
  Rows r = make_rows(B, T, GAP0);
  const VocWs w = voc_ws(*dims, r);
  MGB_REQUIRE(workspace_bytes >= w.total, MGB_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total);
  if (int rc = check_arch()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* status = reinterpret_cast<int*>(ws + w.status);
  __half* meli = reinterpret_cast<__half*>(ws + w.meli);
  __half* prei = reinterpret_cast<__half*>(ws + w.prei);
  float* X = reinterpret_cast<float*>(ws + w.s32[0]);
  float* Y[2] = {reinterpret_cast<float*>(ws + w.s32[1]), reinterpret_cast<float*>(ws + w.s32[2])};
  float* XS = reinterpret_cast<float*>(ws + w.s32[3]);
  __half* A0 = reinterpret_cast<__half*>(ws + w.h16[0]);
  __half* T1 = reinterpret_cast<__half*>(ws + w.h16[1]);
  __half* I[2] = {reinterpret_cast<__half*>(ws + w.h16[2]), reinterpret_cast<__half*>(ws + w.h16[3])};
  const bool sp_edge = dims->split_mode >= 1, sp_all = dims->split_mode >= 2;
  auto lo = [&](size_t off, bool on) { return on ? reinterpret_cast<__half*>(ws + off) : nullptr; };
  __half* meli_lo = lo(w.meli_lo, sp_edge);
  __half* prei_lo = lo(w.prei_lo, sp_edge);
  __half* A0_lo = lo(w.lo16[0], sp_all);
  __half* T1_lo = lo(w.lo16[1], sp_all);
  __half* I_lo[2] = {lo(w.lo16[2], sp_edge), lo(w.lo16[3], sp_all)};
  constexpr float SLOPE = 0.1f;                   // LRELU_SLOPE, models.py:7

  MGB_CUDA_CHECK(cudaMemsetAsync(status, 0, 8192, s));
  if (int rc = pack_rows(mel, nullptr, dims->n_mel, r, meli, 1.f, nullptr, s, meli_lo)) return rc;
  {  // conv_pre, image of leaky_relu(x) for ups[0]   (models.py:152, 154)
    ConvIO io = conv_io(meli, dims->n_mel / 8);
    io.in_lo = meli_lo;
    io.img_out = prei; io.img_slope = SLOPE; io.img_lo_out = prei_lo;
    if (int rc = run_conv(pl.pre, packed, r, io, status, s)) return rc;
  }
  const __half* prev = prei;
  const __half* prev_lo = prei_lo;
  int prev_chunks = dims->initial_channel / 8;
  for (size_t si = 0; si < pl.st.size(); ++si) {
    const StagePlan& S = pl.st[si];
    const int chunks = S.ch / 8;
    {  // x = ups[i](leaky_relu(x))   (models.py:154-155): fp32 x + image of leaky_relu(x) for the resblocks
      ConvIO io = conv_io(prev, prev_chunks);
      io.in_lo = prev_lo;
      io.stream_out = X; io.img_out = A0; io.img_slope = SLOPE; io.img_lo_out = A0_lo;
      if (int rc = run_conv(S.up, packed, r, io, status, s)) return rc;
    }
    r = upsampled(r, S.up.up);
    const int nres = (int)S.res.size();
    const bool last_stage = si + 1 == pl.st.size();
    for (int j = 0; j < nres; ++j) {
      const ResPlan& R = S.res[j];
      const float* cur_s = X;
      const __half* cur_i = A0;
      const __half* cur_lo = A0_lo;
      for (int m = 0; m < 3; ++m) {
        {  // xt = leaky_relu(c1(leaky_relu(x)))   (models.py:98-100)
          ConvIO io = conv_io(cur_i, chunks);
          io.in_lo = cur_lo;
          io.img_out = T1; io.img_slope = SLOPE; io.img_lo_out = T1_lo;
          if (int rc = run_conv(R.c1[m], packed, r, io, status, s)) return rc;
        }
        ConvIO io = conv_io(T1, chunks);   // x = c2(xt) + x   (models.py:101-102)
        io.in_lo = T1_lo;
        io.res1 = cur_s;
        if (m < 2) {
          io.stream_out = Y[m]; io.img_out = I[m]; io.img_slope = SLOPE; io.img_lo_out = sp_all ? I_lo[m] : nullptr;
          cur_s = Y[m]; cur_i = I[m]; cur_lo = sp_all ? I_lo[m] : nullptr;
        } else {
          // the resblock's output joins xs (models.py:157-161); after the last resblock x = xs / num_kernels and only the
          // image of leaky_relu(x) is needed: slope 0.1 in front of the next ups, 0.01 (F.leaky_relu default) in front of
          // conv_post (models.py:162-163)
          if (j > 0) io.res2 = XS;
          if (j + 1 < nres) io.stream_out = XS;
          else { io.scale = 1.f / (float)nres; io.img_out = I[0]; io.img_slope = last_stage ? 0.01f : SLOPE; io.img_lo_out = I_lo[0]; }
        }
        if (int rc = run_conv(R.c2[m], packed, r, io, status, s)) return rc;
      }
    }
    prev = I[0]; prev_lo = I_lo[0]; prev_chunks = chunks;
  }
  {  // tanh(conv_post(.))   (models.py:163-164) -> wav [B][T * hop]
    ConvIO io = conv_io(prev, prev_chunks);
    io.in_lo = prev_lo;
    io.act = ACT_TANH; io.user_out = wav; io.user_ld = 1;
    if (int rc = run_conv(pl.post, packed, r, io, status, s)) return rc;
  }
  return MGB_OK;
}

int mgb_hifigan_debug_status(const mgb_hifigan_dims* dims, int B, int T, const void* workspace, int* host_status) {
  MGB_REQUIRE(voc_dims_ok(dims, nullptr) && workspace && host_status, MGB_E_ARG, "bad argument");
  const VocWs w = voc_ws(*dims, make_rows(B, T, GAP0));
  MGB_CUDA_CHECK(cudaMemcpy(host_status, static_cast<const uint8_t*>(workspace) + w.status, sizeof(int), cudaMemcpyDeviceToHost));
  return MGB_OK;
}

}  // extern "C"

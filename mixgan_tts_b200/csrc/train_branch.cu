// The elementwise steps of GaussianDiffusion.forward's TRAINING branch (reference: model/diffusion.py:201-225) around the
// Denoiser call, as three kernels instead of ~56 torch launches per call:
//   mgb_train_diffuse            :206-207  x_t = diffuse_fn(mel, t) * valid, x_t_prev = diffuse_fn(mel, t - 1) * valid
//                                (diffuse_fn :177-185 = norm_spec :228-229 + transpose + q_sample :147-153, with the
//                                 "t < 0 -> x_start" rule of :180-184)
//   mgb_train_posterior          :210-212,220  x_0_pred = clamp(denoiser_out * valid, -1, 1);
//                                x_t_prev_pred = q_posterior_sample(x_0_pred, x_t, t) * valid   (:104-119)
//   mgb_train_posterior_backward  d/d denoiser_out of the two outputs above (what torch autograd derives for that chain)
// Arithmetic follows the torch expressions operation by operation (explicit round-to-nearest intrinsics, no FMA
// contraction), so forward values are the torch values bit for bit given the same schedule tables.
#include "common.cuh"

namespace mgb {
namespace {

constexpr int TB_TF = 32;          // frames per block of the transposing kernel
constexpr int TB_MAXM = 128;
constexpr int TB_THREADS = 256;

__device__ __forceinline__ float clamp_pm1_nan(float v) {      // torch.clamp propagates NaN
  return v != v ? v : fminf(fmaxf(v, -1.f), 1.f);
}

// mel [B][T][M] -> x_t, x_t_prev [B][M][T]; noise_* [B][M][T]; t int64 [B] in [0, K); sa / sn = sqrt_alphas_cumprod /
// sqrt_one_minus_alphas_cumprod [K]
__global__ void __launch_bounds__(TB_THREADS) train_diffuse_kernel(const float* __restrict__ mel, const float* __restrict__ noise_t,
                                                                   const float* __restrict__ noise_prev, const float* __restrict__ smin,
                                                                   const float* __restrict__ smax, const float* __restrict__ sa,
                                                                   const float* __restrict__ sn, const int64_t* __restrict__ tt,
                                                                   const uint8_t* __restrict__ pad, float* __restrict__ x_t,
                                                                   float* __restrict__ x_prev, int M, int T, int K) {
  __shared__ float tile[TB_MAXM][TB_TF + 1];     // normalised mel, [m][t]
  const int b = blockIdx.y, t0 = blockIdx.x * TB_TF;
  const int nt = min(TB_TF, T - t0);
  const int n = nt * M;
  const float* src = mel + ((size_t)b * T + t0) * M;
  for (int i = threadIdx.x; i < n; i += TB_THREADS) {
    const int t = i / M, m = i - t * M;
    const float lo = smin[m], span = __fsub_rn(smax[m], lo);
    tile[m][t] = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(src[i], lo), span), 2.f), 1.f);
  }
  __syncthreads();
  int64_t ts = tt[b];
  ts = ts < 0 ? 0 : (ts >= K ? K - 1 : ts);              // the host validates the range; never index outside the tables
  const float a1 = sa[ts], n1 = sn[ts];
  const bool has_prev = ts >= 1;
  const float a0 = has_prev ? sa[ts - 1] : 0.f, n0 = has_prev ? sn[ts - 1] : 0.f;
  for (int i = threadIdx.x; i < n; i += TB_THREADS) {
    const int m = i / nt, t = i - m * nt;
    const size_t o = ((size_t)b * M + m) * T + t0 + t;
    const float valid = (pad && pad[(size_t)b * T + t0 + t]) ? 0.f : 1.f;
    const float x0 = tile[m][t];
    x_t[o] = __fmul_rn(__fadd_rn(__fmul_rn(a1, x0), __fmul_rn(n1, noise_t[o])), valid);
    const float xp = has_prev ? __fadd_rn(__fmul_rn(a0, x0), __fmul_rn(n0, noise_prev[o])) : x0;
    x_prev[o] = __fmul_rn(xp, valid);
  }
}

// all tensors [B][M][T]; sched = [3][K]: coef1 | coef2 | sigma (sigma[0] = 0)
__global__ void __launch_bounds__(256) train_posterior_kernel(const float* __restrict__ den, const float* __restrict__ x_t,
                                                              const float* __restrict__ noise, const float* __restrict__ sched,
                                                              const int64_t* __restrict__ tt, const uint8_t* __restrict__ pad,
                                                              int clip, float* __restrict__ x0_out, float* __restrict__ prev_out,
                                                              int M, int T, int K) {
  const int b = blockIdx.y;
  int64_t ts = tt[b];
  ts = ts < 0 ? 0 : (ts >= K ? K - 1 : ts);
  const float c1 = sched[ts], c2 = sched[K + ts], sg = sched[2 * K + ts];
  const size_t base = (size_t)b * M * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * T; i += gridDim.x * blockDim.x) {
    const int t = i % T;
    const float valid = (pad && pad[(size_t)b * T + t]) ? 0.f : 1.f;
    float x0 = __fmul_rn(den[base + i], valid);
    if (clip) x0 = clamp_pm1_nan(x0);
    x0_out[base + i] = x0;
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x_t[base + i]));
    prev_out[base + i] = __fmul_rn(__fadd_rn(mean, __fmul_rn(sg, noise[base + i])), valid);
  }
}

// g_den = valid * [clip passes] * (g_x0 + c1[t] * valid * g_prev)
__global__ void __launch_bounds__(256) train_posterior_bwd_kernel(const float* __restrict__ g_x0, const float* __restrict__ g_prev,
                                                                  const float* __restrict__ den, const float* __restrict__ sched,
                                                                  const int64_t* __restrict__ tt, const uint8_t* __restrict__ pad,
                                                                  int clip, float* __restrict__ g_den, int M, int T, int K) {
  const int b = blockIdx.y;
  int64_t ts = tt[b];
  ts = ts < 0 ? 0 : (ts >= K ? K - 1 : ts);
  const float c1 = sched[ts];
  const size_t base = (size_t)b * M * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * T; i += gridDim.x * blockDim.x) {
    const int t = i % T;
    const bool valid = !(pad && pad[(size_t)b * T + t]);
    float g = 0.f;
    if (valid) {
      g = (g_x0 ? g_x0[base + i] : 0.f) + (g_prev ? c1 * g_prev[base + i] : 0.f);
      if (clip) {
        const float v = den[base + i];
        if (!(v >= -1.f && v <= 1.f)) g = 0.f;       // torch.clamp's backward: gradient where min <= x <= max
      }
    }
    g_den[base + i] = g;
  }
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_train_diffuse(const float* mel, const float* noise_t, const float* noise_prev, const float* spec_min, const float* spec_max,
                      const float* sqrt_acp, const float* sqrt_1m_acp, const int64_t* t, const uint8_t* pad_mask, float* x_t,
                      float* x_t_prev, int B, int T, int n_mel, int K, void* stream) {
  MGB_REQUIRE(mel && noise_t && noise_prev && spec_min && spec_max && sqrt_acp && sqrt_1m_acp && t && x_t && x_t_prev, MGB_E_ARG,
              "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0 && n_mel > 0 && K > 0, MGB_E_ARG, "bad shape");
  MGB_REQUIRE(n_mel <= TB_MAXM, MGB_E_UNSUPPORTED, "n_mel %d > %d", n_mel, TB_MAXM);
  if (int rc = check_arch()) return rc;
  train_diffuse_kernel<<<dim3((T + TB_TF - 1) / TB_TF, B), TB_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      mel, noise_t, noise_prev, spec_min, spec_max, sqrt_acp, sqrt_1m_acp, t, pad_mask, x_t, x_t_prev, n_mel, T, K);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_train_posterior(const float* denoiser_out, const float* x_t, const float* noise, const float* sched, const int64_t* t,
                        const uint8_t* pad_mask, int clip, float* x0_pred, float* x_t_prev_pred, int B, int T, int n_mel, int K,
                        void* stream) {
  MGB_REQUIRE(denoiser_out && x_t && noise && sched && t && x0_pred && x_t_prev_pred, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0 && n_mel > 0 && K > 0, MGB_E_ARG, "bad shape");
  if (int rc = check_arch()) return rc;
  const int per = n_mel * T;
  int gx = (per + 255) / 256;
  if (gx > 64) gx = 64;
  train_posterior_kernel<<<dim3(gx, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(denoiser_out, x_t, noise, sched, t, pad_mask, clip,
                                                                                     x0_pred, x_t_prev_pred, n_mel, T, K);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_train_posterior_backward(const float* grad_x0_pred, const float* grad_x_t_prev_pred, const float* denoiser_out,
                                 const float* sched, const int64_t* t, const uint8_t* pad_mask, int clip, float* grad_denoiser_out,
                                 int B, int T, int n_mel, int K, void* stream) {
  MGB_REQUIRE(denoiser_out && sched && t && grad_denoiser_out, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0 && n_mel > 0 && K > 0, MGB_E_ARG, "bad shape");
  if (int rc = check_arch()) return rc;
  const int per = n_mel * T;
  int gx = (per + 255) / 256;
  if (gx > 64) gx = 64;
  train_posterior_bwd_kernel<<<dim3(gx, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad_x0_pred, grad_x_t_prev_pred, denoiser_out, sched, t, pad_mask, clip, grad_denoiser_out, n_mel, T, K);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

}  // extern "C"

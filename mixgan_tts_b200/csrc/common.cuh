// Shared host-side helpers: error reporting, canonical flat-weight offsets, packed layouts.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/mixgan_b200.h"

namespace mgb {

void set_error(const char* fmt, ...);

#define MGB_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      mgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                           \
      return MGB_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

#define MGB_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      mgb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                     __FILE__, __LINE__);                                           \
      return MGB_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

#define MGB_REQUIRE(cond, code, ...)                                                \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      mgb::set_error(__VA_ARGS__);                                                  \
      return (code);                                                                \
    }                                                                               \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// One-time setup that is PER DEVICE (cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current device only):
//   static PerDeviceOnce once;  if (once.pending()) { ...set attributes...; once.done(); }
// The setup is idempotent, so two host threads racing on a fresh device at worst both run it.
struct PerDeviceOnce {
  unsigned long long mask_[4] = {0, 0, 0, 0};   // 256 device ordinals
  static int current() { int dev = 0; return cudaGetDevice(&dev) == cudaSuccess ? dev & 255 : 0; }
  bool pending() const { const int d = current(); return !(__atomic_load_n(&mask_[d >> 6], __ATOMIC_ACQUIRE) >> (d & 63) & 1ull); }
  void done() { const int d = current(); __atomic_fetch_or(&mask_[d >> 6], 1ull << (d & 63), __ATOMIC_RELEASE); }
};

// Programmatic dependent launch (PDL).  A kernel launched with launch_pdl() may start while its predecessor in the stream
// is still running: its CTAs do their private setup (barrier init, TMEM allocation, cluster sync), then pdl_wait() blocks
// until the predecessor grid has completed and its memory is visible.  pdl_trigger() (issued right away by every CTA)
// lets the successor be scheduled as soon as all of this grid's CTAs are resident.  Both are no-ops in a kernel that was
// launched the ordinary way, so kernels shared with the fp32 path behave as before.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename K, typename... A>
inline cudaError_t launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, const A&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  int n = 0;
  static const int pdl_on = [] { const char* e = getenv("MGB_NO_PDL"); return (e && *e == '1') ? 0 : 1; }();   // A/B switch
  at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[n].val.programmaticStreamSerializationAllowed = pdl_on;
  ++n;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

// ---- canonical flat fp32 weight order (include/mixgan_b200.h) -----------------------------
struct FlatLayer {
  size_t conv_w, conv_b, dproj_w, sproj_w, cproj_w, cproj_b, oproj_w, oproj_b;
};
struct FlatOffsets {
  size_t in_w, in_b, mlp0_w, mlp2_w;
  size_t layer0;        // start of layer 0
  size_t layer_stride;  // floats per layer
  FlatLayer rel;        // offsets relative to a layer's start
  size_t skip_w, skip_b, out_w, out_b;
  size_t total;
};

inline FlatOffsets flat_offsets(const mgb_model_dims& d) {
  const size_t C = d.channels, H = d.d_encoder, M = d.n_mel, L = d.layers;
  FlatOffsets o{};
  size_t p = 0;
  o.in_w = p; p += C * M;
  o.in_b = p; p += C;
  o.mlp0_w = p; p += 4 * C * C;
  o.mlp2_w = p; p += C * 4 * C;
  o.layer0 = p;
  size_t q = 0;
  o.rel.conv_w = q; q += 2 * C * C * 3;
  o.rel.conv_b = q; q += 2 * C;
  o.rel.dproj_w = q; q += C * C;
  o.rel.sproj_w = q; if (d.multi_speaker) q += C * H;
  o.rel.cproj_w = q; q += C * H;
  o.rel.cproj_b = q; q += C;
  o.rel.oproj_w = q; q += 2 * C * C;
  o.rel.oproj_b = q; q += 2 * C;
  o.layer_stride = q;
  p += q * L;
  o.skip_w = p; p += C * C;
  o.skip_b = p; p += C;
  o.out_w = p; p += M * C;
  o.out_b = p; p += M;
  o.total = p;
  return o;
}

inline bool dims_supported(const mgb_model_dims* d) {
  return d && d->n_mel > 0 && d->n_mel <= 128 && d->n_mel % 8 == 0 && d->channels == 256 &&
         d->d_encoder == 256 && d->layers >= 1 && d->layers <= 64 &&
         (d->multi_speaker == 0 || d->multi_speaker == 1);
}

int check_arch();  // MGB_OK or MGB_E_ARCH for the current device

// Launch accounting (bench.py's gpu_launches) and optional CUDA-event timing of the dominant
// kernel on the launching stream (bench.py's roofline.achieved).  See mgb_launch_count / mgb_profile_*.
void note_launch(int n = 1);
void prof_begin(cudaStream_t s);  // no-ops unless profiling is enabled
void prof_end(cudaStream_t s);

// ---- fp32 path (fp32_path.cu) ---------------------------------------------------------------
size_t fp32_packed_bytes(const mgb_model_dims& d);
int fp32_pack(const mgb_model_dims& d, const float* flat, void* packed, cudaStream_t s);
int fp32_pack_tables(const mgb_model_dims& d, const float* flat, void* packed, cudaStream_t s);
size_t fp32_workspace_bytes(const mgb_model_dims& d, int B, int T);
// One Denoiser call.  If sched != nullptr the posterior update is fused (writes x_prev); out_x0
// receives the (optionally clamped) x0 when non-null.
int fp32_denoiser(const mgb_model_dims& d, const void* packed, const float* x, const int64_t* t,
                  const float* cond, const float* spk, const float* noise, const float* sched, int K,
                  int clip, float* x_prev, float* out_x0, int B, int T, void* ws, cudaStream_t s,
                  float* saved = nullptr);

// ---- fp32 training path (fp32_path.cu forward with stash, train_fp32.cu backward) -----------------
// Activations the forward keeps for the backward, all fp32 frames-major (F = B*T rows); offsets in floats.
struct TrainSaved {
  size_t xt;            // [F][n_mel]  input, frames-major
  size_t X0;            // [F][C]      relu(input projection)
  size_t layer0, layer_stride;
  size_t rY, rZ, rG;    // per layer: conv input [F][C], conv pre-activation [F][2C] (gate | filter), gate output [F][C]
  size_t Sn;            // [F][C]      skip sum / sqrt(L)
  size_t P;             // [F][C]      relu(skip projection)
  size_t dvec;          // [B][C]      step MLP output
  size_t h;             // [B][4C]     mish(W0 emb)
  size_t total;
};
inline TrainSaved train_saved_layout(const mgb_model_dims& d, int B, int T) {
  const size_t C = d.channels, F = (size_t)B * T;
  TrainSaved o{};
  size_t p = 0;
  auto take = [&](size_t n) { size_t r = p; p += align_up(n, 64); return r; };
  o.xt = take(F * d.n_mel);
  o.X0 = take(F * C);
  o.layer0 = p;
  {
    size_t q = 0;
    auto tk = [&](size_t n) { size_t r = q; q += align_up(n, 64); return r; };
    o.rY = tk(F * C); o.rZ = tk(F * 2 * C); o.rG = tk(F * C);
    o.layer_stride = q;
  }
  p += o.layer_stride * d.layers;
  o.Sn = take(F * C);
  o.P = take(F * C);
  o.dvec = take((size_t)B * C);
  o.h = take((size_t)B * 4 * C);
  o.total = p;
  return o;
}
int fp32_step_tables(const mgb_model_dims& d, const void* packed, const int64_t* t, const float* spk, int B, float* d_buf,
                     float* h_buf, float* dtab, float* ctab, cudaStream_t s);
size_t train_workspace_bytes(const mgb_model_dims& d, int B, int T);
// Denoiser.forward keeping the activations the backward needs (`saved`, TrainSaved layout); out = [B][n_mel][T].
int fp32_train_forward(const mgb_model_dims& d, const void* packed, const float* x, const int64_t* t, const float* cond,
                       const float* spk, float* out, float* saved, int B, int T, void* ws, cudaStream_t s);
// Backward segments in execution order: 0 = tail (output/skip projections), 1..L = residual layers L-1..0,
// L+1 = head (input projection, step MLP).  Runs segments [seg_begin, seg_end).  Gradients are WRITTEN (not accumulated)
// into grad_flat (canonical flat order) / grad_cond [B][T][H] / grad_spk [B][H] / grad_x [B][n_mel][T] (each optional).
int fp32_train_backward(const mgb_model_dims& d, const float* flat, const float* saved, const int64_t* t,
                        const float* cond, const float* spk, const float* grad_out, float* grad_flat, float* grad_cond,
                        float* grad_spk, float* grad_x, int B, int T, int seg_begin, int seg_end, void* ws, cudaStream_t s);

// ---- bf16 tcgen05 training path (train_bf16.cu): same segments and gradient layout as the fp32 one ----
size_t bf16_train_saved_bytes(const mgb_model_dims& d, int B, int T);
size_t bf16_train_workspace_bytes(const mgb_model_dims& d, int B, int T);
size_t bf16_train_status_offset(const mgb_model_dims& d, int B, int T);
int bf16_train_forward(const mgb_model_dims& d, const void* packed_fp32, const float* flat, const float* x, const int64_t* t,
                       const float* cond, const float* spk, float* out, void* saved, int B, int T, void* ws, cudaStream_t s);
int bf16_train_backward(const mgb_model_dims& d, const float* flat, const void* saved, const int64_t* t, const float* cond,
                        const float* spk, const float* grad_out, float* grad_flat, float* grad_cond, float* grad_spk,
                        float* grad_x, int B, int T, int seg_begin, int seg_end, void* ws, cudaStream_t s);

// ---- tcgen05 sampling path (fused_bf16.cu): bf16 operands, or fp16 operands (f16 = true, MGB_PREC_FP16) ----
size_t bf16_packed_bytes(const mgb_model_dims& d);
int bf16_pack(const mgb_model_dims& d, const float* flat, void* packed, cudaStream_t s, bool f16);
size_t bf16_workspace_bytes(const mgb_model_dims& d, int B, int T, int K, bool f16);
size_t bf16_status_offset(const mgb_model_dims& d, int B, int T);
int bf16_prepare(const mgb_model_dims& d, const void* packed, const int64_t* t, int nsteps, const float* cond,
                 const float* spk, int B, int T, void* ws, cudaStream_t s, bool f16);
int bf16_run(const mgb_model_dims& d, const void* packed, const float* x, const int64_t* t, int t_uniform, int step,
             int nsteps, const float* noise, const float* sched, int K, int clip, float* x_prev, float* out_x0, int B, int T,
             void* ws, cudaStream_t s, bool f16);
int bf16_pack_cond(const mgb_model_dims& d, const float* cond, int B, int T, void* ws, cudaStream_t s, bool f16);
size_t bf16_stamps_offset();   // byte offset (independent of the shape) of the in-situ launch stamps inside the workspace
int bf16_max_stamps();
bool stamp_mode();             // mgb_profile_enable(2): the fused kernel records its own start / end in %globaltimer
long long stamp_next(int add); // launches stamped since mgb_profile_enable(2) (returns the value before adding)
inline bool prec_is_tc(int precision) { return precision == MGB_PREC_BF16 || precision == MGB_PREC_FP16; }

// ---- elementwise (elementwise.cu) -------------------------------------------------------------
int launch_fill_t(int64_t* t, int B, int64_t value, cudaStream_t s);

}  // namespace mgb

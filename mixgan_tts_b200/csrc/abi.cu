// extern "C" entry points of libmixgan_b200.so (see include/mixgan_b200.h) and the small
// elementwise kernels either side of the Denoiser: shallow start, denorm + mask, length regulator.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace mgb {

static thread_local char g_err[512] = "";

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<cudaEvent_t> g_prof_events;  // begin/end pairs

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<long long> g_stamp_seq{0};
bool stamp_mode() { return g_prof_on.load(std::memory_order_relaxed) == 2; }
long long stamp_next(int add) { return g_stamp_seq.fetch_add(add, std::memory_order_relaxed); }

static void prof_record(cudaStream_t s) {
  if (g_prof_on.load(std::memory_order_relaxed) != 1) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_events.push_back(e);
}
void prof_begin(cudaStream_t s) { prof_record(s); }
void prof_end(cudaStream_t s) { prof_record(s); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_arch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device is current (this library has no CPU fallback)");
    return MGB_E_ARCH;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
    return MGB_E_ARCH;
  }
  return MGB_OK;
}

namespace {

__global__ void fill_i64_kernel(int64_t* p, int n, int64_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// The two transposing elementwise kernels are HBM-bound (2 or 3 x 320 B per frame).  One block owns TF consecutive frames
// of one utterance over ALL mel bins: the [frames][M] side of the tile is then one contiguous run of TF * M floats (moved
// as 16-byte vectors), the [M][frames] side is M runs of TF floats (16-byte vectors when T % 4 == 0).  The transpose goes
// through a padded shared-memory tile.  Arithmetic is written with explicit round-to-nearest intrinsics in the reference's
// operation order (no FMA contraction): results are bit-identical to the torch expression ((v + 1) / 2 is computed as
// (v + 1) * 0.5, the same correctly rounded value; the division by (max - min) stays an IEEE division).
// MT = n_mel as a compile-time constant (80: index arithmetic without integer divisions) or 0 = runtime M.
constexpr int EW_TF = 64;        // frames per block
constexpr int EW_MAXM = 128;     // dims_supported: n_mel <= 128
constexpr int EW_THREADS = 256;

// x_T[b][m][t] = (sa * norm(coarse[b][t][m]) + sn * noise[b][m][t]) * valid[b][t]
// norm_spec: (x - min) / (max - min) * 2 - 1  (diffusion.py:228-229); q_sample diffusion.py:150-153
template <int MT>
__global__ void __launch_bounds__(EW_THREADS) shallow_start_kernel(const float* __restrict__ coarse, const float* __restrict__ noise,
                                     const float* __restrict__ smin, const float* __restrict__ smax,
                                     float sa, float sn, const uint8_t* __restrict__ pad,
                                     float* __restrict__ xT, int Mrt, int T) {
  __shared__ float tile[MT ? MT : EW_MAXM][EW_TF + 1];     // [m][t]
  __shared__ float s_lo[MT ? MT : EW_MAXM], s_span[MT ? MT : EW_MAXM], s_valid[EW_TF];
  const int M = MT ? MT : Mrt;
  const int b = blockIdx.y, t0 = blockIdx.x * EW_TF;
  const int nt = min(EW_TF, T - t0);
  for (int i = threadIdx.x; i < M; i += EW_THREADS) { s_lo[i] = smin[i]; s_span[i] = __fsub_rn(smax[i], smin[i]); }
  for (int i = threadIdx.x; i < nt; i += EW_THREADS) s_valid[i] = (pad && pad[(size_t)b * T + t0 + i]) ? 0.f : 1.f;
  __syncthreads();
  const size_t row0 = ((size_t)b * T + t0) * M;      // contiguous run of nt * M floats of coarse
  const int n = nt * M;
  auto norm = [&](float c, int m) { return __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(c, s_lo[m]), s_span[m]), 2.f), 1.f); };
  if constexpr (MT > 0 && (MT * EW_TF / 4) % EW_THREADS == 0) {
    // Full tile, all pointers 16-byte aligned: every thread moves PER float4 of each operand and issues each operand's PER
    // independent 16-byte loads back to back — the kernel is latency-bound (13 us for 49 MB), so bytes in flight per thread
    // are what sets its bandwidth (a plain copy of this size reaches 0.58 of the HBM peak on this box).
    constexpr int PER = (MT * EW_TF / 4) / EW_THREADS, Q = EW_TF / 4;
    if (nt == EW_TF && (T & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(coarse + row0) | reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(xT)) & 15) == 0) {
      const float4* src = reinterpret_cast<const float4*>(coarse + row0);
      float4 cv[PER], nz[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) cv[k] = __ldg(src + threadIdx.x + k * EW_THREADS);
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * EW_THREADS, t = (4 * i) / MT, m = 4 * i - t * MT;
        tile[m][t] = norm(cv[k].x, m); tile[m + 1][t] = norm(cv[k].y, m + 1);
        tile[m + 2][t] = norm(cv[k].z, m + 2); tile[m + 3][t] = norm(cv[k].w, m + 3);
      }
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * EW_THREADS, m = i / Q, tq = (i - m * Q) * 4;
        nz[k] = __ldg(reinterpret_cast<const float4*>(noise + ((size_t)b * MT + m) * T + t0 + tq));
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * EW_THREADS, m = i / Q, tq = (i - m * Q) * 4;
        float4 r;
        r.x = __fadd_rn(__fmul_rn(sa, tile[m][tq]), __fmul_rn(sn, nz[k].x)) * s_valid[tq];
        r.y = __fadd_rn(__fmul_rn(sa, tile[m][tq + 1]), __fmul_rn(sn, nz[k].y)) * s_valid[tq + 1];
        r.z = __fadd_rn(__fmul_rn(sa, tile[m][tq + 2]), __fmul_rn(sn, nz[k].z)) * s_valid[tq + 2];
        r.w = __fadd_rn(__fmul_rn(sa, tile[m][tq + 3]), __fmul_rn(sn, nz[k].w)) * s_valid[tq + 3];
        *reinterpret_cast<float4*>(xT + ((size_t)b * MT + m) * T + t0 + tq) = r;
      }
      return;
    }
  }
  if ((reinterpret_cast<uintptr_t>(coarse + row0) & 15) == 0 && (M & 3) == 0) {   // a float4 never straddles two frames
    const float4* src = reinterpret_cast<const float4*>(coarse + row0);
    for (int i = threadIdx.x; i < n / 4; i += EW_THREADS) {
      const float4 v = __ldg(src + i);
      const int t = (4 * i) / M, m = 4 * i - t * M;
      tile[m][t] = norm(v.x, m); tile[m + 1][t] = norm(v.y, m + 1); tile[m + 2][t] = norm(v.z, m + 2); tile[m + 3][t] = norm(v.w, m + 3);
    }
  } else {
    for (int idx = threadIdx.x; idx < n; idx += EW_THREADS) {
      const int t = idx / M, m = idx - t * M;
      tile[m][t] = norm(coarse[row0 + idx], m);
    }
  }
  __syncthreads();
  const bool vec = (T & 3) == 0 && (nt & 3) == 0 && ((reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(xT)) & 15) == 0;
  if (vec) {
    const int q = nt / 4;
    for (int i = threadIdx.x; i < M * q; i += EW_THREADS) {
      const int m = i / q, tq = (i - m * q) * 4;
      const size_t o = ((size_t)b * M + m) * T + t0 + tq;
      const float4 nz = __ldg(reinterpret_cast<const float4*>(noise + o));
      float4 r;
      r.x = __fadd_rn(__fmul_rn(sa, tile[m][tq]), __fmul_rn(sn, nz.x)) * s_valid[tq];
      r.y = __fadd_rn(__fmul_rn(sa, tile[m][tq + 1]), __fmul_rn(sn, nz.y)) * s_valid[tq + 1];
      r.z = __fadd_rn(__fmul_rn(sa, tile[m][tq + 2]), __fmul_rn(sn, nz.z)) * s_valid[tq + 2];
      r.w = __fadd_rn(__fmul_rn(sa, tile[m][tq + 3]), __fmul_rn(sn, nz.w)) * s_valid[tq + 3];
      *reinterpret_cast<float4*>(xT + o) = r;
    }
  } else {
    for (int i = threadIdx.x; i < M * nt; i += EW_THREADS) {
      const int m = i / nt, t = i - m * nt;
      const size_t o = ((size_t)b * M + m) * T + t0 + t;
      xT[o] = __fadd_rn(__fmul_rn(sa, tile[m][t]), __fmul_rn(sn, noise[o])) * s_valid[t];
    }
  }
}

// mel[b][t][m] = ((x[b][m][t] + 1) / 2 * (max - min) + min) * valid[b][t]   (diffusion.py:231-232)
template <int MT>
__global__ void __launch_bounds__(EW_THREADS) denorm_mask_kernel(const float* __restrict__ x, const float* __restrict__ smin,
                                   const float* __restrict__ smax, const uint8_t* __restrict__ pad,
                                   float* __restrict__ mel, int Mrt, int T) {
  __shared__ float tile[MT ? MT : EW_MAXM][EW_TF + 1];     // [m][t]
  __shared__ float s_lo[MT ? MT : EW_MAXM], s_span[MT ? MT : EW_MAXM], s_valid[EW_TF];
  const int M = MT ? MT : Mrt;
  const int b = blockIdx.y, t0 = blockIdx.x * EW_TF;
  const int nt = min(EW_TF, T - t0);
  for (int i = threadIdx.x; i < M; i += EW_THREADS) { s_lo[i] = smin[i]; s_span[i] = __fsub_rn(smax[i], smin[i]); }
  for (int i = threadIdx.x; i < nt; i += EW_THREADS) s_valid[i] = (pad && pad[(size_t)b * T + t0 + i]) ? 0.f : 1.f;
  const bool vec = (T & 3) == 0 && (nt & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  bool loaded = false;
  if constexpr (MT > 0 && (MT * EW_TF / 4) % EW_THREADS == 0) {
    constexpr int PER = (MT * EW_TF / 4) / EW_THREADS, Q = EW_TF / 4;
    if (vec && nt == EW_TF) {            // full tile: all PER loads of a thread in flight before the first use
      float4 v[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * EW_THREADS, m = i / Q, tq = (i - m * Q) * 4;
        v[k] = __ldg(reinterpret_cast<const float4*>(x + ((size_t)b * MT + m) * T + t0 + tq));
      }
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * EW_THREADS, m = i / Q, tq = (i - m * Q) * 4;
        tile[m][tq] = v[k].x; tile[m][tq + 1] = v[k].y; tile[m][tq + 2] = v[k].z; tile[m][tq + 3] = v[k].w;
      }
      loaded = true;
    }
  }
  if (loaded) {
  } else if (vec) {
    const int q = nt / 4;
    for (int i = threadIdx.x; i < M * q; i += EW_THREADS) {
      const int m = i / q, tq = (i - m * q) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)b * M + m) * T + t0 + tq));
      tile[m][tq] = v.x; tile[m][tq + 1] = v.y; tile[m][tq + 2] = v.z; tile[m][tq + 3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < M * nt; i += EW_THREADS) {
      const int m = i / nt, t = i - m * nt;
      tile[m][t] = x[((size_t)b * M + m) * T + t0 + t];
    }
  }
  __syncthreads();
  const size_t row0 = ((size_t)b * T + t0) * M;
  const int n = nt * M;
  auto value = [&](int m, int t) {
    return __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(tile[m][t], 1.f), 0.5f), s_span[m]), s_lo[m]) * s_valid[t];
  };
  if ((reinterpret_cast<uintptr_t>(mel + row0) & 15) == 0 && (M & 3) == 0) {
    float4* dst = reinterpret_cast<float4*>(mel + row0);
    for (int i = threadIdx.x; i < n / 4; i += EW_THREADS) {
      const int t = (4 * i) / M, m = 4 * i - t * M;
      dst[i] = make_float4(value(m, t), value(m + 1, t), value(m + 2, t), value(m + 3, t));
    }
  } else {
    for (int idx = threadIdx.x; idx < n; idx += EW_THREADS) {
      const int t = idx / M, m = idx - t * M;
      mel[row0 + idx] = value(m, t);
    }
  }
}

// ---- register-transpose forms of the two kernels above (T % 4 == 0, n_mel % 4 == 0, 16-byte aligned pointers) ----------
// A thread owns a 4 (mel) x 4 (frames) block: four 16-byte loads along one axis, a transpose in registers, four 16-byte
// stores along the other — no shared memory, no barrier, no bank conflicts.  Lane = (mb = lane & 7, tb = lane >> 3): a warp
// covers 32 mel bins x 16 frames, so a warp-wide access on the [frames][mel] side is 4 rows x 128 contiguous bytes and on the
// [mel][frames] side 8 rows x 64 contiguous bytes: full 32-byte sectors both ways.  A warp walks two 16-frame chunks (eight
// independent loads in flight per thread per operand).  Same arithmetic, operation by operation, as the tiled kernels.
constexpr int RT_THREADS = 256, RT_FRAMES = 32;       // frames per warp task

// task = (utterance b, 32-frame chunk, 32-mel group); tasks are enumerated mel-group fastest
__device__ __forceinline__ bool rt_task(int B, int T, int M, int frames, int& b, int& t_base, int& m0, int& tb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nmg = (M + 31) / 32, nch = (T + frames - 1) / frames;
  const long long task = (long long)blockIdx.x * (RT_THREADS / 32) + warp;
  if (task >= (long long)B * nch * nmg) return false;
  const int mg = (int)(task % nmg);
  const long long r = task / nmg;
  const int ch = (int)(r % nch);
  b = (int)(r / nch);
  t_base = ch * frames;
  m0 = mg * 32 + (lane & 7) * 4;
  tb = lane >> 3;
  return m0 < M;
}

template <int NH>
__global__ void __launch_bounds__(RT_THREADS) denorm_mask_rt_kernel(const float* __restrict__ x, const float* __restrict__ smin,
                                                                    const float* __restrict__ smax, const uint8_t* __restrict__ pad,
                                                                    float* __restrict__ mel, int B, int M, int T) {
  int b, t_base, m0, tb;
  if (!rt_task(B, T, M, 16 * NH, b, t_base, m0, tb)) return;
  float4 v[NH][4];
  int t0[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    t0[h] = t_base + tb * 4 * NH + h * 4;     // a thread owns 4 NH consecutive frames
#pragma unroll
    for (int r = 0; r < 4; ++r)
      v[h][r] = t0[h] < T ? __ldg(reinterpret_cast<const float4*>(x + ((size_t)b * M + m0 + r) * T + t0[h])) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 lo = __ldg(reinterpret_cast<const float4*>(smin + m0)), hi = __ldg(reinterpret_cast<const float4*>(smax + m0));
  const float span[4] = {__fsub_rn(hi.x, lo.x), __fsub_rn(hi.y, lo.y), __fsub_rn(hi.z, lo.z), __fsub_rn(hi.w, lo.w)};
  const float lov[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    if (t0[h] >= T) continue;
    uint32_t pm = 0u;
    if (pad) pm = *reinterpret_cast<const uint32_t*>(pad + (size_t)b * T + t0[h]);      // 4 mask bytes (T % 4 == 0)
    const float in[4][4] = {{v[h][0].x, v[h][0].y, v[h][0].z, v[h][0].w}, {v[h][1].x, v[h][1].y, v[h][1].z, v[h][1].w},
                            {v[h][2].x, v[h][2].y, v[h][2].z, v[h][2].w}, {v[h][3].x, v[h][3].y, v[h][3].z, v[h][3].w}};   // [r][e]
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float valid = ((pm >> (8 * e)) & 0xFFu) ? 0.f : 1.f;
      float o[4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        o[r] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(in[r][e], 1.f), 0.5f), span[r]), lov[r]) * valid;
      *reinterpret_cast<float4*>(mel + ((size_t)b * T + t0[h] + e) * M + m0) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <int NH>
__global__ void __launch_bounds__(RT_THREADS) shallow_start_rt_kernel(const float* __restrict__ coarse, const float* __restrict__ noise,
                                                                      const float* __restrict__ smin, const float* __restrict__ smax,
                                                                      float sa, float sn, const uint8_t* __restrict__ pad,
                                                                      float* __restrict__ xT, int B, int M, int T) {
  int b, t_base, m0, tb;
  if (!rt_task(B, T, M, 16 * NH, b, t_base, m0, tb)) return;
  float4 c[NH][4], nz[NH][4];
  int t0[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    t0[h] = t_base + tb * 4 * NH + h * 4;     // a thread owns 4 NH consecutive frames
    const bool in = t0[h] < T;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      c[h][e] = in ? __ldg(reinterpret_cast<const float4*>(coarse + ((size_t)b * T + t0[h] + e) * M + m0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      nz[h][r] = in ? __ldg(reinterpret_cast<const float4*>(noise + ((size_t)b * M + m0 + r) * T + t0[h])) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 lo = __ldg(reinterpret_cast<const float4*>(smin + m0)), hi = __ldg(reinterpret_cast<const float4*>(smax + m0));
  const float span[4] = {__fsub_rn(hi.x, lo.x), __fsub_rn(hi.y, lo.y), __fsub_rn(hi.z, lo.z), __fsub_rn(hi.w, lo.w)};
  const float lov[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    if (t0[h] >= T) continue;
    uint32_t pm = 0u;
    if (pad) pm = *reinterpret_cast<const uint32_t*>(pad + (size_t)b * T + t0[h]);
    const float cv[4][4] = {{c[h][0].x, c[h][0].y, c[h][0].z, c[h][0].w}, {c[h][1].x, c[h][1].y, c[h][1].z, c[h][1].w},
                            {c[h][2].x, c[h][2].y, c[h][2].z, c[h][2].w}, {c[h][3].x, c[h][3].y, c[h][3].z, c[h][3].w}};   // [e][r]
    const float nv[4][4] = {{nz[h][0].x, nz[h][0].y, nz[h][0].z, nz[h][0].w}, {nz[h][1].x, nz[h][1].y, nz[h][1].z, nz[h][1].w},
                            {nz[h][2].x, nz[h][2].y, nz[h][2].z, nz[h][2].w}, {nz[h][3].x, nz[h][3].y, nz[h][3].z, nz[h][3].w}};   // [r][e]
    float valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) valid[e] = ((pm >> (8 * e)) & 0xFFu) ? 0.f : 1.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float nrm = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(cv[e][r], lov[r]), span[r]), 2.f), 1.f);
        o[e] = __fadd_rn(__fmul_rn(sa, nrm), __fmul_rn(sn, nv[r][e])) * valid[e];
      }
      *reinterpret_cast<float4*>(xT + ((size_t)b * M + m0 + r) * T + t0[h]) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

inline bool rt_ok(int M, int T, const void* a, const void* b2, const void* c2, const void* d, const void* e) {
  return (M & 3) == 0 && (T & 3) == 0 &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b2) | reinterpret_cast<uintptr_t>(c2) |
           reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e)) & 15) == 0;
}
inline unsigned rt_blocks(int B, int M, int T, int frames = RT_FRAMES) {
  const long long tasks = (long long)B * ((T + frames - 1) / frames) * ((M + 31) / 32);
  return (unsigned)((tasks + RT_THREADS / 32 - 1) / (RT_THREADS / 32));
}

static void launch_denorm_mask(const float* x, const float* smin, const float* smax, const uint8_t* pad, float* mel, int B, int M,
                               int T, cudaStream_t s) {
  if (rt_ok(M, T, x, smin, smax, mel, pad && (reinterpret_cast<uintptr_t>(pad) & 3) ? reinterpret_cast<const void*>(1) : nullptr)) {
    denorm_mask_rt_kernel<1><<<rt_blocks(B, M, T, 16), RT_THREADS, 0, s>>>(x, smin, smax, pad, mel, B, M, T);
    note_launch();
    return;
  }
  dim3 grid((T + EW_TF - 1) / EW_TF, B), block(EW_THREADS);
  if (M == 80) denorm_mask_kernel<80><<<grid, block, 0, s>>>(x, smin, smax, pad, mel, M, T);
  else denorm_mask_kernel<0><<<grid, block, 0, s>>>(x, smin, smax, pad, mel, M, T);
  note_launch();
}

// ---- LengthRegulator: duration rounding, exclusive scan of clamped durations, gather, mask, backward ---------------
// dur = (int64) max(round_half_even(exp(log_d) - 1) * d_control, 0)      (linguistic_encoder.py:310-314)
// exp is evaluated in double and rounded once to fp32 (a correctly rounded fp32 exp), the rest in fp32 as torch does.
__global__ void lr_durations_kernel(const float* __restrict__ log_d, float d_control, int64_t* __restrict__ dur, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = (float)exp((double)log_d[i]);
  const float d = __fmul_rn(rintf(__fsub_rn(e, 1.f)), d_control);
  dur[i] = (int64_t)fmaxf(d, 0.f);                   // .long() truncates towards zero
}

__global__ void lr_scan_kernel(const int64_t* __restrict__ dur, int64_t* __restrict__ cum,
                               int64_t* __restrict__ mel_len, int S) {
  // one block per utterance; S is small (phonemes/words), a serial scan by thread 0 keeps the
  // arithmetic identical to the reference's Python loop (linguistic_encoder.py:402-410)
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int s = 0; s < S; ++s) {
      cum[(size_t)b * (S + 1) + s] = acc;
      const int64_t dv = dur[(size_t)b * S + s];
      acc += dv > 0 ? dv : 0;
    }
    cum[(size_t)b * (S + 1) + S] = acc;
    mel_len[b] = acc;
  }
}

// One block per output frame: binary search of the source row, then a vectorised row copy; also the frame's mask byte.
__global__ void lr_gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ cum,
                                 float* __restrict__ out, uint8_t* __restrict__ mask_valid, int S, int D, int max_len) {
  const int b = blockIdx.y, f = blockIdx.x;
  const int64_t* cb = cum + (size_t)b * (S + 1);
  // binary search: largest s with cum[s] <= f (and f < cum[s+1])
  int src = -1;
  if ((int64_t)f < cb[S]) {
    int lo = 0, hi = S;  // invariant cb[lo] <= f < cb[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cb[mid] <= (int64_t)f) lo = mid; else hi = mid;
    }
    src = lo;
  }
  if (mask_valid && threadIdx.x == 0) mask_valid[(size_t)b * max_len + f] = src >= 0 ? 1 : 0;   // get_mask_from_lengths: f < mel_len
  float* o = out + ((size_t)b * max_len + f) * D;
  const float* xi = src >= 0 ? x + ((size_t)b * S + src) * D : nullptr;
  if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    float4* o4 = reinterpret_cast<float4*>(o);
    const float4* x4 = reinterpret_cast<const float4*>(xi);
    for (int d = threadIdx.x; d < D / 4; d += blockDim.x) o4[d] = xi ? __ldg(x4 + d) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int d = threadIdx.x; d < D; d += blockDim.x) o[d] = xi ? xi[d] : 0.f;
  }
}

// d/dx: grad_x[b][s] = sum of grad_out[b][f] over the frames f that copy source row s (cropped at max_len), summed in
// frame order (deterministic).  One block per (s, b).
__global__ void lr_backward_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ cum,
                                   float* __restrict__ grad_x, int S, int D, int max_len) {
  const int b = blockIdx.y, s = blockIdx.x;
  const int64_t* cb = cum + (size_t)b * (S + 1);
  const int64_t lo = cb[s] < max_len ? cb[s] : max_len, hi = cb[s + 1] < max_len ? cb[s + 1] : max_len;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int64_t f = lo; f < hi; ++f) acc += grad_out[((size_t)b * max_len + f) * D + d];
    grad_x[((size_t)b * S + s) * D + d] = acc;
  }
}

// mask[b][f] = f < lengths[b]  (utils/tools.py:144-153, True = valid)
__global__ void mask_from_lengths_kernel(const int64_t* __restrict__ lengths, uint8_t* __restrict__ mask, int B, int max_len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * max_len) return;
  const int b = i / max_len, f = i - b * max_len;
  mask[i] = (int64_t)f < lengths[b] ? 1 : 0;
}

}  // namespace

int launch_fill_t(int64_t* t, int B, int64_t value, cudaStream_t s) {
  fill_i64_kernel<<<(B + 127) / 128, 128, 0, s>>>(t, B, value);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

static int denoiser_dispatch(const mgb_model_dims* dims, int precision, const void* packed,
                             const float* x, const int64_t* t, const float* cond, const float* spk,
                             const float* noise, const float* sched, int K, int clip, float* x_prev,
                             float* x0_out, int B, int T, void* ws, size_t ws_bytes, cudaStream_t s) {
  MGB_REQUIRE(dims_supported(dims), MGB_E_ARG, "unsupported model dims");
  MGB_REQUIRE(packed && x && t && cond && ws, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0, MGB_E_ARG, "B and T must be positive (got %d, %d)", B, T);
  MGB_REQUIRE(!dims->multi_speaker || spk, MGB_E_ARG,
              "multi_speaker model needs a speaker embedding (reference raises TypeError)");
  if (int rc = check_arch()) return rc;
  const size_t need = mgb_workspace_bytes(dims, precision, B, T, K > 0 ? K : 1);
  MGB_REQUIRE(ws_bytes >= need, MGB_E_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  if (precision == MGB_PREC_FP32)
    return fp32_denoiser(*dims, packed, x, t, cond, spk, noise, sched, K, clip, x_prev, x0_out, B, T, ws, s);
  if (prec_is_tc(precision)) {
    const bool f16 = precision == MGB_PREC_FP16;
    if (int rc = bf16_prepare(*dims, packed, t, 1, cond, spk, B, T, ws, s, f16)) return rc;
    return bf16_run(*dims, packed, x, t, -1, 0, 1, noise, sched, K, clip, x_prev, x0_out, B, T, ws, s, f16);
  }
  set_error("unknown precision %d", precision);
  return MGB_E_ARG;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_abi_version(void) { return MGB_ABI_VERSION; }

long long mgb_launch_count(void) { return g_launches.load(); }
void mgb_note_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void mgb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (cudaEvent_t e : g_prof_events) cudaEventDestroy(e);
  g_prof_events.clear();
  g_stamp_seq.store(0);
  g_prof_on.store(on == 2 ? 2 : on ? 1 : 0);
}

int mgb_profile_read_stamps(const void* workspace, float* total_ms, int* count, float* min_ms, float* max_ms, float* sm_mhz) {
  MGB_REQUIRE(workspace && total_ms && count, MGB_E_ARG, "NULL pointer argument");
  const int n = bf16_max_stamps();
  std::vector<unsigned long long> h(4 * (size_t)n);
  MGB_CUDA_CHECK(cudaMemcpy(h.data(), static_cast<const uint8_t*>(workspace) + bf16_stamps_offset(), h.size() * 8,
                            cudaMemcpyDeviceToHost));
  double total = 0, lo = 1e30, hi = 0;
  int c = 0;
  for (int i = 0; i < n; ++i) {
    if (h[i] == ~0ull || h[n + i] <= h[i]) continue;       // launch slot not used by the last call
    const double ms = (double)(h[n + i] - h[i]) * 1e-6;
    total += ms; lo = ms < lo ? ms : lo; hi = ms > hi ? ms : hi; ++c;
  }
  *total_ms = (float)total; *count = c;
  if (min_ms) *min_ms = c ? (float)lo : 0.f;
  if (max_ms) *max_ms = (float)hi;
  if (sm_mhz) {
    double cyc = 0, ns = 0;
    for (int i = 0; i < n; ++i) { cyc += (double)h[2 * (size_t)n + i]; ns += (double)h[3 * (size_t)n + i]; }
    *sm_mhz = ns > 0 ? (float)(cyc / ns * 1e3) : 0.f;
  }
  return MGB_OK;
}

int mgb_profile_collect(float* total_ms, int* count) {
  MGB_REQUIRE(total_ms && count, MGB_E_ARG, "NULL pointer argument");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  float total = 0.f;
  int n = 0;
  for (size_t i = 0; i + 1 < g_prof_events.size(); i += 2) {
    MGB_CUDA_CHECK(cudaEventSynchronize(g_prof_events[i + 1]));
    float ms = 0.f;
    MGB_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof_events[i], g_prof_events[i + 1]));
    total += ms;
    ++n;
  }
  for (cudaEvent_t e : g_prof_events) cudaEventDestroy(e);
  g_prof_events.clear();
  *total_ms = total;
  *count = n;
  return MGB_OK;
}
const char* mgb_last_error(void) { return g_err; }

int mgb_debug_status(const mgb_model_dims* dims, int precision, int B, int T, const void* workspace, int* host_status) {
  MGB_REQUIRE(dims_supported(dims) && workspace && host_status, MGB_E_ARG, "bad argument");
  *host_status = 0;
  if (!prec_is_tc(precision)) return MGB_OK;
  const char* src = static_cast<const char*>(workspace) + bf16_status_offset(*dims, B, T);
  MGB_CUDA_CHECK(cudaMemcpy(host_status, src, sizeof(int), cudaMemcpyDeviceToHost));
  return MGB_OK;
}

int mgb_device_check(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    set_error("CUDA device %d not available (this library has no CPU fallback)", device);
    return MGB_E_ARCH;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) {
    set_error("device %d has compute capability %d.x; need sm_100", device, major);
    return MGB_E_ARCH;
  }
  return MGB_OK;
}

size_t mgb_flat_weight_count(const mgb_model_dims* dims) {
  return dims_supported(dims) ? flat_offsets(*dims).total : 0;
}

size_t mgb_packed_bytes(const mgb_model_dims* dims, int precision) {
  if (!dims_supported(dims)) return 0;
  return (precision == MGB_PREC_FP32 || precision == MGB_PACK_FP32_TABLES) ? fp32_packed_bytes(*dims)
       : prec_is_tc(precision) ? bf16_packed_bytes(*dims) : 0;
}

int mgb_pack_weights(const mgb_model_dims* dims, int precision, const float* flat, void* packed,
                     size_t packed_bytes, void* stream) {
  MGB_REQUIRE(dims_supported(dims), MGB_E_ARG, "unsupported model dims");
  MGB_REQUIRE(flat && packed, MGB_E_ARG, "NULL pointer argument");
  if (int rc = check_arch()) return rc;
  const size_t need = mgb_packed_bytes(dims, precision);
  MGB_REQUIRE(need > 0, MGB_E_ARG, "unknown precision %d", precision);
  MGB_REQUIRE(packed_bytes >= need, MGB_E_WORKSPACE, "packed buffer too small: %zu < %zu", packed_bytes, need);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == MGB_PACK_FP32_TABLES) return fp32_pack_tables(*dims, flat, packed, s);
  return precision == MGB_PREC_FP32 ? fp32_pack(*dims, flat, packed, s) : bf16_pack(*dims, flat, packed, s, precision == MGB_PREC_FP16);
}

size_t mgb_workspace_bytes(const mgb_model_dims* dims, int precision, int B, int T, int K) {
  if (!dims_supported(dims) || B <= 0 || T <= 0) return 0;
  const size_t tail = align_up((size_t)B * sizeof(int64_t), 256) +            // timestep vector
                      2 * align_up((size_t)B * dims->n_mel * T * sizeof(float), 256);  // x ping-pong
  const size_t core = precision == MGB_PREC_FP32 ? fp32_workspace_bytes(*dims, B, T)
                    : prec_is_tc(precision) ? bf16_workspace_bytes(*dims, B, T, K > 0 ? K : 1, precision == MGB_PREC_FP16) : 0;
  return core ? align_up(core, 256) + tail : 0;
}

int mgb_denoiser_forward(const mgb_model_dims* dims, int precision, const void* packed, const float* x,
                         const int64_t* t, const float* cond, const float* spk, float* out, int B, int T,
                         void* workspace, size_t workspace_bytes, void* stream) {
  MGB_REQUIRE(out, MGB_E_ARG, "NULL output");
  return denoiser_dispatch(dims, precision, packed, x, t, cond, spk, nullptr, nullptr, 0, 0, nullptr, out,
                           B, T, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mgb_reverse_step(const mgb_model_dims* dims, int precision, const void* packed, const float* x_t,
                     const int64_t* t, const float* cond, const float* spk, const float* noise,
                     const float* sched, int K, int clip, float* x_prev, float* x0_out, int B, int T,
                     void* workspace, size_t workspace_bytes, void* stream) {
  MGB_REQUIRE(noise && sched && x_prev && K > 0, MGB_E_ARG, "reverse_step needs noise, sched, x_prev and K > 0");
  return denoiser_dispatch(dims, precision, packed, x_t, t, cond, spk, noise, sched, K, clip, x_prev, x0_out,
                           B, T, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mgb_sample(const mgb_model_dims* dims, int precision, const void* packed, const float* x_T,
               const float* cond, const float* spk, const float* noises, const float* sched, int K, int clip,
               const float* spec_min, const float* spec_max, const uint8_t* pad_mask, float* states_out,
               float* mel_out, float* x0_norm_out, int B, int T, void* workspace, size_t workspace_bytes,
               void* stream) {
  MGB_REQUIRE(dims_supported(dims), MGB_E_ARG, "unsupported model dims");
  MGB_REQUIRE(x_T && noises && sched && spec_min && spec_max && mel_out && workspace && K > 0, MGB_E_ARG,
              "NULL pointer argument or K <= 0");
  MGB_REQUIRE(B > 0 && T > 0, MGB_E_ARG, "B and T must be positive (got %d, %d)", B, T);
  const size_t need = mgb_workspace_bytes(dims, precision, B, T, K);
  MGB_REQUIRE(need > 0 && workspace_bytes >= need, MGB_E_WORKSPACE, "workspace too small: %zu < %zu",
              workspace_bytes, need);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int M = dims->n_mel;
  const size_t xbytes = align_up((size_t)B * M * T * sizeof(float), 256);
  const size_t tbytes = align_up((size_t)B * sizeof(int64_t), 256);
  char* tail = static_cast<char*>(workspace) + (need - tbytes - 2 * xbytes);
  int64_t* tvec = reinterpret_cast<int64_t*>(tail);
  float* xa = reinterpret_cast<float*>(tail + tbytes);
  float* xb = reinterpret_cast<float*>(tail + tbytes + xbytes);
  const size_t mel_elems = (size_t)B * T * M;

  if (states_out)    // sampling() returns the start state too (diffusion.py:160,164)
    launch_denorm_mask(x_T, spec_min, spec_max, nullptr, states_out, B, M, T, s);
  const float* cur = x_T;
  const bool bf16 = prec_is_tc(precision), f16 = precision == MGB_PREC_FP16;
  if (bf16) {   // cond image and the per-step tables of all K steps, once
    MGB_REQUIRE(packed && cond, MGB_E_ARG, "NULL pointer argument");
    MGB_REQUIRE(!dims->multi_speaker || spk, MGB_E_ARG,
                "multi_speaker model needs a speaker embedding (reference raises TypeError)");
    if (int rc = check_arch()) return rc;
    if (int rc = bf16_prepare(*dims, packed, nullptr, K, cond, spk, B, T, workspace, s, f16)) return rc;
  }
  for (int i = K - 1, n = 0; i >= 0; --i, ++n) {
    float* nxt = (n & 1) ? xb : xa;
    const float* nz = noises + (size_t)i * B * M * T;
    int rc;
    if (bf16) {
      rc = bf16_run(*dims, packed, cur, nullptr, i, i, K, nz, sched, K, clip, nxt, nullptr, B, T, workspace, s, f16);
    } else {
      if ((rc = launch_fill_t(tvec, B, i, s))) return rc;
      rc = denoiser_dispatch(dims, precision, packed, cur, tvec, cond, spk, nz, sched, K, clip, nxt, nullptr, B, T,
                             workspace, workspace_bytes, s);
    }
    if (rc) return rc;
    cur = nxt;
    if (states_out) launch_denorm_mask(cur, spec_min, spec_max, nullptr, states_out + (size_t)(n + 1) * mel_elems, B, M, T, s);
  }
  launch_denorm_mask(cur, spec_min, spec_max, pad_mask, mel_out, B, M, T, s);
  if (x0_norm_out)
    MGB_CUDA_CHECK(cudaMemcpyAsync(x0_norm_out, cur, (size_t)B * M * T * sizeof(float),
                                   cudaMemcpyDeviceToDevice, s));
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

static bool train_prec_ok(int precision) { return precision == MGB_PREC_FP32 || precision == MGB_PREC_BF16; }

size_t mgb_train_saved_bytes(const mgb_model_dims* dims, int precision, int B, int T) {
  if (!dims_supported(dims) || B <= 0 || T <= 0 || !train_prec_ok(precision)) return 0;
  return precision == MGB_PREC_FP32 ? train_saved_layout(*dims, B, T).total * sizeof(float)
                                    : bf16_train_saved_bytes(*dims, B, T);
}

size_t mgb_train_workspace_bytes(const mgb_model_dims* dims, int precision, int B, int T) {
  if (!dims_supported(dims) || B <= 0 || T <= 0 || !train_prec_ok(precision)) return 0;
  return precision == MGB_PREC_FP32 ? train_workspace_bytes(*dims, B, T) : bf16_train_workspace_bytes(*dims, B, T);
}

int mgb_train_segments(const mgb_model_dims* dims) { return dims_supported(dims) ? dims->layers + 2 : 0; }

int mgb_train_segment_range(const mgb_model_dims* dims, int seg, size_t* flat_begin, size_t* flat_end) {
  MGB_REQUIRE(dims_supported(dims) && flat_begin && flat_end, MGB_E_ARG, "bad argument");
  const FlatOffsets f = flat_offsets(*dims);
  const int L = dims->layers;
  MGB_REQUIRE(seg >= 0 && seg <= L + 1, MGB_E_ARG, "segment %d out of range [0, %d]", seg, L + 1);
  if (seg == 0) { *flat_begin = f.skip_w; *flat_end = f.total; }
  else if (seg <= L) { *flat_begin = f.layer0 + (size_t)(L - seg) * f.layer_stride; *flat_end = *flat_begin + f.layer_stride; }
  else { *flat_begin = 0; *flat_end = f.layer0; }
  return MGB_OK;
}

int mgb_denoiser_train_forward(const mgb_model_dims* dims, int precision, const void* packed, const float* flat,
                               const float* x, const int64_t* t, const float* cond, const float* spk, float* out,
                               void* saved, size_t saved_bytes, int B, int T, void* workspace, size_t workspace_bytes,
                               void* stream) {
  MGB_REQUIRE(dims_supported(dims), MGB_E_ARG, "unsupported model dims");
  MGB_REQUIRE(packed && flat && x && t && cond && out && saved && workspace, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0, MGB_E_ARG, "B and T must be positive (got %d, %d)", B, T);
  MGB_REQUIRE(!dims->multi_speaker || spk, MGB_E_ARG, "multi_speaker model needs a speaker embedding");
  MGB_REQUIRE(train_prec_ok(precision), MGB_E_ARG, "unknown precision %d", precision);
  if (int rc = check_arch()) return rc;
  MGB_REQUIRE(saved_bytes >= mgb_train_saved_bytes(dims, precision, B, T), MGB_E_WORKSPACE,
              "activation stash too small: %zu < %zu", saved_bytes, mgb_train_saved_bytes(dims, precision, B, T));
  MGB_REQUIRE(workspace_bytes >= mgb_train_workspace_bytes(dims, precision, B, T), MGB_E_WORKSPACE,
              "workspace too small: %zu < %zu", workspace_bytes, mgb_train_workspace_bytes(dims, precision, B, T));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == MGB_PREC_FP32)
    return fp32_train_forward(*dims, packed, x, t, cond, spk, out, static_cast<float*>(saved), B, T, workspace, s);
  return bf16_train_forward(*dims, packed, flat, x, t, cond, spk, out, saved, B, T, workspace, s);
}

int mgb_denoiser_backward(const mgb_model_dims* dims, int precision, const float* flat, const void* saved,
                          size_t saved_bytes, const int64_t* t, const float* cond, const float* spk,
                          const float* grad_out, float* grad_flat, float* grad_cond, float* grad_spk, float* grad_x,
                          int B, int T, int seg_begin, int seg_end, void* workspace, size_t workspace_bytes, void* stream) {
  MGB_REQUIRE(dims_supported(dims), MGB_E_ARG, "unsupported model dims");
  MGB_REQUIRE(flat && saved && t && cond && grad_out && grad_flat && workspace, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0, MGB_E_ARG, "B and T must be positive (got %d, %d)", B, T);
  MGB_REQUIRE(!dims->multi_speaker || spk, MGB_E_ARG, "multi_speaker model needs a speaker embedding");
  MGB_REQUIRE(train_prec_ok(precision), MGB_E_ARG, "unknown precision %d", precision);
  MGB_REQUIRE(seg_begin >= 0 && seg_begin <= seg_end && seg_end <= dims->layers + 2, MGB_E_ARG,
              "bad segment range [%d, %d)", seg_begin, seg_end);
  if (int rc = check_arch()) return rc;
  MGB_REQUIRE(saved_bytes >= mgb_train_saved_bytes(dims, precision, B, T), MGB_E_WORKSPACE, "activation stash too small");
  MGB_REQUIRE(workspace_bytes >= mgb_train_workspace_bytes(dims, precision, B, T), MGB_E_WORKSPACE, "workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == MGB_PREC_FP32)
    return fp32_train_backward(*dims, flat, static_cast<const float*>(saved), t, cond, spk, grad_out, grad_flat, grad_cond,
                               grad_spk, grad_x, B, T, seg_begin, seg_end, workspace, s);
  return bf16_train_backward(*dims, flat, saved, t, cond, spk, grad_out, grad_flat, grad_cond, grad_spk, grad_x, B, T,
                             seg_begin, seg_end, workspace, s);
}

/* Debug: watchdog word of the bf16 training kernels (0 = every kernel completed its barrier protocol). */
int mgb_train_debug_status(const mgb_model_dims* dims, int B, int T, const void* workspace, int* host_status) {
  MGB_REQUIRE(dims_supported(dims) && workspace && host_status && B > 0 && T > 0, MGB_E_ARG, "bad argument");
  MGB_CUDA_CHECK(cudaDeviceSynchronize());
  MGB_CUDA_CHECK(cudaMemcpy(host_status, static_cast<const uint8_t*>(workspace) + bf16_train_status_offset(*dims, B, T),
                            sizeof(int), cudaMemcpyDeviceToHost));
  return MGB_OK;
}

int mgb_pack_cond(const mgb_model_dims* dims, int precision, const float* cond, int B, int T, void* workspace,
                  size_t workspace_bytes, void* stream) {
  MGB_REQUIRE(dims_supported(dims) && cond && workspace && B > 0 && T > 0, MGB_E_ARG, "bad argument");
  MGB_REQUIRE(prec_is_tc(precision), MGB_E_ARG, "mgb_pack_cond: only the tensor-core precisions keep a conditioner image");
  MGB_REQUIRE(workspace_bytes >= mgb_workspace_bytes(dims, precision, B, T, 1), MGB_E_WORKSPACE, "workspace too small");
  if (int rc = check_arch()) return rc;
  return bf16_pack_cond(*dims, cond, B, T, workspace, static_cast<cudaStream_t>(stream), precision == MGB_PREC_FP16);
}

int mgb_shallow_start(const float* coarse, const float* noise, const float* spec_min, const float* spec_max,
                      float sqrt_acp, float sqrt_1m_acp, const uint8_t* pad_mask, float* x_T, int B, int T,
                      int n_mel, void* stream) {
  MGB_REQUIRE(coarse && noise && spec_min && spec_max && x_T, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0 && n_mel > 0, MGB_E_ARG, "bad shape");
  if (int rc = check_arch()) return rc;
  MGB_REQUIRE(n_mel <= EW_MAXM, MGB_E_UNSUPPORTED, "n_mel %d > %d", n_mel, EW_MAXM);
  if (rt_ok(n_mel, T, coarse, noise, x_T, spec_min, spec_max) && !(pad_mask && (reinterpret_cast<uintptr_t>(pad_mask) & 3))) {
    shallow_start_rt_kernel<1><<<rt_blocks(B, n_mel, T, 16), RT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        coarse, noise, spec_min, spec_max, sqrt_acp, sqrt_1m_acp, pad_mask, x_T, B, n_mel, T);
    note_launch();
    MGB_LAUNCH_CHECK();
    return MGB_OK;
  }
  dim3 grid((T + EW_TF - 1) / EW_TF, B), block(EW_THREADS);
  if (n_mel == 80)
    shallow_start_kernel<80><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        coarse, noise, spec_min, spec_max, sqrt_acp, sqrt_1m_acp, pad_mask, x_T, n_mel, T);
  else
    shallow_start_kernel<0><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        coarse, noise, spec_min, spec_max, sqrt_acp, sqrt_1m_acp, pad_mask, x_T, n_mel, T);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_denorm_mask(const float* x, const float* spec_min, const float* spec_max, const uint8_t* pad_mask,
                    float* mel, int B, int T, int n_mel, void* stream) {
  MGB_REQUIRE(x && spec_min && spec_max && mel, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && T > 0 && n_mel > 0, MGB_E_ARG, "bad shape");
  if (int rc = check_arch()) return rc;
  MGB_REQUIRE(n_mel <= EW_MAXM, MGB_E_UNSUPPORTED, "n_mel %d > %d", n_mel, EW_MAXM);
  launch_denorm_mask(x, spec_min, spec_max, pad_mask, mel, B, n_mel, T, static_cast<cudaStream_t>(stream));
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_durations_from_log(const float* log_d, float d_control, int64_t* dur, int n, void* stream) {
  MGB_REQUIRE(log_d && dur && n > 0, MGB_E_ARG, "bad argument");
  if (int rc = check_arch()) return rc;
  lr_durations_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(log_d, d_control, dur, n);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_length_regulate(const float* x, const int64_t* dur, float* out, int64_t* mel_len, uint8_t* mask_valid, int B, int S,
                        int D, int max_len, void* workspace, size_t workspace_bytes, void* stream) {
  MGB_REQUIRE(x && dur && out && mel_len && workspace, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && S > 0 && D > 0 && max_len > 0, MGB_E_ARG, "bad shape");
  MGB_REQUIRE(workspace_bytes >= (size_t)B * (S + 1) * sizeof(int64_t), MGB_E_WORKSPACE,
              "workspace too small for the duration scan");
  if (int rc = check_arch()) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t* cum = static_cast<int64_t*>(workspace);
  lr_scan_kernel<<<B, 32, 0, s>>>(dur, cum, mel_len, S);
  dim3 grid(max_len, B);
  lr_gather_kernel<<<grid, 64, 0, s>>>(x, cum, out, mask_valid, S, D, max_len);
  note_launch(2);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_length_regulate_backward(const float* grad_out, const void* workspace, float* grad_x, int B, int S, int D,
                                 int max_len, void* stream) {
  MGB_REQUIRE(grad_out && workspace && grad_x, MGB_E_ARG, "NULL pointer argument");
  MGB_REQUIRE(B > 0 && S > 0 && D > 0 && max_len > 0, MGB_E_ARG, "bad shape");
  if (int rc = check_arch()) return rc;
  dim3 grid(S, B);
  lr_backward_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, static_cast<const int64_t*>(workspace),
                                                                           grad_x, S, D, max_len);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int mgb_mask_from_lengths(const int64_t* lengths, uint8_t* mask_valid, int B, int max_len, void* stream) {
  MGB_REQUIRE(lengths && mask_valid && B > 0 && max_len > 0, MGB_E_ARG, "bad argument");
  if (int rc = check_arch()) return rc;
  const int n = B * max_len;
  mask_from_lengths_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(lengths, mask_valid, B, max_len);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

}  // extern "C"

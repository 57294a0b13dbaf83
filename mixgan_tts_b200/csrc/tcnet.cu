// Tensor-core Conv1d / Linear engine + self-attention kernel (tcnet.cuh header comment has the layouts).
#include "tcnet.cuh"

#include "tc05.cuh"
#include "tmap.cuh"

namespace mgb {
namespace tcnet {
namespace {

constexpr long long kTimeout = 400000000LL;   // ~0.2 s of SM cycles: a protocol bug traps instead of hanging

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void tmem_ld_f32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tc::tmem_ld32(taddr, r);
  tc::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void load_f32x32(const float* p, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(p + q * 4);
    v[q * 4] = a.x; v[q * 4 + 1] = a.y; v[q * 4 + 2] = a.z; v[q * 4 + 3] = a.w;
  }
}
// fp32 stream [C/4][Rp][4]: a thread owns one row, so a warp's 16-byte access to chunk q is 512 contiguous bytes
__device__ __forceinline__ void load_s32(const float* base, size_t Rp, size_t row, int ch0, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(base + ((size_t)(ch0 / 4 + q) * Rp + row) * 4);
    v[q * 4] = a.x; v[q * 4 + 1] = a.y; v[q * 4 + 2] = a.z; v[q * 4 + 3] = a.w;
  }
}
__device__ __forceinline__ void store_s32(float* base, size_t Rp, size_t row, int ch0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(base + ((size_t)(ch0 / 4 + q) * Rp + row) * 4) =
        make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
}
// fp16 image [C/8][Rp][8]
__device__ __forceinline__ void store_img32(__half* img, size_t Rp, size_t row, int ch0, const float (&v)[32], float slope) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float x = v[q * 8 + e]; w[e] = x > 0.f ? x : x * slope; }
    uint4 u = make_uint4(pack_h2(w[0], w[1]), pack_h2(w[2], w[3]), pack_h2(w[4], w[5]), pack_h2(w[6], w[7]));
    *reinterpret_cast<uint4*>(img + ((size_t)(ch0 / 8 + q) * Rp + row) * 8) = u;
  }
}

struct KArgs {
  const __half* Wpk; const float* bias;
  int taps, dil, kspt;
  int B, T, Tg;
  const int* lens; int len_mul;
  int Cout, Cout32;            // channels of the output tensor, rounded up to 32
  int act; float scale;
  const float* res1; const float* res2;
  const float* ln_g; const float* ln_b;
  float* stream_out; __half* img_out; float img_slope;
  float* user_out; int user_ld;
  int up; long long Rp_out;
  int* status;
};

template <int NT, int KC>
struct Smem {
  static constexpr int A_BYTES = TILE * KC * 2;
  static constexpr int B_BYTES = NT * KC * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int RAW = (100 * 1024) / STAGE;
  static constexpr int STAGES = RAW < 2 ? 2 : (RAW > 6 ? 6 : RAW);
  static constexpr int TOTAL = STAGES * STAGE + 1024;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

template <int NT>
__device__ __forceinline__ void epilogue(const KArgs& p, uint32_t trow, int tile, int ntile, int i) {
  const long long row = (long long)tile * TILE + i;
  const int b = (int)(row / p.Tg), t = (int)(row - (long long)b * p.Tg);
  const bool inrange = b < p.B && t < p.T;
  bool valid = inrange;
  if (valid && p.lens) valid = t < p.lens[b] * p.len_mul;
  const long long orow = p.up > 1 ? row * p.up + ntile : row;
  const int nbase = p.up > 1 ? 0 : ntile * NT;
  const int bcol = ntile * NT;
  const size_t Rp = (size_t)p.Rp_out;
  float* uo = nullptr;
  if (p.user_out && inrange) uo = p.user_out + ((size_t)((long long)b * p.T + t) * p.up + (p.up > 1 ? ntile : 0)) * p.user_ld;

  if (p.ln_g != nullptr) {
    if constexpr (NT == 256) {
      // LayerNorm(acc + bias + residual) over the 256 channels of this thread's row (nn.LayerNorm: biased variance, eps 1e-5)
      float mean = 0.f;
#pragma unroll 1
      for (int cg = 0; cg < 8; ++cg) {
        float v[32], bv[32];
        tmem_ld_f32x32(trow + cg * 32, v);
        load_f32x32(p.bias + cg * 32, bv);
        if (valid && p.res1) {
          float r[32];
          load_s32(p.res1, Rp, (size_t)orow, cg * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) mean += v[j] + bv[j];
      }
      mean *= (1.f / 256.f);
      float var = 0.f;
#pragma unroll 1
      for (int cg = 0; cg < 8; ++cg) {
        float v[32], bv[32];
        tmem_ld_f32x32(trow + cg * 32, v);
        load_f32x32(p.bias + cg * 32, bv);
        if (valid && p.res1) {
          float r[32];
          load_s32(p.res1, Rp, (size_t)orow, cg * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = v[j] + bv[j] - mean; var += d * d; }
      }
      const float rstd = rsqrtf(var * (1.f / 256.f) + 1e-5f);
#pragma unroll 1
      for (int cg = 0; cg < 8; ++cg) {
        float v[32], bv[32], g[32], be[32];
        tmem_ld_f32x32(trow + cg * 32, v);
        load_f32x32(p.bias + cg * 32, bv);
        load_f32x32(p.ln_g + cg * 32, g);
        load_f32x32(p.ln_b + cg * 32, be);
        if (valid && p.res1) {
          float r[32];
          load_s32(p.res1, Rp, (size_t)orow, cg * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = valid ? (v[j] + bv[j] - mean) * rstd * g[j] + be[j] : 0.f;
        if (p.stream_out) store_s32(p.stream_out, Rp, (size_t)orow, cg * 32, v);
        if (p.img_out) store_img32(p.img_out, Rp, (size_t)orow, cg * 32, v, p.img_slope);
        if (uo) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(uo + cg * 32 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
      }
    }
    return;
  }

#pragma unroll 1
  for (int cg = 0; cg < NT / 32; ++cg) {
    const int n0 = nbase + cg * 32;
    if (n0 >= p.Cout32) break;                       // uniform over the warp
    float v[32], bv[32];
    tmem_ld_f32x32(trow + cg * 32, v);
    load_f32x32(p.bias + bcol + cg * 32, bv);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j] + bv[j], p.act);
    if (valid && p.res1) {
      float r[32];
      load_s32(p.res1, Rp, (size_t)orow, n0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r[j];
    }
    if (valid && p.res2) {
      float r[32];
      load_s32(p.res2, Rp, (size_t)orow, n0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = valid ? v[j] * p.scale : 0.f;
    if (p.stream_out) store_s32(p.stream_out, Rp, (size_t)orow, n0, v);
    if (p.img_out) store_img32(p.img_out, Rp, (size_t)orow, n0, v, p.img_slope);
    if (uo) {
      if (n0 + 32 <= p.Cout && (p.user_ld & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(uo + n0 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.Cout) uo[n0 + j] = v[j];
      }
    }
  }
}

template <int NT, int KC>
__global__ void __launch_bounds__(192) tcconv_kernel(const KArgs p, const __grid_constant__ CUtensorMap tmA) {
  using S = Smem<NT, KC>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_full[S::STAGES], bar_empty[S::STAGES], bar_acc;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, ntile = blockIdx.y;

  if (warp == 1) tc::tmem_alloc<NT>(&tmem_slot);
  if (tid == 0) {
    for (int i = 0; i < S::STAGES; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_empty[i], 1); }
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int nsteps = p.taps * p.kspt;
  const uint32_t smem_base = tc::smem_u32(smem);

  if (warp == 0) {
    if (lane == 0) {
      const __half* bsrc = p.Wpk + (size_t)ntile * nsteps * (size_t)(NT * KC);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % S::STAGES, ph = (s / S::STAGES) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&bar_empty[stage]), ph ^ 1, kTimeout, p.status, 1);
        const int tap = s / p.kspt, kc = s - tap * p.kspt;
        const int off = (tap - (p.taps >> 1)) * p.dil;
        const uint32_t sa = smem_base + stage * S::STAGE, sb = sa + S::A_BYTES;
        const uint32_t fb = tc::smem_u32(&bar_full[stage]);
        tc::mbar_arrive_expect_tx_addr(fb, S::STAGE);
        tc::tma_load_2d(sa, &tmA, 2 * (tile * TILE + off), kc * (KC / 8), fb);   // 128 rows x KC channels, zero-filled outside
        tc::bulk_g2s_addr(sb, bsrc + (size_t)s * (NT * KC), S::B_BYTES, fb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_16(128, NT, true);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % S::STAGES, ph = (s / S::STAGES) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&bar_full[stage]), ph, kTimeout, p.status, 2);
        tc::tc_fence_after();
        const uint32_t sa = smem_base + stage * S::STAGE, sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
          const uint64_t ad = tc::make_smem_desc(sa + k * 2 * (TILE * 16), TILE * 16, 128);
          const uint64_t bd = tc::make_smem_desc(sb + k * 2 * (NT * 16), NT * 16, 128);
          tc::umma_bf16(tmem, ad, bd, idesc, (s | k) ? 1u : 0u);
        }
        tc::umma_commit(&bar_empty[stage]);
      }
      tc::umma_commit(&bar_acc);
    }
  } else {
    tc::mbar_wait_trap(tc::smem_u32(&bar_acc), 0, kTimeout, p.status, 4);
    tc::tc_fence_after();
    const int i = (warp & 3) * 32 + lane;                  // TMEM lane = tile row
    epilogue<NT>(p, tmem + ((uint32_t)((warp & 3) * 32) << 16), tile, ntile, i);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<NT>(tmem);
}

template <int NT, int KC>
int launch_conv(const KArgs& a, const __half* in, int in_chunks, int Rp_in, int ntiles, int ntn, cudaStream_t s) {
  using S = Smem<NT, KC>;
  static PerDeviceOnce once;
  if (once.pending()) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(tcconv_kernel<NT, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    once.done();
  }
  CUtensorMap m;
  if (int rc = make_image_map(&m, in, in_chunks, Rp_in, KC / 8)) return rc;
  MGB_CUDA_CHECK(launch_pdl(tcconv_kernel<NT, KC>, dim3(ntiles, ntn), dim3(192), S::TOTAL, s, 1, a, m));
  note_launch();
  return MGB_OK;
}

// ---- weight packing: torch [Cout][Cin][k] (conv / linear) or [Cin][Cout][k] (transposed conv) -> streamed fp16 tiles
//      wp[ntile][step = tap*kspt + kc][KC/8][NT][8] -------------------------------------------------------------------
__global__ void pack_w_kernel(const float* __restrict__ w, const float* __restrict__ oscale, __half* __restrict__ wp,
                              int Cin, int Cout, int k, int NT, int KC, int kspt, int taps, int up, int tpad, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx % 8);
  const int nl = (int)((idx / 8) % NT);
  const int c8 = (int)((idx / (8LL * NT)) % (KC / 8));
  const long long per_step = (long long)NT * KC;
  const int nsteps = taps * kspt;
  const int s = (int)((idx / per_step) % nsteps);
  const int nt = (int)(idx / (per_step * nsteps));
  const int tap = s / kspt, kc = s - tap * kspt;
  const int ci = kc * KC + c8 * 8 + e;
  float val = 0.f;
  if (up > 1) {
    const int co = nl, j = nt + tpad - up * (tap - 1);       // output phase nt, input offset tap - 1
    if (ci < Cin && co < Cout && j >= 0 && j < k) val = w[((size_t)ci * Cout + co) * k + j];
  } else {
    const int co = nt * NT + nl;
    if (ci < Cin && co < Cout) val = w[((size_t)co * Cin + ci) * k + tap] * (oscale ? oscale[co] : 1.f);
  }
  wp[idx] = __float2half_rn(val);
}
__global__ void pack_b_kernel(const float* __restrict__ bias, const float* __restrict__ oscale, const float* __restrict__ oshift,
                              float* __restrict__ bp, int Cout, int NT, int ntn, int up) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntn * NT) return;
  const int co = up > 1 ? i % NT : i;
  float v = 0.f;
  if (co < Cout) {
    v = bias ? bias[co] : 0.f;
    if (oscale) v = v * oscale[co] + oshift[co];
  }
  bp[i] = v;
}

// thread = (8-channel chunk, row), rows fastest: a warp writes 512 contiguous bytes of one image chunk and reads one
// full 32-byte sector per thread
__global__ void pack_rows_kernel(const float* __restrict__ user, const float* __restrict__ pos, int C, int B, int T, int Tg,
                                 long long Rp, __half* __restrict__ img, float slope, float* __restrict__ stream) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nch = C / 8;
  if (idx >= Rp * nch) return;
  const long long row = idx % Rp;
  const int ch = (int)(idx / Rp);
  const int b = (int)(row / Tg), t = (int)(row - (long long)b * Tg);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (b < B && t < T) {
    const float* src = user + ((size_t)b * T + t) * C + ch * 8;
    const float4 a = *reinterpret_cast<const float4*>(src), c = *reinterpret_cast<const float4*>(src + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    if (pos) {
      const float* ps = pos + (size_t)t * C + ch * 8;
      const float4 pa = *reinterpret_cast<const float4*>(ps), pc = *reinterpret_cast<const float4*>(ps + 4);
      v[0] += pa.x; v[1] += pa.y; v[2] += pa.z; v[3] += pa.w; v[4] += pc.x; v[5] += pc.y; v[6] += pc.z; v[7] += pc.w;
    }
  }
  if (stream) {
    *reinterpret_cast<float4*>(stream + ((size_t)(ch * 2) * Rp + row) * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(stream + ((size_t)(ch * 2 + 1) * Rp + row) * 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (img) {
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = v[e] > 0.f ? v[e] : v[e] * slope;
    *reinterpret_cast<uint4*>(img + ((size_t)ch * Rp + row) * 8) =
        make_uint4(pack_h2(w[0], w[1]), pack_h2(w[2], w[3]), pack_h2(w[4], w[5]), pack_h2(w[6], w[7]));
  }
}

}  // namespace

int pack_rows(const float* user, const float* pos, int C, const Rows& r, __half* img, float img_slope, float* stream,
              cudaStream_t s) {
  MGB_REQUIRE(C % 8 == 0, MGB_E_UNSUPPORTED, "pack_rows: channels must be a multiple of 8");
  const long long total = (long long)r.Rp * (C / 8);
  pack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(user, pos, C, r.B, r.T, r.Tg, r.Rp, img, img_slope, stream);
  note_launch();
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int pack_conv(const Layer& l, void* packed, const float* w, const float* bias, const float* oscale, const float* oshift,
              int ntile0, int Cout_part, cudaStream_t s) {
  const int nsteps = l.taps * l.kspt;
  const int ntn = l.up > 1 ? l.ntn : (Cout_part + l.NT - 1) / l.NT;
  const long long total = (long long)ntn * nsteps * l.NT * l.KC;
  __half* wp = static_cast<__half*>(packed) + l.w_off + (size_t)ntile0 * nsteps * l.NT * l.KC;
  pack_w_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, oscale, wp, l.Cin, Cout_part, l.k, l.NT, l.KC, l.kspt,
                                                                 l.taps, l.up, l.tpad, total);
  float* bp = reinterpret_cast<float*>(packed) + l.b_off + (size_t)ntile0 * l.NT;
  pack_b_kernel<<<(ntn * l.NT + 255) / 256, 256, 0, s>>>(bias, oscale, oshift, bp, Cout_part, l.NT, ntn, l.up);
  note_launch(2);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

int run_conv(const Layer& l, const void* packed, const Rows& rin, const ConvIO& io, int* status, cudaStream_t s) {
  KArgs a{};
  a.Wpk = static_cast<const __half*>(packed) + l.w_off;
  a.bias = reinterpret_cast<const float*>(packed) + l.b_off;
  a.taps = l.taps; a.dil = l.dil; a.kspt = l.kspt;
  a.B = rin.B; a.T = rin.T; a.Tg = rin.Tg;
  a.lens = io.lens; a.len_mul = io.len_mul;
  a.Cout = l.Cout; a.Cout32 = (l.Cout + 31) / 32 * 32;
  a.act = io.act; a.scale = io.scale;
  a.res1 = io.res1; a.res2 = io.res2; a.ln_g = io.ln_g; a.ln_b = io.ln_b;
  a.stream_out = io.stream_out; a.img_out = io.img_out; a.img_slope = io.img_slope;
  a.user_out = io.user_out; a.user_ld = io.user_ld;
  a.up = l.up; a.Rp_out = (long long)rin.Rp * l.up;
  a.status = status;
  MGB_REQUIRE(io.ln_g == nullptr || (l.NT == 256 && l.Cout == 256 && l.up == 1), MGB_E_UNSUPPORTED,
              "the fused LayerNorm epilogue needs a 256-channel output");
  MGB_REQUIRE(io.in_chunks * 8 >= l.Cin, MGB_E_ARG, "input image narrower than the layer's input channels");
#define MGB_TC_CASE(NT_, KC_)                                                                       \
  if (l.NT == NT_ && l.KC == KC_) return launch_conv<NT_, KC_>(a, io.in, io.in_chunks, rin.Rp, rin.ntiles, l.ntn, s);
  MGB_TC_CASE(256, 64) MGB_TC_CASE(256, 32) MGB_TC_CASE(128, 64) MGB_TC_CASE(128, 32)
  MGB_TC_CASE(64, 64) MGB_TC_CASE(64, 32) MGB_TC_CASE(32, 64) MGB_TC_CASE(32, 32)
#undef MGB_TC_CASE
  MGB_REQUIRE(false, MGB_E_UNSUPPORTED, "no tensor-core tile for NT=%d KC=%d", l.NT, l.KC);
}

}  // namespace tcnet
}  // namespace mgb

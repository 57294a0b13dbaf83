// bf16 tensor-core training path of the Denoiser (BASELINE configs[4]): forward with an activation stash and the
// backward, every GEMM on tcgen05 (bf16 operands staged by the TMA engine's bulk copies, fp32 accumulation in TMEM),
// fp32 residual / skip / gradient streams and fp32 weight gradients.  Same algebra and segment structure as the fp32
// path (train_fp32.cu header); reference: autograd through model/modules.py:420-446, model/blocks.py:1157-1176.
//
// One activation layout serves both GEMM families.  Activations live as bf16 "images" [C/8][Rp][8] on a row axis that
// holds all utterances (utterance b owns rows RLEAD + b*(T+1) + [0,T); the row between two utterances and the margins
// are zero, which is the convolution's zero padding):
//   * frames GEMMs (forward, data gradients): a 128-row x 8-channel block is one contiguous 2 KB bulk copy that lands
//     in the K-major no-swizzle core-matrix order (LBO = 2048, SBO = 128); a convolution tap is the same copy one
//     row earlier/later;
//   * weight-gradient GEMMs (reduction over frames): the very same block is an MN-major operand whose K axis is the
//     frame axis (LBO = 128, SBO = 2048; pinned by tests/test_umma_probe.py::test_mn_major_operands), so dW = P^T Q
//     needs no transposed copy of any activation.  The frame axis is split over CTAs; fp32 partials are summed in a
//     fixed order (train_small.cuh) straight into the flat gradient.
// Weights are re-packed to bf16 K-major tiles once per step (they change every step).
#include "common.cuh"
#include "tc05.cuh"
#include "tmap.cuh"
#include "train_small.cuh"

#include <cstdlib>

namespace mgb {

namespace {

using namespace trainsmall;
using bf16 = __nv_bfloat16;

constexpr int C = 256;            // channels == d_encoder (dims_supported)
constexpr int RLEAD = 8;          // zero rows in front of the first utterance
constexpr int TILE = 128;         // rows (frames) per CTA tile
constexpr int FCL = 4;            // frames-GEMM cluster: 4 row tiles share every weight tile (one multicast quarter each)
constexpr long long kTimeout = 200000000LL;    // ~0.1 s of SM cycles: a protocol bug traps instead of hanging
constexpr float RSQRT2 = 0.70710678118654752440f;

// Debug library only (MGB_DEBUG_BUILD): MGB_TRAIN_TRACE=1 names every launch on stderr and synchronises after it
#ifdef MGB_DEBUG_BUILD
bool trace_on() {
  static const bool on = [] { const char* e = getenv("MGB_TRAIN_TRACE"); return e && *e == '1'; }();
  return on;
}
#else
constexpr bool trace_on() { return false; }   // the product library never synchronises
#endif
void trace(const char* what, cudaStream_t s) {
#ifdef MGB_DEBUG_BUILD
  if (!trace_on()) return;
  fprintf(stderr, "[mgb train] %s ...", what);
  fflush(stderr);
  const cudaError_t e = cudaStreamSynchronize(s);
  fprintf(stderr, " %s\n", cudaGetErrorString(e));
  fflush(stderr);
#else
  (void)what; (void)s;
#endif
}

struct RowSpace { int T, Tg, R, ntiles, Rp; };
RowSpace row_space(int B, int T) {
  RowSpace r{};
  r.T = T; r.Tg = T + 1; r.R = B * r.Tg;
  r.ntiles = ((r.R + TILE - 1) / TILE + FCL - 1) / FCL * FCL;
  r.Rp = RLEAD + r.ntiles * TILE + 8;
  return r;
}

__device__ __forceinline__ float tanh_mufu(float x) {   // MUFU.TANH, as the inference kernel's gate
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
// 32 consecutive channels [ch0, ch0+32) of image row `rho` <-> registers
__device__ __forceinline__ void store_img32(bf16* img, int Rp, int rho, int ch0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u = make_uint4(pack2(v[q * 8], v[q * 8 + 1]), pack2(v[q * 8 + 2], v[q * 8 + 3]),
                         pack2(v[q * 8 + 4], v[q * 8 + 5]), pack2(v[q * 8 + 6], v[q * 8 + 7]));
    *reinterpret_cast<uint4*>(img + ((size_t)(ch0 / 8 + q) * Rp + rho) * 8) = u;
  }
}
__device__ __forceinline__ void load_img32(const bf16* img, int Rp, int rho, int ch0, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(img + ((size_t)(ch0 / 8 + q) * Rp + rho) * 8);
    float2 a = unpack2(u.x), b = unpack2(u.y), c = unpack2(u.z), d = unpack2(u.w);
    v[q * 8] = a.x; v[q * 8 + 1] = a.y; v[q * 8 + 2] = b.x; v[q * 8 + 3] = b.y;
    v[q * 8 + 4] = c.x; v[q * 8 + 5] = c.y; v[q * 8 + 6] = d.x; v[q * 8 + 7] = d.y;
  }
}
// 32 consecutive floats of a row-major row (user tensors, bias vectors, per-utterance tables)
__device__ __forceinline__ void load_f32x32(const float* p, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(p + q * 4);
    v[q * 4] = a.x; v[q * 4 + 1] = a.y; v[q * 4 + 2] = a.z; v[q * 4 + 3] = a.w;
  }
}
__device__ __forceinline__ void store_f32x32(float* p, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(p + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
}
// fp32 streams (x, skip sum, e, X0) and the (sigmoid, tanh) pairs live in a chunked layout [C/4][Rp][4]: one thread owns
// one row, so a warp's 16-byte access to chunk q is 512 contiguous bytes (a row-major [Rp][C] array would cost 32
// separate sectors per warp instruction)
__device__ __forceinline__ void load_s32(const float* base, int Rp, int rho, int ch0, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(base + ((size_t)(ch0 / 4 + q) * Rp + rho) * 4);
    v[q * 4] = a.x; v[q * 4 + 1] = a.y; v[q * 4 + 2] = a.z; v[q * 4 + 3] = a.w;
  }
}
__device__ __forceinline__ void store_s32(float* base, int Rp, int rho, int ch0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(base + ((size_t)(ch0 / 4 + q) * Rp + rho) * 4) =
        make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
}
__device__ __forceinline__ void tmem_ld_f32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tc::tmem_ld32(taddr, r);
  tc::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// =====================================================================================================
// frames GEMM: D[128 rows][NT cols] = sum over k-steps of A[128][64] * B[NT][64]^T
// =====================================================================================================
enum FMode {
  F_IN = 0, F_COND, F_GATE, F_OUT, F_SKIP, F_FINAL,              // forward
  B_PMASK, B_DS, B_GATE, B_DX, B_COND, B_DXT                      // backward (data gradients)
};

struct FArgs {
  const bf16* A0; const bf16* A1;   // activation images; K = taps x (steps0 from A0, then steps1 from A1) x 64 channels
  int steps0, steps1, taps;
  const bf16* Bpk;                  // packed weights [n-tile][ksteps_b][8][NT][8]
  int ksteps_b, kstep_b0;
  int Rp, B, T, L;
  int* status;
  // epilogue operands (meaning depends on the mode; see fgemm_epilogue)
  const float* bias;                // flat-order bias vector
  const float* dtab; const float* ctab;   // [B][L][C] tables, already offset to the layer
  const float* dtab2; const float* ctab2; // F_IN / F_OUT: the tables of the layer whose conv input y this epilogue also writes
  const float* call;                      // F_IN / F_OUT: that layer's slot of the hoisted conditioner projection (fp32 stream)
  const float* fin;                 // fp32 [Rp][C] input stream (x / e / X0)
  float* fout;                      // fp32 [Rp][C] output stream (x / e) or [B][n_mel][T] / [B][T][C] user tensors
  float* fout2;                     // fp32 [Rp][C] skip sum
  bf16* img; bf16* img2;            // output images
  const bf16* aux_img;              // input image (P for the ReLU mask)
  uint32_t* sgth; const uint32_t* sgth_in;   // [Rp][C] (sigmoid, tanh) bf16 pairs of the gate
  int first, last, relu0, n_mel;
  float scale;
};

template <int NT>
struct FSmem {
  static constexpr int A_BYTES = 8 * TILE * 16;       // 64 channels x 128 rows
  static constexpr int B_BYTES = NT * 128;            // NT rows x 64 k
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = NT == 256 ? 4 : 5;
  static constexpr int TOTAL = STAGES * STAGE + 1024;
};

template <int MODE>
__device__ __forceinline__ void fgemm_epilogue(const FArgs& p, uint32_t tmem_row, int rho, int b, int t, bool valid,
                                               int ntile) {
  const int L = p.L;
  if constexpr (MODE == F_IN || MODE == F_SKIP) {
    // relu(acc + bias): F_IN -> fp32 X0 stream (+ the first block's conv input y_0); F_SKIP -> P image
#pragma unroll 2
    for (int cg = 0; cg < 8; ++cg) {
      float v[32];
      tmem_ld_f32x32(tmem_row + cg * 32, v);
      float bv[32];
      load_f32x32(p.bias + cg * 32, bv);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = valid ? fmaxf(v[j] + bv[j], 0.f) : 0.f;
      if constexpr (MODE == F_IN) {
        store_s32(p.fout, p.Rp, rho, cg * 32, v);
        if (valid) {            // y_0 = (x_0 + d_0) + (Wc_0 cond + bc_0 [+ s_0])   (blocks.py:1166-1168)
          float dt[32], ct[32], cc[32];
          load_s32(p.call, p.Rp, rho, cg * 32, cc);
          load_f32x32(p.dtab2 + (size_t)b * L * C + cg * 32, dt);
          load_f32x32(p.ctab2 + (size_t)b * L * C + cg * 32, ct);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (v[j] + dt[j]) + (cc[j] + ct[j]);
        }
        store_img32(p.img2, p.Rp, rho, cg * 32, v);
      } else {
        store_img32(p.img, p.Rp, rho, cg * 32, v);
      }
    }
  } else if constexpr (MODE == F_COND) {
    // hoisted conditioner projection: ONE launch computes Wc_l cond for every layer l = ntile (the projection does not
    // depend on the residual stream); fp32, chunked layout, consumed by the F_IN / F_OUT epilogues that form y_l
    float* dst = p.fout + (size_t)ntile * p.Rp * C;
#pragma unroll 2
    for (int cg = 0; cg < 8; ++cg) {
      float v[32];
      tmem_ld_f32x32(tmem_row + cg * 32, v);
      store_s32(dst, p.Rp, rho, cg * 32, v);
    }
  } else if constexpr (MODE == F_GATE) {
    // tile = 128 gate + 128 filter columns of channels [ntile*128, +128): g = sigmoid(a) tanh(f)  (blocks.py:1170-1171)
#pragma unroll 2
    for (int cg = 0; cg < 4; ++cg) {
      const int ch0 = ntile * 128 + cg * 32;
      float a[32], f[32];
      tmem_ld_f32x32(tmem_row + cg * 32, a);
      tmem_ld_f32x32(tmem_row + 128 + cg * 32, f);
      uint32_t st[32];
      float ba[32], bf[32];
      load_f32x32(p.bias + ch0, ba);
      load_f32x32(p.bias + C + ch0, bf);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float sg = fmaf(0.5f, tanh_mufu(0.5f * (a[j] + ba[j])), 0.5f);   // sigmoid(a) = (1 + tanh(a / 2)) / 2
        const float th = tanh_mufu(f[j] + bf[j]);
        st[j] = pack2(sg, th);
        a[j] = valid ? sg * th : 0.f;
      }
      store_img32(p.img, p.Rp, rho, ch0, a);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(p.sgth + ((size_t)(ch0 / 4 + q) * p.Rp + rho) * 4) =
            make_uint4(st[q * 4], st[q * 4 + 1], st[q * 4 + 2], st[q * 4 + 3]);
    }
  } else if constexpr (MODE == F_OUT) {
    // tile = 128 x-columns + 128 skip columns: x' = (o_x + (x + d_l)) / sqrt(2), skip sum += o_s   (blocks.py:1173-1176)
#pragma unroll 2
    for (int cg = 0; cg < 4; ++cg) {
      const int ch0 = ntile * 128 + cg * 32;
      float ox[32], os[32], x[32];
      tmem_ld_f32x32(tmem_row + cg * 32, ox);
      tmem_ld_f32x32(tmem_row + 128 + cg * 32, os);
      if (valid) {
        float dt[32], bx[32], bs[32], sk[32];
        load_s32(p.fin, p.Rp, rho, ch0, x);
        if (!p.first) load_s32(p.fout2, p.Rp, rho, ch0, sk);
        load_f32x32(p.dtab + (size_t)b * L * C + ch0, dt);
        load_f32x32(p.bias + ch0, bx);
        load_f32x32(p.bias + C + ch0, bs);
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = ((ox[j] + bx[j]) + (x[j] + dt[j])) * RSQRT2;
        store_s32(p.fout, p.Rp, rho, ch0, x);
        if (!p.last) {          // conv input of the next block: y_{l+1} = (x_{l+1} + d_{l+1}) + (Wc_{l+1} cond + bc_{l+1} [+ s])
          float dn[32], cn[32], cc[32], y[32];
          load_s32(p.call, p.Rp, rho, ch0, cc);
          load_f32x32(p.dtab2 + (size_t)b * L * C + ch0, dn);
          load_f32x32(p.ctab2 + (size_t)b * L * C + ch0, cn);
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = (x[j] + dn[j]) + (cc[j] + cn[j]);
          store_img32(p.img2, p.Rp, rho, ch0, y);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          os[j] = (p.first ? 0.f : sk[j]) + (os[j] + bs[j]);
          if (p.last) os[j] *= p.scale;
        }
        store_s32(p.fout2, p.Rp, rho, ch0, os);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) os[j] = 0.f;
        if (!p.last) store_img32(p.img2, p.Rp, rho, ch0, os);      // zero rows between utterances = the conv's padding
      }
      if (p.last) store_img32(p.img, p.Rp, rho, ch0, os);
    }
  } else if constexpr (MODE == F_FINAL || MODE == B_DXT) {
    // [B][n_mel][T] user tensor, coalesced over frames: out = acc (+ bias)
#pragma unroll 1
    for (int cg = 0; cg < 4; ++cg) {
      if (cg * 32 >= p.n_mel) break;
      float v[32];
      tmem_ld_f32x32(tmem_row + cg * 32, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int m = cg * 32 + j;
          if (m < p.n_mel) p.fout[((size_t)b * p.n_mel + m) * p.T + t] = v[j] + (MODE == F_FINAL ? p.bias[m] : 0.f);
        }
      }
    }
  } else if constexpr (MODE == B_PMASK || MODE == B_DS) {
#pragma unroll 2
    for (int cg = 0; cg < 8; ++cg) {
      float v[32];
      tmem_ld_f32x32(tmem_row + cg * 32, v);
      if constexpr (MODE == B_PMASK) {
        float a[32];
        load_img32(p.aux_img, p.Rp, rho, cg * 32, a);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (valid && a[j] > 0.f) ? v[j] : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = valid ? v[j] * p.scale : 0.f;
      }
      store_img32(p.img, p.Rp, rho, cg * 32, v);
    }
  } else if constexpr (MODE == B_GATE) {
    // dza = dg th sg (1 - sg), dzb = dg sg (1 - th^2) -> dZ image [gate C | filter C], zero outside utterances
#pragma unroll 2
    for (int cg = 0; cg < 8; ++cg) {
      float dg[32], df[32];
      tmem_ld_f32x32(tmem_row + cg * 32, dg);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 u = *reinterpret_cast<const uint4*>(p.sgth_in + ((size_t)(cg * 8 + q) * p.Rp + rho) * 4);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = q * 4 + e;
          const float2 s = unpack2(w[e]);
          const float g = valid ? dg[j] : 0.f;
          dg[j] = g * s.y * s.x * (1.0f - s.x);
          df[j] = g * s.x * (1.0f - s.y * s.y);
        }
      }
      store_img32(p.img, p.Rp, rho, cg * 32, dg);
      store_img32(p.img, p.Rp, rho, C + cg * 32, df);
    }
  } else if constexpr (MODE == B_DX) {
    // dY = acc; dx_l = e_l + dY; e_{l-1} = dx_l / sqrt(2)   (layer 0: ReLU mask of the input projection instead)
#pragma unroll 2
    for (int cg = 0; cg < 8; ++cg) {
      float dy[32], e[32];
      tmem_ld_f32x32(tmem_row + cg * 32, dy);
      if (valid) {
        if (!p.first) load_s32(p.fout, p.Rp, rho, cg * 32, e);
        float x0[32];
        if (p.relu0) load_s32(p.fin, p.Rp, rho, cg * 32, x0);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dx = (p.first ? 0.f : e[j]) + dy[j];
          e[j] = p.relu0 ? (x0[j] > 0.f ? dx : 0.f) : dx * RSQRT2;
        }
        store_s32(p.fout, p.Rp, rho, cg * 32, e);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) { dy[j] = 0.f; e[j] = 0.f; }
      }
      store_img32(p.img, p.Rp, rho, cg * 32, dy);
      store_img32(p.img2, p.Rp, rho, cg * 32, e);
    }
  } else if constexpr (MODE == B_COND) {
    // grad_cond[b][t][:] (+)= dY Wc.  The TMEM load is warp-collective (.sync.aligned): every lane issues it, only the
    // global accesses are predicated on the row being a real frame.
    float* o = p.fout + ((size_t)(valid ? b : 0) * p.T + (valid ? t : 0)) * C;
#pragma unroll 2
    for (int cg = 0; cg < 8; ++cg) {
      float v[32], a[32];
      tmem_ld_f32x32(tmem_row + cg * 32, v);
      if (valid) {
        if (!p.first) {
          load_f32x32(o + cg * 32, a);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += a[j];
        }
        store_f32x32(o + cg * 32, v);
      }
    }
  }
}

// A cluster of CL CTAs (CL consecutive row tiles, same column tile) shares every weight tile: CTA r fetches part r
// and multicasts it into all CL shared memories, so the L2 serves each weight byte once per cluster instead of once
// per CTA (measured: 102 CTAs pulling the same 32 KB tile at once are bound by the few L2 slices that hold it).
// EMPTY barriers count CL arrivals (every CTA's MMA warp multicasts its commit), so a producer overwrites a stage in its
// peers only after all of them have consumed it.  CL = 4 for GEMMs with one column tile; CL = 2 for the N = 512 GEMMs
// (two column tiles): with one 193 KB CTA per SM only ~32 clusters of 4 fit on the 148 SMs at once (sum over GPCs of
// floor(SMs / 4)), and 2 x 26 clusters of 4 would run as two waves.
template <int NT, int MODE, int CL>
__global__ void __launch_bounds__(192, 1) fgemm_kernel(const FArgs p, const __grid_constant__ CUtensorMap tmA0,
                                                       const __grid_constant__ CUtensorMap tmA1) {
  using S = FSmem<NT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_full[S::STAGES], bar_empty[S::STAGES], bar_acc;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, ntile = blockIdx.y;
  const uint32_t rank = tc::cluster_ctarank();
  constexpr uint16_t kMask = (1u << CL) - 1;

  if (warp == 1) tc::tmem_alloc<NT>(&tmem_slot);
  if (tid == 0) {
    for (int i = 0; i < S::STAGES; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_empty[i], CL); }
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();            // every CTA's barriers exist before any multicast / remote arrive
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_trigger();                     // setup done: the next kernel may start its own setup on free SMs ...
  pdl_wait();                        // ... and this one touches global memory only after its predecessor has completed

  const int per_tap = p.steps0 + p.steps1;
  const int nsteps = p.taps * per_tap;
  const uint32_t smem_base = tc::smem_u32(smem);

  if (warp == 0) {
    if (lane == 0) {
      const bf16* bsrc = p.Bpk + ((size_t)ntile * p.ksteps_b + p.kstep_b0) * (size_t)(NT * 64) + (size_t)rank * (NT * 64 / CL);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % S::STAGES, ph = (s / S::STAGES) & 1;
        tc::mbar_wait_cluster_trap(tc::smem_u32(&bar_empty[stage]), ph ^ 1, kTimeout, p.status, 1);
        const int tap = s / per_tap, j = s - tap * per_tap;
        const void* tm = j < p.steps0 ? &tmA0 : &tmA1;
        const int ch0 = (j < p.steps0 ? j : j - p.steps0) * 8;
        const int row = RLEAD + tile * TILE + tap - (p.taps >> 1);
        const uint32_t sa = smem_base + stage * S::STAGE, sb = sa + S::A_BYTES;
        const uint32_t fb = tc::smem_u32(&bar_full[stage]);
        tc::mbar_arrive_expect_tx_addr(fb, S::STAGE);
        tc::tma_load_2d(sa, tm, 2 * row, ch0, fb);          // 128 rows x 64 channels in ONE 16 KB box
        tc::bulk_g2s_multicast_addr(sb + rank * (S::B_BYTES / CL), bsrc + (size_t)s * (NT * 64), S::B_BYTES / CL, fb, kMask);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, NT);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % S::STAGES, ph = (s / S::STAGES) & 1;
        tc::mbar_wait_cluster_trap(tc::smem_u32(&bar_full[stage]), ph, kTimeout, p.status, 2);
        tc::tc_fence_after();
        const uint32_t sa = smem_base + stage * S::STAGE, sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = tc::make_smem_desc(sa + k * 2 * (TILE * 16), TILE * 16, 128);
          const uint64_t bd = tc::make_smem_desc(sb + k * 2 * (NT * 16), NT * 16, 128);
          tc::umma_bf16(tmem, ad, bd, idesc, (s | k) ? 1u : 0u);
        }
        tc::umma_commit_mc_addr(tc::smem_u32(&bar_empty[stage]), kMask);
      }
      tc::umma_commit(&bar_acc);
    }
  } else {
    tc::mbar_wait_trap(tc::smem_u32(&bar_acc), 0, kTimeout, p.status, 4);
    tc::tc_fence_after();
    const int i = (warp & 3) * 32 + lane;                  // TMEM lane = tile row (a warp reaches lanes 32*(warp%4)..)
    const int rho = RLEAD + tile * TILE + i;
    const int r = rho - RLEAD, Tg = p.T + 1;
    const int b = r / Tg, t = r - b * Tg;
    const bool valid = b < p.B && t < p.T;
    fgemm_epilogue<MODE>(p, tmem + ((uint32_t)((warp & 3) * 32) << 16), rho, b, t, valid, ntile);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<NT>(tmem);
  tc::cluster_sync_all();            // no CTA leaves while a peer may still multicast into it or arrive on its barriers
}

template <typename K, typename... A>
cudaError_t launch_cluster(K kernel, dim3 grid, int cluster_x, size_t smem, cudaStream_t s, const A&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int NT, int MODE>
int launch_fgemm(const FArgs& a, int ntiles, int n_tiles_n, cudaStream_t s) {
  constexpr int CL = (MODE == F_GATE || MODE == F_OUT) ? 2 : FCL;
  static PerDeviceOnce once;
  if (once.pending()) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(fgemm_kernel<NT, MODE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        FSmem<NT>::TOTAL));
    once.done();
  }
  CUtensorMap m0, m1;
  if (int rc = make_image_map(&m0, a.A0, a.steps0 * 8, a.Rp, 8)) return rc;
  if (a.steps1 > 0) { if (int rc = make_image_map(&m1, a.A1, a.steps1 * 8, a.Rp, 8)) return rc; }
  else m1 = m0;
  MGB_CUDA_CHECK(launch_pdl(fgemm_kernel<NT, MODE, CL>, dim3(ntiles, n_tiles_n), dim3(192), FSmem<NT>::TOTAL, s, CL, a, m0, m1));
  note_launch();
  if (trace_on()) { char b[64]; snprintf(b, sizeof b, "fgemm<%d, mode %d>", NT, MODE); trace(b, s); }
  return MGB_OK;
}

// =====================================================================================================
// weight-gradient GEMM: part[z][m][n] = sum over the split's frames of P[f][m] * Q[f + sh][ci]
// =====================================================================================================
// Up to three independent problems per launch (a block's dWo, dW3 and dWc share one launch and one reduction):
// blockIdx.x runs over the concatenated (m-tile, n-tile) lists, blockIdx.y over the frame splits.
struct WProb {
  int mtiles, mtiles0;           // M tiles (128 out-channels); tiles < mtiles0 come from P0, the rest from P1
  int taps, ntiles_per_tap, Kin, Mo, N;
  int tile0;                     // first blockIdx.x of this problem
  long long part_off;            // floats: this problem's partials start here; layout [split][n/4][m][4]
};
constexpr int WPROB_MAX = 3;
struct WgArgs {
  WProb pr[WPROB_MAX];
  int nprob;
  float* part;
  int Rp, nblocks, blocks_per_split;
  int* status;
};
struct WgMaps { CUtensorMap P0[WPROB_MAX], P1[WPROB_MAX], Q[WPROB_MAX]; };
template <int NT>
struct WSmem {
  static constexpr int A_BYTES = 16 * TILE * 16;        // 128 channels x 128 frames
  static constexpr int B_BYTES = (NT / 8) * TILE * 16;  // NT channels x 128 frames
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = 2;
  static constexpr int TOTAL = STAGES * STAGE + 1024;
};

template <int NT>
__global__ void __launch_bounds__(192, 1) wgemm_kernel(const WgArgs p, const __grid_constant__ WgMaps maps) {
  using S = WSmem<NT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_full[S::STAGES], bar_empty[S::STAGES], bar_acc;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int k = 0;
  while (k + 1 < p.nprob && (int)blockIdx.x >= p.pr[k + 1].tile0) ++k;
  const WProb& q = p.pr[k];
  const int local = blockIdx.x - q.tile0;
  const int mtile = local % q.mtiles, ny = local / q.mtiles, z = blockIdx.y;
  const int tap = ny / q.ntiles_per_tap, nt = ny - tap * q.ntiles_per_tap;
  const int kb0 = z * p.blocks_per_split;
  const int kb1 = min(p.nblocks, kb0 + p.blocks_per_split);
  const int nsteps = kb1 - kb0;

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
  const uint32_t smem_base = tc::smem_u32(smem);
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const void* tmP = mtile < q.mtiles0 ? &maps.P0[k] : &maps.P1[k];
      const void* tmQ = &maps.Q[k];
      const int mchunk0 = (mtile < q.mtiles0 ? mtile : mtile - q.mtiles0) * 16;
      const int nchunk0 = nt * (NT / 8);
      const int sh = tap - (q.taps >> 1);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % S::STAGES, ph = (s / S::STAGES) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&bar_empty[stage]), ph ^ 1, kTimeout, p.status, 1);
        const uint32_t sa = smem_base + stage * S::STAGE, sb = sa + S::A_BYTES;
        const uint32_t fb = tc::smem_u32(&bar_full[stage]);
        tc::mbar_arrive_expect_tx_addr(fb, S::STAGE);
        const int row = RLEAD + (kb0 + s) * TILE;
        tc::tma_load_2d(sa, tmP, 2 * row, mchunk0, fb);              // 128 channels x 128 frames: one 32 KB box
        tc::tma_load_2d(sb, tmQ, 2 * (row + sh), nchunk0, fb);       // NT channels x 128 frames: one 32/64 KB box
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // both operands MN-major (the frame axis is K): instruction-descriptor bits 15 and 16
      const uint32_t idesc = tc::make_idesc_bf16(128, NT) | (1u << 15) | (1u << 16);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % S::STAGES, ph = (s / S::STAGES) & 1;
        tc::mbar_wait_trap(tc::smem_u32(&bar_full[stage]), ph, kTimeout, p.status, 2);
        tc::tc_fence_after();
        const uint32_t sa = smem_base + stage * S::STAGE, sb = sa + S::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // 16 frames per MMA = 256 B along the frame axis
          const uint64_t ad = tc::make_smem_desc(sa + kk * 256, 128, TILE * 16);
          const uint64_t bd = tc::make_smem_desc(sb + kk * 256, 128, TILE * 16);
          tc::umma_bf16(tmem, ad, bd, idesc, (s | kk) ? 1u : 0u);
        }
        tc::umma_commit(&bar_empty[stage]);
      }
      tc::umma_commit(&bar_acc);
    }
  } else {
    tc::mbar_wait_trap(tc::smem_u32(&bar_acc), 0, kTimeout, p.status, 4);
    tc::tc_fence_after();
    const int i = (warp & 3) * 32 + lane;
    const int m = mtile * 128 + i;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    // partials in the chunked layout part[z][n/4][m][4]: thread = output row m, so every 16-byte store of a warp is
    // 512 contiguous bytes (a row-major [m][n] partial costs 32 sectors per store instruction)
    float* dst = p.part + q.part_off + (size_t)z * q.Mo * q.N;
    const int nbase = tap * q.Kin;
#pragma unroll 1
    for (int cg = 0; cg < NT / 32; ++cg) {
      const int ci0 = nt * NT + cg * 32;
      if (ci0 >= q.Kin) break;                       // uniform over the warp
      float v[32];
      if (nsteps > 0) {
        tmem_ld_f32x32(trow + cg * 32, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (m < q.Mo) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          if (ci0 + c4 * 4 < q.Kin)                  // Kin % 4 == 0
            *reinterpret_cast<float4*>(dst + ((size_t)((nbase + ci0) / 4 + c4) * q.Mo + m) * 4) =
                make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<NT>(tmem);
}

// dst[k][(m*Kin + ci)*taps + tap] = sum over splits of problem k's partials (fixed order), for up to three problems
struct WRedSeg { long long part_off, begin; float* dst; int Mo, N, Kin, taps; };
struct WRedArgs { WRedSeg seg[WPROB_MAX]; int nseg, S; long long total; const float* part; };
__global__ void __launch_bounds__(256) wgrad_reduce_multi_kernel(const WRedArgs a) {
  pdl_trigger();
  pdl_wait();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.total) return;
  int k = 0;
  while (k + 1 < a.nseg && j >= a.seg[k + 1].begin) ++k;
  const WRedSeg& g = a.seg[k];
  const long long i = j - g.begin, tot = (long long)g.Mo * g.N;
  const float* src = a.part + g.part_off + i;
  float s = 0.f;
  for (int z = 0; z < a.S; ++z) s += src[(size_t)z * tot];
  const int nq = (int)(i / ((long long)g.Mo * 4));
  const int rem = (int)(i - (long long)nq * g.Mo * 4);
  const int m = rem >> 2, n = nq * 4 + (rem & 3);
  const int tap = n / g.Kin, ci = n - tap * g.Kin;
  g.dst[((size_t)m * g.Kin + ci) * g.taps + tap] = s;
}

struct WPlan { int S, blocks_per_split; };
WPlan plan_wg(int nblocks, int tiles, int sm_budget = 148) {
  int S = sm_budget / tiles;                                   // one wave: at most one CTA per SM ...
  const int cap = (nblocks + 1) / 2;                           // ... but at least 2 frame blocks per CTA: every split costs a
  if (S > cap) S = cap;                                        // 128 x NT fp32 partial that the reduction has to read back
  if (S > nblocks) S = nblocks;
  if (S < 1) S = 1;
  const int per = (nblocks + S - 1) / S;
  return {(nblocks + per - 1) / per, per};
}
// upper bound of the partial buffer: the three per-block problems together
size_t wg_part_floats_max(int nblocks) {
  const size_t out_elems = (size_t)512 * 768 + (size_t)512 * 256 + (size_t)256 * 256;
  return (size_t)plan_wg(nblocks, 18).S * out_elems + (size_t)nblocks * 256 * 256;   // + room for the single-problem launches
}

struct WgSpec {                   // one weight-gradient problem: dst (flat gradient, state_dict layout) = P^T Q over frames
  const bf16* P0; const bf16* P1; int mtiles0; int Mo; const bf16* Q; int Kin; int taps; float* dst;
};
template <int NT>
int launch_wg_multi(const WgSpec* sp, int n, const RowSpace& rs, float* part, int* status, cudaStream_t s, int sm_budget = 148) {
  static PerDeviceOnce once;
  if (once.pending()) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(wgemm_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, WSmem<NT>::TOTAL));
    once.done();
  }
  WgArgs a{};
  WgMaps maps{};
  WRedArgs r{};
  a.nprob = n; a.part = part; a.Rp = rs.Rp; a.nblocks = rs.ntiles; a.status = status;
  int tiles = 0;
  for (int k = 0; k < n; ++k) {
    WProb& q = a.pr[k];
    q.mtiles = (sp[k].Mo + 127) / 128; q.mtiles0 = sp[k].mtiles0; q.taps = sp[k].taps;
    q.ntiles_per_tap = (sp[k].Kin + NT - 1) / NT; q.Kin = sp[k].Kin; q.Mo = sp[k].Mo; q.N = sp[k].Kin * sp[k].taps;
    q.tile0 = tiles;
    tiles += q.mtiles * q.taps * q.ntiles_per_tap;
    if (int rc = make_image_map(&maps.P0[k], sp[k].P0, q.mtiles0 * 16, rs.Rp, 16)) return rc;
    if (q.mtiles > q.mtiles0) { if (int rc = make_image_map(&maps.P1[k], sp[k].P1, (q.mtiles - q.mtiles0) * 16, rs.Rp, 16)) return rc; }
    else maps.P1[k] = maps.P0[k];
    if (int rc = make_image_map(&maps.Q[k], sp[k].Q, q.ntiles_per_tap * (NT / 8), rs.Rp, NT / 8)) return rc;
  }
  for (int k = n; k < WPROB_MAX; ++k) { maps.P0[k] = maps.P0[0]; maps.P1[k] = maps.P0[0]; maps.Q[k] = maps.Q[0]; }
  const WPlan pl = plan_wg(rs.ntiles, tiles, sm_budget);
  a.blocks_per_split = pl.blocks_per_split;
  long long off = 0, begin = 0;
  for (int k = 0; k < n; ++k) {
    a.pr[k].part_off = off;
    r.seg[k].part_off = off; r.seg[k].begin = begin; r.seg[k].dst = sp[k].dst; r.seg[k].Mo = a.pr[k].Mo; r.seg[k].N = a.pr[k].N;
    r.seg[k].Kin = a.pr[k].Kin; r.seg[k].taps = a.pr[k].taps;
    off += (long long)pl.S * a.pr[k].Mo * a.pr[k].N;
    begin += (long long)a.pr[k].Mo * a.pr[k].N;
  }
  r.nseg = n; r.S = pl.S; r.total = begin; r.part = part;
  MGB_CUDA_CHECK(launch_pdl(wgemm_kernel<NT>, dim3(tiles, pl.S), dim3(192), WSmem<NT>::TOTAL, s, 1, a, maps));
  if (trace_on()) { char b[96]; snprintf(b, sizeof b, "wgemm<%d> %d problems, grid %d x %d", NT, n, tiles, pl.S); trace(b, s); }
  MGB_CUDA_CHECK(launch_pdl(wgrad_reduce_multi_kernel, dim3((unsigned)((begin + 255) / 256)), dim3(256), 0, s, 1, r));
  note_launch(2);
  return MGB_OK;
}
template <int NT>
int launch_wg(const bf16* P0, const bf16* P1, int mtiles0, int Mo, const bf16* Q, int Kin, int taps, const RowSpace& rs,
              float* part, float* dst, int* status, cudaStream_t s) {
  const WgSpec sp{P0, P1, mtiles0, Mo, Q, Kin, taps, dst};
  return launch_wg_multi<NT>(&sp, 1, rs, part, status, s);
}

// =====================================================================================================
// weight packing: fp32 flat parameters -> bf16 K-major tiles [n-tile][k-step][8][NT][8]
// W_eff[n][tap*Kin + ci] = flat[base + n*sn + ci*sc + tap'*st],  tap' = rev ? taps-1-tap : tap
// =====================================================================================================
struct PackDesc {
  long long dst, base;
  int N, NT, Kin, Kin_pad, taps, sn, sc, st, rev, perm;
};
constexpr int PACK_MAX = 42;
struct PackArgs { PackDesc d[PACK_MAX]; int n; const float* flat; bf16* out; };

__global__ void __launch_bounds__(256) pack_tiles_kernel(const PackArgs a) {
  const PackDesc& d = a.d[blockIdx.y];
  const int ksteps = d.taps * (d.Kin_pad / 64);
  const int n_tiles = (d.N + d.NT - 1) / d.NT;
  const long long total = (long long)n_tiles * ksteps * 8 * d.NT;      // 16-byte units
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
    const int nn = (int)(u % d.NT);
    long long r = u / d.NT;
    const int kc = (int)(r % 8); r /= 8;
    const int s = (int)(r % ksteps);
    const int ntile = (int)(r / ksteps);
    int n = ntile * d.NT + nn;
    if (d.perm) {
      const int half = d.NT / 2;
      n = nn < half ? ntile * half + nn : d.N / 2 + ntile * half + (nn - half);
    }
    const int per_tap = d.Kin_pad / 64;
    const int tap = s / per_tap, j = s - tap * per_tap;
    const int tp = d.rev ? d.taps - 1 - tap : tap;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ci = j * 64 + kc * 8 + e;
      v[e] = (n < d.N && ci < d.Kin) ? a.flat[d.base + (long long)n * d.sn + (long long)ci * d.sc + (long long)tp * d.st] : 0.f;
    }
    *reinterpret_cast<uint4*>(a.out + d.dst + u * 8) =
        make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
  }
}

// =====================================================================================================
// input packers, margins, column sums
// =====================================================================================================
// src [B][M][T] fp32 -> image [nchunks][Rp][8] (channels >= M, separator rows and margins zero)
__global__ void __launch_bounds__(256) bmt_to_image_kernel(const float* __restrict__ src, bf16* __restrict__ img, int M,
                                                           int nchunks, int B, int T, int Rp) {
  const int rho = blockIdx.x * 256 + threadIdx.x, chunk = blockIdx.y;
  if (rho >= Rp) return;
  const int r = rho - RLEAD, Tg = T + 1;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (r >= 0) {
    const int b = r / Tg, t = r - b * Tg;
    if (b < B && t < T) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int m = chunk * 8 + e;
        if (m < M) v[e] = src[((size_t)b * M + m) * T + t];
      }
    }
  }
  *reinterpret_cast<uint4*>(img + ((size_t)chunk * Rp + rho) * 8) =
      make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
// cond [B][T][H] fp32 -> image [H/8][Rp][8]
__global__ void __launch_bounds__(256) bth_to_image_kernel(const float* __restrict__ src, bf16* __restrict__ img, int H,
                                                           int B, int T, int Rp) {
  const int rho = blockIdx.x * 8 + (threadIdx.x >> 5), chunk = threadIdx.x & 31;   // H == 256: 32 chunks per row
  if (rho >= Rp) return;
  const int r = rho - RLEAD, Tg = T + 1;
  uint4 u = make_uint4(0u, 0u, 0u, 0u);
  if (r >= 0) {
    const int b = r / Tg, t = r - b * Tg;
    if (b < B && t < T) {
      const float4* s4 = reinterpret_cast<const float4*>(src + ((size_t)b * T + t) * H + chunk * 8);
      const float4 a = __ldg(s4), c = __ldg(s4 + 1);
      u = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(c.x, c.y), pack2(c.z, c.w));
    }
  }
  *reinterpret_cast<uint4*>(img + ((size_t)chunk * Rp + rho) * 8) = u;
}
// zero the RLEAD leading and 8 trailing rows of `nimg` images of `nchunks` chunks each (stride img_stride elements)
__global__ void zero_margins_kernel(bf16* img, size_t img_stride, int nchunks, int Rp) {
  const int chunk = blockIdx.x, which = blockIdx.y, i = threadIdx.x;   // 16 threads: 8 lead + 8 trail rows
  const int rho = i < 8 ? i : Rp - 16 + i;
  *reinterpret_cast<uint4*>(img + (size_t)which * img_stride + ((size_t)chunk * Rp + rho) * 8) = make_uint4(0u, 0u, 0u, 0u);
}
// out[b][chunk*8 + e] = sum_t img[chunk][row(b,t)][e]
__global__ void __launch_bounds__(256) img_colsum_kernel(const bf16* __restrict__ img, int Rp, int T, float* __restrict__ out,
                                                         int ldo) {
  __shared__ float red[256][9];
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.x, b = blockIdx.y;
  const bf16* base = img + ((size_t)chunk * Rp + RLEAD + (size_t)b * (T + 1)) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = threadIdx.x; t < T; t += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(base + (size_t)t * 8);
    const float2 a = unpack2(u.x), c = unpack2(u.y), d = unpack2(u.z), e = unpack2(u.w);
    acc[0] += a.x; acc[1] += a.y; acc[2] += c.x; acc[3] += c.y; acc[4] += d.x; acc[5] += d.y; acc[6] += e.x; acc[7] += e.y;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int k = 0; k < 256; ++k) s += red[k][threadIdx.x];
    out[(size_t)b * ldo + chunk * 8 + threadIdx.x] = s;
  }
}
// up to three images in one launch (blockIdx.x runs over the concatenated chunk lists)
struct Colsum3 { const bf16* img[3]; float* out[3]; int nchunks[3], ldo[3]; };
__global__ void __launch_bounds__(256) img_colsum3_kernel(const Colsum3 a, int Rp, int T) {
  __shared__ float red[256][9];
  pdl_trigger();
  pdl_wait();
  int chunk = blockIdx.x, w = 0;
  while (w < 2 && chunk >= a.nchunks[w]) { chunk -= a.nchunks[w]; ++w; }
  const int b = blockIdx.y;
  const bf16* base = a.img[w] + ((size_t)chunk * Rp + RLEAD + (size_t)b * (T + 1)) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = threadIdx.x; t < T; t += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(base + (size_t)t * 8);
    const float2 p0 = unpack2(u.x), p1 = unpack2(u.y), p2 = unpack2(u.z), p3 = unpack2(u.w);
    acc[0] += p0.x; acc[1] += p0.y; acc[2] += p1.x; acc[3] += p1.y; acc[4] += p2.x; acc[5] += p2.y; acc[6] += p3.x; acc[7] += p3.y;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int k = 0; k < 256; ++k) s += red[k][threadIdx.x];
    a.out[w][(size_t)b * a.ldo[w] + chunk * 8 + threadIdx.x] = s;
  }
}
void launch_colsum_img(const bf16* img, int nchunks, const RowSpace& rs, int B, float* out, int ldo, cudaStream_t s) {
  launch_pdl(img_colsum_kernel, dim3(nchunks, B), dim3(256), 0, s, 1, img, rs.Rp, rs.T, out, ldo);
  note_launch();
  trace("img_colsum", s);
}

// =====================================================================================================
// buffers
// =====================================================================================================
struct PackedB16 {   // offsets in bf16 elements into the packed-weight buffer (re-packed every step)
  size_t in_f, skip_f, out_f;                       // forward
  size_t out_b, skip_b, in_b;                       // backward (transposed)
  size_t layer0, layer_stride;
  size_t r_conv_f, r_oproj_f, r_oproj_b, r_conv_b;
  size_t cond_f_all;                                // Wc of all layers as the column tiles of ONE N = 256 L GEMM
  size_t cond_b_all;                                // Wc^T of all layers, 4 k-steps each, contiguous (one K = 256 L GEMM)
  size_t total;
};
PackedB16 packed16_layout(const mgb_model_dims& d) {
  PackedB16 o{};
  size_t p = 0;
  auto take = [&](size_t n) { size_t r = p; p += n; return r; };
  o.in_f = take((size_t)256 * 128);        // N=256, K=80 -> 128
  o.skip_f = take((size_t)256 * 256);
  o.out_f = take((size_t)128 * 256);       // N=80 -> 128
  o.out_b = take((size_t)256 * 128);       // N=256 (channels), K=80 -> 128
  o.skip_b = take((size_t)256 * 256);
  o.in_b = take((size_t)128 * 256);        // N=80 -> 128, K=256
  o.layer0 = p;
  size_t q = 0;
  auto tk = [&](size_t n) { size_t r = q; q += n; return r; };
  o.r_conv_f = tk((size_t)512 * 768);
  o.r_oproj_f = tk((size_t)512 * 256);
  o.r_oproj_b = tk((size_t)256 * 512);
  o.r_conv_b = tk((size_t)256 * 1536);
  o.layer_stride = q;
  p += q * d.layers;
  o.cond_f_all = take((size_t)256 * 256 * d.layers);
  o.cond_b_all = take((size_t)256 * 256 * d.layers);
  o.total = p;
  return o;
}

struct Saved16 {     // byte offsets into the activation stash
  size_t xt, cond, X0, layer0, layer_stride, rY, rG, rSGTH, Sn, P, dvec, h, wpk, total;
};
Saved16 saved16_layout(const mgb_model_dims& d, int B, int T) {
  const RowSpace rs = row_space(B, T);
  const size_t Rp = rs.Rp;
  Saved16 o{};
  size_t p = 0;
  auto take = [&](size_t n) { size_t r = p; p += align_up(n, 256); return r; };
  o.xt = take(16 * Rp * 16);
  o.cond = take(32 * Rp * 16);
  o.X0 = take(Rp * C * 4);
  o.layer0 = p;
  {
    size_t q = 0;
    auto tk = [&](size_t n) { size_t r = q; q += align_up(n, 256); return r; };
    o.rY = tk(32 * Rp * 16); o.rG = tk(32 * Rp * 16); o.rSGTH = tk(Rp * C * 4);
    o.layer_stride = q;
  }
  p += o.layer_stride * d.layers;
  o.Sn = take(32 * Rp * 16);
  o.P = take(32 * Rp * 16);
  o.dvec = take((size_t)B * C * 4);
  o.h = take((size_t)B * 4 * C * 4);
  o.wpk = take(packed16_layout(d).total * 2);     // the step's bf16 weight tiles (the backward reuses them)
  o.total = p;
  return o;
}

struct Work16 {      // byte offsets into the workspace
  size_t status, X, S, Call, dtab, ctab, dout, dPre, dS, E, Eimg, dZ, dY, part, usumE, usumE2, usumZ, usumY, usumS, usumT, ddvec,
      dspk, dpre, dd_all, ds_all, lpart, total;
};
Work16 work16_layout(const mgb_model_dims& d, int B, int T) {
  const RowSpace rs = row_space(B, T);
  const size_t Rp = rs.Rp;
  Work16 w{};
  size_t p = 0;
  auto take = [&](size_t n) { size_t r = p; p += align_up(n, 256); return r; };
  w.status = take(256);
  w.X = take(Rp * C * 4);
  w.S = take(Rp * C * 4);
  w.Call = take((size_t)d.layers * Rp * C * 4);   // Wc_l cond of every layer (fp32 stream), written by one launch
  w.dtab = take((size_t)B * d.layers * C * 4);
  w.ctab = take((size_t)B * d.layers * C * 4);
  w.dout = take(16 * Rp * 16);
  w.dPre = take(32 * Rp * 16);
  w.dS = take(32 * Rp * 16);
  w.E = take(Rp * C * 4);
  w.Eimg = take(3 * 32 * Rp * 16);   // e images rotate over three buffers and dZ over two: the weight-gradient work of block
  w.dZ = take(2 * 64 * Rp * 16);     // l (side stream) still reads e_l / dZ_l while blocks l-1, l-2 run their data-gradient chain
  w.dY = take((size_t)d.layers * 32 * Rp * 16);   // every layer's dY image: d loss / d cond is ONE GEMM over K = 256 L at the head
  w.part = take(wg_part_floats_max(rs.ntiles) * 4);
  w.usumE = take((size_t)B * C * 4);
  w.usumE2 = take((size_t)B * C * 4);
  w.usumZ = take((size_t)B * 2 * C * 4);
  w.usumY = take((size_t)B * C * 4);
  w.usumS = take((size_t)B * C * 4);
  w.usumT = take((size_t)B * C * 4);
  w.ddvec = take((size_t)B * C * 4);
  w.dspk = take((size_t)B * C * 4);
  w.dpre = take((size_t)B * 4 * C * 4);
  w.dd_all = take((size_t)d.layers * B * C * 4);
  w.ds_all = take((size_t)d.layers * B * C * 4);
  w.lpart = take((size_t)d.layers * B * C * 4);
  w.total = p;
  return w;
}

int pack_step_weights(const mgb_model_dims& d, const float* flat, bf16* out, cudaStream_t s) {
  const FlatOffsets f = flat_offsets(d);
  const PackedB16 o = packed16_layout(d);
  const int M = d.n_mel;
  auto desc = [](size_t dst, size_t base, int N, int NT, int Kin, int taps, int sn, int sc, int st, int rev, int perm) {
    PackDesc p{};
    p.dst = (long long)dst; p.base = (long long)base; p.N = N; p.NT = NT; p.Kin = Kin; p.Kin_pad = (Kin + 63) / 64 * 64;
    p.taps = taps; p.sn = sn; p.sc = sc; p.st = st; p.rev = rev; p.perm = perm;
    return p;
  };
  auto run = [&](PackArgs& a) {
    a.flat = flat; a.out = out;
    pack_tiles_kernel<<<dim3(48, a.n), 256, 0, s>>>(a);
    note_launch();
  };
  PackArgs a{};
  auto push = [&](const PackDesc& pd) {
    a.d[a.n++] = pd;
    if (a.n == PACK_MAX) { run(a); a.n = 0; }
  };
  push(desc(o.in_f, f.in_w, C, 256, M, 1, M, 1, 0, 0, 0));          // W[c][m]
  push(desc(o.skip_f, f.skip_w, C, 256, C, 1, C, 1, 0, 0, 0));
  push(desc(o.out_f, f.out_w, M, 128, C, 1, C, 1, 0, 0, 0));        // W[m][c]
  push(desc(o.out_b, f.out_w, C, 256, M, 1, 1, C, 0, 0, 0));        // W_eff[n=c][k=m] = Wout[m][c]
  push(desc(o.skip_b, f.skip_w, C, 256, C, 1, 1, C, 0, 0, 0));      // Wsk^T
  push(desc(o.in_b, f.in_w, M, 128, C, 1, 1, M, 0, 0, 0));          // W_eff[n=m][k=c] = Win[c][m]
  for (int l = 0; l < d.layers; ++l) {
    const size_t fl = f.layer0 + (size_t)l * f.layer_stride, ol = o.layer0 + (size_t)l * o.layer_stride;
    push(desc(o.cond_f_all + (size_t)l * 256 * 256, fl + f.rel.cproj_w, C, 256, C, 1, C, 1, 0, 0, 0));
    push(desc(ol + o.r_conv_f, fl + f.rel.conv_w, 2 * C, 256, C, 3, 3 * C, 3, 1, 0, 1));      // W3[co][ci][tap], gate|filter tiles
    push(desc(ol + o.r_oproj_f, fl + f.rel.oproj_w, 2 * C, 256, C, 1, C, 1, 0, 0, 1));        // Wo[co][ci], x|skip tiles
    push(desc(ol + o.r_oproj_b, fl + f.rel.oproj_w, C, 256, 2 * C, 1, 1, C, 0, 0, 0));        // W_eff[n=ci][k=co]
    push(desc(ol + o.r_conv_b, fl + f.rel.conv_w, C, 256, 2 * C, 3, 3, 3 * C, 1, 1, 0));      // W_eff[n=ci][tap', co] = W3[co][ci][2-tap']
    push(desc(o.cond_b_all + (size_t)l * 256 * 256, fl + f.rel.cproj_w, C, 256, C, 1, 1, C, 0, 0, 0));   // Wc^T
  }
  if (a.n) run(a);
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

}  // namespace

size_t bf16_train_saved_bytes(const mgb_model_dims& d, int B, int T) { return saved16_layout(d, B, T).total; }
size_t bf16_train_workspace_bytes(const mgb_model_dims& d, int B, int T) { return work16_layout(d, B, T).total; }

int bf16_train_forward(const mgb_model_dims& d, const void* packed_fp32, const float* flat, const float* x, const int64_t* t,
                       const float* cond, const float* spk, float* out, void* saved, int B, int T, void* ws, cudaStream_t s) {
  const RowSpace rs = row_space(B, T);
  const Saved16 sv = saved16_layout(d, B, T);
  const Work16 w = work16_layout(d, B, T);
  const PackedB16 o = packed16_layout(d);
  const FlatOffsets f = flat_offsets(d);
  uint8_t* SV = static_cast<uint8_t*>(saved);
  uint8_t* W = static_cast<uint8_t*>(ws);
  const int L = d.layers, M = d.n_mel;
  auto img = [](uint8_t* p) { return reinterpret_cast<bf16*>(p); };
  auto f32 = [](uint8_t* p) { return reinterpret_cast<float*>(p); };
  int* status = reinterpret_cast<int*>(W + w.status);
  bf16* wpk = img(SV + sv.wpk);

  MGB_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));
  if (int rc = pack_step_weights(d, flat, wpk, s)) return rc;
  if (int rc = fp32_step_tables(d, packed_fp32, t, spk, B, f32(SV + sv.dvec), f32(SV + sv.h), f32(W + w.dtab),
                                f32(W + w.ctab), s)) return rc;
  bmt_to_image_kernel<<<dim3((rs.Rp + 255) / 256, 16), 256, 0, s>>>(x, img(SV + sv.xt), M, 16, B, T, rs.Rp);
  bth_to_image_kernel<<<(rs.Rp + 7) / 8, 256, 0, s>>>(cond, img(SV + sv.cond), C, B, T, rs.Rp);
  zero_margins_kernel<<<dim3(32, L), 16, 0, s>>>(img(SV + sv.layer0 + sv.rY), sv.layer_stride / 2, 32, rs.Rp);
  note_launch(3);

  FArgs base{};
  base.Rp = rs.Rp; base.B = B; base.T = T; base.L = L; base.status = status; base.taps = 1; base.n_mel = M;
  {
    FArgs a = base;                                     // conditioner projection of ALL layers in one launch (N = 256 L)
    a.A0 = img(SV + sv.cond); a.steps0 = 4; a.Bpk = wpk + o.cond_f_all; a.ksteps_b = 4; a.fout = f32(W + w.Call);
    if (int rc = launch_fgemm<256, F_COND>(a, rs.ntiles, L, s)) return rc;
  }
  {
    FArgs a = base;                                     // input projection + ReLU (modules.py:429-431), and y_0
    a.A0 = img(SV + sv.xt); a.steps0 = 2; a.Bpk = wpk + o.in_f; a.ksteps_b = 2;
    a.bias = flat + f.in_b; a.fout = f32(SV + sv.X0);
    a.call = f32(W + w.Call); a.dtab2 = f32(W + w.dtab); a.ctab2 = f32(W + w.ctab); a.img2 = img(SV + sv.layer0 + sv.rY);
    if (int rc = launch_fgemm<256, F_IN>(a, rs.ntiles, 1, s)) return rc;
  }
  for (int l = 0; l < L; ++l) {
    uint8_t* sl = SV + sv.layer0 + (size_t)l * sv.layer_stride;
    const bf16* wl = wpk + o.layer0 + (size_t)l * o.layer_stride;
    const float* fl = flat + f.layer0 + (size_t)l * f.layer_stride;
    const float* xin = l == 0 ? f32(SV + sv.X0) : f32(W + w.X);
    {
      FArgs a = base;                                   // k=3 conv + gate
      a.A0 = img(sl + sv.rY); a.steps0 = 4; a.taps = 3; a.Bpk = wl + o.r_conv_f; a.ksteps_b = 12;
      a.bias = fl + f.rel.conv_b; a.img = img(sl + sv.rG); a.sgth = reinterpret_cast<uint32_t*>(sl + sv.rSGTH);
      if (int rc = launch_fgemm<256, F_GATE>(a, rs.ntiles, 2, s)) return rc;
    }
    {
      FArgs a = base;                                   // output projection, residual + skip
      a.A0 = img(sl + sv.rG); a.steps0 = 4; a.Bpk = wl + o.r_oproj_f; a.ksteps_b = 4;
      a.bias = fl + f.rel.oproj_b; a.fin = xin; a.fout = f32(W + w.X); a.fout2 = f32(W + w.S);
      a.dtab = f32(W + w.dtab) + (size_t)l * C; a.first = l == 0; a.last = l == L - 1;
      a.scale = 1.0f / sqrtf((float)L); a.img = img(SV + sv.Sn);
      if (l + 1 < L) {                                  // this epilogue also writes the next block's conv input y_{l+1}
        a.call = f32(W + w.Call) + (size_t)(l + 1) * rs.Rp * C;
        a.dtab2 = f32(W + w.dtab) + (size_t)(l + 1) * C; a.ctab2 = f32(W + w.ctab) + (size_t)(l + 1) * C;
        a.img2 = img(sl + sv.layer_stride + sv.rY);
      }
      if (int rc = launch_fgemm<256, F_OUT>(a, rs.ntiles, 2, s)) return rc;
    }
  }
  {
    FArgs a = base;                                     // skip projection + ReLU
    a.A0 = img(SV + sv.Sn); a.steps0 = 4; a.Bpk = wpk + o.skip_f; a.ksteps_b = 4;
    a.bias = flat + f.skip_b; a.img = img(SV + sv.P);
    if (int rc = launch_fgemm<256, F_SKIP>(a, rs.ntiles, 1, s)) return rc;
  }
  {
    FArgs a = base;                                     // output projection -> [B][n_mel][T]
    a.A0 = img(SV + sv.P); a.steps0 = 4; a.Bpk = wpk + o.out_f; a.ksteps_b = 4;
    a.bias = flat + f.out_b; a.fout = out;
    if (int rc = launch_fgemm<128, F_FINAL>(a, rs.ntiles, 1, s)) return rc;
  }
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

// The backward of one block is a data-gradient chain (gate backward -> conv^T: what the NEXT block waits for) plus weight
// gradients, column sums and small per-utterance terms that nothing downstream needs before the head.  The second group
// runs on a library-owned side stream, one block behind the chain: events fork it after the block's conv^T GEMM and join
// it back before the buffers it reads are overwritten (dZ two blocks later, the e image three blocks later), before the
// head segment, and at the end of every call (so a caller — eager, graph capture, bucketed all-reduce — sees a finished
// gradient slice on its own stream).  The weight-gradient GEMM is planned for 96 SMs there, so that the 52-CTA chain
// kernels of the next block find free SMs beside it.  One stream + 8 events per device, created at first use.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t main_ev[4] = {}, side_ev[4] = {}; };
SideStream* side_stream() {
  static SideStream ctx[256];
  static PerDeviceOnce once;
  const int dev = PerDeviceOnce::current();
  if (once.pending()) {
    SideStream& c = ctx[dev];
    if (cudaStreamCreateWithFlags(&c.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < 4; ++i) {
      if (cudaEventCreateWithFlags(&c.main_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&c.side_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    once.done();
  }
  return &ctx[dev];
}
#ifdef MGB_DEBUG_BUILD
bool overlap_enabled() {                          // debug library only: MGB_TRAIN_NO_OVERLAP=1 keeps everything on one stream (A/B)
  static const bool on = [] { const char* e = getenv("MGB_TRAIN_NO_OVERLAP"); return !(e && e[0] == '1'); }();
  return on;
}
int side_wg_sms() {                               // debug library only: MGB_TRAIN_WG_SMS overrides the weight-gradient SM budget
  static const int v = [] { const char* e = getenv("MGB_TRAIN_WG_SMS"); return e ? atoi(e) : 96; }();
  return v;
}
#else
constexpr bool overlap_enabled() { return true; }
constexpr int side_wg_sms() { return 96; }
#endif

int bf16_train_backward(const mgb_model_dims& d, const float* flat, const void* saved, const int64_t* t, const float* cond,
                        const float* spk, const float* grad_out, float* grad_flat, float* grad_cond, float* grad_spk,
                        float* grad_x, int B, int T, int seg_begin, int seg_end, void* ws, cudaStream_t s) {
  (void)cond;
  const RowSpace rs = row_space(B, T);
  const Saved16 sv = saved16_layout(d, B, T);
  const Work16 w = work16_layout(d, B, T);
  const PackedB16 o = packed16_layout(d);
  const FlatOffsets f = flat_offsets(d);
  const uint8_t* SV = static_cast<const uint8_t*>(saved);
  uint8_t* W = static_cast<uint8_t*>(ws);
  const int L = d.layers, M = d.n_mel, H = d.d_encoder;
  auto img = [](uint8_t* p) { return reinterpret_cast<bf16*>(p); };
  auto cimg = [](const uint8_t* p) { return reinterpret_cast<const bf16*>(p); };
  auto f32 = [](uint8_t* p) { return reinterpret_cast<float*>(p); };
  auto cf32 = [](const uint8_t* p) { return reinterpret_cast<const float*>(p); };
  int* status = reinterpret_cast<int*>(W + w.status);
  const bf16* wpk = cimg(SV + sv.wpk);
  float* part = f32(W + w.part);

  FArgs base{};
  base.Rp = rs.Rp; base.B = B; base.T = T; base.L = L; base.status = status; base.taps = 1; base.n_mel = M;

  SideStream* side = overlap_enabled() ? side_stream() : nullptr;
  cudaStream_t ss = side ? side->s : s;           // stream of the off-chain work
  int first_side_seg = -1, last_side_seg = -1;    // blocks whose off-chain work was issued by THIS call
  auto join_side = [&]() -> int {
    if (side && last_side_seg >= 0) MGB_CUDA_CHECK(cudaStreamWaitEvent(s, side->side_ev[last_side_seg & 3], 0));
    first_side_seg = last_side_seg = -1;
    return MGB_OK;
  };
  const size_t dz_bytes = (size_t)64 * rs.Rp * 16, eimg_bytes = (size_t)32 * rs.Rp * 16;

  for (int seg = seg_begin; seg < seg_end; ++seg) {
    if (seg == 0) {
      bmt_to_image_kernel<<<dim3((rs.Rp + 255) / 256, 16), 256, 0, s>>>(grad_out, img(W + w.dout), M, 16, B, T, rs.Rp);
      zero_margins_kernel<<<dim3(64, 2), 16, 0, s>>>(img(W + w.dZ), dz_bytes / 2, 64, rs.Rp);
      note_launch(2);
      // dWout = dout^T P, dbout
      if (int rc = launch_wg<256>(cimg(W + w.dout), nullptr, 1, M, cimg(SV + sv.P), C, 1, rs, part, grad_flat + f.out_w,
                                  status, s)) return rc;
      launch_colsum_img(cimg(W + w.dout), 10, rs, B, f32(W + w.usumT), C, s);
      bias_from_usum_kernel<<<1, 128, 0, s>>>(f32(W + w.usumT), B, M, C, grad_flat + f.out_b);
      note_launch();
      {
        FArgs a = base;                                 // dPre = (dout Wout) masked by P > 0
        a.A0 = cimg(W + w.dout); a.steps0 = 2; a.Bpk = wpk + o.out_b; a.ksteps_b = 2;
        a.aux_img = cimg(SV + sv.P); a.img = img(W + w.dPre);
        if (int rc = launch_fgemm<256, B_PMASK>(a, rs.ntiles, 1, s)) return rc;
      }
      if (int rc = launch_wg<256>(cimg(W + w.dPre), nullptr, 2, C, cimg(SV + sv.Sn), C, 1, rs, part, grad_flat + f.skip_w,
                                  status, s)) return rc;
      launch_colsum_img(cimg(W + w.dPre), 32, rs, B, f32(W + w.usumT), C, s);
      bias_from_usum_kernel<<<2, 128, 0, s>>>(f32(W + w.usumT), B, C, C, grad_flat + f.skip_b);
      note_launch();
      {
        FArgs a = base;                                 // dS = (dPre Wsk) / sqrt(L)
        a.A0 = cimg(W + w.dPre); a.steps0 = 4; a.Bpk = wpk + o.skip_b; a.ksteps_b = 4;
        a.scale = 1.0f / sqrtf((float)L); a.img = img(W + w.dS);
        if (int rc = launch_fgemm<256, B_DS>(a, rs.ntiles, 1, s)) return rc;
      }
      launch_colsum_img(cimg(W + w.dS), 32, rs, B, f32(W + w.usumS), C, s);
    } else if (seg <= L) {
      const int l = L - seg;
      const bool top = (l == L - 1);
      const uint8_t* sl = SV + sv.layer0 + (size_t)l * sv.layer_stride;
      const bf16* wl = wpk + o.layer0 + (size_t)l * o.layer_stride;
      const float* fl = flat + f.layer0 + (size_t)l * f.layer_stride;
      float* gl = grad_flat + f.layer0 + (size_t)l * f.layer_stride;
      bf16* dYl = img(W + w.dY) + (size_t)l * 32 * rs.Rp * 8;
      // e images rotate over three buffers: e_cur = e_l (input of this block, written by the previous one), e_new = e_{l-1};
      // dZ alternates between two
      bf16* e_cur = img(W + w.Eimg + (size_t)((seg + 2) % 3) * eimg_bytes);
      bf16* e_new = img(W + w.Eimg + (size_t)(seg % 3) * eimg_bytes);
      bf16* dZl = img(W + w.dZ + (size_t)(seg & 1) * dz_bytes);
      // the off-chain work of block seg - 2 read the dZ buffer this block overwrites (and the e buffer the NEXT chain step
      // of this block's successor overwrites): wait for it, if this call issued it (an earlier call joined before returning)
      if (side && seg - 2 >= first_side_seg && first_side_seg >= 0 && seg - 2 <= last_side_seg)
        MGB_CUDA_CHECK(cudaStreamWaitEvent(s, side->side_ev[(seg - 2) & 3], 0));
      {
        FArgs a = base;                                 // dG = [e | dS] Wo; gate backward -> dZ
        if (top) { a.A0 = cimg(W + w.dS); a.steps0 = 4; a.kstep_b0 = 4; }
        else { a.A0 = e_cur; a.steps0 = 4; a.A1 = cimg(W + w.dS); a.steps1 = 4; }
        a.Bpk = wl + o.r_oproj_b; a.ksteps_b = 8;
        a.sgth_in = reinterpret_cast<const uint32_t*>(sl + sv.rSGTH); a.img = dZl;
        if (int rc = launch_fgemm<256, B_GATE>(a, rs.ntiles, 1, s)) return rc;
      }
      {
        FArgs a = base;                                 // dY = conv3^T(dZ); dx = e + dY; e' = dx / sqrt(2)
        a.A0 = dZl; a.steps0 = 8; a.taps = 3; a.Bpk = wl + o.r_conv_b; a.ksteps_b = 24;
        a.fout = f32(W + w.E); a.fin = cf32(SV + sv.X0); a.img = dYl; a.img2 = e_new;
        a.first = top ? 1 : 0; a.relu0 = (l == 0) ? 1 : 0;
        if (int rc = launch_fgemm<256, B_DX>(a, rs.ntiles, 1, s)) return rc;
      }
      if (side) {                                       // fork: everything below runs behind the chain
        MGB_CUDA_CHECK(cudaEventRecord(side->main_ev[seg & 3], s));
        MGB_CUDA_CHECK(cudaStreamWaitEvent(ss, side->main_ev[seg & 3], 0));
      }
      {
        // the block's three weight gradients in ONE launch (18 output tiles x 8 frame splits = one wave) + one reduction:
        //   dWo = [e_l | dS]^T g_l (top block: e = 0, only the skip half), dW3 = dZ^T shift(y_l), dWc = dY^T cond
        WgSpec sp[3];
        if (top) {
          MGB_CUDA_CHECK(cudaMemsetAsync(gl + f.rel.oproj_w, 0, sizeof(float) * (size_t)C * C, ss));
          sp[0] = WgSpec{cimg(W + w.dS), nullptr, 2, C, cimg(sl + sv.rG), C, 1, gl + f.rel.oproj_w + (size_t)C * C};
        } else {
          sp[0] = WgSpec{e_cur, cimg(W + w.dS), 2, 2 * C, cimg(sl + sv.rG), C, 1, gl + f.rel.oproj_w};
        }
        sp[1] = WgSpec{dZl, nullptr, 4, 2 * C, cimg(sl + sv.rY), C, 3, gl + f.rel.conv_w};
        sp[2] = WgSpec{dYl, nullptr, 2, C, cimg(SV + sv.cond), H, 1, gl + f.rel.cproj_w};
        if (int rc = launch_wg_multi<256>(sp, 3, rs, part, status, ss, side ? side_wg_sms() : 148)) return rc;
      }
      // per-utterance column sums of dZ_l, dY_l and of the NEW e image (= e_{l-1}, used by the next block's small terms):
      // one launch; the two e-sum buffers alternate between blocks
      float* usumE_cur = f32(W + ((seg & 1) ? w.usumE : w.usumE2));
      float* usumE_next = f32(W + ((seg & 1) ? w.usumE2 : w.usumE));
      {
        Colsum3 c3{};
        c3.img[0] = dZl; c3.nchunks[0] = 64; c3.out[0] = f32(W + w.usumZ); c3.ldo[0] = 2 * C;
        c3.img[1] = dYl; c3.nchunks[1] = 32; c3.out[1] = f32(W + w.usumY); c3.ldo[1] = C;
        c3.img[2] = e_new; c3.nchunks[2] = 32; c3.out[2] = usumE_next; c3.ldo[2] = C;
        launch_pdl(img_colsum3_kernel, dim3(128, B), dim3(256), 0, ss, 1, c3, rs.Rp, rs.T);
        note_launch();
        trace("img_colsum3", ss);
      }
      LayerSmallArgs q{};
      q.usumE = top ? nullptr : usumE_cur; q.usumZ = f32(W + w.usumZ); q.usumY = f32(W + w.usumY);
      q.usumS = f32(W + w.usumS); q.dvec = cf32(SV + sv.dvec); q.spk = d.multi_speaker ? spk : nullptr;
      q.Wd = fl + f.rel.dproj_w; q.Ws = d.multi_speaker ? fl + f.rel.sproj_w : nullptr;
      q.g_conv_b = gl + f.rel.conv_b; q.g_oproj_b = gl + f.rel.oproj_b; q.g_cproj_b = gl + f.rel.cproj_b;
      q.g_dproj_w = gl + f.rel.dproj_w; q.g_sproj_w = d.multi_speaker ? gl + f.rel.sproj_w : nullptr;
      q.dd_l = f32(W + w.dd_all) + (size_t)l * B * C; q.ds_l = d.multi_speaker ? f32(W + w.ds_all) + (size_t)l * B * C : nullptr;
      q.B = B; q.C = C; q.H = H;
      launch_pdl(layer_small_kernel, dim3(C + B + 1), dim3(256), 0, ss, 1, q);
      note_launch();
      if (side) {
        MGB_CUDA_CHECK(cudaEventRecord(side->side_ev[seg & 3], ss));
        if (first_side_seg < 0) first_side_seg = seg;
        last_side_seg = seg;
      }
    } else if (seg == L + 1) {
      if (int rc = join_side()) return rc;
      // head: E and the e image written by block 0 hold the ReLU-masked gradient of the input projection's pre-activation
      const bf16* e_last = cimg(W + w.Eimg + (size_t)(L % 3) * eimg_bytes);
      if (int rc = launch_wg<128>(e_last, nullptr, 2, C, cimg(SV + sv.xt), M, 1, rs, part, grad_flat + f.in_w,
                                  status, s)) return rc;
      launch_colsum_img(e_last, 32, rs, B, f32(W + w.usumT), C, s);
      bias_from_usum_kernel<<<2, 128, 0, s>>>(f32(W + w.usumT), B, C, C, grad_flat + f.in_b);
      note_launch();
      if (grad_cond) {
        // d loss / d cond = sum_l dY_l Wc_l: the L dY images are one image of 32 L chunks, so this is a single GEMM over
        // K = 256 L whose epilogue writes the user tensor once (no read-modify-write per layer)
        FArgs a = base;
        a.A0 = cimg(W + w.dY); a.steps0 = 4 * L; a.Bpk = wpk + o.cond_b_all; a.ksteps_b = 4 * L;
        a.fout = grad_cond; a.first = 1;
        if (int rc = launch_fgemm<256, B_COND>(a, rs.ntiles, 1, s)) return rc;
      }
      if (grad_x) {
        FArgs a = base;                                 // d loss / d mel = dPre0 Win
        a.A0 = e_last; a.steps0 = 4; a.Bpk = wpk + o.in_b; a.ksteps_b = 4; a.fout = grad_x;
        if (int rc = launch_fgemm<128, B_DXT>(a, rs.ntiles, 1, s)) return rc;
      }
      launch_dvec_contraction(f32(W + w.dd_all), f32(W + w.ds_all), flat + f.layer0 + f.rel.dproj_w,
                              d.multi_speaker ? flat + f.layer0 + f.rel.sproj_w : nullptr, f.layer_stride, f32(W + w.lpart),
                              f32(W + w.ddvec), f32(W + w.dspk), B, L, s);
      mlp_bwd_w2_kernel<<<C, 256, 0, s>>>(f32(W + w.ddvec), cf32(SV + sv.h), grad_flat + f.mlp2_w, B, C);
      mlp_bwd_pre_kernel<<<dim3(4 * C / 256, B), 256, 0, s>>>(t, f32(W + w.ddvec), flat + f.mlp0_w, flat + f.mlp2_w,
                                                             f32(W + w.dpre), C);
      mlp_bwd_w0_kernel<<<4 * C, C, 0, s>>>(t, f32(W + w.dpre), grad_flat + f.mlp0_w, B, C);
      note_launch(3);
      if (grad_spk && d.multi_speaker)
        MGB_CUDA_CHECK(cudaMemcpyAsync(grad_spk, W + w.dspk, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, s));
    }
  }
  if (int rc = join_side()) return rc;
  MGB_LAUNCH_CHECK();
  return MGB_OK;
}

size_t bf16_train_status_offset(const mgb_model_dims& d, int B, int T) { return work16_layout(d, B, T).status; }

}  // namespace mgb

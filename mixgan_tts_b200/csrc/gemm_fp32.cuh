// fp32 CUDA-core GEMM building blocks shared by the fp32 parity path (fp32_path.cu) and the fp32
// training path (train_fp32.cu): a 128x128x8 register-tiled main loop over a frames-major activation
// matrix (one row per mel frame) with optional row-shifted taps (k=3 convolutions as three shifted GEMMs,
// zero rows at utterance edges), and the transposed form that reduces over frames (weight gradients).
#pragma once

#include "common.cuh"

namespace mgb {
namespace gemm32 {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256, BMP = BM + 4;

// out[rows][N] tile = A[rows][taps*Kin] (implicit: tap shifts rows) * Wt[taps*Kin][ldw]
struct FrameGemm {
  const float* A;    // [rows][lda]
  const float* A2;   // optional second activation matrix for k >= ksplit (taps == 1 only), [rows][lda2]
  const float* Wt;   // [taps*Kin][ldw]
  int lda, lda2, ksplit, ldw, rows, T, Kin, taps;
};

// Thread (tx = tid & 15, ty = tid >> 4) ends up with acc[i][j] for rows m0 + {ty*4+i | 64+ty*4+i-4} and
// columns n0 + {tx*4+j | 64+tx*4+j-4}.
__device__ __forceinline__ void frame_gemm_mainloop(const FrameGemm& p, int m0, int n0, float (&acc)[8][8],
                                                    float (&As)[2][BK][BMP], float (&Bs)[2][BK][BN]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int a_row = tid >> 1, a_kq = (tid & 1) * 4;
  const int a_m = m0 + a_row;
  const int a_t = a_m % p.T;
  const bool a_in = a_m < p.rows;
  const int b_k = tid >> 5, b_n = (tid & 31) * 4;

#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int Ktot = p.taps * p.Kin;
  const int nk = Ktot / BK;
  const int half = p.taps >> 1;

  float4 ra, rb;
  auto gload = [&](int kt) {
    const int kk = kt * BK;
    const int tap = kk / p.Kin;
    const int k0 = kk - tap * p.Kin;
    const int sh = tap - half;
    const int ts = a_t + sh;
    if (a_in && ts >= 0 && ts < p.T) {
      if (p.A2 && k0 >= p.ksplit)
        ra = *reinterpret_cast<const float4*>(p.A2 + (size_t)a_m * p.lda2 + (k0 - p.ksplit) + a_kq);
      else
        ra = *reinterpret_cast<const float4*>(p.A + (size_t)(a_m + sh) * p.lda + k0 + a_kq);
    } else {
      ra = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rb = *reinterpret_cast<const float4*>(p.Wt + (size_t)(kk + b_k) * p.ldw + n0 + b_n);
  };
  auto sstore = [&](int buf) {
    As[buf][a_kq + 0][a_row] = ra.x;
    As[buf][a_kq + 1][a_row] = ra.y;
    As[buf][a_kq + 2][a_row] = ra.z;
    As[buf][a_kq + 3][a_row] = ra.w;
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
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
}

__device__ __forceinline__ int acc_row(int ty, int i) { return i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
__device__ __forceinline__ int acc_col(int tx, int j) { return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

}  // namespace gemm32
}  // namespace mgb

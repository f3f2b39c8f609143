// Strict-fp32 SIMT GEMM with functor operand gather and functor epilogue.
//   C[m][n] = epi( sum_k A(m,k) * W[n][k] )        W is K-major ([N][ldw]), as PyTorch stores weights.
// This is the mode-0 ("strict fp32") engine: every GEMM-shaped op of the path can run through it, and it is the
// on-device fp32 yardstick for the tcgen05 kernels.  It is not the fast path.
#pragma once
#include <cuda_runtime.h>

namespace cbx {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <class AF, class EF>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, AF af, const float* __restrict__ W, int ldw, EF ef) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int kk = e & 15, mm = e >> 4;
      const int gk = k0 + kk, gm = m0 + mm, gn = n0 + mm;
      As[kk][mm] = (gm < M && gk < K) ? af(gm, gk) : 0.f;
      Bs[kk][mm] = (gn < N && gk < K) ? __ldg(W + (size_t)gn * ldw + gk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) ef(gm, gn, acc[i][j]);
    }
  }
}

template <class AF, class EF>
inline void sgemm(Launches& L, cudaStream_t st, const char* tag, int M, int N, int K, AF af, const float* W, int ldw, EF ef) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((M + SG_BM - 1) / SG_BM, (N + SG_BN - 1) / SG_BN);
  Scope sc(L, st, tag, 2.0 * M * N * K);
  sgemm_kernel<AF, EF><<<grid, 256, 0, st>>>(M, N, K, af, W, ldw, ef);
}

}  // namespace cbx

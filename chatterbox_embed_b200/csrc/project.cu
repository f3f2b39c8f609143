// Consumer projections of the two speaker embeddings, so a voice bank can store generator-ready conditioning:
//   T3CondEnc.spkr_enc                 Linear(256 -> 1024) on the VoiceEncoder embedding        (t3/modules/cond_enc.py:50,70)
//   F.normalize + spk_embed_affine_layer  Linear(192 -> 80) on the CAMPPlus x-vector             (s3gen/flow.py:252-253, :73)
// y[i] = W . (normalize ? x[i] / max(||x[i]||_2, 1e-12) : x[i]) + b, fp32 throughout.
//
// HBM/L2-bound row work, no tensor cores: a block stages 32 embedding rows in shared memory (normalised once), each warp walks
// its share of the output features with lanes along K (coalesced weight reads, one pass over W per 32 rows), keeps the 32
// per-row partial sums in registers and folds them with a transposed butterfly (31 shuffles instead of 160) so that lane r
// ends up owning row r's dot product.
#include "cbx_internal.h"

namespace cbx {
namespace proj {

constexpr int ROWS = 32;
constexpr int MAX_K = 1024;

__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ x, long long n, int K, const float* __restrict__ w,
                                                      const float* __restrict__ b, int N, int normalize, float* __restrict__ y) {
  extern __shared__ float sx[];                        // [ROWS][K + 1]
  const int ld = K + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * ROWS;
  // stage + normalise: warp w owns rows w, w + 8, ...
  for (int r = warp; r < ROWS; r += 8) {
    const long long row = r0 + r;
    float ss = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float v = row < n ? x[row * K + k] : 0.f;
      sx[r * ld + k] = v;
      ss = fmaf(v, v, ss);
    }
    if (normalize) {
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);           // F.normalize(dim=1, eps=1e-12)
      for (int k = lane; k < K; k += 32) sx[r * ld + k] *= inv;
    }
  }
  __syncthreads();
  for (int o = blockIdx.y * 8 + warp; o < N; o += gridDim.y * 8) {
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(w + (size_t)o * K + k);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(wv, sx[r * ld + k], acc[r]);
    }
    // transposed butterfly: after the step with offset s a lane keeps the half of the rows whose bit s matches its own
#pragma unroll
    for (int s = 16; s; s >>= 1) {
      const bool up = lane & s;
#pragma unroll
      for (int j = 0; j < s; ++j) {
        const float keep = up ? acc[j + s] : acc[j];
        const float send = up ? acc[j] : acc[j + s];
        acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    const long long row = r0 + lane;                   // lane r now holds row r
    if (row < n) y[row * N + o] = acc[0] + (b ? __ldg(b + o) : 0.f);
  }
}

}  // namespace proj
}  // namespace cbx

using namespace cbx;

extern "C" int cbx_project(cbx_ctx* c, const float* x_dev, int64_t n, int in_dim, const float* w_dev, const float* b_dev, int out_dim,
                           int normalize, float* y_dev, void* stream) {
  if (!c) return CBX_ERR_ARG;
  if (!x_dev || !w_dev || !y_dev || n < 0 || in_dim <= 0 || in_dim > proj::MAX_K || out_dim <= 0) { c->err = "cbx_project: bad argument"; return CBX_ERR_ARG; }
  if (n == 0) return CBX_OK;
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  enter_stream(c, st);
  const size_t smem = (size_t)proj::ROWS * (in_dim + 1) * sizeof(float);
  ensure_max_smem(proj::project_kernel, proj::ROWS * (proj::MAX_K + 1) * 4);
  const long long blocks = (n + proj::ROWS - 1) / proj::ROWS;
  if (blocks > 0x7fffffffLL) { c->err = "cbx_project: too many rows"; return CBX_ERR_ARG; }
  // few rows: spread the output features over more blocks so a single profile still fills the SMs
  int gy = 1;
  while (blocks * gy < 2 * 148 && gy * 8 < out_dim) gy *= 2;
  {
    Scope sc(c->launches, st, "project_kernel", 2.0 * n * in_dim * out_dim, 4.0 * n * (in_dim + out_dim));
    proj::project_kernel<<<dim3((unsigned)blocks, gy), 256, smem, st>>>(x_dev, n, in_dim, w_dev, b_dev, out_dim, normalize, y_dev);
  }
  CBX_CUDA_OK(c, cudaGetLastError());
  return CBX_OK;
}

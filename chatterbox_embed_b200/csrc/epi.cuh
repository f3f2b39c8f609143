// Epilogue functors of the tcgen05 GEMM engine.  Each call owns one output row m and 32 consecutive columns [n0, n0+32)
// held in registers (fp32, straight from TMEM).
#pragma once
#include "cbx_internal.h"

namespace cbx {
namespace tc {

__device__ __forceinline__ void store32(float* dst, const float* v) {
  float4* o = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

struct EpiBias {                 // out = acc + bias                                  (LSTM input projections)
  float* out; int ld; const float* bias; int M;
  __device__ void operator()(int m, int n0, float* v) const {
    if (m >= M) return;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __ldg(bias + n0 + i);
    store32(out + (size_t)m * ld + n0, v);
  }
};

struct EpiBiasReluMask {         // out = row is a real frame ? relu(acc + bias) : 0   (TDNN, dense-layer bottleneck)
  float* out; int ld; const float* bias; const int32_t* row_clip; int M;
  __device__ void operator()(int m, int n0, float* v) const {
    if (m >= M) return;
    const bool live = row_clip[m] >= 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = live ? fmaxf(v[i] + __ldg(bias + n0 + i), 0.f) : 0.f;
    store32(out + (size_t)m * ld + n0, v);
  }
};

struct EpiMask {                 // out = row is a real frame ? acc : 0                (transit layers)
  float* out; int ld; const int32_t* row_clip; int M;
  __device__ void operator()(int m, int n0, float* v) const {
    if (m >= M) return;
    const bool live = row_clip[m] >= 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = live ? v[i] : 0.f;
    store32(out + (size_t)m * ld + n0, v);
  }
};

struct EpiGate {                 // out[:, col0 + n] = acc * gate[segment(row)][n]     (CAM local conv, N = 32)
  float* out; int ld; int col0; const float* gate; const int32_t* row_seg; int M;
  __device__ void operator()(int m, int n0, float* v) const {
    if (m >= M) return;
    const int s = row_seg[m];
    const float4* g = reinterpret_cast<const float4*>(gate + (size_t)max(s, 0) * kGrowth + n0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 gg = __ldg(g + i);
      v[4 * i] = s >= 0 ? v[4 * i] * gg.x : 0.f;
      v[4 * i + 1] = s >= 0 ? v[4 * i + 1] * gg.y : 0.f;
      v[4 * i + 2] = s >= 0 ? v[4 * i + 2] * gg.z : 0.f;
      v[4 * i + 3] = s >= 0 ? v[4 * i + 3] * gg.w : 0.f;
    }
    store32(out + (size_t)m * ld + col0 + n0, v);
  }
};

}  // namespace tc
}  // namespace cbx

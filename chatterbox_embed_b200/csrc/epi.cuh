// Epilogue functors of the tcgen05 GEMM engine.  Each call owns one output row m and 32 consecutive columns [n0, n0+32)
// held in registers (fp32, straight from TMEM).
#pragma once
#include "cbx_internal.h"
#include "tc.cuh"

namespace cbx {
namespace tc {

__device__ __forceinline__ void store32(float* dst, const float* v) {
  float4* o = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// 32 values -> 32 bf16 (round to nearest even), 64 contiguous bytes: the bf16 copy of the concatenation buffer (option cat_bf16)
// (pack_bf16x2 lives in tc.cuh)
__device__ __forceinline__ void store32_bf16(uint16_t* dst, const float* v) {
  uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                      pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
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
  uint16_t* shadow = nullptr; int ldh = 0;                     // bf16 copy of the same tile (option cat_bf16)
  __device__ void operator()(int m, int n0, float* v) const {
    if (m >= M) return;
    const bool live = row_clip[m] >= 0;
    const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bb = __ldg(b4 + j);
      v[4 * j] = live ? fmaxf(v[4 * j] + bb.x, 0.f) : 0.f;
      v[4 * j + 1] = live ? fmaxf(v[4 * j + 1] + bb.y, 0.f) : 0.f;
      v[4 * j + 2] = live ? fmaxf(v[4 * j + 2] + bb.z, 0.f) : 0.f;
      v[4 * j + 3] = live ? fmaxf(v[4 * j + 3] + bb.w, 0.f) : 0.f;
    }
    store32(out + (size_t)m * ld + n0, v);
    if (shadow) store32_bf16(shadow + (size_t)m * ldh + n0, v);
  }
};

// Bottleneck of a CAM dense layer: u = row is a real frame ? relu(acc + bias) : 0, stored, AND the per-segment column sums of
// u that the CAM context needs (seg_pooling / mean over T, xvector.py:214-231) accumulated on the fly: the 32 rows a warp
// holds are summed with a transposed butterfly (31 shuffles per 32 columns) and one 128-byte red.add per segment present.
// Warp-collective: every lane of the warp must call it (rows >= M take part as "no segment").
// The sums are accumulated in 40.24 FIXED POINT with 64-bit integer reductions: integer addition is associative, so the
// result does not depend on the order in which the warps' partial sums arrive and the x-vector is bit-reproducible from
// run to run (with float atomics every run drew a different sample of the TF32 rounding noise: the max-abs error of the
// ragged W1 batch wandered between 5e-4 and 2.3e-3).  Resolution 6e-8, range +-5e11 per column and segment.
constexpr float kSegFix = 16777216.f;                 // 2^24
// kExact (option "batch_invariant"): the warp's partial sums are exact too -- every value is split into its integer part and a
// 24-bit fraction and each part is added across the warp with redux.sync (32-bit integer adds) -- so the sum of a segment's rows
// does not depend on which rows share a warp and a clip's x-vector is bit-identical alone, in any batch and at any chunking
// (the VoiceEncoder embedding already is).  64 redux.sync per segment and chunk cost ~0.5 ms per step (a 64-bit shuffle butterfly
// spills and costs 1.1 ms), so the default is the fp32 butterfly: reproducible run to run, position dependent within 1.2e-4.
template <bool kExact>
struct EpiBiasReluMaskSegsumT {
  float* out; int ld; const float* bias; const int32_t* row_seg; unsigned long long* seg_sum; int M;     // out == nullptr: the kernel stores (TMA)
  uint16_t* u16 = nullptr; int ldu = 0;        // bf16 mode: u goes out as bf16 from here and the kernel's fp32 TMA store is skipped
  __device__ bool skip_c() const { return u16 != nullptr; }
  __device__ void operator()(int m, int n0, float* v) const {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int seg = m < M ? row_seg[m] : -1;
    // bias as eight 16-byte loads, unconditionally (one LSU request each instead of 32 predicated scalar ones: the LSU queue
    // is shared with the other CTA's prefetch stream)
    const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bb = __ldg(b4 + j);
      v[4 * j] = seg >= 0 ? fmaxf(v[4 * j] + bb.x, 0.f) : 0.f;
      v[4 * j + 1] = seg >= 0 ? fmaxf(v[4 * j + 1] + bb.y, 0.f) : 0.f;
      v[4 * j + 2] = seg >= 0 ? fmaxf(v[4 * j + 2] + bb.z, 0.f) : 0.f;
      v[4 * j + 3] = seg >= 0 ? fmaxf(v[4 * j + 3] + bb.w, 0.f) : 0.f;
    }
    if (out && m < M) store32(out + (size_t)m * ld + n0, v);
    if (u16 && m < M) store32_bf16(u16 + (size_t)m * ldu + n0, v);
    unsigned rem = __ballot_sync(full, seg >= 0);
    while (rem) {
      const int s = __shfl_sync(full, seg, __ffs(rem) - 1);
      rem &= ~__ballot_sync(full, seg == s);
      const bool mine = seg == s;
      if constexpr (kExact) {
        int keep_hi = 0; unsigned keep_lo = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = mine ? fminf(v[i], 1.0e9f) : 0.f;                  // activations are >= 0 and far below 2^31
          const int hi = __float2int_rd(x);
          const unsigned lo = __float2uint_rn((x - (float)hi) * kSegFix);    // exact: < 2^24, or 2^24 after rounding (still exact in the sum)
          const int sh = __reduce_add_sync(full, hi);
          const unsigned sl = __reduce_add_sync(full, lo);
          if (lane == i) { keep_hi = sh; keep_lo = sl; }
        }
        atomicAdd(seg_sum + (size_t)s * kBnC + n0 + lane, (unsigned long long)(((long long)keep_hi << 24) + (long long)keep_lo));
        continue;
      }
      float a[16], b[8], c[4], d[2];
      {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float lo = mine ? v[i] : 0.f, hi = mine ? v[16 + i] : 0.f;
          a[i] = (up ? hi : lo) + __shfl_xor_sync(full, up ? lo : hi, 16);
        }
      }
      {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = (up ? a[8 + i] : a[i]) + __shfl_xor_sync(full, up ? a[i] : a[8 + i], 8);
      }
      {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = (up ? b[4 + i] : b[i]) + __shfl_xor_sync(full, up ? b[i] : b[4 + i], 4);
      }
      {
        const bool up = lane & 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) d[i] = (up ? c[2 + i] : c[i]) + __shfl_xor_sync(full, up ? c[i] : c[2 + i], 2);
      }
      const bool up = lane & 1;
      const float e = (up ? d[1] : d[0]) + __shfl_xor_sync(full, up ? d[0] : d[1], 1);
      // lane l holds the sum of column n0 + l (a fixed-order tree over the warp's rows)
      atomicAdd(seg_sum + (size_t)s * kBnC + n0 + lane, (unsigned long long)__float2ll_rn(e * kSegFix));
    }
  }
};

using EpiBiasReluMaskSegsum = EpiBiasReluMaskSegsumT<false>;
using EpiBiasReluMaskSegsumExact = EpiBiasReluMaskSegsumT<true>;

struct EpiMask {                 // out = row is a real frame ? acc : 0                (transit layers)
  float* out; int ld; const int32_t* row_clip; int M;           // out == nullptr: the kernel stores (TMA)
  uint16_t* shadow = nullptr; int ldh = 0;                     // bf16 copy of the same tile (option cat_bf16)
  __device__ void operator()(int m, int n0, float* v) const {
    const bool live = m < M && row_clip[m] >= 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = live ? v[i] : 0.f;
    if (out && m < M) store32(out + (size_t)m * ld + n0, v);
    if (shadow && m < M) store32_bf16(shadow + (size_t)m * ldh + n0, v);
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

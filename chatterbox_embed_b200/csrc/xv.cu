// CAMPPlus path: Kaldi 80-bin log-fbank + CMN -> FCM 2-D ResNet head -> TDNN(k5,s2) -> 3 CAM-dense-TDNN blocks
// + transits -> BN/ReLU -> statistics pooling -> dense + BN.   Reference: xvector.py:45-58, 61-127, 146-428.
//
// Layout: every activation is time-major with channels contiguous.  Clips are laid back to back on one row
// axis, separated by zero "guard" rows (kGuardTd rows in the T' domain, twice that in the fbank domain), so a
// time-shifted read implements the convolutions' zero padding and no kernel needs per-clip bounds.
#include <math.h>

#include "cbx_internal.h"
#include "sgemm.cuh"
#include "tc.cuh"
#include "epi.cuh"

namespace cbx {

// ---- row maps ------------------------------------------------------------------------------------------
__global__ void xv_maps_kernel(const ClipPlan* __restrict__ plan, int32_t* __restrict__ fb_row_clip,
                               int32_t* __restrict__ td_row_clip, int32_t* __restrict__ td_row_seg,
                               int32_t* __restrict__ seg_clip) {
  const int c = blockIdx.x;
  const ClipPlan cp = plan[c];
  for (int t = threadIdx.x; t < cp.xv_frames; t += blockDim.x) fb_row_clip[cp.fb_row + t] = c;
  for (int t = threadIdx.x; t < cp.xv_tdnn; t += blockDim.x) {
    td_row_clip[cp.td_row + t] = c;
    td_row_seg[cp.td_row + t] = cp.seg0 + t / kSegLen;
  }
  for (int s = threadIdx.x; s < cp.xv_segs; s += blockDim.x) seg_clip[cp.seg0 + s] = c;
}

// ---- K9: Kaldi fbank.  Frames are raw 400-sample windows (snip_edges); DC removal, pre-emphasis and the Povey
// window are folded into the [514][400] DFT matrix, so the GEMM reads PCM directly. ---------------------------
struct KaldiFrameGather {
  const float* pcm; const ClipPlan* plan; const int32_t* row_clip; int row0;
  __device__ float operator()(int m, int k) const {
    const int r = row0 + m;
    const int c = row_clip[r];
    if (c < 0) return 0.f;
    return __ldg(pcm + plan[c].pcm_off + (size_t)(r - plan[c].fb_row) * kKHop + k);
  }
};
struct StoreRM {
  float* out; int ld;
  __device__ void operator()(int m, int n, float v) const { out[(size_t)m * ld + n] = v; }
};

__global__ void __launch_bounds__(256) fbank_from_spec_kernel(const float* __restrict__ spec, const float* __restrict__ bank,
                                                              const int32_t* __restrict__ row_clip, float* __restrict__ fbank,
                                                              int row0, int rows) {
  __shared__ float pw[8][kKBins + 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = blockIdx.x * 8 + warp;
  if (lr >= rows) return;
  const int r = row0 + lr;
  if (row_clip[r] < 0) {
    for (int m = lane; m < kKMels; m += 32) fbank[(size_t)r * kKMels + m] = 0.f;
    return;
  }
  const float2* sp = reinterpret_cast<const float2*>(spec + (size_t)lr * kKSpecN);
  for (int k = lane; k < kKBins; k += 32) { float2 v = sp[k]; pw[warp][k] = v.x * v.x + v.y * v.y; }
  __syncwarp();
  for (int m = lane; m < kKMels; m += 32) {
    const float* b = bank + (size_t)m * kKBins;
    float a = 0.f;
    for (int k = 0; k < kKBins; ++k) a = fmaf(__ldg(b + k), pw[warp][k], a);
    fbank[(size_t)r * kKMels + m] = logf(fmaxf(a, 1.1920928955078125e-07f));
  }
}

// K10: per-clip column mean of the log-fbank (xvector.py:51); deterministic two-level sum: 12 stripes of frames per column, summed in
// a fixed order (960 threads per clip: with 4 stripes the kernel was a 0.7 TB/s latency chain).
constexpr int kCmnStripes = 12;
__global__ void __launch_bounds__(kCmnStripes * kKMels) cmn_mean_kernel(const float* __restrict__ fbank, const ClipPlan* __restrict__ plan,
                                                                        float* __restrict__ cmn_mean) {
  __shared__ float part[kCmnStripes][kKMels];
  const ClipPlan cp = plan[blockIdx.x];
  const int col = threadIdx.x % kKMels, stripe = threadIdx.x / kKMels;
  float a = 0.f;
  const float* src = fbank + (size_t)cp.fb_row * kKMels + col;
  int t = stripe;
  // eight loads in flight per thread; the additions keep their order (same bits as the plain loop)
  for (; t + 7 * kCmnStripes < cp.xv_frames; t += 8 * kCmnStripes) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = src[(size_t)(t + k * kCmnStripes) * kKMels];
#pragma unroll
    for (int k = 0; k < 8; ++k) a += v[k];
  }
  for (; t < cp.xv_frames; t += kCmnStripes) a += src[(size_t)t * kKMels];
  part[stripe][col] = a;
  __syncthreads();
  if (stripe == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kCmnStripes; ++i) s += part[i][col];
    cmn_mean[blockIdx.x * kKMels + col] = cp.xv_frames > 0 ? s / (float)cp.xv_frames : 0.f;
  }
}

// ---- K11: FCM head -------------------------------------------------------------------------------------
// conv1: Conv2d(1->32, 3x3, pad 1) + BN + ReLU on the CMN'd fbank.  One warp = the 32 channels of one (row, f).
__global__ void __launch_bounds__(256) fcm_conv1_kernel(const float* __restrict__ fbank, const float* __restrict__ cmn_mean,
                                                        const int32_t* __restrict__ row_clip, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int row0, int rows, int fb_rows) {
  const int co = threadIdx.x & 31;
  const long long pos = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pos >= (long long)rows * kKMels) return;
  const int lr = (int)(pos / kKMels), f = (int)(pos - (long long)lr * kKMels);
  const int r = row0 + lr;
  float v = 0.f;
  if (row_clip[r] >= 0) {
    float acc = __ldg(bias + co);
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int rr = r + kw - 1;
      if (rr < 0 || rr >= fb_rows) continue;
      const int cc = row_clip[rr];
      if (cc < 0) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ff = f + kh - 1;
        if (ff < 0 || ff >= kKMels) continue;
        const float x = fbank[(size_t)rr * kKMels + ff] - cmn_mean[cc * kKMels + ff];
        acc = fmaf(x, __ldg(w + co * 9 + kh * 3 + kw), acc);
      }
    }
    v = fmaxf(acc, 0.f);
  }
  out[((size_t)lr * kKMels + f) * kFcmC + co] = v;
}

// conv1, one warp per fbank row: the three CMN'd input rows go to shared memory once, lane = output channel (weights in
// registers), 128-byte coalesced stores.
__global__ void __launch_bounds__(256) fcm_conv1_rows_kernel(const float* __restrict__ fbank, const float* __restrict__ cmn_mean,
                                                             const int32_t* __restrict__ row_clip, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int row0, int rows, int fb_rows) {
  __shared__ float xin[8][3][kKMels + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = blockIdx.x * 8 + warp;
  tc::pdl_trigger();            // the first residual conv may load its weights meanwhile (this kernel itself is launched serialised)
  if (lr >= rows) return;
  const int r = row0 + lr;
  const int c = row_clip[r];
  float* o = out + (size_t)lr * kKMels * kFcmC + lane;
  if (c < 0) {
    for (int f = 0; f < kKMels; ++f) o[(size_t)f * kFcmC] = 0.f;
    return;
  }
  for (int kw = 0; kw < 3; ++kw) {
    const int rr = r + kw - 1;
    const bool ok = rr >= 0 && rr < fb_rows && row_clip[rr] == c;
    for (int ff = lane; ff < kKMels; ff += 32)
      xin[warp][kw][ff + 1] = ok ? fbank[(size_t)rr * kKMels + ff] - cmn_mean[c * kKMels + ff] : 0.f;
    if (lane == 0) { xin[warp][kw][0] = 0.f; xin[warp][kw][kKMels + 1] = 0.f; }
  }
  float wr[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wr[i] = __ldg(w + lane * 9 + i);
  const float b = __ldg(bias + lane);
  __syncwarp();
#pragma unroll 4
  for (int f = 0; f < kKMels; ++f) {
    float acc = b;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) acc = fmaf(xin[warp][kw][f + kh], wr[kh * 3 + kw], acc);
    o[(size_t)f * kFcmC] = fmaxf(acc, 0.f);
  }
}

// 3x3 conv over [row][F_in][32] as an implicit GEMM (k = (kh*3+kw)*32 + ci), optionally with the block's
// 1x1 stride-2 shortcut conv appended as 32 extra K columns (k >= 288).
struct FcmConvA {
  const float* in; int F_in, F_out, sf;
  const float* sc; int F_sc;
  __device__ float operator()(int m, int k) const {
    const int lr = m / F_out, f = m - lr * F_out;
    if (k < 288) {
      const int tap = k >> 5, ci = k & 31;
      const int kh = tap / 3, kw = tap - kh * 3;
      const int fi = f * sf + kh - 1;
      if (fi < 0 || fi >= F_in) return 0.f;
      return in[((long long)(lr + kw - 1) * F_in + fi) * kFcmC + ci];     // lr-1 / lr+1 hit the buffer's pad rows at the ends
    }
    return sc[((long long)lr * F_sc + 2 * f) * kFcmC + (k - 288)];
  }
};
struct FcmEpi {
  float* out; const float* bias; const float* res; const int32_t* row_clip; int F_out;
  __device__ void operator()(int m, int n, float acc) const {
    const int lr = m / F_out;
    float v = acc + __ldg(bias + n);
    if (res) v += res[(size_t)m * kFcmC + n];
    v = fmaxf(v, 0.f);
    out[(size_t)m * kFcmC + n] = row_clip[lr] >= 0 ? v : 0.f;
  }
};

// ---- K12: TDNN Conv1d(320->128, k5, stride 2, pad 2) + BN + ReLU ---------------------------------------------
struct TdnnA {
  const float* fcm; int fb_rows;
  __device__ float operator()(int m, int k) const {
    const int tap = k / kFcmOut, j = k - tap * kFcmOut;
    const int r = 2 * m + tap - 2;
    if (r < 0 || r >= fb_rows) return 0.f;
    return fcm[(size_t)r * kFcmOut + j];
  }
};
struct BiasReluMaskEpi {
  float* out; int ld; const float* bias; const int32_t* row_clip;
  __device__ void operator()(int m, int n, float acc) const {
    out[(size_t)m * ld + n] = row_clip[m] >= 0 ? fmaxf(acc + __ldg(bias + n), 0.f) : 0.f;
  }
};

// ---- K13: CAM dense TDNN layer ---------------------------------------------------------------------------
struct BnReluA {   // BN + ReLU applied while gathering the A operand (pre-activation order, xvector.py:266-271)
  const float* x; int ld; const float* a; const float* b;
  __device__ float operator()(int m, int k) const { return fmaxf(fmaf(x[(size_t)m * ld + k], __ldg(a + k), __ldg(b + k)), 0.f); }
};
struct MaskEpi {
  float* out; int ld; const int32_t* row_clip;
  __device__ void operator()(int m, int n, float acc) const { out[(size_t)m * ld + n] = row_clip[m] >= 0 ? acc : 0.f; }
};

// per-segment column sums of u (seg_pooling, xvector.py:221-231): one CTA per 100-frame segment
__global__ void __launch_bounds__(128) seg_sum_kernel(const float* __restrict__ u, const ClipPlan* __restrict__ plan,
                                                      const int32_t* __restrict__ seg_clip, float* __restrict__ seg_sum) {
  const int s = blockIdx.x, ch = threadIdx.x;
  const ClipPlan cp = plan[seg_clip[s]];
  const int t0 = (s - cp.seg0) * kSegLen, t1 = min(t0 + kSegLen, cp.xv_tdnn);
  float a = 0.f;
  for (int t = t0; t < t1; ++t) a += u[(size_t)(cp.td_row + t) * kBnC + ch];
  seg_sum[(size_t)s * kBnC + ch] = a;
}

// gate m = sigmoid(W2 relu(W1 (mean_T u + segmean u) + b1) + b2), constant over a segment (xvector.py:214-219)
__global__ void __launch_bounds__(128) cam_gate_kernel(const float* __restrict__ seg_sum, const ClipPlan* __restrict__ plan,
                                                       const int32_t* __restrict__ seg_clip, DenseLayerW L,
                                                       float* __restrict__ gate) {
  __shared__ float ctx[kBnC];
  __shared__ float hid[kCamHid];
  const int s = blockIdx.x, ch = threadIdx.x;
  const ClipPlan cp = plan[seg_clip[s]];
  float tot = 0.f;
  for (int i = 0; i < cp.xv_segs; ++i) tot += seg_sum[(size_t)(cp.seg0 + i) * kBnC + ch];
  const int t0 = (s - cp.seg0) * kSegLen, len = min(kSegLen, cp.xv_tdnn - t0);
  ctx[ch] = tot / (float)cp.xv_tdnn + seg_sum[(size_t)s * kBnC + ch] / (float)len;
  __syncthreads();
  if (ch < kCamHid) {
    float a = __ldg(L.bc1 + ch);
    for (int k = 0; k < kBnC; ++k) a = fmaf(__ldg(L.wc1 + ch * kBnC + k), ctx[k], a);
    hid[ch] = fmaxf(a, 0.f);
  }
  __syncthreads();
  if (ch < kGrowth) {
    float a = __ldg(L.bc2 + ch);
    for (int k = 0; k < kCamHid; ++k) a = fmaf(__ldg(L.wc2 + ch * kCamHid + k), hid[k], a);
    gate[(size_t)s * kGrowth + ch] = 1.f / (1.f + expf(-a));
  }
}

constexpr int kMaxSegsSm = 8;      // segments per pass of the gate kernel
// Gate from the segment sums the bottleneck epilogue accumulated (tensor-core mode), one CTA per clip; the sums are zeroed
// after use so the next layer's epilogue can accumulate into the same buffer.
__global__ void __launch_bounds__(256) cam_gate_clip_kernel(unsigned long long* __restrict__ seg_sum, const ClipPlan* __restrict__ plan,
                                                            DenseLayerW L, float* __restrict__ gate) {
  __shared__ float w1[kBnC * kCamHid];        // [128][64] transposed
  __shared__ float w2[kCamHid * kGrowth];     // [64][32] transposed
  __shared__ float ctx[kMaxSegsSm][kBnC];
  __shared__ float hid[kMaxSegsSm][kCamHid];
  __shared__ float tot[kBnC];
  const int tid = threadIdx.x;
  // weights -> shared memory (coalesced 16-byte loads, all in flight at once)
  for (int i = tid; i < kBnC * kCamHid / 4; i += 256) reinterpret_cast<float4*>(w1)[i] = __ldg(reinterpret_cast<const float4*>(L.wc1T) + i);
  for (int i = tid; i < kCamHid * kGrowth / 4; i += 256) reinterpret_cast<float4*>(w2)[i] = __ldg(reinterpret_cast<const float4*>(L.wc2T) + i);
  tc::pdl_trigger();            // short kernel: the local conv behind it may set up right away
  tc::pdl_wait();               // only the layer's weights (static since load time) are read before this point: everything written
                                // during this call -- the plan table included -- is visible only from here on
  const ClipPlan cp = plan[blockIdx.x];
  const int T = cp.xv_tdnn, S = cp.xv_segs;
  unsigned long long* ss = seg_sum + (size_t)cp.seg0 * kBnC;      // 40.24 fixed point (EpiBiasReluMaskSegsum)
  auto seg_val = [&](int s, int ch) { return (float)((double)(long long)ss[(size_t)s * kBnC + ch] * (1.0 / (double)tc::kSegFix)); };
  if (tid < kBnC) {
    float t = 0.f;
    for (int s = 0; s < S; ++s) t += seg_val(s, tid);
    tot[tid] = t / (float)T;
  }
  __syncthreads();
  for (int s0 = 0; s0 < S; s0 += kMaxSegsSm) {
    const int ns = min(kMaxSegsSm, S - s0);
    for (int o = tid; o < ns * kBnC; o += 256) {
      const int s = o >> 7, ch = o & 127;
      const int len = min(kSegLen, T - (s0 + s) * kSegLen);
      ctx[s][ch] = tot[ch] + seg_val(s0 + s, ch) / (float)len;
    }
    __syncthreads();
    for (int o = tid; o < ns * kCamHid; o += 256) {
      const int s = o / kCamHid, hch = o - s * kCamHid;
      float a = __ldg(L.bc1 + hch);
#pragma unroll 8
      for (int k = 0; k < kBnC; ++k) a = fmaf(w1[k * kCamHid + hch], ctx[s][k], a);
      hid[s][hch] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int o = tid; o < ns * kGrowth; o += 256) {
      const int s = o / kGrowth, g = o - s * kGrowth;
      float a = __ldg(L.bc2 + g);
#pragma unroll 8
      for (int k = 0; k < kCamHid; ++k) a = fmaf(w2[k * kGrowth + g], hid[s][k], a);
      gate[(size_t)(cp.seg0 + s0 + s) * kGrowth + g] = 1.f / (1.f + expf(-a));
    }
    __syncthreads();
  }
  for (int o = tid; o < S * kBnC; o += 256) ss[o] = 0ull;
}

struct LocalConvA {   // Conv1d(128->32, k3, dilation d, zero pad d): k = tap*128 + c
  const float* u; int dil; int td_rows;
  __device__ float operator()(int m, int k) const {
    const int tap = k >> 7, c = k & 127;
    const int r = m + (tap - 1) * dil;
    if (r < 0 || r >= td_rows) return 0.f;
    return u[(size_t)r * kBnC + c];
  }
};
struct GateEpi {
  float* out; int ld; int col0; const float* gate; const int32_t* row_seg;
  __device__ void operator()(int m, int n, float acc) const {
    const int s = row_seg[m];
    out[(size_t)m * ld + col0 + n] = s >= 0 ? acc * gate[(size_t)s * kGrowth + n] : 0.f;
  }
};

// ---- K15/K16: out_nonlinear BN+ReLU, statistics pooling (mean, unbiased std), dense + BN -----------------
// Warp-reduction kernel (xvector.py:146-152): one CTA per (clip, 32-channel slab) -- 16 slabs x clips CTAs fill the GPU even for a
// handful of clips.  A lane is (time phase 0..3) x (channel quad 0..7): one warp instruction reads four consecutive rows x 128
// contiguous bytes; the four phases are folded with two shuffles, the eight warps through shared memory in a fixed order
// (deterministic).  Two passes like torch.std (mean first, then squared deviations; the second pass hits L2).
constexpr int kSpSlab = 32, kSpWarps = 8;
__device__ __forceinline__ float4 sp_fold(float4 v) {
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  return v;
}
__global__ void __launch_bounds__(32 * kSpWarps) stats_pool_kernel(const float* __restrict__ x, const ClipPlan* __restrict__ plan,
                                                                   const float* __restrict__ a, const float* __restrict__ b,
                                                                   float* __restrict__ stats) {
  __shared__ float4 red[kSpWarps][8];
  __shared__ float4 mean_s[8];
  const ClipPlan cp = plan[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, quad = lane & 7, phase = lane >> 3;
  const int c0 = blockIdx.y * kSpSlab + quad * 4;
  const float4 sa = *reinterpret_cast<const float4*>(a + c0), sb = *reinterpret_cast<const float4*>(b + c0);
  const int T = cp.xv_tdnn;
  const float* base = x + (size_t)cp.td_row * kStatsC + c0;
  auto act = [&](int t) {
    float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)t * kStatsC));
    v.x = fmaxf(fmaf(v.x, sa.x, sb.x), 0.f); v.y = fmaxf(fmaf(v.y, sa.y, sb.y), 0.f);
    v.z = fmaxf(fmaf(v.z, sa.z, sb.z), 0.f); v.w = fmaxf(fmaf(v.w, sa.w, sb.w), 0.f);
    return v;
  };
  auto across_warps = [&](float4 v) {      // -> the CTA-wide sum in threads 0..7 (quad = threadIdx.x)
    v = sp_fold(v);
    if (phase == 0) red[warp][quad] = v;
    __syncthreads();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < 8) {
#pragma unroll
      for (int w = 0; w < kSpWarps; ++w) { const float4 r = red[w][threadIdx.x]; t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
    }
    return t;
  };
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int t = warp * 4 + phase; t < T; t += 4 * kSpWarps) { const float4 v = act(t); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
  s = across_warps(s);
  const float nanv = nanf("");
  if (threadIdx.x < 8) {
    const float ft = (float)T;
    mean_s[threadIdx.x] = T > 0 ? make_float4(s.x / ft, s.y / ft, s.z / ft, s.w / ft) : make_float4(nanv, nanv, nanv, nanv);
  }
  __syncthreads();
  const float4 m = mean_s[quad];
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int t = warp * 4 + phase; t < T; t += 4 * kSpWarps) {
    const float4 v = act(t);
    const float dx = v.x - m.x, dy = v.y - m.y, dz = v.z - m.z, dw = v.w - m.w;
    q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
  }
  q = across_warps(q);
  if (threadIdx.x < 8) {
    const float d = (float)(T - 1);          // T'=1 -> 0/0 = NaN like torch.std
    float* o = stats + (size_t)blockIdx.x * 2 * kStatsC + blockIdx.y * kSpSlab + threadIdx.x * 4;
    *reinterpret_cast<float4*>(o) = mean_s[threadIdx.x];
    *reinterpret_cast<float4*>(o + kStatsC) = make_float4(sqrtf(q.x / d), sqrtf(q.y / d), sqrtf(q.z / d), sqrtf(q.w / d));
  }
}

__global__ void __launch_bounds__(256) xv_final_kernel(const float* __restrict__ stats, const ClipPlan* __restrict__ plan,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ xv_out, int32_t* __restrict__ status) {
  __shared__ float st[2 * kStatsC];
  const ClipPlan cp = plan[blockIdx.x];
  for (int i = threadIdx.x; i < 2 * kStatsC; i += 256) st[i] = stats[(size_t)blockIdx.x * 2 * kStatsC + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < kXvEmbed; n += 8) {
    float acc = 0.f;
    for (int k = lane; k < 2 * kStatsC; k += 32) acc = fmaf(__ldg(w + (size_t)n * 2 * kStatsC + k), st[k], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) xv_out[(size_t)cp.out_index * kXvEmbed + n] = cp.xv_frames > 0 ? acc + __ldg(bias + n) : nanf("");
  }
  if (threadIdx.x == 0 && status && cp.xv_frames <= 0) atomicOr(status + cp.out_index, CBX_CLIP_XV_TOO_SHORT);
}

// ---------------------------------------------------------------------------------------------------------
void run_xv_chunk(cbx_ctx* c, const float* pcm, const XvChunk& ch, float* xv_out, int32_t* status, cudaStream_t st,
                  const float* feats, const int64_t* feat_off) {
  const XvWeights& W = c->xv;
  Launches& L = c->launches;
  const ClipPlan* hp = ch.hplan;

  cudaMemsetAsync(ch.fb_row_clip, 0xff, sizeof(int32_t) * ch.fb_rows, st);
  cudaMemsetAsync(ch.td_row_clip, 0xff, sizeof(int32_t) * ch.td_rows, st);
  cudaMemsetAsync(ch.td_row_seg, 0xff, sizeof(int32_t) * ch.td_rows, st);
  { Scope sc(L, st, "xv_maps_kernel"); xv_maps_kernel<<<ch.n_clips, 256, 0, st>>>(ch.plan, ch.fb_row_clip, ch.td_row_clip, ch.td_row_seg, ch.seg_clip); }

  // sub-chunks of whole clips, each covering its leading guard rows too
  struct Sub { int r0, r1; };
  std::vector<Sub> subs;
  {
    int i = 0;
    while (i < ch.n_clips) {
      const int r0 = hp[i].fb_row - 2 * kGuardTd;
      int j = i, r1 = r0;
      while (j < ch.n_clips) {
        const int end = hp[j].fb_row + 2 * (hp[j].xv_tdnn + kGuardTd);
        if (j > i && end - r0 > ch.fcm_rows) break;
        r1 = end; ++j;
      }
      subs.push_back({r0, r1});
      i = j;
    }
  }
  if (feats) {
    // CAMPPlus.forward (xvector.py:417-423): the caller's feature rows go where the fbank kernel would have written them; no CMN
    cudaMemsetAsync(ch.fbank, 0, sizeof(float) * (size_t)ch.fb_rows * kKMels, st);
    cudaMemsetAsync(ch.cmn_sum, 0, sizeof(float) * (size_t)ch.n_clips * kKMels, st);
    for (int i = 0; i < ch.n_clips; ++i)
      if (hp[i].xv_frames > 0)
        cudaMemcpyAsync(ch.fbank + (size_t)hp[i].fb_row * kKMels, feats + (size_t)feat_off[hp[i].out_index] * kKMels,
                        sizeof(float) * (size_t)hp[i].xv_frames * kKMels, cudaMemcpyDeviceToDevice, st);
  } else if (c->mode == 1) run_kaldi_fbank_tc(c, pcm, ch, st);
  else for (const Sub& s : subs) {
    const int rows = s.r1 - s.r0;
    sgemm(L, st, "kaldi_dft_gemm", rows, kKSpecN, kKWin, KaldiFrameGather{pcm, ch.plan, ch.fb_row_clip, s.r0}, c->ft.k_dft, kKWin, StoreRM{ch.spec, kKSpecN});
    { Scope sc(L, st, "fbank_from_spec_kernel"); fbank_from_spec_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ch.spec, c->ft.k_mel, ch.fb_row_clip, ch.fbank, s.r0, rows); }
  }
  if (!feats) { Scope sc(L, st, "cmn_mean_kernel", 0.0, 4.0 * kKMels * ((double)ch.fb_rows + ch.n_clips)); cmn_mean_kernel<<<ch.n_clips, kCmnStripes * kKMels, 0, st>>>(ch.fbank, ch.plan, ch.cmn_sum); }

  for (const Sub& s : subs) {
    const int rows = s.r1 - s.r0;
    const int32_t* rc = ch.fb_row_clip + s.r0;
    // every sub-chunk buffer has one pad row in front (index -1 is readable)
    float* b0 = ch.b0 + 80 * kFcmC; float* b1 = ch.b1 + 40 * kFcmC; float* b2 = ch.b2 + 40 * kFcmC;
    float* b4 = ch.b4 + 20 * kFcmC; float* b5 = ch.b5 + 20 * kFcmC;
    float* b3 = ch.b3 + 40 * kFcmC; float* b6 = ch.b6 + 20 * kFcmC;
    {
      const long long npos = (long long)rows * kKMels;
      if (c->mode == 1) { Scope sc(L, st, "fcm_conv1_kernel", 2.0 * rows * 80 * kFcmC * 9, 4.0 * rows * (kKMels + 80.0 * kFcmC)); fcm_conv1_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ch.fbank, ch.cmn_sum, ch.fb_row_clip, W.conv1_w, W.conv1_b, b0, s.r0, rows, ch.fb_rows); }
      else { Scope sc(L, st, "fcm_conv1_kernel"); fcm_conv1_kernel<<<(unsigned)((npos + 7) / 8), 256, 0, st>>>(ch.fbank, ch.cmn_sum, ch.fb_row_clip, W.conv1_w, W.conv1_b, b0, s.r0, rows, ch.fb_rows); }
    }
    // in: [row][F_in][32] (pad row in front), optional shortcut source sc [row][F_sc][32], residual res / out [row][F_out][32]
    int conv_idx = 0;
    static const char* kConvTags[9] = {"fcm_conv_gemm:l1b0c1", "fcm_conv_gemm:l1b0c2", "fcm_conv_gemm:l1b1c1", "fcm_conv_gemm:l1b1c2", "fcm_conv_gemm:l2b0c1",
                                       "fcm_conv_gemm:l2b0c2", "fcm_conv_gemm:l2b1c1", "fcm_conv_gemm:l2b1c2", "fcm_conv_gemm:head2"};
    auto conv = [&](const ConvW& w, const CUtensorMap& tmw, const float* in, int F_in, int F_out, int sf, const float* sc, int F_sc,
                    const float* res, float* out) {
      if (c->mode == 1) {
        const char* tag = c->launches.prof ? kConvTags[conv_idx++ % 9] : "fcm_conv_gemm";
        run_fcm_conv_tc(c, st, tmw, w.bias, in, F_in, F_out, sf, sc, F_sc, res, out, rc, rows, ch.fcm_rows + 2, 2.0 * rows * F_out * kFcmC * w.K, tag, c->pdl != 0 && !c->launches.prof);
      } else {
        sgemm(L, st, "fcm_conv_gemm", rows * F_out, kFcmC, w.K, FcmConvA{in, F_in, F_out, sf, sc, F_sc}, w.w, w.K, FcmEpi{out, w.bias, res, rc, F_out});
      }
    };
    // layer1: 80 -> 40
    conv(W.res[0][0][0], W.tm_res[0][0][0], b0, 80, 40, 2, nullptr, 0, nullptr, b1);
    conv(W.res[0][0][1], W.tm_res[0][0][1], b1, 40, 40, 1, b0, 80, nullptr, b2);
    // identity residual blocks: ONE fused kernel (the intermediate stays in shared memory, fcm_block_tc.cu) unless fcm_fuse = 0
    const bool fuse = c->mode == 1 && c->fcm_fuse != 0;
    const bool fpdl = c->pdl != 0 && !c->launches.prof;
    if (fuse) run_fcm_block_tc(c, st, W.tm_res[0][1][0], W.res[0][1][0].bias, W.tm_res[0][1][1], W.res[0][1][1].bias, b2, 40, b3, rc, rows, ch.fcm_rows + 2,
                               c->launches.prof ? "fcm_block:l1b1" : "fcm_block", fpdl);
    else {
      conv_idx = 2;
      conv(W.res[0][1][0], W.tm_res[0][1][0], b2, 40, 40, 1, nullptr, 0, nullptr, b1);
      conv(W.res[0][1][1], W.tm_res[0][1][1], b1, 40, 40, 1, nullptr, 0, b2, b3);
    }
    conv_idx = 4;
    // layer2: 40 -> 20
    conv(W.res[1][0][0], W.tm_res[1][0][0], b3, 40, 20, 2, nullptr, 0, nullptr, b4);
    conv(W.res[1][0][1], W.tm_res[1][0][1], b4, 20, 20, 1, b3, 40, nullptr, b5);
    if (fuse) run_fcm_block_tc(c, st, W.tm_res[1][1][0], W.res[1][1][0].bias, W.tm_res[1][1][1], W.res[1][1][1].bias, b5, 20, b6, rc, rows, ch.fcm_rows + 2,
                               c->launches.prof ? "fcm_block:l2b1" : "fcm_block", fpdl);
    else {
      conv_idx = 6;
      conv(W.res[1][1][0], W.tm_res[1][1][0], b5, 20, 20, 1, nullptr, 0, nullptr, b4);
      conv(W.res[1][1][1], W.tm_res[1][1][1], b4, 20, 20, 1, nullptr, 0, b5, b6);
    }
    conv_idx = 8;
    // head.conv2: 20 -> 10, written straight into the chunk-level [row][f*32+c] buffer
    conv(W.head_conv2, W.tm_head2, b6, 20, 10, 2, nullptr, 0, nullptr, ch.fcm_out + (size_t)s.r0 * kFcmOut);
  }

  const int M = ch.td_rows;
  const bool tcm = c->mode == 1;
  const bool pdl = c->pdl != 0 && !c->launches.prof;      // event-bracketed launches (profiling) serialise anyway
  if (tcm) {
    // stride-2 conv: view the FCM output as [fb_rows/2][640] so that output row m reads the row pairs m-1, m, m+1
    CUtensorMap tmA = tc::make_map_2d(ch.fcm_out, ch.fb_rows / 2, 2 * kFcmOut, 2 * kFcmOut, tc::BM, true);
    tc::TapMap tap{};
    tap.cpb = kFcmOut / tc::BK;
    const int sh[5] = {-1, -1, 0, 0, 1}, c0[5] = {0, kFcmOut, 0, kFcmOut, 0};
    for (int i = 0; i < 5; ++i) { tap.shift[i] = sh[i]; tap.col0[i] = c0[i]; }
    tc::tgemm<128, 3>(L, st, "tdnn_gemm", tmA, W.tm_tdnn, M, kTdnnC, W.tdnn.K, tap, 5, tc::NoPrologue{},
                      tc::EpiBiasReluMask{ch.cat1, 512, W.tdnn.bias, ch.td_row_clip, M, ch.cat1h, 512});
  } else {
    sgemm(L, st, "tdnn_gemm", M, kTdnnC, W.tdnn.K, TdnnA{ch.fcm_out, ch.fb_rows}, W.tdnn.w, W.tdnn.K, BiasReluMaskEpi{ch.cat1, 512, W.tdnn.bias, ch.td_row_clip});
  }

  // the D-TDNN phase starts here: its GEMMs are many small CTAs (two per SM, 75 KB of shared memory each) that fill whatever SMs are
  // free -- this is the phase the VoiceEncoder's recurrence (112 SMs, one 226 KB CTA each) should share the GPU with (api.cu)
  if (c->xv_mark_dtdnn) cudaEventRecord(c->ev_dtdnn, st);
  if (tcm && ch.segs > 0) cudaMemsetAsync(ch.seg_sum, 0, sizeof(unsigned long long) * (size_t)ch.segs * kBnC, st);
  static const int kLayers[3] = {12, 24, 16};
  static const int kDil[3] = {1, 2, 2};
  float* cats[3] = {ch.cat1, ch.cat2, ch.cat3};
  uint16_t* cath[3] = {ch.cat1h, ch.cat2h, ch.cat3h};      // bf16 copies (option cat_bf16, tensor-core mode): what the GEMMs then read
  const bool bf = tcm && ch.cat1h != nullptr;
  const int lds[3] = {512, 1024, 1024};
  int li = 0;
  for (int b = 0; b < 3; ++b) {
    float* cat = cats[b];
    const int ld = lds[b];
    CUtensorMap tm_cat, tm_u;
    tc::TapMap tap_local{};
    if (tcm) {
      // halo block: 128 + 2 d frames (bf16 mode: of the bf16 u the bottleneck GEMM wrote)
      tm_u = ch.u16 ? tc::make_map_2d_bf16(ch.u16, M, kBnC, kBnC, tc::BM + 2 * kDil[b]) : tc::make_map_2d(ch.u, M, kBnC, kBnC, tc::BM + 2 * kDil[b], true);
      tm_cat = tc::make_map_2d(cat, M, ld, ld, tc::BM, false);                           // store target of the 32 new channels
      tap_local.cpb = kBnC / tc::BK;
      for (int t = 0; t < 3; ++t) tap_local.shift[t] = (t - 1) * kDil[b];
    }
    for (int i = 0; i < kLayers[b]; ++i, ++li) {
      const DenseLayerW& D = W.dense[li];
      if (tcm && c->batch_invariant)
        tc::tgemm_bnrelu<128, 2>(L, st, "dense_bottleneck_gemm", cat, ld, D.a1, D.b1, W.tm_w1[li], ch.u, kBnC, M, kBnC, D.cin,
                                 tc::EpiBiasReluMaskSegsumExact{nullptr, kBnC, D.t2, ch.td_row_seg, reinterpret_cast<unsigned long long*>(ch.seg_sum), M}, pdl);
      else if (bf && c->cat_bf16 == 2)
        tc::tgemm_bnrelu<128, 2, tc::EpiBiasReluMaskSegsum, 2>(L, st, "dense_bottleneck_gemm", cath[b], ld, D.a1, D.b1, W.tm_w1h[li], ch.u, kBnC, M, kBnC, D.cin,
                                 tc::EpiBiasReluMaskSegsum{nullptr, kBnC, D.t2, ch.td_row_seg, reinterpret_cast<unsigned long long*>(ch.seg_sum), M, ch.u16, kBnC}, pdl);
      else if (bf)
        tc::tgemm_bnrelu<128, 2, tc::EpiBiasReluMaskSegsum, 1>(L, st, "dense_bottleneck_gemm", cath[b], ld, D.a1, D.b1, W.tm_w1[li], ch.u, kBnC, M, kBnC, D.cin,
                                 tc::EpiBiasReluMaskSegsum{nullptr, kBnC, D.t2, ch.td_row_seg, reinterpret_cast<unsigned long long*>(ch.seg_sum), M, ch.u16, kBnC}, pdl);
      else if (tcm)
        tc::tgemm_bnrelu<128, 2>(L, st, "dense_bottleneck_gemm", cat, ld, D.a1, D.b1, W.tm_w1[li], ch.u, kBnC, M, kBnC, D.cin,
                                 tc::EpiBiasReluMaskSegsum{nullptr, kBnC, D.t2, ch.td_row_seg, reinterpret_cast<unsigned long long*>(ch.seg_sum), M, ch.u16, kBnC}, pdl);
      else
        sgemm(L, st, "dense_bottleneck_gemm", M, kBnC, D.cin, BnReluA{cat, ld, D.a1, D.b1}, D.w1, D.cin, BiasReluMaskEpi{ch.u, kBnC, D.t2, ch.td_row_clip});
#ifdef CBX_DEV_TOOLS
      if (ch.segs > 0 && tcm && (c->probe & 1)) {
        // timing probe: what the step costs without this kernel on the chain (stale gates, sums never zeroed: wrong results)
      } else
#endif
      if (ch.segs > 0 && tcm) {
        // latency-bound (one CTA per clip); bytes: the fixed-point segment sums in, the gates out, the two small weight matrices per CTA (L2)
        Scope sc(L, st, "cam_gate_kernel", 2.0 * (ch.segs + ch.n_clips) * (kBnC * kCamHid + kCamHid * kGrowth),
                 8.0 * ch.segs * kBnC + 4.0 * ch.segs * kGrowth + 4.0 * ch.n_clips * (kBnC * kCamHid + kCamHid * kGrowth));
        tc::launch_pdl(cam_gate_clip_kernel, dim3(ch.n_clips), dim3(256), 0, st, pdl, reinterpret_cast<unsigned long long*>(ch.seg_sum), ch.plan, D, ch.gate);
      } else if (ch.segs > 0) {
        { Scope sc(L, st, "seg_sum_kernel"); seg_sum_kernel<<<ch.segs, 128, 0, st>>>(ch.u, ch.plan, ch.seg_clip, ch.seg_sum); }
        { Scope sc(L, st, "cam_gate_kernel"); cam_gate_kernel<<<ch.segs, 128, 0, st>>>(ch.seg_sum, ch.plan, ch.seg_clip, D, ch.gate); }
      }
      if (tcm)
        run_local_conv_tc(c, st, tm_u, ch.u16 ? W.tm_wlh[li] : W.tm_wl[li], tm_cat, M, kDil[b], D.cin, ch.gate, ch.td_row_seg, pdl, bf ? cath[b] : nullptr, ld, ch.u16 != nullptr);
      else
        sgemm(L, st, "dense_local_gemm", M, kGrowth, 3 * kBnC, LocalConvA{ch.u, kDil[b], M}, D.wl, 3 * kBnC, GateEpi{cat, ld, D.cin, ch.gate, ch.td_row_seg});
    }
    const TransitW& T = W.transit[b];
    float* out = b == 0 ? ch.cat2 : (b == 1 ? ch.cat3 : ch.tr3);
    const int ldo = b == 2 ? kStatsC : 1024;
    if (bf && c->cat_bf16 == 2)
      tc::tgemm_bnrelu<128, 2, tc::EpiMask, 2>(L, st, "transit_gemm", cath[b], ld, T.a, T.b, W.tm_trh[b], out, ldo, M, T.cout, T.cin,
                               tc::EpiMask{nullptr, ldo, ch.td_row_clip, M, b < 2 ? cath[b + 1] : nullptr, ldo}, pdl);
    else if (bf)
      tc::tgemm_bnrelu<128, 2, tc::EpiMask, 1>(L, st, "transit_gemm", cath[b], ld, T.a, T.b, W.tm_tr[b], out, ldo, M, T.cout, T.cin,
                               tc::EpiMask{nullptr, ldo, ch.td_row_clip, M, b < 2 ? cath[b + 1] : nullptr, ldo}, pdl);
    else if (tcm && c->transit_n256 && T.cout % 256 == 0)
      // N = 256 output tiles: the X tile (BN + ReLU producers, the stage stores) is made once per 256 output channels instead of once
      // per 128, and the A operand is read from shared memory once per 256: 48 KB instead of 64 KB through the array per 128 x 128 x 32
      tc::tgemm_bnrelu<256, 2>(L, st, "transit_gemm", cat, ld, T.a, T.b, W.tm_tr256[b], out, ldo, M, T.cout, T.cin,
                               tc::EpiMask{nullptr, ldo, ch.td_row_clip, M}, pdl);
    else if (tcm)
      tc::tgemm_bnrelu<128, 2>(L, st, "transit_gemm", cat, ld, T.a, T.b, W.tm_tr[b], out, ldo, M, T.cout, T.cin,
                               tc::EpiMask{nullptr, ldo, ch.td_row_clip, M}, pdl);
    else
      sgemm(L, st, "transit_gemm", M, T.cout, T.cin, BnReluA{cat, ld, T.a, T.b}, T.w, T.cin, MaskEpi{out, ldo, ch.td_row_clip});
  }
  { Scope sc(L, st, "stats_pool_kernel", 0.0, 4.0 * kStatsC * ((double)ch.td_rows + 2.0 * ch.n_clips)); stats_pool_kernel<<<dim3(ch.n_clips, kStatsC / kSpSlab), 32 * kSpWarps, 0, st>>>(ch.tr3, ch.plan, W.out_a, W.out_b, ch.stats); }
  { Scope sc(L, st, "xv_final_kernel", 2.0 * ch.n_clips * 2 * kStatsC * 192, 4.0 * ((double)ch.n_clips * (2 * kStatsC + 192) + 2.0 * kStatsC * 192)); xv_final_kernel<<<ch.n_clips, 256, 0, st>>>(ch.stats, ch.plan, W.fin_w, W.fin_b, xv_out, status); }
}

}  // namespace cbx

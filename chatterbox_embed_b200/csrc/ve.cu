// VoiceEncoder path: trim -> 40-bin power mel -> overlapping 160-frame partials -> 3x256 LSTM ->
// proj + ReLU + L2 -> per-clip mean + L2.   Reference: voice_encoder.py:139-199, 246-274; melspec.py:26-64.
#include <math.h>

#include "cbx_internal.h"
#include "sgemm.cuh"
#include "tc.cuh"
#include "epi.cuh"

namespace cbx {

// =================================================================================================
// K1  librosa.effects.trim(top_db)  (voice_encoder.py:267) + device-side window plan.
// One CTA per clip.  Block energies over 512-sample hops; an RMS frame (2048, centred, zero padded)
// is the sum of four consecutive hop blocks.
// =================================================================================================
__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, sh[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ int block_reduce_min_i(int v, int* sh) {
  for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = min(r, sh[i]);
  __syncthreads();
  return r;
}

// 1024 threads per clip: a streaming kernel needs the loads of many warps in flight (with 256 threads it ran at 1.8 TB/s); the
// summation order inside a 512-sample block is unchanged (the trim indices are bit-exact against the reference's algorithm)
constexpr int kTrimThreads = 1024;
__global__ void __launch_bounds__(kTrimThreads) trim_plan_kernel(const float* __restrict__ pcm, const ClipPlan* __restrict__ plan,
                                                        ClipDyn* __restrict__ dyn, float* __restrict__ scratch,
                                                        float top_db, int no_trim, int step, double min_cov) {
  __shared__ float shf[kTrimThreads / 32];
  __shared__ int shi[kTrimThreads / 32];
  const ClipPlan cp = plan[blockIdx.x];
  const int n = cp.n_samples;
  const float* y = pcm + cp.pcm_off;
  int s = 0, e = n;
  if (!no_trim && n > 0) {
    float* E = scratch + cp.trim_blk0;
    const int nblk = (n + kTrimHop - 1) / kTrimHop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < nblk; b += kTrimThreads / 32) {
      const int i0 = b * kTrimHop, i1 = min(n, i0 + kTrimHop);
      float a = 0.f;
      for (int i = i0 + lane; i < i1; i += 32) { float v = __ldg(y + i); a = fmaf(v, v, a); }
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) E[b] = a;
    }
    __syncthreads();
    const int nfr = 1 + n / kTrimHop;
    auto frame_power = [&](int j) {
      float p = 0.f;
      for (int b = j - 2; b <= j + 1; ++b) if (b >= 0 && b < nblk) p += E[b];
      return p * (1.0f / kTrimFrame);
    };
    float pmax = 0.f;
    for (int j = threadIdx.x; j < nfr; j += blockDim.x) pmax = fmaxf(pmax, frame_power(j));
    pmax = block_reduce_max(pmax, shf);
    const float ref_db = 10.f * log10f(fmaxf(1e-10f, pmax));
    int first = 0x7fffffff, last = -1;
    for (int j = threadIdx.x; j < nfr; j += blockDim.x) {
      const float db = 10.f * log10f(fmaxf(1e-10f, frame_power(j))) - ref_db;
      if (db > -top_db) { first = min(first, j); last = max(last, j); }
    }
    first = block_reduce_min_i(first, shi);
    last = -block_reduce_min_i(-last, shi);
    if (last < 0) { s = 0; e = 0; }
    else { s = first * kTrimHop; e = min(n, (last + 1) * kTrimHop); }
  }
  if (threadIdx.x == 0) {
    ClipDyn d;
    d.trim_s = s; d.trim_e = e;
    const int nt = e - s;
    d.status = 0;
    if (nt < kVeNfft / 2 + 1) {        // reflect pad of 200 needs at least 201 samples
      d.status |= CBX_CLIP_VE_TOO_SHORT;
      d.ve_frames = 0; d.ve_frames_eff = 0; d.ve_parts = 0;
    } else {
      const int T = 1 + nt / kVeHop;
      int x = T - kVePartial + step; if (x < 0) x = 0;
      int wins = x / step, rem = x % step;
      if (wins == 0 || (double)(rem + (kVePartial - step)) / (double)kVePartial >= min_cov) wins += 1;
      const int target = kVePartial + step * (wins - 1);
      d.ve_frames = T; d.ve_frames_eff = min(T, target); d.ve_parts = wins;
    }
    if (n < kKWin) d.status |= CBX_CLIP_XV_TOO_SHORT;
    dyn[blockIdx.x] = d;
  }
}

// Row -> clip map of the mel buffer and the partial-slot tables (one CTA per clip).
__global__ void ve_maps_kernel(const ClipPlan* __restrict__ plan, const ClipDyn* __restrict__ dyn, int step,
                               int32_t* __restrict__ mel_row_clip, int32_t* __restrict__ slot_clip,
                               int32_t* __restrict__ slot_row) {
  const int c = blockIdx.x;
  const ClipPlan cp = plan[c];
  const ClipDyn d = dyn[c];
  for (int r = threadIdx.x; r < cp.mel_rows; r += blockDim.x) mel_row_clip[cp.mel_row + r] = c;
  for (int p = threadIdx.x; p < cp.slots; p += blockDim.x) {
    slot_clip[cp.slot0 + p] = p < d.ve_parts ? c : -1;
    slot_row[cp.slot0 + p] = cp.mel_row + step * p;
  }
}

// =================================================================================================
// K2/K3  framing + reflect pad + (Hann folded) 400-point DFT as a GEMM, then |X|^2 and the 40x201 mel bank.
// =================================================================================================
struct VeFrameGather {
  const float* pcm; const ClipPlan* plan; const ClipDyn* dyn; const int32_t* row_clip;
  __device__ float operator()(int m, int k) const {
    const int c = row_clip[m];
    if (c < 0) return 0.f;
    const int t = m - plan[c].mel_row;
    const ClipDyn d = dyn[c];
    if (t >= d.ve_frames_eff) return 0.f;
    const int n = d.trim_e - d.trim_s;
    int i = t * kVeHop + k - kVeNfft / 2;
    if (i < 0) i = -i; else if (i >= n) i = 2 * (n - 1) - i;     // np.pad(mode="reflect"), single fold (n >= 201)
    return __ldg(pcm + plan[c].pcm_off + d.trim_s + i);
  }
};
struct StoreRowMajor {
  float* out; int ld;
  __device__ void operator()(int m, int n, float v) const { out[(size_t)m * ld + n] = v; }
};

// one warp per mel row: power spectrum in registers/smem, dense 40x201 filterbank
__global__ void __launch_bounds__(256) ve_mel_kernel(const float* __restrict__ spec, const float* __restrict__ bank,
                                                     const ClipPlan* __restrict__ plan, const ClipDyn* __restrict__ dyn,
                                                     const int32_t* __restrict__ row_clip, float* __restrict__ mel, int rows) {
  __shared__ float pw[8][kVeBins + 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const int c = row_clip[r];
  const bool live = c >= 0 && (r - plan[c].mel_row) < dyn[c].ve_frames_eff;
  if (!live) {
    for (int m = lane; m < kVeMels; m += 32) mel[(size_t)r * kVeMels + m] = 0.f;   // zero rows past the clip (voice_encoder.py:176-179)
    return;
  }
  const float2* sp = reinterpret_cast<const float2*>(spec + (size_t)r * kVeSpecN);
  for (int k = lane; k < kVeBins; k += 32) { float2 v = sp[k]; pw[warp][k] = v.x * v.x + v.y * v.y; }
  __syncwarp();
  for (int m = lane; m < kVeMels; m += 32) {
    const float* b = bank + (size_t)m * kVeBins;
    float a = 0.f;
    for (int k = 0; k < kVeBins; ++k) a = fmaf(__ldg(b + k), pw[warp][k], a);
    mel[(size_t)r * kVeMels + m] = a;
  }
}

// =================================================================================================
// K5  LSTM.  Input projections are dense GEMMs over all rows; the recurrence is a persistent kernel per
// tile of MT partial slots (thread j owns hidden unit j and its four gates).
// =================================================================================================
struct PlainA {
  const float* a; int ld;
  __device__ float operator()(int m, int k) const { return __ldg(a + (size_t)m * ld + k); }
};
struct StoreBias {
  float* out; int ld; const float* bias;
  __device__ void operator()(int m, int n, float v) const { out[(size_t)m * ld + n] = v + __ldg(bias + n); }
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int LSTM_MT = 16;

// xw_row(q,t) = xw + (xw_base[q] + t) * 1024   (layer 0: base = first mel row of the partial; else q*160)
template <bool kLayer0>
__global__ void __launch_bounds__(256) lstm_rec_kernel(const float* __restrict__ xw, const int32_t* __restrict__ slot_row,
                                                       const float* __restrict__ whhT, float* __restrict__ hseq, int n_slots) {
  __shared__ float h[LSTM_MT][kVeHidden];
  const int j = threadIdx.x;
  const int q0 = blockIdx.x * LSTM_MT;
  float cst[LSTM_MT];
  size_t base[LSTM_MT];
#pragma unroll
  for (int m = 0; m < LSTM_MT; ++m) {
    const int q = min(q0 + m, n_slots - 1);
    base[m] = kLayer0 ? (size_t)slot_row[q] : (size_t)q * kVePartial;
    cst[m] = 0.f;
    h[m][j] = 0.f;
  }
  __syncthreads();
  for (int t = 0; t < kVePartial; ++t) {
    float acc[4][LSTM_MT];
#pragma unroll
    for (int m = 0; m < LSTM_MT; ++m) {
      const float* x = xw + (base[m] + t) * kVeGates + j;
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g][m] = __ldg(x + g * kVeHidden);
    }
#pragma unroll 4
    for (int k = 0; k < kVeHidden; ++k) {
      const float* w = whhT + (size_t)k * kVeGates + j;
      const float w0 = __ldg(w), w1 = __ldg(w + 256), w2 = __ldg(w + 512), w3 = __ldg(w + 768);
#pragma unroll
      for (int m = 0; m < LSTM_MT; ++m) {
        const float hv = h[m][k];
        acc[0][m] = fmaf(hv, w0, acc[0][m]);
        acc[1][m] = fmaf(hv, w1, acc[1][m]);
        acc[2][m] = fmaf(hv, w2, acc[2][m]);
        acc[3][m] = fmaf(hv, w3, acc[3][m]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < LSTM_MT; ++m) {
      const float ig = sigmoidf_(acc[0][m]), fg = sigmoidf_(acc[1][m]), gg = tanhf(acc[2][m]), og = sigmoidf_(acc[3][m]);
      cst[m] = fg * cst[m] + ig * gg;
      const float hv = og * tanhf(cst[m]);
      h[m][j] = hv;
      if (q0 + m < n_slots) hseq[((size_t)(q0 + m) * kVePartial + t) * kVeHidden + j] = hv;
    }
    __syncthreads();
  }
}

// K6  proj + ReLU + L2 on the final hidden state of layer 3 (one CTA per slot)
__global__ void __launch_bounds__(256) ve_proj_kernel(const float* __restrict__ hfin, size_t stride, const float* __restrict__ wpT,
                                                      const float* __restrict__ bp, float* __restrict__ pemb) {
  __shared__ float h[kVeHidden];
  __shared__ float red[8];
  const int q = blockIdx.x, j = threadIdx.x;
  h[j] = hfin[(size_t)q * stride + j];
  __syncthreads();
  float a = __ldg(bp + j);
  for (int k = 0; k < kVeHidden; ++k) a = fmaf(h[k], __ldg(wpT + (size_t)k * kVeEmbed + j), a);
  a = fmaxf(a, 0.f);
  float s = a * a;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((j & 31) == 0) red[j >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  pemb[(size_t)q * kVeEmbed + j] = a / sqrtf(tot);       // 0/0 -> NaN exactly like voice_encoder.py:160
}

// K7  per-clip mean over its partial embeddings + L2 (voice_encoder.py:194-197)
__global__ void __launch_bounds__(256) ve_clip_mean_kernel(const float* __restrict__ pemb, const ClipPlan* __restrict__ plan,
                                                           ClipDyn* __restrict__ dyn, float* __restrict__ ve_out,
                                                           int32_t* __restrict__ status) {
  __shared__ float red[8];
  const int c = blockIdx.x, j = threadIdx.x;
  const ClipPlan cp = plan[c];
  const ClipDyn d = dyn[c];
  float a = 0.f;
  for (int p = 0; p < d.ve_parts; ++p) a += pemb[(size_t)(cp.slot0 + p) * kVeEmbed + j];
  a = d.ve_parts > 0 ? a / (float)d.ve_parts : nanf("");
  float s = a * a;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((j & 31) == 0) red[j >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float v = a / sqrtf(tot);
  ve_out[(size_t)cp.out_index * kVeEmbed + j] = v;
  if (j == 0 && status) {
    int stt = d.status;
    if (!(tot > 0.f) || isnan(tot)) stt |= CBX_CLIP_VE_NAN;
    atomicOr(status + cp.out_index, stt);
  }
}

// -------------------------------------------------------------------------------------------------
void run_ve_lstm(cbx_ctx* c, const VeChunk& ch, cudaStream_t st) {
  const VeWeights& W = c->ve;
  Launches& L = c->launches;
  const int rows = ch.slots * kVePartial;
  if (c->mode == 1) {
    // tensor-core path: input projections as dense tcgen05 GEMMs (gate columns in the recurrence's permuted order), the
    // recurrence as the persistent cluster kernel of lstm_tc.cu.  Layer 0 projects once per mel frame (partials overlap:
    // hop 77 < 160) and the recurrence gathers rows through slot_row.
    CUtensorMap tmA = tc::make_map_2d(ch.mel, ch.mel_rows, kVeMels, kVeMels, tc::BM, true);
    const bool x16 = c->xw_bf16 != 0;        // bf16 mode: the projections are stored as bf16 (the GEMM is bound by its C writes)
    if (x16) tc::pgemm_bias_tma<256, 4, true>(L, st, "lstm_xw0_gemm", tmA, W.tm_wih_p256[0], ch.xw0, kVeGates, W.bias_p[0], ch.mel_rows, kVeGates, kVeMels);
    else tc::pgemm_bias_tma<256, 4>(L, st, "lstm_xw0_gemm", tmA, W.tm_wih_p256[0], ch.xw0, kVeGates, W.bias_p[0], ch.mel_rows, kVeGates, kVeMels);
#ifdef CBX_DEV_TOOLS
    if (c->lstm_impl == 2)
#endif
    {
      // L2-exchange recurrence: hseq / xw of layers 1, 2 in the tiled time-major row order over whole 224-partial tiles
      const int prow = lstm_padded_slots(ch.slots) * kVePartial;
      const size_t hl = (size_t)ch.slots * kVeHidden;
      // two-stream overlap: hold the recurrence back until the CAMPPlus chain is in its D-TDNN phase (see cbx_embed)
      if (c->ve_wait_dtdnn) cudaStreamWaitEvent(st, c->ev_dtdnn, 0);
      run_lstm_rec_tc2(c, ch.xw0, x16, ch.slot_row, W.whh_p[0], ch.hseq, ch.hlast, ch.slots, st);
      for (int l = 1; l < 3; ++l) {
        CUtensorMap tmH = tc::make_map_2d(ch.hseq, prow, kVeHidden, kVeHidden, tc::BM, true);
        if (x16) tc::pgemm_bias_tma<256, 4, true>(L, st, "lstm_xw_gemm", tmH, W.tm_wih_p256[l], ch.xw, kVeGates, W.bias_p[l], prow, kVeGates, kVeHidden);
        else tc::pgemm_bias_tma<256, 4>(L, st, "lstm_xw_gemm", tmH, W.tm_wih_p256[l], ch.xw, kVeGates, W.bias_p[l], prow, kVeGates, kVeHidden);
        run_lstm_rec_tc2(c, ch.xw, x16, nullptr, W.whh_p[l], ch.hseq, ch.hlast + l * hl, ch.slots, st);
      }
      { Scope sc(L, st, "ve_proj_kernel", 2.0 * ch.slots * kVeHidden * kVeEmbed, 4.0 * ((double)ch.slots * (kVeHidden + kVeEmbed) + (double)kVeHidden * kVeEmbed)); ve_proj_kernel<<<ch.slots, 256, 0, st>>>(ch.hlast + 2 * hl, (size_t)kVeHidden, W.wpT, W.bp, ch.pemb); }
      return;
    }
#ifdef CBX_DEV_TOOLS   // v1 recurrence (DSMEM pushes), tools/ comparisons only
    run_lstm_rec_tc(c, ch.xw0, ch.slot_row, W.whh_p[0], ch.hseq, nullptr, ch.slots, st);
    for (int l = 1; l < 3; ++l) {
      CUtensorMap tmH = tc::make_map_2d(ch.hseq, rows, kVeHidden, kVeHidden, tc::BM, true);
      tc::tgemm<128, 3>(L, st, "lstm_xw_gemm", tmH, W.tm_wih_p[l], rows, kVeGates, kVeHidden, tc::plain_map(kVeHidden), 1, tc::NoPrologue{},
                        tc::EpiBias{ch.xw, kVeGates, W.bias_p[l], rows});
      run_lstm_rec_tc(c, ch.xw, nullptr, W.whh_p[l], l == 2 ? nullptr : ch.hseq, l == 2 ? ch.hlast + 2 * (size_t)ch.slots * kVeHidden : nullptr, ch.slots, st);
    }
    { Scope sc(L, st, "ve_proj_kernel", 2.0 * ch.slots * kVeHidden * kVeEmbed, 4.0 * ((double)ch.slots * (kVeHidden + kVeEmbed) + (double)kVeHidden * kVeEmbed)); ve_proj_kernel<<<ch.slots, 256, 0, st>>>(ch.hlast + 2 * (size_t)ch.slots * kVeHidden, (size_t)kVeHidden, W.wpT, W.bp, ch.pemb); }
    return;
#endif
  }
  // strict-fp32 path
  const int nb = (ch.slots + LSTM_MT - 1) / LSTM_MT;
  sgemm(L, st, "lstm_xw0_gemm", ch.mel_rows, kVeGates, kVeMels, PlainA{ch.mel, kVeMels}, W.wih0, kVeMels, StoreBias{ch.xw0, kVeGates, W.bias[0]});
  { Scope sc(L, st, "lstm_rec_kernel", 2.0 * ch.slots * kVePartial * kVeHidden * kVeGates); lstm_rec_kernel<true><<<nb, 256, 0, st>>>(ch.xw0, ch.slot_row, W.whhT[0], ch.hseq, ch.slots); }
  // stage tap: last hidden state of every layer -> hlast[l] (hseq is reused by the next layer)
  auto tap_last = [&](int l) {
    cudaMemcpy2DAsync(ch.hlast + (size_t)l * ch.slots * kVeHidden, sizeof(float) * kVeHidden, ch.hseq + (size_t)(kVePartial - 1) * kVeHidden,
                      sizeof(float) * kVePartial * kVeHidden, sizeof(float) * kVeHidden, ch.slots, cudaMemcpyDeviceToDevice, st);
  };
  tap_last(0);
  for (int l = 1; l < 3; ++l) {
    sgemm(L, st, "lstm_xw_gemm", rows, kVeGates, kVeHidden, PlainA{ch.hseq, kVeHidden}, W.wih[l], kVeHidden, StoreBias{ch.xw, kVeGates, W.bias[l]});
    { Scope sc(L, st, "lstm_rec_kernel", 2.0 * ch.slots * kVePartial * kVeHidden * kVeGates); lstm_rec_kernel<false><<<nb, 256, 0, st>>>(ch.xw, ch.slot_row, W.whhT[l], ch.hseq, ch.slots); }
    tap_last(l);
  }
  { Scope sc(L, st, "ve_proj_kernel", 2.0 * ch.slots * kVeHidden * kVeEmbed, 4.0 * ((double)ch.slots * (kVeHidden + kVeEmbed) + (double)kVeHidden * kVeEmbed)); ve_proj_kernel<<<ch.slots, 256, 0, st>>>(ch.hseq + (size_t)(kVePartial - 1) * kVeHidden, (size_t)kVePartial * kVeHidden, W.wpT, W.bp, ch.pemb); }
}

void run_ve_chunk(cbx_ctx* c, const float* pcm, const VeChunk& ch, float trim_top_db, bool no_trim, int step,
                  double min_cov, float* ve_out, int32_t* status, cudaStream_t st) {
  Launches& L = c->launches;
  { Scope sc(L, st, "trim_plan_kernel", 0.0, 4.0 * (double)ch.pcm_samples); trim_plan_kernel<<<ch.n_clips, kTrimThreads, 0, st>>>(pcm, ch.plan, ch.dyn, ch.trim_scratch, trim_top_db, no_trim ? 1 : 0, step, min_cov); }
  cudaMemsetAsync(ch.mel_row_clip, 0xff, sizeof(int32_t) * ch.mel_rows, st);
  { Scope sc(L, st, "ve_maps_kernel"); ve_maps_kernel<<<ch.n_clips, 256, 0, st>>>(ch.plan, ch.dyn, step, ch.mel_row_clip, ch.slot_clip, ch.slot_row); }
  if (c->mode == 1) {
    run_ve_mel_tc(c, pcm, ch, st);
  } else {
    sgemm(L, st, "ve_dft_gemm", ch.mel_rows, kVeSpecN, kVeNfft, VeFrameGather{pcm, ch.plan, ch.dyn, ch.mel_row_clip}, c->ft.ve_dft, kVeNfft,
          StoreRowMajor{ch.spec, kVeSpecN});
    { Scope sc(L, st, "ve_mel_kernel"); ve_mel_kernel<<<(ch.mel_rows + 7) / 8, 256, 0, st>>>(ch.spec, c->ft.ve_mel, ch.plan, ch.dyn, ch.mel_row_clip, ch.mel, ch.mel_rows); }
  }
  run_ve_lstm(c, ch, st);
  { Scope sc(L, st, "ve_clip_mean_kernel", 0.0, 4.0 * kVeEmbed * ((double)ch.slots + ch.n_clips)); ve_clip_mean_kernel<<<ch.n_clips, 256, 0, st>>>(ch.pemb, ch.plan, ch.dyn, ve_out, status); }
}

// VoiceEncoder.forward on pre-cut partials: every partial is its own 160-row "clip".
__global__ void ve_identity_slots_kernel(int32_t* slot_row, int n) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) slot_row[q] = q * kVePartial;
}

void run_ve_forward_partials(cbx_ctx* c, const float* mels, int n, float* out, void* ws, cudaStream_t st) {
  Carver cv(ws, INT64_MAX);
  VeChunk ch{};
  ch.n_clips = 0; ch.slots = n; ch.mel_rows = n * kVePartial;
  ch.mel = const_cast<float*>(mels);
  ch.slot_row = cv.take<int32_t>(n);
  ch.xw0 = cv.take<float>((int64_t)lstm_padded_slots(n) * kVePartial * kVeGates);
  ch.xw = ch.xw0;                       // layer-0 rows are exactly slot*160+t here, so one buffer serves both
  ch.hseq = cv.take<float>((int64_t)lstm_padded_slots(n) * kVePartial * kVeHidden);
  ch.hlast = cv.take<float>((int64_t)3 * n * kVeHidden);
  ch.pemb = out;
  { Scope sc(c->launches, st, "ve_identity_slots_kernel"); ve_identity_slots_kernel<<<(n + 255) / 256, 256, 0, st>>>(ch.slot_row, n); }
  run_ve_lstm(c, ch, st);
}

}  // namespace cbx

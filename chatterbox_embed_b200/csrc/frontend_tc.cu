// Audio front-ends on the sm_100a tensor cores: framing + window + DFT as ONE GEMM per 128-frame tile, then |X|^2, the
// mel filterbank and (Kaldi) the log, fused in the epilogue.  Nothing but the mel/fbank rows is written to HBM.
//
//   VoiceEncoder mel  (melspec.py:26-64):   reflect-padded frames [T x 400] . (Hann-folded DFT)^T [400 x 2*199]  -> 40 Slaney mels
//   Kaldi fbank       (xvector.py:50, torchaudio kaldi.py:514-645):
//                                           raw frames [T x 400] . (DC-removal/pre-emphasis/Povey-folded 512-pt DFT)^T [400 x 2*255]
//                                           -> 80 HTK mels -> log(max(., eps))
//   S3Tokenizer log-mel (s3tokenizer.py:128-168): the VoiceEncoder's frames and DFT (torch.stft 400/160, centred, reflect),
//                                           last frame dropped -> 128 Slaney mels -> log10(max(., 1e-10)), per-clip max - 8 floor
//                                           and (x + 4) / 4 in a second, transposing pass (the consumer wants [128][T])
//
// Precision: plain TF32 is not acceptable here (the log exposes near-empty bins: errors of 4.5 in the log domain on chirps,
// SURVEY.md 8d hazard 3), so both operands are split hi + lo and the product is 3 MMAs: A_hi.B_hi + A_lo.B_hi + A_hi.B_lo.
//
// Warp roles (192 threads): warp 0 = TMA producer of the DFT-matrix tiles (B operand, hi and lo maps), warp 1 = MMA issuer,
// warps 2..5 = frame producers (gather PCM -> split hi/lo -> 128B-swizzled K-major A stages in shared memory) and, after the
// last K block, the epilogue (TMEM -> power -> 2-sparse mel accumulation in shared memory -> log -> coalesced store).
// The accumulator is the whole TMEM: [128 frames x up to 512 columns] fp32 (re, im interleaved).
#include <math.h>

#include <type_traits>

#include "cbx_internal.h"
#include "tc.cuh"

namespace cbx {
namespace fe {

using namespace tc;

constexpr int KTOT = 400;                       // both front-ends: 400-sample frames
constexpr int NKB = (KTOT + BK - 1) / BK;       // 13 K blocks of 32 (the last holds 16)
constexpr int SA = 2;                           // A stages: hi + lo, 16 KB each
constexpr int SB = 4;                           // B slots of 32 KB: [256 x 32] of either the hi or the lo matrix
constexpr int A_BYTES = BM * BK * 4;            // 16 KB
constexpr int B_BYTES = 256 * BK * 4;           // 32 KB
constexpr int PGRP = 2;                         // producer groups (4 warps each) on alternate K blocks
constexpr int THREADS = 32 * (2 + 4 * PGRP + 4);   // warp 0 TMA, warp 1 MMA, 8 producer warps, 4 epilogue warps
constexpr int MAX_BINS = 256;
constexpr int SMEM_BYTES = SA * 2 * A_BYTES + SB * B_BYTES + 1024 + 2 * 128 * 16 + MAX_BINS * 16 + 256;

struct RowDesc { long long base; int start; int n; };   // element k of the frame = pcm[base + reflect(start + k, n)]; n == 0: zero row
constexpr int kNoReflect = 0x7fffffff;                  // n of a frame that lies inside its clip: no index ever reflects (fast path of the producers)
// a frame whose 401 samples [start, start + 400] lie inside [0, n) never reflects: mark it so that the producers take the fast path
__device__ __forceinline__ RowDesc mark_interior(RowDesc d) {
  if (d.n > 0 && d.n != kNoReflect && d.start >= 0 && d.start + KTOT < d.n) d.n = kNoReflect;
  return d;
}

// Kaldi frames: snip_edges, no padding (kaldi.py:63-67): frame t of clip c = pcm[off_c + 160 t ...]
struct KaldiRows {
  const ClipPlan* plan; const int32_t* row_clip;
  __device__ RowDesc operator()(int row) const {
    const int c = row_clip[row];
    if (c < 0) return RowDesc{0, 0, 0};
    const ClipPlan cp = plan[c];
    return RowDesc{cp.pcm_off, (row - cp.fb_row) * kKHop, kNoReflect};
  }
  __device__ bool live(int row) const { return row_clip[row] >= 0; }
};
// VoiceEncoder frames: centred, reflect-padded by 200 on the trimmed clip (melspec.py:57-64, voice_encoder.py:267)
struct VeRows {
  const ClipPlan* plan; const ClipDyn* dyn; const int32_t* row_clip;
  __device__ RowDesc operator()(int row) const {
    const int c = row_clip[row];
    if (c < 0) return RowDesc{0, 0, 0};
    const ClipPlan cp = plan[c];
    const ClipDyn d = dyn[c];
    const int t = row - cp.mel_row;
    if (t >= d.ve_frames_eff) return RowDesc{0, 0, 0};
    return RowDesc{cp.pcm_off + d.trim_s, t * kVeHop - kVeNfft / 2, d.trim_e - d.trim_s};
  }
  __device__ bool live(int) const { return true; }      // rows past the clip are zero frames -> zero mels (voice_encoder.py:176-179)
};
// S3Tokenizer frames: ragged clips back to back, centred reflect pad 200, frames 0 .. L/160 - 1 (s3tokenizer.py:158-163)
struct S3Clip { long long pcm_off; int n; int row0; int frames; int pad; };
struct S3Rows {
  const S3Clip* clips; int n_clips; float* cmax;
  __device__ int clip_of(int row) const {
    int lo = 0, hi = n_clips - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (clips[mid].row0 <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
  }
  __device__ RowDesc operator()(int row) const {
    const S3Clip c = clips[clip_of(row)];
    const int t = row - c.row0;
    if (t >= c.frames) return RowDesc{0, 0, 0};
    return RowDesc{c.pcm_off, t * kVeHop - kVeNfft / 2, c.n};
  }
  __device__ bool live(int) const { return true; }
  // float max through integer atomics (init -inf): positive values by signed max, negative ones by unsigned min
  __device__ void note_max(int row, float v) const {
    float* a = cmax + clip_of(row);
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
  }
};

// LOG: 0 = power mel, 1 = ln(max(., eps)) (Kaldi), 2 = log10(max(., 1e-10)) + per-clip maximum (S3Tokenizer)
//
// Persistent: one CTA per SM loops over 128-frame tiles.  Round 1 launched one CTA per tile with the frame producers doubling as
// the epilogue: with the whole TMEM as one accumulator only one CTA fits an SM, so the tensor pipe idled through every tile's
// set-up, pipeline fill and epilogue (tensor pipe 40 % active, ~45 us per tile of which 16 us of MMAs).  Now
//   * warps 2..9 are two producer groups on alternate K blocks (the gather's L2 latency of one block hides behind the other's
//     stores) and run straight on into the next tile while
//   * warps 10..13 are the epilogue: TMEM -> power -> mel with TWO running accumulators in registers (bin table of weights.cu:
//     no shared-memory read-modify-write chain) -> transform -> global store; they release the accumulator as soon as they have
//     read it, so the MMAs of the next tile start while the last mel values are still being written;
//   * the DFT-matrix tiles (TMA) and the first A stages of the next tile are already in flight when the accumulator is released.
template <class Rows, int NB1, int NMEL, int LOG>
__global__ void __launch_bounds__(THREADS, 1)
dftmel_kernel(const __grid_constant__ CUtensorMap tmHi0, const __grid_constant__ CUtensorMap tmLo0,
              const __grid_constant__ CUtensorMap tmHi1, const __grid_constant__ CUtensorMap tmLo1,
              const float* __restrict__ pcm, Rows rows_fn, const float4* __restrict__ bintab, float* __restrict__ out, int rows, int ntiles) {
  constexpr int NB0 = 256;
  constexpr int NTOT = NB0 + NB1;
  static_assert(NTOT / 2 <= MAX_BINS, "bin table");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [SA][hi | lo][128 x 128 B]
  uint8_t* sB = smem + SA * 2 * A_BYTES;                // [SB][256 x 128 B]
  RowDesc* rdesc = reinterpret_cast<RowDesc*>(sB + SB * B_BYTES);     // [2][128]: the tile's row descriptors, double-buffered
  float4* sbins = reinterpret_cast<float4*>(rdesc + 2 * BM);          // [NTOT / 2] bin table
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbins + MAX_BINS);
  uint64_t* a_full = bars;                 // [SA] 128 producer arrivals
  uint64_t* a_empty = bars + SA;           // [SA] MMA commit
  uint64_t* b_full = bars + 2 * SA;        // [SB] TMA bytes
  uint64_t* b_empty = b_full + SB;         // [SB] MMA commit
  uint64_t* accum = b_empty + SB;          // accumulator ready (MMA commit)
  uint64_t* drained = accum + 1;           // accumulator read out (4 epilogue warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmHi0); tma_prefetch_desc(&tmLo0); tma_prefetch_desc(&tmHi1); tma_prefetch_desc(&tmLo1);
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(accum, 1);
    mbar_init(drained, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < NTOT / 2; i += THREADS) sbins[i] = __ldg(bintab + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: per K block the slots are  hi(cols 0..255), lo(0..255), hi(256..), lo(256..); the ring runs on across tiles
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int kb = 0; kb < NKB; ++kb)
          for (int q = 0; q < 4; ++q, ++it) {
            const int s = it % SB, ph = (it / SB) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            const int nb = q >> 1;
            mbar_expect_tx(&b_full[s], (nb ? NB1 : NB0) * BK * 4);
            const CUtensorMap* tm = nb ? ((q & 1) ? &tmLo1 : &tmHi1) : ((q & 1) ? &tmLo0 : &tmHi0);
            tma_load_2d(sB + s * B_BYTES, tm, &b_full[s], kb * BK, nb * NB0);
          }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues
    constexpr uint32_t idesc0 = make_idesc_tf32(BM, NB0), idesc1 = make_idesc_tf32(BM, NB1);
    int it = 0, ia = 0, ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      mbar_wait(drained, (ti & 1) ^ 1);               // the epilogue has read the previous tile's accumulator
      tc_fence_after();
      for (int kb = 0; kb < NKB; ++kb, ++ia) {
        const int sa = ia % SA, pa = (ia / SA) & 1;
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint64_t ahi = make_desc_sw128(smem_u32(sA + sa * 2 * A_BYTES));
        const uint64_t alo = make_desc_sw128(smem_u32(sA + sa * 2 * A_BYTES + A_BYTES));
        const int ksteps = (kb == NKB - 1) ? (KTOT - (NKB - 1) * BK) / UMMA_K : BK / UMMA_K;
        for (int q = 0; q < 4; ++q, ++it) {
          const int s = it % SB, ph = (it / SB) & 1;
          mbar_wait(&b_full[s], ph);
          tc_fence_after();
          const uint64_t bd = make_desc_sw128(smem_u32(sB + s * B_BYTES));
          const int nb = q >> 1;
          const uint32_t d = tmem_base + nb * NB0;
          const uint32_t idesc = nb ? idesc1 : idesc0;
          if (elect_one()) {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ko = (uint64_t)(k * UMMA_K * 4 >> 4);
              if ((q & 1) == 0) {       // B_hi: A_hi.B_hi + A_lo.B_hi
                umma_tf32(d, ahi + ko, bd + ko, idesc, (kb | k) != 0);
                umma_tf32(d, alo + ko, bd + ko, idesc, 1);
              } else {                  // B_lo: A_hi.B_lo
                umma_tf32(d, ahi + ko, bd + ko, idesc, 1);
              }
            }
            umma_commit(&b_empty[s]);
            if (q == 3) {
              umma_commit(&a_empty[sa]);
              if (kb == NKB - 1) umma_commit(accum);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 2 + 4 * PGRP) {
    // ===== frame producers: group g takes the K blocks kb = g, g + 2, ...; warp w of a group fills rows [32 w, 32 w + 32) of the A
    // stage, lanes along K (coalesced PCM reads).  The row descriptors of a tile are computed by group 0's threads one tile ahead
    // (double-buffered), published to the other group through a named barrier.
    const int g = (warp - 2) >> 2, wq = (warp - 2) & 3;
    const int r_own = (threadIdx.x - 64) & 127;        // row whose descriptor this thread computes (both groups: 2 x 128 threads, same values)
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int m0 = tile * BM;
      RowDesc* rd = rdesc + (ti & 1) * BM;
      if (g == 0) rd[r_own] = (m0 + r_own < rows) ? mark_interior(rows_fn(m0 + r_own)) : RowDesc{0, 0, 0};
      asm volatile("bar.sync 1, 256;" ::: "memory");   // all 8 producer warps: the descriptors of this tile are in place (and the
                                                       // previous tile's reads of the other buffer are over before it is rewritten next time)
      // whole 32-row group interior (9 of 10 groups): a loop without reflection arithmetic or per-row branches
      const bool all_fast = __all_sync(0xffffffffu, rd[wq * 32 + lane].n == kNoReflect);
      auto run = [&](auto fastc) {
        constexpr bool FAST = decltype(fastc)::value;
        for (int kb = g; kb < NKB; kb += PGRP) {
          const int ia = ti * NKB + kb;
          const int sa = ia % SA, pa = (ia / SA) & 1;
          const int k = kb * BK + lane;
          float v[32];
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const RowDesc d = rd[wq * 32 + rr];
            if constexpr (FAST) {
              v[rr] = k < KTOT ? __ldg(pcm + d.base + d.start + k) : 0.f;
            } else {
              int i = d.start + k;
              if (i < 0) i = -i; else if (i >= d.n) i = 2 * (d.n - 1) - i;
              v[rr] = (d.n > 0 && k < KTOT) ? __ldg(pcm + d.base + i) : 0.f;
            }
          }
          mbar_wait(&a_empty[sa], pa ^ 1);
          uint8_t* hi = sA + sa * 2 * A_BYTES;
          uint8_t* lo = hi + A_BYTES;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const int r = wq * 32 + rr;
            const float vh = to_tf32(v[rr]);
            const float vl = to_tf32(v[rr] - vh);
            const uint32_t o = r * 128 + ((((uint32_t)lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4;
            *reinterpret_cast<float*>(hi + o) = vh;
            *reinterpret_cast<float*>(lo + o) = vl;
          }
          fence_proxy_async();
          mbar_arrive(&a_full[sa]);
        }
      };
      if (all_fast) run(std::true_type{}); else run(std::false_type{});
    }
  } else {
    // ===== epilogue: thread = frame row; D columns 2b, 2b+1 = re, im of bin b.  Two running accumulators: filter `cur` (A) and
    // `cur + 1` (B); the bin table says how many finished filters to retire before a bin and the bin's two weights.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int gr = tile * BM + row;
      const bool in_range = gr < rows;
      const bool live = in_range && rows_fn.live(gr);
      float* orow = out + (size_t)(in_range ? gr : 0) * NMEL;
      float accA = 0.f, accB = 0.f, mx = -INFINITY;
      int cur = 0;
      auto retire = [&]() {
        float val = accA;
        if (LOG == 1) val = logf(fmaxf(val, 1.1920928955078125e-07f));
        if (LOG == 2) { val = log10f(fmaxf(val, 1e-10f)); mx = fmaxf(mx, val); }
        if (in_range) orow[cur] = live ? val : 0.f;
        accA = accB; accB = 0.f; ++cur;
      };
      mbar_wait(accum, ti & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < NTOT / 16; ++c) {
        float v[16];
        {
          uint32_t* rv = reinterpret_cast<uint32_t*>(v);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(rv[0]), "=r"(rv[1]), "=r"(rv[2]), "=r"(rv[3]), "=r"(rv[4]), "=r"(rv[5]), "=r"(rv[6]), "=r"(rv[7]), "=r"(rv[8]),
                "=r"(rv[9]), "=r"(rv[10]), "=r"(rv[11]), "=r"(rv[12]), "=r"(rv[13]), "=r"(rv[14]), "=r"(rv[15])
              : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16))
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (c == NTOT / 16 - 1) {            // the last columns are in registers: the next tile's MMAs may overwrite the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(drained);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float pw = fmaf(v[2 * i], v[2 * i], v[2 * i + 1] * v[2 * i + 1]);
          const float4 tb = sbins[c * 8 + i];                      // warp-uniform: {wA, wB, retire count, -}
          for (int n = __float_as_int(tb.z); n > 0; --n) retire();
          accA = fmaf(tb.x, pw, accA);
          accB = fmaf(tb.y, pw, accB);
        }
      }
      while (cur < NMEL) retire();
      if constexpr (LOG == 2) {
        if (in_range) rows_fn.note_max(gr, mx);
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <class Rows, int NB1, int NMEL, int LOG>
void launch(cbx_ctx* c, cudaStream_t st, const char* tag, const CUtensorMap& hi0, const CUtensorMap& lo0, const CUtensorMap& hi1,
            const CUtensorMap& lo1, const float* pcm, Rows rf, const float* bintab, float* out, int rows) {
  if (rows <= 0) return;
  auto kern = dftmel_kernel<Rows, NB1, NMEL, LOG>;
  ensure_max_smem(kern, SMEM_BYTES);
  const int ntiles = (rows + BM - 1) / BM;
  const int grid = ntiles < sm_count() ? ntiles : sm_count();
  Scope sc(c->launches, st, tag, 2.0 * rows * (256 + NB1) * KTOT, 4.0 * rows * (160 + NMEL));   // ALGORITHMIC: one fp32 DFT per frame (the 3xTF32 split executes 3x this on the tensor pipe); one hop of PCM in, one feature row out
  kern<<<grid, THREADS, SMEM_BYTES, st>>>(hi0, lo0, hi1, lo1, pcm, rf, reinterpret_cast<const float4*>(bintab), out, rows, ntiles);
}

// ---------------------------------------------------------------------------------------------------------------------
// Even / odd form of the VoiceEncoder / S3Tokenizer DFT (tables: weights.cu, ve_eo_hi / ve_eo_lo).  The Hann window is
// symmetric, so re X[k] needs only e[j] = s[j] + s[400-j] and im X[k] only o[j] = s[j] - s[400-j] (j = 1..200): two K = 200
// GEMMs into two accumulator blocks (re: TMEM columns [0, 208), im: [208, 416)) instead of one K = 400 GEMM -- half the MMAs and
// half the DFT-matrix bytes streamed from L2 per tile, which is what bounds the K loop.  Same persistent structure as
// dftmel_kernel; the unit of the pipeline is a HALF block: hb = 2 kb + part, part 0 = (e, cos rows), part 1 = (o, sin rows).
// Producer group 0 makes the e stages, group 1 the o stages (14 half blocks per tile: the stage of a half block is its part).
namespace eo {
constexpr int NHB = 14;                         // 7 K blocks of 32 columns (200 valid) x {re, im}
constexpr int NROWS = 208;                      // 199 bins + padding to a legal UMMA N
constexpr int SBE = 5;                          // B slots
constexpr int BE_BYTES = 27 * 1024;             // 208 x 128 B, rounded up to the 1 KB swizzle atom
constexpr int SMEM_BYTES = SA * 2 * A_BYTES + SBE * BE_BYTES + 1024 + 2 * 128 * 16 + MAX_BINS * 16 + 256;
}  // namespace eo

template <class Rows, int NMEL, int LOG>
__global__ void __launch_bounds__(THREADS, 1)
dftmel_eo_kernel(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo,
                 const float* __restrict__ pcm, Rows rows_fn, const float4* __restrict__ bintab, float* __restrict__ out, int rows, int ntiles) {
  using namespace eo;
  constexpr int NBINS = 200;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [2 parts][hi | lo][128 x 128 B]
  uint8_t* sB = smem + SA * 2 * A_BYTES;                // [SBE][208 x 128 B]
  RowDesc* rdesc = reinterpret_cast<RowDesc*>(sB + SBE * BE_BYTES);   // [2][128]
  float4* sbins = reinterpret_cast<float4*>(rdesc + 2 * BM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbins + MAX_BINS);
  uint64_t* a_full = bars;                 // [2] 128 producer arrivals
  uint64_t* a_empty = bars + SA;           // [2] MMA commit
  uint64_t* b_full = bars + 2 * SA;        // [SBE] TMA bytes
  uint64_t* b_empty = b_full + SBE;        // [SBE] MMA commit
  uint64_t* accum = b_empty + SBE;
  uint64_t* drained = accum + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmHi); tma_prefetch_desc(&tmLo);
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SBE; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(accum, 1);
    mbar_init(drained, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < NBINS; i += THREADS) sbins[i] = __ldg(bintab + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int hb = 0; hb < NHB; ++hb)
          for (int q = 0; q < 2; ++q, ++it) {
            const int s = it % SBE, ph = (it / SBE) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_expect_tx(&b_full[s], NROWS * BK * 4);
            tma_load_2d(sB + s * BE_BYTES, q ? &tmLo : &tmHi, &b_full[s], (hb >> 1) * BK, (hb & 1) * NROWS);
          }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(BM, NROWS);
    int it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      mbar_wait(drained, (ti & 1) ^ 1);
      tc_fence_after();
      for (int hb = 0; hb < NHB; ++hb) {
        const int sa = hb & 1, pa = (ti * (NHB / 2) + (hb >> 1)) & 1;      // stage = part; its use count = ti * 7 + kb
        const int kb = hb >> 1;
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint64_t ahi = make_desc_sw128(smem_u32(sA + sa * 2 * A_BYTES));
        const uint64_t alo = make_desc_sw128(smem_u32(sA + sa * 2 * A_BYTES + A_BYTES));
        const int ksteps = (kb == NHB / 2 - 1) ? (NBINS - (NHB / 2 - 1) * BK) / UMMA_K : BK / UMMA_K;     // 200 = 6 x 32 + 8
        const uint32_t d = tmem_base + (uint32_t)(hb & 1) * NROWS;
        for (int q = 0; q < 2; ++q, ++it) {
          const int s = it % SBE, ph = (it / SBE) & 1;
          mbar_wait(&b_full[s], ph);
          tc_fence_after();
          const uint64_t bd = make_desc_sw128(smem_u32(sB + s * BE_BYTES));
          if (elect_one()) {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ko = (uint64_t)(k * UMMA_K * 4 >> 4);
              if (q == 0) {             // B_hi: A_hi.B_hi + A_lo.B_hi
                umma_tf32(d, ahi + ko, bd + ko, idesc, (kb | k) != 0);
                umma_tf32(d, alo + ko, bd + ko, idesc, 1);
              } else {                  // B_lo: A_hi.B_lo
                umma_tf32(d, ahi + ko, bd + ko, idesc, 1);
              }
            }
            umma_commit(&b_empty[s]);
            if (q == 1) {
              umma_commit(&a_empty[sa]);
              if (hb == NHB - 1) umma_commit(accum);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 2 + 4 * PGRP) {
    // ===== frame producers: group 0 -> e = s[j] + s[400 - j], group 1 -> o = s[j] - s[400 - j]; column c of a stage is j = 32 kb + c + 1
    const int g = (warp - 2) >> 2, wq = (warp - 2) & 3;
    const int r_own = (threadIdx.x - 64) & 127;
    const float sgn = g == 0 ? 1.f : -1.f;
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int m0 = tile * BM;
      RowDesc* rd = rdesc + (ti & 1) * BM;
      if (g == 0) rd[r_own] = (m0 + r_own < rows) ? mark_interior(rows_fn(m0 + r_own)) : RowDesc{0, 0, 0};
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool all_fast = __all_sync(0xffffffffu, rd[wq * 32 + lane].n == kNoReflect);
      auto run = [&](auto fastc) {
        constexpr bool FAST = decltype(fastc)::value;
        for (int kb = 0; kb < NHB / 2; ++kb) {
          const int pa = (ti * (NHB / 2) + kb) & 1;
          const int j = kb * BK + lane + 1;                 // 1 .. 224; valid up to 200
          float v[32];
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const RowDesc d = rd[wq * 32 + rr];
            float a, b;
            if constexpr (FAST) {                           // interior frames: no reflection arithmetic
              const float* p = pcm + d.base + d.start;
              a = j <= KTOT / 2 ? __ldg(p + j) : 0.f;
              b = j < KTOT / 2 ? __ldg(p + (KTOT - j)) : 0.f;       // j = 200: the centre sample stands alone
            } else {
              int i1 = d.start + j, i2 = d.start + KTOT - j;
              if (i1 < 0) i1 = -i1; else if (i1 >= d.n) i1 = 2 * (d.n - 1) - i1;
              if (i2 < 0) i2 = -i2; else if (i2 >= d.n) i2 = 2 * (d.n - 1) - i2;
              const bool on = d.n > 0 && j <= KTOT / 2;
              a = on ? __ldg(pcm + d.base + i1) : 0.f;
              b = (on && j < KTOT / 2) ? __ldg(pcm + d.base + i2) : 0.f;
            }
            v[rr] = (g == 1 && j == KTOT / 2) ? 0.f : fmaf(sgn, b, a);
          }
          mbar_wait(&a_empty[g], pa ^ 1);
          uint8_t* hi = sA + g * 2 * A_BYTES;
          uint8_t* lo = hi + A_BYTES;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const int r = wq * 32 + rr;
            const float vh = to_tf32(v[rr]);
            const float vl = to_tf32(v[rr] - vh);
            const uint32_t o = r * 128 + ((((uint32_t)lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4;
            *reinterpret_cast<float*>(hi + o) = vh;
            *reinterpret_cast<float*>(lo + o) = vl;
          }
          fence_proxy_async();
          mbar_arrive(&a_full[g]);
        }
      };
      if (all_fast) run(std::true_type{}); else run(std::false_type{});
    }
  } else {
    // ===== epilogue: thread = frame row; re of bin b in column b, im in column 208 + b
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int gr = tile * BM + row;
      const bool in_range = gr < rows;
      const bool live = in_range && rows_fn.live(gr);
      float* orow = out + (size_t)(in_range ? gr : 0) * NMEL;
      float accA = 0.f, accB = 0.f, mx = -INFINITY;
      int cur = 0;
      auto retire = [&]() {
        float val = accA;
        if (LOG == 1) val = logf(fmaxf(val, 1.1920928955078125e-07f));
        if (LOG == 2) { val = log10f(fmaxf(val, 1e-10f)); mx = fmaxf(mx, val); }
        if (in_range) orow[cur] = live ? val : 0.f;
        accA = accB; accB = 0.f; ++cur;
      };
      mbar_wait(accum, ti & 1);
      tc_fence_after();
      // (Software-pipelining these reads -- the loads of chunk c + 1 in flight while chunk c is processed, two register buffers -- was
      // measured and is not faster: 0.53 -> 0.58 ms.)
#pragma unroll 1
      for (int c = 0; c < NBINS / 8; ++c) {
        float vr[8], vi[8];
        {
          uint32_t* a = reinterpret_cast<uint32_t*>(vr);
          uint32_t* b = reinterpret_cast<uint32_t*>(vi);
          const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 8);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "r"(t0) : "memory");
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(t0 + NROWS) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (c == NBINS / 8 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(drained);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float pw = fmaf(vr[i], vr[i], vi[i] * vi[i]);
          const float4 tb = sbins[c * 8 + i];
          for (int n = __float_as_int(tb.z); n > 0; --n) retire();
          accA = fmaf(tb.x, pw, accA);
          accB = fmaf(tb.y, pw, accB);
        }
      }
      while (cur < NMEL) retire();
      if constexpr (LOG == 2) {
        if (in_range) rows_fn.note_max(gr, mx);
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <class Rows, int NMEL, int LOG>
void launch_eo(cbx_ctx* c, cudaStream_t st, const char* tag, const CUtensorMap& hi, const CUtensorMap& lo, const float* pcm, Rows rf,
               const float* bintab, float* out, int rows) {
  if (rows <= 0) return;
  auto kern = dftmel_eo_kernel<Rows, NMEL, LOG>;
  ensure_max_smem(kern, eo::SMEM_BYTES);
  const int ntiles = (rows + BM - 1) / BM;
  const int grid = ntiles < sm_count() ? ntiles : sm_count();
  // ALGORITHMIC flops: the fp32 DFT of a real 400-sample frame as the reference computes it (2 x 400 x 400 per frame), so that the
  // figure stays comparable across rounds; this kernel executes 3 x 2 x 2 x 200 x 208 per frame on the tensor pipe
  Scope sc(c->launches, st, tag, 2.0 * rows * 400 * KTOT, 4.0 * rows * (160 + NMEL));
  kern<<<grid, THREADS, eo::SMEM_BYTES, st>>>(hi, lo, pcm, rf, reinterpret_cast<const float4*>(bintab), out, rows, ntiles);
}

// second pass of the S3Tokenizer log-mel: floor at (clip max - 8), (x + 4) / 4, [T][128] -> [128][T] through a 32 x 32 tile
__global__ void __launch_bounds__(256) s3_finish_kernel(const float* __restrict__ tmp, const S3Clip* __restrict__ clips, const float* __restrict__ cmax,
                                                        float* __restrict__ out) {
  __shared__ float tile[32][33];
  const S3Clip c = clips[blockIdx.z];
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  if (t0 >= c.frames) return;
  const float floor_v = cmax[blockIdx.z] - 8.0f;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i;
    tile[i][tx] = t < c.frames ? tmp[((size_t)c.row0 + t) * kS3Mels + m0 + tx] : 0.f;
  }
  __syncthreads();
  float* o = out + (size_t)c.row0 * kS3Mels;            // clip block: [128][frames]
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + tx;
    if (t < c.frames) o[(size_t)(m0 + i) * c.frames + t] = (fmaxf(tile[tx][i], floor_v) + 4.0f) / 4.0f;
  }
}

}  // namespace fe

void run_ve_mel_tc(cbx_ctx* c, const float* pcm, const VeChunk& ch, cudaStream_t st) {
  const FrontendTables& F = c->ft;
  if (c->dft_eo)
    fe::launch_eo<fe::VeRows, kVeMels, 0>(c, st, "ve_dftmel_tc_kernel", F.tm_ve_eo_hi, F.tm_ve_eo_lo, pcm, fe::VeRows{ch.plan, ch.dyn, ch.mel_row_clip},
                                          F.ve_bins, ch.mel, ch.mel_rows);
  else
    fe::launch<fe::VeRows, 2 * kVeTcBins - 256, kVeMels, 0>(c, st, "ve_dftmel_tc_kernel", F.tm_ve_hi[0], F.tm_ve_lo[0], F.tm_ve_hi[1], F.tm_ve_lo[1],
                                                            pcm, fe::VeRows{ch.plan, ch.dyn, ch.mel_row_clip}, F.ve_bins, ch.mel, ch.mel_rows);
}

void run_kaldi_fbank_tc(cbx_ctx* c, const float* pcm, const XvChunk& ch, cudaStream_t st) {
  const FrontendTables& F = c->ft;
  fe::launch<fe::KaldiRows, 2 * kKTcBins - 256, kKMels, 1>(c, st, "kaldi_dftmel_tc_kernel", F.tm_k_hi, F.tm_k_lo, F.tm_k_hi, F.tm_k_lo,
                                                             pcm, fe::KaldiRows{ch.plan, ch.fb_row_clip}, F.k_bins, ch.fbank, ch.fb_rows);
}

}  // namespace cbx

using namespace cbx;

extern "C" {

int64_t cbx_s3_log_mel_frames(int64_t n_samples) {
  if (n_samples <= kVeNfft / 2) return CBX_ERR_ARG;      // torch.stft(center=True) reflect pad needs more than n_fft / 2 samples
  return n_samples / kVeHop;                             // 1 + L / 160 frames, the last one dropped (s3tokenizer.py:163)
}

int cbx_s3_log_mel(cbx_ctx* c, const float* pcm_dev, const int64_t* offsets_host, int n_clips, float* out_dev, void* stream) {
  if (!c) return CBX_ERR_ARG;
  if (!pcm_dev || !out_dev || !offsets_host || n_clips <= 0) { c->err = "bad argument"; return CBX_ERR_ARG; }
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  enter_stream(c, st);
  // host image of the device table: n_clips S3Clip records followed by n_clips floats of -inf (the identity of note_max)
  const size_t table_bytes = (sizeof(fe::S3Clip) + sizeof(float)) * (size_t)n_clips;
  std::vector<uint8_t> host(table_bytes);
  fe::S3Clip* clips = reinterpret_cast<fe::S3Clip*>(host.data());
  float* ninf = reinterpret_cast<float*>(clips + n_clips);
  int64_t rows = 0; int max_frames = 0;
  for (int i = 0; i < n_clips; ++i) {
    const int64_t len = offsets_host[i + 1] - offsets_host[i];
    if (len <= kVeNfft / 2) { c->err = "s3 log-mel: clip of " + std::to_string(len) + " samples; the centred STFT's reflect padding needs more than 200"; return CBX_ERR_ARG; }
    if (len > ((int64_t)1 << 30)) { c->err = "s3 log-mel: clip too long"; return CBX_ERR_ARG; }
    const int fr = (int)(len / kVeHop);
    clips[i] = fe::S3Clip{(long long)offsets_host[i], (int)len, (int)rows, fr, 0};
    ninf[i] = -INFINITY;
    rows += fr;
    max_frames = std::max(max_frames, fr);
    if (rows > 0x7fffffffLL / kS3Mels) { c->err = "s3 log-mel: too many frames in one call"; return CBX_ERR_ARG; }
  }
  if (rows == 0) return CBX_OK;
  if (c->s3_clips_cap < n_clips) {
    if (c->s3_clips) cudaFree(c->s3_clips);
    c->s3_clips_cap = n_clips + n_clips / 2 + 16;
    CBX_CUDA_OK(c, cudaMalloc(&c->s3_clips, (sizeof(fe::S3Clip) + sizeof(float)) * c->s3_clips_cap));
  }
  if (c->s3_tmp_cap < rows * kS3Mels) {
    if (c->s3_tmp) cudaFree(c->s3_tmp);
    c->s3_tmp_cap = rows * kS3Mels + rows * kS3Mels / 4;
    CBX_CUDA_OK(c, cudaMalloc((void**)&c->s3_tmp, sizeof(float) * c->s3_tmp_cap));
  }
  fe::S3Clip* dclips = (fe::S3Clip*)c->s3_clips;
  float* cmax = reinterpret_cast<float*>(dclips + n_clips);
  CBX_CUDA_OK(c, cudaMemcpyAsync(dclips, host.data(), table_bytes, cudaMemcpyHostToDevice, st));
  const FrontendTables& F = c->ft;
  if (c->dft_eo)
    fe::launch_eo<fe::S3Rows, kS3Mels, 2>(c, st, "s3_dftmel_tc_kernel", F.tm_ve_eo_hi, F.tm_ve_eo_lo, pcm_dev, fe::S3Rows{dclips, n_clips, cmax}, F.s3_bins,
                                          c->s3_tmp, (int)rows);
  else
    fe::launch<fe::S3Rows, 2 * kVeTcBins - 256, kS3Mels, 2>(c, st, "s3_dftmel_tc_kernel", F.tm_ve_hi[0], F.tm_ve_lo[0], F.tm_ve_hi[1], F.tm_ve_lo[1],
                                                            pcm_dev, fe::S3Rows{dclips, n_clips, cmax}, F.s3_bins, c->s3_tmp, (int)rows);
  for (int z0 = 0; z0 < n_clips; z0 += 65535) {
    const int nz = std::min(65535, n_clips - z0);
    Scope sc(c->launches, st, "s3_finish_kernel", 0.0, 8.0 * rows * kS3Mels * nz / n_clips);
    fe::s3_finish_kernel<<<dim3((max_frames + 31) / 32, kS3Mels / 32, nz), 256, 0, st>>>(c->s3_tmp, dclips + z0, cmax + z0, out_dev);
  }
  CBX_CUDA_OK(c, cudaGetLastError());
  return CBX_OK;
}

}  // extern "C"

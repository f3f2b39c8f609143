// S3Gen prompt mel: mel_spectrogram() of s3gen/utils/mel.py:33-81 with CosyVoice's defaults (:20-29) -- 24 kHz, n_fft = win =
// 1920, hop 480, 80 Slaney mels 0..8 kHz, reflect pad 720, magnitude sqrt(re^2 + im^2 + 1e-9), log(clamp(., 1e-5)) -- the
// "prompt_feat" half of S3Token2Mel.embed_ref (s3gen.py:177).
//
// Same construction as frontend_tc.cu (framing + Hann + DFT as one 3xTF32 GEMM per 128-frame tile, magnitude / mel / log in the
// epilogue) but the DFT is 4.8x longer and 2.5x wider: K = 1920 (60 K blocks), N = 2 * 640 columns (bins 1..639, the only
// ones with mel weight below 8 kHz, plus one zero pair).  TMEM holds 512 fp32 columns, so a tile makes three passes over its
// frames (512 + 512 + 256 columns); the 80 mel accumulators of every frame live in shared memory across the passes.
// Nothing but the [frames x 80] result is written to HBM; the PCM is read 4x (hop / n_fft) from L2.
//
// Warp roles (192 threads): warp 0 = TMA producer of the DFT tiles (hi / lo), warp 1 = MMA issuer (elect-one), warps 2..5 =
// frame producers (gather -> hi/lo split -> swizzled K-major A stages) and, at the end of each pass, the epilogue.
#include <math.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "cbx_internal.h"
#include "tc.cuh"

namespace cbx {
namespace pm {

using namespace tc;

constexpr int SR = 24000, NFFT = 1920, HOP = 480, NMEL = 80, PAD = (NFFT - HOP) / 2;
constexpr int NBINS = 640;                      // table bins: DFT bins 1..639 + one zero pair
constexpr int NCOLS = 2 * NBINS;                // 1280 = 5 x 256
constexpr int NKB = NFFT / BK;                  // 60
constexpr int NPASS = 3;
constexpr int SA = 2;                           // A stages: hi + lo, 16 KB each
constexpr int SB = 3;                           // B slots of 32 KB
constexpr int A_BYTES = BM * BK * 4;
constexpr int B_BYTES = 256 * BK * 4;
constexpr int MELLD = NMEL + 1;
constexpr int MEL_BYTES = BM * MELLD * 4;
constexpr int SMEM_BYTES = SA * 2 * A_BYTES + SB * B_BYTES + MEL_BYTES + BM * 16 + 1024 + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Clip { long long pcm_off; int n; int row0; int frames; int pad; };
struct RowDesc { long long base; int start; int n; };

__global__ void __launch_bounds__(192, 1)
promptmel_kernel(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo, const float* __restrict__ pcm,
                 const Clip* __restrict__ clips, int n_clips, const float4* __restrict__ bintab, float* __restrict__ out, int rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [SA][hi | lo][128 x 128 B]
  uint8_t* sB = smem + SA * 2 * A_BYTES;                // [SB][256 x 128 B]
  float* melacc = reinterpret_cast<float*>(sB + SB * B_BYTES);        // [128][MELLD]
  RowDesc* rdesc = reinterpret_cast<RowDesc*>(reinterpret_cast<uint8_t*>(melacc) + MEL_BYTES);   // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rdesc + BM);
  uint64_t* a_full = bars;                 // [SA] 128 producer arrivals
  uint64_t* a_empty = bars + SA;           // [SA] MMA commit
  uint64_t* b_full = bars + 2 * SA;        // [SB] TMA bytes
  uint64_t* b_empty = b_full + SB;         // [SB] MMA commit
  uint64_t* accum = b_empty + SB;          // one phase per pass
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmHi); tma_prefetch_desc(&tmLo);
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (warp >= 2) {
    // frame row -> (clip, t): clips are packed back to back, row0 ascending
    const int r = threadIdx.x - 64;
    const int row = m0 + r;
    RowDesc d{0, 0, 0};
    if (row < rows) {
      int lo = 0, hi = n_clips - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (clips[mid].row0 <= row) lo = mid; else hi = mid - 1;
      }
      const Clip c = clips[lo];
      const int t = row - c.row0;
      if (t < c.frames) d = RowDesc{c.pcm_off, t * HOP - PAD, c.n};
    }
    rdesc[r] = d;
#pragma unroll 1
    for (int m = 0; m < MELLD; ++m) melacc[r * MELLD + m] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: per K block of a pass the slots are hi(256 cols), lo(256 cols) [, hi(next 256), lo(next 256)]
    if (lane == 0) {
      int it = 0;
      for (int pass = 0; pass < NPASS; ++pass) {
        const int col0 = pass * 512, nq = pass == NPASS - 1 ? 2 : 4;
        for (int kb = 0; kb < NKB; ++kb)
          for (int q = 0; q < nq; ++q, ++it) {
            const int s = it % SB, ph = (it / SB) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_expect_tx(&b_full[s], B_BYTES);
            tma_load_2d(sB + s * B_BYTES, (q & 1) ? &tmLo : &tmHi, &b_full[s], kb * BK, col0 + (q >> 1) * 256);
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    constexpr uint32_t idesc = make_idesc_tf32(BM, 256);
    int it = 0, ia = 0;
    for (int pass = 0; pass < NPASS; ++pass) {
      const int nq = pass == NPASS - 1 ? 2 : 4;
      for (int kb = 0; kb < NKB; ++kb, ++ia) {
        const int sa = ia % SA, pa = (ia / SA) & 1;
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint64_t ahi = make_desc_sw128(smem_u32(sA + sa * 2 * A_BYTES));
        const uint64_t alo = make_desc_sw128(smem_u32(sA + sa * 2 * A_BYTES + A_BYTES));
        for (int q = 0; q < nq; ++q, ++it) {
          const int s = it % SB, ph = (it / SB) & 1;
          mbar_wait(&b_full[s], ph);
          tc_fence_after();
          const uint64_t bd = make_desc_sw128(smem_u32(sB + s * B_BYTES));
          const uint32_t d = tmem_base + (q >> 1) * 256;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t ko = (uint64_t)(k * UMMA_K * 4 >> 4);
              if ((q & 1) == 0) {       // B_hi: A_hi.B_hi + A_lo.B_hi
                umma_tf32(d, ahi + ko, bd + ko, idesc, (kb | k) != 0);
                umma_tf32(d, alo + ko, bd + ko, idesc, 1);
              } else {                  // B_lo: A_hi.B_lo
                umma_tf32(d, ahi + ko, bd + ko, idesc, 1);
              }
            }
            umma_commit(&b_empty[s]);
            if (q == nq - 1) {
              umma_commit(&a_empty[sa]);
              if (kb == NKB - 1) umma_commit(accum);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int wq = warp - 2;
    const int q4 = warp & 3;                       // TMEM lane quadrant this warp may read
    int ia = 0;
    for (int pass = 0; pass < NPASS; ++pass) {
      // ===== frame producers: warp fills rows [32 wq, 32 wq + 32) of every A stage, lanes along K (coalesced PCM reads)
      for (int kb = 0; kb < NKB; ++kb, ++ia) {
        const int sa = ia % SA, pa = (ia / SA) & 1;
        const int k = kb * BK + lane;
        float v[32];
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const RowDesc d = rdesc[wq * 32 + rr];
          int i = d.start + k;
          if (i < 0) i = -i; else if (i >= d.n) i = 2 * (d.n - 1) - i;
          v[rr] = d.n > 0 ? __ldg(pcm + d.base + i) : 0.f;
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
      // ===== epilogue of the pass: thread = frame row; D columns 2b, 2b+1 = re, im of table bin (pass * 256 + b)
      mbar_wait(accum, pass & 1);
      tc_fence_after();
      const int row = q4 * 32 + lane;
      float* acc = melacc + row * MELLD;
      const int nchunk = (pass == NPASS - 1 ? 256 : 512) / 16;
      const float4* bt = bintab + pass * 256;
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        float v[16];
        {
          uint32_t* rv = reinterpret_cast<uint32_t*>(v);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(rv[0]), "=r"(rv[1]), "=r"(rv[2]), "=r"(rv[3]), "=r"(rv[4]), "=r"(rv[5]), "=r"(rv[6]), "=r"(rv[7]), "=r"(rv[8]),
                "=r"(rv[9]), "=r"(rv[10]), "=r"(rv[11]), "=r"(rv[12]), "=r"(rv[13]), "=r"(rv[14]), "=r"(rv[15])
              : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c * 16))
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float mag = sqrtf(fmaf(v[2 * i], v[2 * i], v[2 * i + 1] * v[2 * i + 1]) + 1e-9f);       // mel.py:76
          const float4 tb = __ldg(bt + c * 8 + i);                  // warp-uniform: {w0, w1, m0, m1}
          const int ma = __float_as_int(tb.z), mb = __float_as_int(tb.w);
          acc[ma] = fmaf(tb.x, mag, acc[ma]);
          acc[mb] = fmaf(tb.y, mag, acc[mb]);
        }
      }
      tc_fence_before();          // the next pass overwrites TMEM: its first MMA waits for all 128 a_full arrivals below
    }
    __syncwarp();
    // each warp owns the 32 rows it accumulated: write them out row by row, lanes along the mel axis
    for (int rr = 0; rr < 32; ++rr) {
      const int r = q4 * 32 + rr;
      const int gr = m0 + r;
      if (gr >= rows) break;
      for (int m = lane; m < NMEL; m += 32) out[(size_t)gr * NMEL + m] = logf(fmaxf(melacc[r * MELLD + m], 1e-5f));   // mel.py:12-13
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- tables (float64 on the host, rounded once) -------------------------------------------------------------------------
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : m * f_sp;
}
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static float tf32_round(float x) {
  uint32_t u; std::memcpy(&u, &x, 4);
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  std::memcpy(&x, &u, 4);
  return x;
}

static int build_tables(cbx_ctx* c) {
  PromptMelTables& T = c->pm;
  const double PI = 3.14159265358979323846;
  const size_t nmat = (size_t)NCOLS * NFFT;
  std::vector<float> host(2 * nmat + 4 * NBINS, 0.f);
  float* hi = host.data();
  float* lo = hi + nmat;
  float* bins = lo + nmat;
  // torch.hann_window(1920) (periodic) folded into the DFT rows of bins 1..639; row 2b = re, 2b+1 = im of bin b+1
  for (int b = 0; b < NBINS - 1; ++b)
    for (int n = 0; n < NFFT; ++n) {
      const double w = 0.5 - 0.5 * std::cos(2.0 * PI * n / NFFT);
      const double ang = 2.0 * PI * (double)(((long long)(b + 1) * n) % NFFT) / NFFT;
      const double v[2] = {w * std::cos(ang), -w * std::sin(ang)};
      for (int part = 0; part < 2; ++part) {
        const float h = tf32_round((float)v[part]);
        hi[(size_t)(2 * b + part) * NFFT + n] = h;
        lo[(size_t)(2 * b + part) * NFFT + n] = tf32_round((float)(v[part] - (double)h));
      }
    }
  // librosa.filters.mel(sr=24000, n_fft=1920, n_mels=80, fmin=0, fmax=8000): Slaney scale, area norm; 2-sparse per bin
  std::vector<double> edges(NMEL + 2);
  const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
  for (int i = 0; i < NMEL + 2; ++i) edges[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (NMEL + 1));
  for (int b = 0; b < NBINS - 1; ++b) {
    const double f = (double)SR * (b + 1) / NFFT;
    int cnt = 0; int mm[2] = {0, 0}; float ww[2] = {0.f, 0.f};
    for (int m = 0; m < NMEL; ++m) {
      const double up = (f - edges[m]) / (edges[m + 1] - edges[m]);
      const double dn = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
      const float tri = (float)std::fmax(0.0, std::fmin(up, dn));
      const float wv = (float)((double)tri * (2.0 / (edges[m + 2] - edges[m])));
      if (wv != 0.f) {
        if (cnt >= 2) { c->err = "prompt mel bank is not 2-sparse"; return CBX_ERR_STATE; }
        mm[cnt] = m; ww[cnt] = wv; ++cnt;
      }
    }
    bins[4 * b] = ww[0]; bins[4 * b + 1] = ww[1];
    std::memcpy(&bins[4 * b + 2], &mm[0], 4); std::memcpy(&bins[4 * b + 3], &mm[1], 4);
  }
  CBX_CUDA_OK(c, cudaMalloc((void**)&T.blob, host.size() * sizeof(float)));
  CBX_CUDA_OK(c, cudaMemcpy(T.blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  T.hi = T.blob; T.lo = T.blob + nmat; T.bins = T.blob + 2 * nmat;
  T.tm_hi = tc::make_map_2d(T.hi, NCOLS, NFFT, NFFT, 256, false);
  T.tm_lo = tc::make_map_2d(T.lo, NCOLS, NFFT, NFFT, 256, false);
  ensure_max_smem(promptmel_kernel, SMEM_BYTES);
  T.ready = true;
  return CBX_OK;
}

}  // namespace pm
}  // namespace cbx

using namespace cbx;

extern "C" {

int64_t cbx_prompt_mel_frames(int64_t n_samples) {
  if (n_samples <= pm::PAD) return CBX_ERR_ARG;          // torch reflect padding needs pad < length (mel.py:56-58)
  return 1 + (n_samples + 2 * pm::PAD - pm::NFFT) / pm::HOP;
}

int cbx_prompt_mel(cbx_ctx* c, const float* pcm_dev, const int64_t* offsets_host, int n_clips, float* out_dev, void* stream) {
  if (!c) return CBX_ERR_ARG;
  if (!pcm_dev || !out_dev || !offsets_host || n_clips <= 0) { c->err = "bad argument"; return CBX_ERR_ARG; }
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  enter_stream(c, st);
  PromptMelTables& T = c->pm;
  if (!T.ready) { int rc = pm::build_tables(c); if (rc) return rc; }
  std::vector<pm::Clip> clips(n_clips);
  int64_t rows = 0;
  for (int i = 0; i < n_clips; ++i) {
    const int64_t len = offsets_host[i + 1] - offsets_host[i];
    if (len <= pm::PAD) { c->err = "prompt mel: clip of " + std::to_string(len) + " samples; reflect padding by 720 needs more (mel.py:56-58)"; return CBX_ERR_ARG; }
    if (len > ((int64_t)1 << 30)) { c->err = "prompt mel: clip too long"; return CBX_ERR_ARG; }
    const int64_t fr = cbx_prompt_mel_frames(len);
    clips[i] = pm::Clip{(long long)offsets_host[i], (int)len, (int)rows, (int)fr, 0};
    rows += fr;
    if (rows > 0x7fff0000LL) { c->err = "prompt mel: too many frames in one call"; return CBX_ERR_ARG; }
  }
  if (T.clips_cap < n_clips) {
    if (T.clips) cudaFree(T.clips);
    T.clips_cap = n_clips + n_clips / 2 + 16;
    CBX_CUDA_OK(c, cudaMalloc(&T.clips, sizeof(pm::Clip) * T.clips_cap));
  }
  CBX_CUDA_OK(c, cudaMemcpyAsync(T.clips, clips.data(), sizeof(pm::Clip) * n_clips, cudaMemcpyHostToDevice, st));
  {
    Scope sc(c->launches, st, "promptmel_tc_kernel", 2.0 * rows * pm::NCOLS * pm::NFFT,   /* algorithmic; executed = 3x (3xTF32) */ 4.0 * rows * (pm::HOP + pm::NMEL));
    pm::promptmel_kernel<<<(unsigned)((rows + tc::BM - 1) / tc::BM), 192, pm::SMEM_BYTES, st>>>(
        T.tm_hi, T.tm_lo, pcm_dev, (const pm::Clip*)T.clips, n_clips, reinterpret_cast<const float4*>(T.bins), out_dev, (int)rows);
  }
  CBX_CUDA_OK(c, cudaGetLastError());
  return CBX_OK;
}

}  // extern "C"

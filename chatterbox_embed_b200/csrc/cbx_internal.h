// Internal declarations shared by the translation units of libcbx.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/cbx.h"

namespace cbx {

// ---- fixed hyper-parameters of the path ------------------------------------------------------
// VoiceEncConfig (voice_encoder/config.py:1-18)
constexpr int kSR = 16000;
constexpr int kVeNfft = 400, kVeHop = 160, kVeBins = 201, kVeMels = 40;
constexpr int kVePartial = 160, kVeHidden = 256, kVeEmbed = 256, kVeGates = 1024;
constexpr int kVeSpecN = 2 * kVeBins;      // interleaved re/im columns of the DFT GEMM
constexpr int kS3Mels = 128;                // S3Tokenizer log-mel (s3tokenizer.py:39-47): the VoiceEncoder STFT with 128 mels
constexpr int kVeTcBins = 200, kKTcBins = 256;   // bins (incl. one zero pad) of the tensor-core front-end GEMMs
// Kaldi fbank as called by xvector.py:50 (torchaudio kaldi.py defaults, 80 bins)
constexpr int kKWin = 400, kKHop = 160, kKPad = 512, kKBins = 257, kKMels = 80;
constexpr int kKSpecN = 2 * kKBins;
// librosa.effects.trim defaults (voice_encoder.py:267)
constexpr int kTrimFrame = 2048, kTrimHop = 512;
// CAMPPlus (xvector.py:340-405)
constexpr int kFcmC = 32, kFcmOut = 320, kTdnnC = 128, kBnC = 128, kGrowth = 32, kCamHid = 64;
constexpr int kSegLen = 100, kXvEmbed = 192, kStatsC = 512;
constexpr int kGuardTd = 2;                // zero guard rows (T' domain) between clips; x2 in the fbank domain
constexpr float kBnEps = 1e-5f;

// ---- per-clip tables ----------------------------------------------------------------------------
struct ClipPlan {            // host-computed, uploaded once per chunk
  int64_t pcm_off;           // first sample in the flat PCM buffer
  int32_t n_samples;
  int32_t out_index;         // row of the caller's output arrays
  // VoiceEncoder, upper bounds from the untrimmed length
  int32_t mel_row, mel_rows; // first row / rows allocated (= target of the untrimmed length)
  int32_t slot0, slots;      // first partial slot / slots allocated
  int32_t trim_blk0;         // first entry of the trim scratch (one float per 512-sample block)
  // CAMPPlus
  int32_t xv_frames, xv_tdnn, xv_segs;
  int32_t fb_row;            // first fbank/FCM row (even), = 2*td_row
  int32_t td_row;            // first row in the T' domain
  int32_t seg0;              // first CAM segment
};

struct ClipDyn {             // device-computed (depends on the trim result)
  int32_t trim_s, trim_e;    // [start,end) samples kept by librosa.effects.trim
  int32_t ve_frames;         // 1 + (trim_e-trim_s)/160
  int32_t ve_frames_eff;     // min(ve_frames, target): rows of mel actually computed
  int32_t ve_parts;          // partials of the trimmed clip
  int32_t status;
};

// Launch accounting; with profiling on, every launch is bracketed by CUDA events on its own stream.
struct ProfRec { const char* tag; double flops; double bytes; cudaEvent_t e0, e1; };
struct Launches {
  int64_t count = 0;
  bool prof = false;
  int bn_prefetch = 1;          // option bn_prefetch: successor-tile L2 prefetch of the pre-activation GEMM (tc.cuh)
  std::vector<ProfRec> recs;
};
struct Scope {
  Launches& L; cudaStream_t st; bool on;
  Scope(Launches& l, cudaStream_t s, const char* tag, double flops = 0.0, double bytes = 0.0) : L(l), st(s), on(l.prof) {
    if (on) {
      ProfRec r{tag, flops, bytes, nullptr, nullptr};
      cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
      cudaEventRecord(r.e0, st);
      L.recs.push_back(r);
    }
  }
  ~Scope() {
    if (on) cudaEventRecord(L.recs.back().e1, st);
    L.count++;
  }
};

// ---- packed weights (device pointers into one allocation each) ----------------------------------
struct ConvW { const float* w; const float* bias; int K; };   // w [Cout][K], K-major
struct DenseLayerW {
  int cin;
  const float *a1, *b1;      // BN1 scale/shift on the layer input (prologue)
  const float *w1, *t2;      // [128][cin] with BN2 scale folded, BN2 shift
  const float *w1h;          // the same as bf16 [128][cin] (option cat_bf16 = 2)
  const float *wl;           // [32][3*128] local conv, k = tap*128 + c
  const float *wlh;          // the same as bf16 (bf16 mode)
  const float *wc1, *bc1;    // [64][128], [64]
  const float *wc2, *bc2;    // [32][64], [32]
  const float *wc1T, *wc2T;  // [128][64], [64][32] transposed copies
};
struct TransitW { int cin, cout; const float *a, *b, *w, *wh; };   // wh: w as bf16 (option cat_bf16 = 2)

struct VeWeights {
  bool loaded = false;
  float* blob = nullptr;
  const float *wih0, *wih[3], *whhT[3], *bias[3], *wpT, *bp;   // whhT [256][1024], wpT [256][256]
  CUtensorMap tm_wih[3];     // TMA maps of W_ih (B operand of the input-projection GEMMs)
  // tensor-core recurrence (lstm_tc.cu): gate rows permuted to row 128 j + 4 u + g  <->  gate g of unit 32 j + u
  const float *wih_p[3], *whh_p[3], *bias_p[3];   // whh_p tf32-rounded [1024][256]
  CUtensorMap tm_wih_p[3], tm_wih_p256[3];     // boxes of 128 / 256 gate rows
};
struct XvWeights {
  bool loaded = false;
  float* blob = nullptr;
  const float *conv1_w, *conv1_b;                 // [32][9], [32]
  ConvW res[2][2][2];                             // [layer][block][conv]; conv2 of block 0 has the shortcut appended (K=320)
  ConvW head_conv2;
  ConvW tdnn;                                     // [128][5*320]
  DenseLayerW dense[52];
  TransitW transit[3];
  const float *out_a, *out_b;                     // out_nonlinear BN
  const float *fin_w, *fin_b;                     // [192][1024] (BN folded), [192]
  CUtensorMap tm_tdnn, tm_w1[52], tm_wl[52], tm_tr[3], tm_tr256[3];   // TMA maps of the GEMM B operands (tm_tr256: 256-row boxes for the N = 256 transit tiles)
  CUtensorMap tm_w1h[52], tm_trh[3];                     // bf16 copies: box {64 channels, 128 rows}
  CUtensorMap tm_wlh[52];                                // bf16 local-conv weights: box {64 channels, 32 rows}
  CUtensorMap tm_res[2][2][2], tm_head2;
};
struct FrontendTables {
  float* blob = nullptr;
  const float *ve_dft;     // [402][400]  window folded
  const float *ve_mel;     // [40][201]
  const float *k_dft;      // [514][400]  dc-removal, pre-emphasis, povey window folded
  const float *k_mel;      // [80][257]
  // tensor-core front end (frontend_tc.cu): DFT rows of bins 1.. only (bin 0 and the Nyquist bin carry zero mel weight in
  // both banks), split hi/lo for 3xTF32, and the 2-sparse bin -> mel tables
  const float *ve_dft_hi, *ve_dft_lo;   // [400][400]  bins 1..199 + one zero pair
  const float *k_dft_hi, *k_dft_lo;     // [512][400]  bins 1..255 + one zero pair
  const float *ve_bins, *k_bins;        // [200][4], [256][4]
  const float *s3_bins;                 // [200][4]
  CUtensorMap tm_ve_hi[2], tm_ve_lo[2], tm_k_hi, tm_k_lo;
  // even/odd form of the VoiceEncoder / S3 DFT (frontend_tc.cu, dftmel_eo_kernel): rows [0, 208) = cos part of bins 1..199 against
  // e[j] = s[j] + s[400-j], rows [208, 416) = -sin part against o[j] = s[j] - s[400-j], j = 1..200 in columns 0..199 of 224
  const float *ve_eo_hi, *ve_eo_lo;     // [416][224]
  CUtensorMap tm_ve_eo_hi, tm_ve_eo_lo;
};
// S3Gen prompt mel (promptmel_tc.cu), built on first use: Hann-folded 1920-point DFT rows of bins 1..639 split hi/lo, bin -> mel table
struct PromptMelTables {
  bool ready = false;
  float* blob = nullptr;
  const float *hi = nullptr, *lo = nullptr, *bins = nullptr;   // [1280][1920] x2, [640][4]
  CUtensorMap tm_hi, tm_lo;
  void* clips = nullptr; int clips_cap = 0;
};

}  // namespace cbx

struct cbx_ctx {
  int device = 0;
  std::string err;
  cbx::VeWeights ve;
  cbx::XvWeights xv;
  cbx::FrontendTables ft;
  cbx::Launches launches;
  // options
  int64_t xv_chunk_rows = 300000;     // fbank rows per CAMPPlus chunk
  int64_t fcm_chunk_rows = 65536;     // fbank rows per FCM sub-chunk
  int64_t lstm_chunk_slots = 3072;    // partial slots per VoiceEncoder chunk
  int64_t mode = 1;                   // 1: tcgen05 (TF32) kernels, the product path; 0: strict-fp32 SIMT yardstick
  int64_t lstm_trace = 0;             // device pointer of the clock trace buffer
  int64_t lstm_impl = 2;              // 1: DSMEM-push recurrence, 2: L2 multicast-TMA recurrence
  int64_t lstm_dbg = 0;               // timing experiments (lstm_tc.cu)
  int64_t lstm_gate_warps = 4;        // gate warps per TMEM lane quadrant of the recurrence kernel: 4 (16 gate warps) or 2 (8, round 1)
  int64_t lstm_late = 1;              // with both encoders on two streams: start the recurrence when CAMPPlus enters its D-TDNN phase
  int64_t dft_eo = 1;                 // VoiceEncoder / S3 front-end DFT in its even / odd form (two K = 200 GEMMs); 0 = one K = 400 GEMM
  int64_t transit_n256 = 1;           // transit GEMMs with 128 x 256 output tiles
  int64_t fcm_fuse = 1;               // identity residual blocks of the FCM head as one fused kernel (fcm_block_tc.cu); 0 = two convolution kernels
  int64_t u_bf16 = 0;                 // 1 = the bottleneck output u is stored as bf16 and the local convolution runs on bf16 operands (bf16 mode)
  int64_t xw_bf16 = 0;                // 1 = the LSTM input projections are stored as bf16 (bf16 mode)
  int64_t cat_bf16 = 0;               // 1 = the D-TDNN GEMMs read a bf16 copy of the concatenation buffers, 2 = and run on bf16 operands (kind::f16)
  int64_t probe = 0;                  // timing experiments, results are WRONG while set: bit 0 = no CAM gate kernel (tools/probe_bounds.py)
  int64_t batch_invariant = 0;        // 1: exact warp-level segment sums: x-vectors bit-identical whatever the batch (about 0.5 ms per step)
  int64_t pdl = 1;                    // programmatic dependent launch along the dense-layer chain
  // buffers owned by the library (cbx_embed_host / cbx_embed_host_submit): two slots so that the host<->device copies of
  // one batch overlap the kernels of the other; the workspace is shared (kernels of both slots run on one compute stream)
  void* own_ws = nullptr; int64_t own_ws_bytes = 0;
  struct HostSlot {
    float* dev_pcm = nullptr; int64_t dev_pcm_bytes = 0;
    float* dev_out = nullptr; int64_t dev_out_bytes = 0;
    float* pin_pcm = nullptr; int64_t pin_pcm_bytes = 0;
    float* pin_out = nullptr; int64_t pin_out_bytes = 0;
    cudaEvent_t h2d = nullptr, comp = nullptr, d2h = nullptr;
    int n = 0, flags = 0; bool busy = false;
  } slot[2];
  cudaStream_t own_stream = nullptr;     // compute
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaStream_t aux_stream = nullptr;     // CAMPPlus chain when it runs beside the VoiceEncoder chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int64_t overlap = 1;
  // resampler (resample.cu): filter banks per (orig, new) and the clip table
  std::map<long long, float*> resample_banks;
  void* resample_clips = nullptr; int resample_clips_cap = 0;
  cbx::PromptMelTables pm;
  // S3Tokenizer log-mel (frontend_tc.cu): clip table, per-clip maxima, [frames][128] scratch
  void* s3_clips = nullptr; int s3_clips_cap = 0;
  float* s3_tmp = nullptr; int64_t s3_tmp_cap = 0;
  // Stream discipline (include/cbx.h "Conventions"): the context's tables and scratch buffers (clip tables of the resampler /
  // prompt mel / S3 front-end, their scratch, the shared workspace of the host path, the aux stream's fork / join events) are
  // single buffers.  Every stream-ordered entry point calls cbx::enter_stream(): when the caller's stream differs from the
  // one the previous call used, the new stream first waits for everything the previous stream had been given, so two calls on
  // different streams can never overlap on those buffers (they serialise instead of corrupting each other).
  // recurrence placement (api.cu, embed_core): event recorded by the CAMPPlus chain when it enters the D-TDNN phase; the two flags are
  // set for the duration of one call
  cudaEvent_t ev_dtdnn = nullptr; bool xv_mark_dtdnn = false, ve_wait_dtdnn = false;
  cudaStream_t last_stream = nullptr; bool last_stream_valid = false;
  cudaEvent_t ev_order = nullptr;
  // last-run bookkeeping for the stage taps
  std::vector<cbx::ClipPlan> last_plan;
  std::map<std::string, std::vector<int64_t>> taps;   // name -> {byte_offset, rows, cols, ld}
};

namespace cbx {

// RAII: make the context's device current for the duration of an entry point and put the caller's device back afterwards
// (a single-process multi-GPU caller must not find its current device switched under it)
struct DeviceGuard {
  int prev = -1; bool changed = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) { cudaSetDevice(dev); changed = true; }
  }
  ~DeviceGuard() { if (changed && prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// see cbx_ctx::last_stream
inline void enter_stream(cbx_ctx* c, cudaStream_t st) {
  if (c->last_stream_valid && c->last_stream != st && c->ev_order) {
    cudaEventRecord(c->ev_order, c->last_stream);
    cudaStreamWaitEvent(st, c->ev_order, 0);
  }
  c->last_stream = st; c->last_stream_valid = true;
}

// host integer logic (host_plan.cpp)
int ve_frame_step(double overlap, double rate);
void ve_num_wins(int64_t n_frames, int step, double min_cov, int64_t* n, int64_t* target);

// weights.cu
int load_ve(cbx_ctx* c, const std::map<std::string, std::pair<const float*, int64_t>>& t);
int load_xv(cbx_ctx* c, const std::map<std::string, std::pair<const float*, int64_t>>& t);
int build_frontend_tables(cbx_ctx* c);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: set it once per (kernel, device), so that several
// contexts on several GPUs of one process all get it (tc.cu)
void ensure_max_smem(const void* kernel, int bytes);
template <class K> inline void ensure_max_smem(K kernel, int bytes) { ensure_max_smem(reinterpret_cast<const void*>(kernel), bytes); }

// ---- workspace carving ------------------------------------------------------------------------
struct Carver {
  char* base; int64_t off = 0; int64_t cap;
  Carver(void* b, int64_t c) : base((char*)b), cap(c) {}
  template <class T> T* take(int64_t n) {
    off = (off + 255) & ~int64_t(255);
    T* p = (T*)(base ? base + off : nullptr);
    off += n * (int64_t)sizeof(T);
    return p;
  }
};

struct VeChunk {            // device buffers of one VoiceEncoder chunk
  int n_clips; int mel_rows; int slots; int trim_blocks;
  int64_t pcm_samples;      // samples of the chunk's clips (algorithmic bytes of the trim kernel)
  ClipPlan* plan; ClipDyn* dyn;
  float* trim_scratch;
  int32_t* mel_row_clip;    // [mel_rows]
  int32_t* slot_clip;       // [slots]  clip of the slot, -1 if unused
  int32_t* slot_row;        // [slots]  first mel row of the partial
  float* spec;              // [mel_rows][402]
  float* mel;               // [mel_rows][40]
  float* xw0;               // [mel_rows][1024]
  float* xw;                // [slots*160][1024]
  float* hseq;              // [slots*160][256]
  float* hlast;             // [3][slots][256]  final hidden state of each layer (stage tap; layer 3 feeds the projection)
  float* pemb;              // [slots][256]
};

struct XvChunk {
  int n_clips; int fb_rows; int td_rows; int segs; int fcm_rows;   // fcm_rows: rows of the FCM sub-chunk buffers
  ClipPlan* plan;             // device copy
  const ClipPlan* hplan;      // host copy (sub-chunk planning)
  int32_t* fb_row_clip;     // [fb_rows] clip or -1 (guard)
  int32_t* td_row_clip;     // [td_rows]
  int32_t* td_row_seg;      // [td_rows] global segment index or -1
  int32_t* seg_clip;        // [segs]
  float* spec;              // [fcm_rows][514]   (sub-chunk)
  float* fbank;             // [fb_rows][80]
  float* cmn_sum;           // [n_clips][80]
  float *b0, *b1, *b2, *b3, *b4, *b5, *b6;   // FCM activations of one sub-chunk
  float* fcm_out;           // [fb_rows][320]
  float *cat1, *cat2, *cat3;   // [td_rows][512|1024|1024]
  uint16_t *cat1h, *cat2h, *cat3h;   // bf16 copies of the same (option cat_bf16; null otherwise)
  float* u;                 // [td_rows][128]
  uint16_t* u16;            // the same as bf16 (bf16 mode: the bottleneck GEMM writes it INSTEAD of u, the local convolution reads it)
  float* tr3;               // [td_rows][512]
  float* seg_sum;           // [segs][128] fp32 (strict mode) or 64-bit fixed point (tensor-core mode: 2 floats per entry)
  float* gate;              // [segs][32]
  float* stats;             // [n_clips][1024]
};

// kernels (frontend.cu / lstm.cu / campplus.cu)
void run_ve_chunk(cbx_ctx* c, const float* pcm, const VeChunk& ch, float trim_top_db, bool no_trim, int step,
                  double min_cov, float* ve_out, int32_t* status, cudaStream_t st);
void run_ve_lstm(cbx_ctx* c, const VeChunk& ch, cudaStream_t st);   // xw0 .. pemb
void run_ve_forward_partials(cbx_ctx* c, const float* mels, int n, float* out, void* ws, cudaStream_t st);
// frontend_tc.cu: frames -> 3xTF32 DFT GEMM -> power -> sparse mel (-> log) in one kernel
void run_ve_mel_tc(cbx_ctx* c, const float* pcm, const VeChunk& ch, cudaStream_t st);
void run_kaldi_fbank_tc(cbx_ctx* c, const float* pcm, const XvChunk& ch, cudaStream_t st);
void run_local_conv_tc(cbx_ctx* c, cudaStream_t st, const CUtensorMap& tmU, const CUtensorMap& tmW, const CUtensorMap& tmOut, int M, int dil,
                       int col0, const float* gate, const int32_t* row_seg, bool pdl = false, uint16_t* shadow = nullptr, int ldh = 0, bool u16 = false);   // local_tc.cu
int lstm_padded_slots(int n_slots);     // slots rounded up to whole 224-partial cluster tiles (lstm_tc.cu)
void run_lstm_rec_tc2(cbx_ctx* c, const void* xw, bool xw_bf16, const int32_t* slot_row, const float* whh_perm, float* hseq, float* hlast,
                      int n_slots, cudaStream_t st);
void run_fcm_conv_tc(cbx_ctx* c, cudaStream_t st, const CUtensorMap& tmW, const float* bias, const float* in, int F_in, int F_out, int sf,
                     const float* sc, int F_sc, const float* res, float* out, const int32_t* row_clip, int rows, int prows, double flops,
                     const char* tag = "fcm_conv_gemm", bool pdl = false);
// fcm_block_tc.cu: an identity residual block (conv 3x3 -> BN -> ReLU -> conv 3x3 -> BN -> + x -> ReLU) as one kernel
void run_fcm_block_tc(cbx_ctx* c, cudaStream_t st, const CUtensorMap& tmW1, const float* bias1, const CUtensorMap& tmW2, const float* bias2,
                      const float* in, int F, float* out, const int32_t* row_clip, int rows, int prows, const char* tag, bool pdl);   // fcm_tc.cu
void run_lstm_rec_tc(cbx_ctx* c, const float* xw, const int32_t* slot_row, const float* whh_perm, float* hseq, float* hlast,
                     int n_slots, cudaStream_t st);   // lstm_tc.cu
// feats != nullptr: CAMPPlus.forward on precomputed (already mean-normalised) features [sum T][80]; feat_off (host, frames, indexed by the
// clip's position in the call) says where each clip's rows start; the fbank / CMN kernels are skipped
void run_xv_chunk(cbx_ctx* c, const float* pcm, const XvChunk& ch, float* xv_out, int32_t* status, cudaStream_t st,
                  const float* feats = nullptr, const int64_t* feat_off = nullptr);

}  // namespace cbx

#define CBX_CUDA_OK(ctx, expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
      return CBX_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

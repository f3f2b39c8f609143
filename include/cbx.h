/*
 * cbx.h -- C ABI of libcbx.so: the B200-native speaker-embedding path.
 *
 * This is the drop-in boundary for the voice-clone conditioning hot path of
 * chrijaque/chatterbox_embed.  The reference has no FFI of its own (it is 100 %
 * Python); the boundary it exposes is the Python method surface
 *   VoiceEncoder.embeds_from_wavs   src/chatterbox/models/voice_encoder/voice_encoder.py:246-274
 *   VoiceEncoder.inference/forward  voice_encoder.py:139-199
 *   CAMPPlus.inference              src/chatterbox/models/s3gen/xvector.py:425-428
 *   S3Token2Mel.save_voice_clone    src/chatterbox/models/s3gen/s3gen.py:107-119
 * The Python classes in chatterbox_embed_b200/ keep those signatures and bind the
 * entry points below through ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions: plain pointers and sizes only; every function returning int
 * returns 0 on success and a negative code on failure with a message available
 * from cbx_last_error(); device pointers are caller-owned; work is stream
 * ordered on the stream passed in (a cudaStream_t cast to void*; NULL = the
 * legacy default stream) and the calls do not synchronise unless stated; one
 * context per GPU, not thread-safe (one host thread at a time per context).
 * The context's device tables and scratch buffers are single buffers: when a
 * call arrives on a different stream than the previous call, the library makes
 * the new stream wait (event) for everything queued on the previous one, so
 * calls on two streams serialise instead of overwriting each other's tables.
 * Every entry point restores the caller's current device before returning.
 * There is NO CPU fallback: every compute
 * entry point fails with CBX_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CBX_H_
#define CBX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBX_OK 0
#define CBX_ERR_ARG (-1)
#define CBX_ERR_CUDA (-2)
#define CBX_ERR_STATE (-3)
#define CBX_ERR_WORKSPACE (-4)

/* per-clip status bits written next to the embeddings */
#define CBX_CLIP_OK 0
#define CBX_CLIP_VE_TOO_SHORT 1   /* < 201 samples left after trim: librosa reflect-pad cannot be formed (melspec.py:57-64) */
#define CBX_CLIP_XV_TOO_SHORT 2   /* < 400 samples: Kaldi window does not fit (torchaudio kaldi.py:142-144) */
#define CBX_CLIP_VE_NAN 4         /* all-zero post-ReLU projection -> 0/0, as in voice_encoder.py:160 */

/* flags for cbx_embed* */
#define CBX_DO_VE 1        /* VoiceEncoder 256-d embedding */
#define CBX_DO_XV 2        /* CAMPPlus 192-d x-vector      */
#define CBX_NO_TRIM 4      /* trim_top_db=None (voice_encoder.py:266) */
#define CBX_PCM_PINNED 8   /* cbx_embed_host: pcm_host is page-locked already, copy straight from it */

typedef struct cbx_ctx cbx_ctx;

/* Integer plan of one clip of n_samples at 16 kHz (header formulas).
 * VoiceEncoder side: melspec.py:50 (T_ve), voice_encoder.py:54-66 (partials).
 * CAMPPlus side: torchaudio kaldi.py:63-67 (T_k), xvector.py:364-372 (stride-2 TDNN),
 * xvector.py:221-231 (100-frame CAM segments). */
typedef struct cbx_clip_plan {
  int64_t n_samples;
  int64_t ve_frames;     /* 1 + n/160                                   */
  int64_t ve_partials;   /* get_num_wins(ve_frames, step, min_coverage) */
  int64_t ve_target;     /* 160 + step*(partials-1)                      */
  int64_t xv_frames;     /* 1 + (n-400)/160, 0 if n < 400                */
  int64_t xv_tdnn;       /* (xv_frames-1)/2 + 1                          */
  int64_t xv_segments;   /* ceil(xv_tdnn/100)                            */
} cbx_clip_plan;

/* ---- integer host logic (no GPU needed) ------------------------------------------------ */
/* voice_encoder.py:69-81; rate <= 0 means rate=None (use overlap). Returns the step or <0. */
int cbx_ve_frame_step(double overlap, double rate);
/* voice_encoder.py:54-66 */
int cbx_ve_num_wins(int64_t n_frames, int step, double min_coverage, int64_t* n_wins, int64_t* target_n);
int cbx_plan_clip(int64_t n_samples, int step, double min_coverage, cbx_clip_plan* out);
/* librosa.effects.trim frame arithmetic used at voice_encoder.py:267: number of RMS frames. */
int64_t cbx_trim_num_frames(int64_t n_samples);
/* Cost model for rank balancing, in FLOP. */
double cbx_clip_cost(int64_t n_samples);
/* Ragged-batch scheduler (north_star (d)): deals n clips of n_samples[i] samples to `world` ranks
 * so that clip counts differ by at most one and the summed cbx_clip_cost is near equal (clips sorted by cost, most
 * expensive first, dealt out and back over the ranks).  rank_of[n] receives every clip's rank; row_of[n] (may be NULL)
 * its row inside that rank's shard, a shard keeping its clips in ascending clip order; rank_cost[world] (may be NULL)
 * the summed cost per rank.  The reference has no counterpart: its worker embeds one clip per message
 * (worker_redis.py:162). */
int cbx_partition(const int64_t* n_samples, int64_t n, int world, int32_t* rank_of, int64_t* row_of, double* rank_cost);

/* ---- context --------------------------------------------------------------------------- */
int cbx_create(int device, cbx_ctx** out);
void cbx_destroy(cbx_ctx* ctx);
const char* cbx_last_error(const cbx_ctx* ctx);   /* ctx may be NULL: last error of cbx_create */
const char* cbx_version(void);

/* Tuning.  key / values:
 * "mode": 1 (default) = tensor-core kernels (tcgen05, TF32 / 3xTF32; the fp32 / TF32 parity mode); 0 = strict fp32 SIMT kernels
 *   everywhere (the on-device fp32 yardstick); 2 = the bf16 mode (BASELINE config 5): mode 1 with cat_bf16 = 2, u_bf16 = 1 and
 *   xw_bf16 = 1 -- its own, looser tolerance (tests/test_gpu_parity.py MODE2_TOL).  Setting 0 / 1 switches the
 *   three bf16 options off again; cbx_get_option("mode") returns 2 only while all three are on.
 * "cat_bf16" (0 | 1 | 2): 1 = the D-TDNN bottleneck / transit GEMMs read a bf16 copy of the concatenation buffers that the
 *   producing epilogues write beside the fp32 one (activations rounded once to bf16; weights and MMAs stay TF32); 2 = those GEMMs
 *   also run on bf16 operands (kind::f16 MMAs, bf16 weight copies, fp32 accumulation).
 * "u_bf16" (0 | 1): the bottleneck output u is stored as bf16 only and the CAM local convolution runs on bf16 operands.
 * "xw_bf16" (0 | 1): the LSTM input projections are stored as bf16 (the recurrence reads them as such).
 * "batch_invariant": 1 = the x-vector of a clip is bit-identical whatever else is in the batch and however the call is chunked
 *   (exact warp-level segment sums, ~3 % slower; u_bf16 is ignored while it is set); 0 (default) = reproducible from run to run,
 *   position dependent within ~1e-4 (the VoiceEncoder embedding is batch invariant either way).
 * Scheduling / layout switches -- results are bit-identical with either value (tested), defaults are the fast setting:
 *   "overlap" (1): with both encoders requested, CAMPPlus runs on an internal second stream beside the VoiceEncoder chain (forked
 *   from / joined into the caller's stream); "lstm_late" (1): on that path the recurrence starts when the CAMPPlus chain enters its
 *   D-TDNN phase; "pdl" (1): the CAMPPlus convolution and dense-layer chains use programmatic dependent launch; "transit_n256" (1):
 *   transit GEMMs on 128 x 256 output tiles; "lstm_gate_warps" (4 | 2): gate warps per TMEM lane quadrant of the recurrence kernel;
 *   "bn_prefetch" (1): every CTA of the pre-activation GEMM ends its K loop with an L2 prefetch of the first K blocks of the row tile
 *   that the next CTA in its slot will take (a hint: no data reaches the SM).
 * Same arithmetic up to rounding order (tested against each other and the oracle): "fcm_fuse" (1): the identity residual blocks of the
 *   FCM head as one fused kernel; "dft_eo" (1): the VoiceEncoder / S3 front-end DFT in its even / odd form.
 * Chunking (results invariant): "xv_chunk_rows", "fcm_chunk_rows", "lstm_chunk_partials".
 * ("lstm_trace" takes a device pointer for the recurrence kernel's clock trace; "lstm_impl", "lstm_dbg" and "probe" exist only in the
 * CBX_DEV_TOOLS build of the library -- "probe" removes kernels from the chain to time what is left, results are WRONG while set.) */
int cbx_set_option(cbx_ctx* ctx, const char* key, int64_t value);
int64_t cbx_get_option(const cbx_ctx* ctx, const char* key);

/* ---- weights: reference state_dict tensors, fp32, host memory, C-contiguous ------------ */
/* Generic name -> tensor table (keys exactly as in the reference state_dict).
 * which = 0: VoiceEncoder (ve.safetensors keys), 1: CAMPPlus (speaker_encoder.* keys without prefix).
 * BatchNorm folding, gate re-layout and DFT/mel tables are built once here. */
int cbx_load_weights(cbx_ctx* ctx, int which, int n_tensors, const char* const* names,
                     const float* const* data, const int64_t* numel);

/* ---- batched embedding ------------------------------------------------------------------ */
/* Bytes of device workspace cbx_embed needs for these clips (lengths in samples) with this partial step. */
int64_t cbx_workspace_bytes(cbx_ctx* ctx, int n_clips, const int64_t* lengths, int ve_step, double min_coverage, int flags);

/* pcm_dev: all clips back to back, fp32 16 kHz mono in device memory; offsets_host[n_clips+1] (in
 * samples, host memory).  ve_out_dev [n_clips,256], xv_out_dev [n_clips,192], status_dev [n_clips]
 * int32 (CBX_CLIP_* bits) are device buffers (either output may be NULL if its flag is off).
 * trim_top_db as in embeds_from_wavs (ignored with CBX_NO_TRIM); ve_step / min_coverage as in
 * VoiceEncoder.inference (77 / 0.8 for the reference's callers). */
int cbx_embed(cbx_ctx* ctx, const float* pcm_dev, const int64_t* offsets_host, int n_clips,
              float trim_top_db, int ve_step, double min_coverage,
              float* ve_out_dev, float* xv_out_dev, int32_t* status_dev,
              void* workspace_dev, int64_t workspace_bytes, void* stream, int flags);

/* Same through HOST buffers: copies pcm host->device, runs, copies results back, synchronises.
 * The library owns the staging and workspace buffers (grown on demand). */
int cbx_embed_host(cbx_ctx* ctx, const float* pcm_host, const int64_t* offsets_host, int n_clips,
                   float trim_top_db, int ve_step, double min_coverage,
                   float* ve_out_host, float* xv_out_host, int32_t* status_host, int flags);

/* Streaming form for voice-bank extraction (many batches): two slots.  submit() enqueues the host->device copy (own
 * stream), the kernels (one compute stream shared by both slots) and the device->host copy of the results and returns
 * without waiting; wait() blocks until that slot's results are on the host and copies them out.  With two batches in
 * flight the copies of one overlap the kernels of the other.  pcm_host must stay valid (and, with CBX_PCM_PINNED,
 * unmodified) until the slot's wait() returns.  cbx_embed_host == submit(slot 0) + wait(slot 0). */
int cbx_embed_host_submit(cbx_ctx* ctx, int slot, const float* pcm_host, const int64_t* offsets_host, int n_clips,
                          float trim_top_db, int ve_step, double min_coverage, int flags);
int cbx_embed_host_wait(cbx_ctx* ctx, int slot, float* ve_out_host, float* xv_out_host, int32_t* status_host);

/* ---- resampling ("next" row 1) ------------------------------------------------------------------- */
/* torchaudio.transforms.Resample(src_sr, dst_sr) with its defaults, the resampler behind get_resampler()
 * (s3gen/s3gen.py:41-44, called at :116 and :175-183).  Ragged batch: clips back to back, offsets in samples (host
 * arrays of n_clips+1); out_offsets[i+1]-out_offsets[i] must equal cbx_resample_out_len(src, dst, len_i)
 * = ceil(len_i * dst / src) on the reduced ratio.  Stream ordered. */
int64_t cbx_resample_out_len(int src_sr, int dst_sr, int64_t n_samples);
int cbx_resample(cbx_ctx* ctx, const float* x_dev, const int64_t* in_offsets_host, int n_clips, int src_sr, int dst_sr,
                 float* y_dev, const int64_t* out_offsets_host, void* stream);

/* ---- S3Gen prompt mel ("next" row 1) ------------------------------------------------------------- */
/* mel_spectrogram() of s3gen/utils/mel.py:33-81 with its defaults (24 kHz, n_fft = win = 1920, hop 480, 80 Slaney mels
 * 0..8 kHz, reflect pad 720, sqrt(power + 1e-9), log(clamp(., 1e-5))): the prompt_feat of S3Token2Mel.embed_ref
 * (s3gen.py:177).  Ragged batch of 24 kHz clips back to back (offsets in samples, host array of n_clips+1); out_dev
 * is [sum_i frames_i][80] fp32, clip after clip, frames_i = cbx_prompt_mel_frames(len_i) = 1 + (len_i - 480) / 480 --
 * the (T, 80) layout embed_ref hands on after its transpose(1, 2).  Clips of <= 720 samples are refused like torch's
 * reflect pad refuses them.  Stream ordered. */
int64_t cbx_prompt_mel_frames(int64_t n_samples);
int cbx_prompt_mel(cbx_ctx* ctx, const float* pcm_dev, const int64_t* offsets_host, int n_clips, float* out_dev, void* stream);

/* ---- S3Tokenizer front-end ("next" row 3) ------------------------------------------------------------ */
/* S3Tokenizer.log_mel_spectrogram (s3tokenizer/s3tokenizer.py:128-168): torch.stft(400, 160, Hann, centred/reflect) with
 * the last frame dropped, power, 128 Slaney mels, log10(clamp(., 1e-10)), floor at (clip maximum - 8), (x + 4) / 4.
 * Ragged batch of 16 kHz clips back to back; out_dev holds one [128][T_i] block per clip, clip after clip,
 * T_i = cbx_s3_log_mel_frames(len_i) = len_i / 160 -- the [F, T] layout quantize() takes.  The maximum is per clip (the
 * reference calls the method one clip at a time, s3tokenizer.py:107-113).  Clips of <= 200 samples are refused like
 * torch.stft's reflect pad refuses them.  The tokenizer network itself (third-party s3tokenizer package) is not part of
 * this library.  Stream ordered. */
int64_t cbx_s3_log_mel_frames(int64_t n_samples);
int cbx_s3_log_mel(cbx_ctx* ctx, const float* pcm_dev, const int64_t* offsets_host, int n_clips, float* out_dev, void* stream);

/* ---- consumer projections ("next" row 4) ------------------------------------------------------------ */
/* y[i] = W . (normalize ? x[i] / max(||x[i]||, 1e-12) : x[i]) + b for n embedding rows: T3CondEnc.spkr_enc
 * Linear(256 -> 1024) on the VoiceEncoder embedding (t3/modules/cond_enc.py:50,70; normalize = 0) and F.normalize +
 * spk_embed_affine_layer Linear(192 -> 80) on the x-vector (s3gen/flow.py:252-253; normalize = 1).  x_dev [n][in_dim],
 * w_dev [out_dim][in_dim] (torch Linear layout), b_dev [out_dim] or NULL, y_dev [n][out_dim]; in_dim <= 1024. */
int cbx_project(cbx_ctx* ctx, const float* x_dev, int64_t n, int in_dim, const float* w_dev, const float* b_dev, int out_dim,
                int normalize, float* y_dev, void* stream);

/* VoiceEncoder.forward on already-cut partials (voice_encoder.py:139-160):
 * mels_dev [n_partials,160,40] -> out_dev [n_partials,256] (L2-normed). */
int cbx_ve_forward_partials(cbx_ctx* ctx, const float* mels_dev, int n_partials, float* out_dev,
                            void* workspace_dev, int64_t workspace_bytes, void* stream);
int64_t cbx_ve_forward_workspace_bytes(cbx_ctx* ctx, int n_partials);

/* CAMPPlus.forward on precomputed features (s3gen/xvector.py:417-423; what CAMPPlus.inference hands it after
 * extract_feature): feats_dev [sum T_i][80] fp32 (already mean-normalised, clip i = rows frame_offsets_host[i] ..
 * frame_offsets_host[i+1]) -> xv_out_dev [n_clips][192].  Skips the fbank / CMN kernels; everything after them is the
 * path of cbx_embed(CBX_DO_XV).  status_dev [n_clips] as in cbx_embed (T_i = 0 sets CBX_CLIP_XV_TOO_SHORT). */
int cbx_campplus_forward_feats(cbx_ctx* ctx, const float* feats_dev, const int64_t* frame_offsets_host, int n_clips,
                               float* xv_out_dev, int32_t* status_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);
int64_t cbx_campplus_forward_workspace_bytes(cbx_ctx* ctx, const int64_t* frame_offsets_host, int n_clips);

/* ---- stage taps for parity tests --------------------------------------------------------- */
/* After cbx_embed (single chunk), locate an intermediate inside the caller's workspace.
 * name in {"ve_trim","ve_mel","ve_partial_emb","xv_fbank","xv_fcm","xv_cat1","xv_cat2","xv_cat3","xv_stats",...};
 * returns byte offset, row count, row length (floats) and leading dimension (floats). */
int cbx_locate(cbx_ctx* ctx, const char* name, int64_t* byte_offset, int64_t* rows, int64_t* cols, int64_t* ld);
/* Row offsets of clip i inside the stage buffers of the last cbx_embed call. */
int cbx_clip_rows(cbx_ctx* ctx, int clip, int64_t* ve_mel_row, int64_t* ve_partial_slot,
                  int64_t* xv_fbank_row, int64_t* xv_tdnn_row);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t cbx_launch_count(const cbx_ctx* ctx);

/* Per-kernel timing: with profiling on, every launch is bracketed by CUDA events on its stream.
 * cbx_profile_report synchronises the device and writes one line per kernel family:
 *   "<tag> <launches> <total_ms> <flops> <bytes>\n"; returns the bytes needed (call with buf=NULL to size). */
int cbx_profile_enable(cbx_ctx* ctx, int on);
int64_t cbx_profile_report(cbx_ctx* ctx, char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* CBX_H_ */

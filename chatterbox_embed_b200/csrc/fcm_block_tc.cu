// FCM head of CAMPPlus (xvector.py:61-91, BasicResBlock): an identity residual block
//     out = relu( bn2(conv2( relu(bn1(conv1(x))) )) + x )          (two 3x3 convolutions, 32 -> 32 channels, stride 1)
// as ONE persistent kernel.  The intermediate activation never goes to HBM: the block reads its input once and writes its
// output once (10 KB per frame at F = 40 instead of the 25 KB the two separate convolution kernels of fcm_tc.cu move: the
// intermediate's write and read, and the residual's second read of the input).
//
// A CTA owns a contiguous range of time tiles (BR rows x (F + 2) frequency positions = one M = 128 accumulator) and streams
// along it.  Per step it issues conv 1 of tile k + 2 and conv 2 of tile k:
//   * conv 1 is the implicit GEMM of fcm_tc.cu: ONE halo block of the input per tile by TMA (zero padding = out-of-bounds
//     fill), nine taps = nine row-shifted UMMA descriptors of that block;
//   * its epilogue group (bias, ReLU, guard-row mask, tf32 rounding) writes the tile into a RING of four tile slots in shared
//     memory, in exactly the layout conv 2 wants as ITS halo block: position (row, f + 1), with the two padding positions of
//     every row written as zeros, 128-byte swizzled rows.  The ring has one mirrored row in front of slot 0 (a copy of the last
//     row of slot 3) and one behind slot 3 (a copy of the first row of slot 0), so the three time taps of conv 2 stay
//     row-shifted views of CONTIGUOUS shared memory across the wrap;
//   * conv 2 of tile k reads rows [k BR - 1, k BR + BR + 1) of the ring: tiles k - 1 (last row), k, k + 1 (first row) -- conv 1
//     runs two tiles ahead so that its epilogue's latency (TMEM -> registers -> shared memory, ~1 us) is off the tensor pipe's
//     critical path; a slot is rewritten only after the MMAs that read it have completed, which the in-order tensor pipe
//     guarantees by the time the accumulator of the tile that overwrites it is committed;
//   * the second epilogue group adds the residual (the block's input, prefetched from global memory -- an L2 hit, the CTA
//     loaded the same rows two tiles earlier -- before it waits for the accumulator), bias, ReLU, mask, and stores through a
//     staging tile and one TMA tensor store.
// The first and last tile of a CTA's range also need the intermediate of the neighbouring range's edge tile: conv 1 is simply
// run for one extra tile on either side (2 of ~580 tiles per CTA).
#include <algorithm>

#include "cbx_internal.h"
#include "tc.cuh"

namespace cbx {
namespace fcmb {

using namespace tc;

constexpr int W_TAPS = 9;
constexpr int W_BYTES = W_TAPS * 32 * 128;         // 36 KB per convolution
constexpr int OUT_BYTES = 128 * 128;               // output staging tile
constexpr int SLOTS = 4;
constexpr int MAX_STAGES = 3;

struct Params {
  int F, pitch, BR, rows, ntiles, per_cta, nstages;
  uint32_t stage_bytes, ring_bytes, tail_bytes;
  const float* bias1; const float* bias2; const float* res; const int32_t* row_clip;
};

__global__ void __launch_bounds__(320, 1)
fcm_block_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + W_BYTES;
  uint8_t* sOut = sW2 + W_BYTES;
  uint8_t* sRing = sOut + OUT_BYTES;                    // [1 + SLOTS * BR + 1 (+ slack)][pitch][128 B]; physical row 0 = logical row -1
  uint8_t* sIn = sRing + p.ring_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sIn + p.nstages * p.stage_bytes + p.tail_bytes);    // tail: over-read slack of the last stage
  uint64_t* full = bars;                    // [MAX_STAGES] input stage landed
  uint64_t* empty = full + MAX_STAGES;      // [MAX_STAGES] conv-1 MMAs of the stage completed
  uint64_t* a1full = empty + MAX_STAGES;    // [2] conv-1 accumulator ready
  uint64_t* a1empty = a1full + 2;           // [2] drained (4 warp arrivals)
  uint64_t* a2full = a1empty + 2;           // [2]
  uint64_t* a2empty = a2full + 2;           // [2]
  uint64_t* midfull = a2empty + 2;          // [SLOTS] ring slot written (128 arrivals)
  uint64_t* wfull = midfull + SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's tiles: [t_lo, t_hi); conv 1 runs over [t_lo - 1, t_hi + 1)
  const int t_lo = blockIdx.x * p.per_cta;
  const int t_hi = min(t_lo + p.per_cta, p.ntiles);
  const int n1 = t_hi - t_lo + 2;            // conv-1 tiles (local index j = 0 .. n1 - 1 is global tile t_lo - 1 + j)
  const int pitch = p.pitch, BR = p.BR;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmOut);
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&a1full[a], 1); mbar_init(&a1empty[a], 4); mbar_init(&a2full[a], 1); mbar_init(&a2empty[a], 4); }
    for (int s = 0; s < SLOTS; ++s) mbar_init(&midfull[s], 128);
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (t_lo >= t_hi) {                        // (grid is sized so that this does not happen; keep the exit collective)
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
    return;
  }

  if (warp == 0) {
    // ===== TMA producer: both weight sets once, then one input halo block per conv-1 tile
    if (lane == 0) {
      mbar_expect_tx(wfull, 2 * W_BYTES);
      for (int t = 0; t < W_TAPS; ++t) { tma_load_2d(sW1 + t * 4096, &tmW1, wfull, t * 32, 0); tma_load_2d(sW2 + t * 4096, &tmW2, wfull, t * 32, 0); }
      pdl_wait();                            // the input comes from the kernel in front (the weights above do not)
      for (int j = 0; j < n1; ++j) {
        const int s = j % p.nstages, ph = (j / p.nstages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], (uint32_t)((BR + 2) * pitch * 128));
        // tensor {32, 1, F, prows} starting one pad row before row 0: row coordinate r + 1 - 1 = first halo row of the tile
        tma_load_4d(sIn + s * p.stage_bytes, &tmIn, &full[s], 0, 0, -1, (t_lo - 1 + j) * BR);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp in the loop, one elected lane issues): conv 1 of local tile j + 1, then conv 2 of local tile j - 1
    constexpr uint32_t idesc = make_idesc_tf32(128, 32);
    uint32_t aoff[W_TAPS];                   // tap offsets in 16-byte units: (kw * pitch + kh) positions of 128 bytes
#pragma unroll
    for (int t = 0; t < W_TAPS; ++t) aoff[t] = (uint32_t)(((t % 3) * pitch + (t / 3)) * 128) >> 4;     // t = kh * 3 + kw
    const uint64_t dhi = make_desc_sw128(0);
    const uint32_t w1 = smem_u32(sW1) >> 4, w2 = smem_u32(sW2) >> 4, ring16 = smem_u32(sRing) >> 4;
    mbar_wait(wfull, 0);
    auto conv1 = [&](int j) {
      const int s = j % p.nstages, ph = (j / p.nstages) & 1, a = j & 1, pa = (j >> 1) & 1;
      mbar_wait(&a1empty[a], pa ^ 1);
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t d = tmem_base + a * 32;
      const uint32_t in16 = smem_u32(sIn + s * p.stage_bytes) >> 4;
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < W_TAPS; ++t) {
          const uint64_t ad = dhi | (uint64_t)((in16 + aoff[t]) & 0x3FFFu);
          const uint64_t bd = dhi | (uint64_t)((w1 + t * 256) & 0x3FFFu);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (t | k) != 0);
        }
        umma_commit(&empty[s]);
        umma_commit(&a1full[a]);
      }
      __syncwarp();
    };
    auto conv2 = [&](int j) {                // local tile j (1 .. n1 - 2) = ring slot j % SLOTS; needs the intermediate of tiles j - 1 .. j + 1
      const int a = (j - 1) & 1, pa = ((j - 1) >> 1) & 1;        // accumulator 2 is first used by tile 1: the phase counts from there
      mbar_wait(&a2empty[a], pa ^ 1);
      mbar_wait(&midfull[(j + 1) % SLOTS], ((j + 1) / SLOTS) & 1);      // epilogue 1 works in order: tiles j - 1 and j are in place too
      tc_fence_after();
      const uint32_t d = tmem_base + 64 + a * 32;
      // first halo row = logical row (j % SLOTS) * BR - 1 = physical row (j % SLOTS) * BR
      const uint32_t r16 = ring16 + (uint32_t)(((j % SLOTS) * BR * pitch * 128) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < W_TAPS; ++t) {
          const uint64_t ad = dhi | (uint64_t)((r16 + aoff[t]) & 0x3FFFu);
          const uint64_t bd = dhi | (uint64_t)((w2 + t * 256) & 0x3FFFu);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (t | k) != 0);
        }
        umma_commit(&a2full[a]);
      }
      __syncwarp();
    };
    // conv 1 runs TWO tiles ahead of conv 2: conv 2 of tile j needs the intermediate of tile j + 1, whose epilogue (~1 us) must not
    // sit between two consecutive entries of the tensor pipe's queue
    conv1(0);
    conv1(1);
    if (n1 > 2) conv1(2);
    for (int j = 1; j <= n1 - 2; ++j) {
      if (j + 2 < n1) conv1(j + 2);
      conv2(j);
    }
  } else if (warp < 6) {
    // ===== epilogue 1: conv-1 accumulator -> bias, ReLU, guard-row mask, tf32 -> ring slot (position (t, f + 1); zero padding positions)
    const int q = warp & 3;
    const int i = q * 32 + lane;             // accumulator row = position (t, f') of the tile
    const int t = i / pitch, fr = i - t * pitch;
    float bias[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bias[c] = __ldg(p.bias1 + c);
    const bool in_tile = t < BR;
    // where this thread's 128 bytes go inside a slot: f' < F -> (t, f' + 1); f' = F -> (t, F + 1) zero; f' = F + 1 -> (t, 0) zero
    const int col = fr < p.F ? fr + 1 : (fr == p.F ? p.F + 1 : 0);
    const bool is_pad = fr >= p.F;
    pdl_wait();
    for (int j = 0; j < n1; ++j) {
      const int a = j & 1, pa = (j >> 1) & 1;
      const int slot = j % SLOTS;
      mbar_wait(&a1full[a], pa);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * 32, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a1empty[a]);
      const int row = (t_lo - 1 + j) * BR + t;                          // global time row of this position
      const bool live = in_tile && !is_pad && row >= 0 && row < p.rows && p.row_clip[min(max(row, 0), p.rows - 1)] >= 0;
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = live ? to_tf32(fmaxf(v[c] + bias[c], 0.f)) : 0.f;
      if (in_tile) {
        const int lrow = slot * BR + t;                                 // logical ring row
        auto put = [&](int prow) {
          const uint32_t pos = (uint32_t)(prow * pitch + col);
          float4* dst = reinterpret_cast<float4*>(sRing) + (size_t)pos * 8;
#pragma unroll
          for (int c = 0; c < 8; ++c) dst[c ^ (pos & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        };
        put(lrow + 1);
        if (slot == SLOTS - 1 && t == BR - 1) put(0);                   // mirror: logical row -1
        if (slot == 0 && t == 0) put(SLOTS * BR + 1);                   // mirror: logical row SLOTS * BR
      }
      fence_proxy_async();
      mbar_arrive(&midfull[slot]);
    }
  } else {
    // ===== epilogue 2: conv-2 accumulator + residual + bias -> ReLU, mask -> staging tile -> one TMA tensor store per tile
    const int q = warp & 3;
    const int i = q * 32 + lane;
    const int t = i / pitch, fr = i - t * pitch;
    float bias[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bias[c] = __ldg(p.bias2 + c);
    float4* so = reinterpret_cast<float4*>(sOut) + i * 8;
    const bool issuer = warp == 6 && lane == 0;
    pdl_wait();
    for (int j = 1; j <= n1 - 2; ++j) {
      const int a = (j - 1) & 1, pa = ((j - 1) >> 1) & 1;
      const int tile = t_lo - 1 + j;
      const int row = tile * BR + t;
      const bool has = t < BR && fr < p.F && row < p.rows;
      float4 r[8];
      if (has) {
        const float4* rp = reinterpret_cast<const float4*>(p.res + ((size_t)row * p.F + fr) * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) r[c] = __ldg(rp + c);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) r[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(&a2full[a], pa);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 64 + a * 32, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a2empty[a]);
      const bool live = has && p.row_clip[min(row, p.rows - 1)] >= 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        v[4 * c] = live ? fmaxf(v[4 * c] + r[c].x + bias[4 * c], 0.f) : 0.f;
        v[4 * c + 1] = live ? fmaxf(v[4 * c + 1] + r[c].y + bias[4 * c + 1], 0.f) : 0.f;
        v[4 * c + 2] = live ? fmaxf(v[4 * c + 2] + r[c].z + bias[4 * c + 2], 0.f) : 0.f;
        v[4 * c + 3] = live ? fmaxf(v[4 * c + 3] + r[c].w + bias[4 * c + 3], 0.f) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");                    // the previous store has read the staging tile (its issuer waited)
#pragma unroll
      for (int c = 0; c < 8; ++c) so[c ^ (i & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      fence_proxy_async();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (issuer) {
        tma_store_3d(&tmOut, sOut, 0, 0, tile * BR);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    pdl_trigger();
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// [prows][F][32] fp32 with one pad row in front, viewed as {32, 1, F, prows}; box {32, 1, pitch, nrows} (see fcm_tc.cu)
static CUtensorMap make_map_in(const float* base, int prows, int F, int pitch, int nrows) {
  CUtensorMap m;
  memset(&m, 0, sizeof m);
  cuuint64_t dims[4] = {32, 1, (cuuint64_t)F, (cuuint64_t)prows};
  cuuint64_t strides[3] = {128, (cuuint64_t)128, (cuuint64_t)128 * F};
  cuuint32_t box[4] = {32, 1, (cuuint32_t)pitch, (cuuint32_t)nrows};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  CUresult r = fn ? fn(&m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                  : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled(fcm block in) failed (%d) prows=%d F=%d pitch=%d nrows=%d\n", (int)r, prows, F, pitch, nrows);
  return m;
}

static CUtensorMap make_map_out(const float* base, int rows, int F, int pitch, int BR) {
  CUtensorMap m;
  memset(&m, 0, sizeof m);
  cuuint64_t dims[3] = {32, (cuuint64_t)F, (cuuint64_t)rows};
  cuuint64_t strides[2] = {128, (cuuint64_t)128 * F};
  cuuint32_t box[3] = {32, (cuuint32_t)pitch, (cuuint32_t)BR};
  cuuint32_t estr[3] = {1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  CUresult r = fn ? fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                  : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled(fcm block out) failed (%d) rows=%d F=%d pitch=%d BR=%d\n", (int)r, rows, F, pitch, BR);
  return m;
}

}  // namespace fcmb

// One identity residual block: in [rows][F][32] (one pad row in front and behind, prows rows in all) -> out [rows][F][32].
void run_fcm_block_tc(cbx_ctx* c, cudaStream_t st, const CUtensorMap& tmW1, const float* bias1, const CUtensorMap& tmW2, const float* bias2,
                      const float* in, int F, float* out, const int32_t* row_clip, int rows, int prows, const char* tag, bool pdl) {
  using namespace fcmb;
  Params p{};
  const int pitch = F + 2;
  const int BR = 128 / pitch;
  p.F = F; p.pitch = pitch; p.BR = BR; p.rows = rows; p.ntiles = (rows + BR - 1) / BR;
  p.bias1 = bias1; p.bias2 = bias2; p.res = in; p.row_clip = row_clip;
  auto align1k = [](uint32_t x) { return (x + 1023u) & ~1023u; };
  p.stage_bytes = align1k((uint32_t)((BR + 2) * pitch * 128));
  // An A operand is always 128 positions: with BR * pitch < 128 the junk accumulator rows read past the halo block (their results
  // are dropped).  The deepest tap starts at position 2 * pitch + 2, so a block of (BR + 2) * pitch positions is over-read by
  // `over` bytes: the ring is sized for the last slot's reads, and the input stages are followed by that much slack.
  const uint32_t over = (uint32_t)std::max(0, (2 * pitch + 2 + 128) - (BR + 2) * pitch) * 128u;
  p.ring_bytes = align1k((uint32_t)((SLOTS * BR + 2) * pitch * 128) + over);
  const int tail = (int)align1k(over + 128);
  constexpr int kSmemMax = 227 * 1024;
  p.nstages = MAX_STAGES;
  p.tail_bytes = (uint32_t)tail;
  auto need = [&]() { return 2 * W_BYTES + OUT_BYTES + (int)p.ring_bytes + p.nstages * (int)p.stage_bytes + tail + 256 + 1024; };
  if (need() > kSmemMax) p.nstages = 2;
  const int smem = need();
  if (smem > kSmemMax) { fprintf(stderr, "libcbx: fcm block kernel does not fit shared memory (F=%d)\n", F); return; }
  ensure_max_smem(fcm_block_kernel, kSmemMax);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  p.per_cta = (p.ntiles + nsm - 1) / nsm;
  const int grid = (p.ntiles + p.per_cta - 1) / p.per_cta;
  CUtensorMap tmIn = make_map_in(in - (size_t)F * kFcmC, prows, F, pitch, BR + 2);
  CUtensorMap tmOut = make_map_out(out, rows, F, pitch, BR);
  // algorithmic bytes: the input once, the output once (the residual is the input); FLOPs of both convolutions
  Scope scp(c->launches, st, tag, 2.0 * 2.0 * rows * F * kFcmC * 9 * kFcmC, 128.0 * rows * (F + F));
  tc::launch_pdl(fcm_block_kernel, dim3(grid), dim3(320), smem, st, pdl, tmIn, tmW1, tmW2, tmOut, p);
}

}  // namespace cbx

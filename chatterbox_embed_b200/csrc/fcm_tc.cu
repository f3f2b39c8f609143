// FCM head of CAMPPlus (xvector.py:61-127): the 3x3 convolutions as implicit GEMMs on tcgen05, one persistent kernel per
// convolution.
//
// Activations are [time row][F][32 channels] fp32.  A CTA works on tiles of BR time rows.  For every tile ONE halo block of
// the input -- (BR+2) rows x (F+2) frequencies x 32 channels, zero padding supplied by TMA out-of-bounds fill -- is brought
// into shared memory (128B-swizzled rows of 32 channels); the nine taps are then nine tcgen05.mma groups whose A
// descriptors are ROW-SHIFTED VIEWS of that one block (a shift of kw*(F+2)+kh rows), so the input crosses L2->SM once instead
// of nine times.  The GEMM's M index therefore runs over (t, f') with f' in [0, F+2): the two extra positions per time row
// are computed and dropped.  Frequency-stride-2 convs read two parity planes ({32, parity, F/2, rows} view of the tensor) so
// the stride disappears into the TMA coordinates; a residual block's 1x1 stride-2 shortcut conv is a tenth tap on its own
// plane.  An identity residual tile comes by its own tensor load into a buffer of the epilogue group, with its own mbarrier that
// the epilogue threads themselves wait on.  Weights ([32][taps*32], BN folded) stay resident in shared memory for the CTA's lifetime; the fp32 accumulator
// [128 x 32] is double-buffered in TMEM so the epilogue (bias, residual, ReLU, guard-row mask, 128-byte stores) of one tile
// overlaps the MMAs of the next.  The epilogue is a latency-bound dependent chain (TMEM load, residual, bias, ReLU, staging,
// barrier, TMA store: ~1 us per tile for one warp per SMSP), and it -- not HBM or the tensor pipe -- bounded this kernel, so
// TWO epilogue groups of four warps work on alternate tiles, each on its own accumulator and staging buffer.
#include "cbx_internal.h"
#include "tc.cuh"

namespace cbx {
namespace fcm {

using namespace tc;

constexpr int STAGES = 3;                         // at most; a conv whose stage is too large for three runs with two (Params::nstages)
constexpr int MAX_TAPS = 10;
constexpr int W_BYTES = MAX_TAPS * 32 * 128;       // 40 KB: [tap][32 out channels][32 in channels]
constexpr int RES_BYTES = 128 * 128;               // one residual tile (BR rows x F_out positions x 128 B <= 16 KB)
constexpr int OUT_BYTES = 2 * 128 * 128;           // 2 x 16 KB: one output staging tile per epilogue group

struct Plane { int par, f0, dr, nrows; uint32_t bytes, offset; };   // TMA box {32, 1, pitch, nrows} at (0, par, f0, row + dr)
struct Tap { int plane, start; };                  // A rows start at row `start` of the plane
struct Params {
  int nplanes, ntaps;
  Plane plane[3];
  Tap tap[MAX_TAPS];
  int pitch, BR, F_out, rows, row_base, ntiles, nstages, ngroups;
  uint32_t stage_bytes;
  const float* bias; const float* res; float* out; const int32_t* row_clip;
};

__global__ void __launch_bounds__(320, 1)
fcm_conv_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
                const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sOut = smem + W_BYTES;                       // [2][128 positions][128 B] output staging (swizzled), one tile per group
  uint8_t* sIn = sOut + p.ngroups * (OUT_BYTES / 2);
  uint8_t* sRes = sIn + p.nstages * p.stage_bytes + 2048;       // [2 groups][2][16 KB] residual tiles (only when p.res)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRes + (p.res ? 4 * RES_BYTES : 0));
  uint64_t* full = bars;              // [STAGES]
  uint64_t* empty = bars + STAGES;    // [STAGES]
  uint64_t* tfull = empty + STAGES;   // [2] accumulator ready
  uint64_t* tempty = tfull + 2;       // [2] accumulator drained (4 warp arrivals)
  uint64_t* wfull = tempty + 2;
  uint64_t* rfull = wfull + 1;        // [2 groups][2] residual tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm0); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }     // a stage is free when its MMAs have completed
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_init(wfull, 1);
    for (int r = 0; r < 4; ++r) mbar_init(&rfull[r], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, p.ntaps * 32 * 128);
      for (int t = 0; t < p.ntaps; ++t) tma_load_2d(sW + t * 4096, &tmW, wfull, t * 32, 0);
      pdl_wait();                   // the activations come from the convolution in front (the weights above do not)
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int s = it % p.nstages, ph = (it / p.nstages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint32_t bytes = 0;
        for (int q = 0; q < p.nplanes; ++q) bytes += p.plane[q].bytes;
        mbar_expect_tx(&full[s], bytes);
        const int r0 = p.row_base + tile * p.BR;
        for (int q = 0; q < p.nplanes; ++q) {
          const Plane& pl = p.plane[q];
          const CUtensorMap* tm = q == 0 ? &tm0 : (q == 1 ? &tm1 : &tm2);
          tma_load_4d(sIn + s * p.stage_bytes + pl.offset, tm, &full[s], 0, pl.par, pl.f0, r0 + pl.dr);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop (waits included), one elected lane issues
    constexpr uint32_t idesc = make_idesc_tf32(128, 32);
    // per-tap A offsets (16-byte units) once, in registers: the issue loop below is then adds and MMAs only -- with N = 32
    // an MMA occupies the tensor pipe for ~16 cycles, so descriptor arithmetic on the issuing thread is what would bound it
    uint32_t aoff[MAX_TAPS];
#pragma unroll
    for (int t = 0; t < MAX_TAPS; ++t) aoff[t] = t < p.ntaps ? (p.plane[p.tap[t].plane].offset + (uint32_t)p.tap[t].start * 128u) >> 4 : 0u;
    const uint64_t dhi = make_desc_sw128(0);                      // descriptor bits other than the start address
    const uint32_t w16 = smem_u32(sW) >> 4;
    const int ntaps = p.ntaps;
    mbar_wait(wfull, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.nstages, ph = (it / p.nstages) & 1;
      const int a = it & 1, pa = (it >> 1) & 1;
      mbar_wait(&tempty[a], pa ^ 1);
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t d = tmem_base + a * 32;
      const uint32_t in16 = smem_u32(sIn + s * p.stage_bytes) >> 4;
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < MAX_TAPS; ++t) {
          if (t < ntaps) {
            const uint64_t ad = dhi | (uint64_t)((in16 + aoff[t]) & 0x3FFFu);
            const uint64_t bd = dhi | (uint64_t)((w16 + t * 256) & 0x3FFFu);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (t | k) != 0);
          }
        }
        umma_commit(&empty[s]);
        umma_commit(&tfull[a]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: thread = output position i = (t, f') of the tile (f' >= F_out are the padding positions).  The output
    // tile leaves through a staging buffer and ONE tensor store whose box is clipped by the tensor's bounds (padding
    // positions / rows past the end).  (An earlier version brought the residual tile in as one more TMA plane of the input
    // stage, read after the accumulator barrier only; its results were not reproducible from run to run -- tests/tools/determinism.py.)
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;                    // epilogue group = accumulator = parity of the CTA's tile counter
    const int i = q * 32 + lane;
    const int t = i / p.pitch;
    float bias[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bias[c] = __ldg(p.bias + c);
    uint8_t* const stage_out = sOut + grp * (128 * 128);
    float4* so = reinterpret_cast<float4*>(stage_out) + i * 8;
    const bool issuer = (warp == 2 || warp == 6) && lane == 0;
    pdl_wait();                     // the output buffer may still be an input of the kernel in front
    const int ng = p.ngroups;       // 2, or 1 when three input stages only fit beside ONE staging tile (group 1 then idles)
    // Identity residual: the group's issuer brings the tile's [BR][F_out][32] block in with ONE tensor load (128B swizzle,
    // box clipped at the last row) into the group's own pair of buffers, one tile ahead, and every epilogue thread waits on
    // that load's own mbarrier before it reads its 128-byte row.
    const int fr = i - t * p.pitch;
    uint8_t* const res_buf = sRes + grp * (2 * RES_BYTES);
    uint64_t* const res_bar = rfull + grp * 2;
    auto res_issue = [&](int tile, int n) {              // n = ordinal of the tile within this group
      if (p.res != nullptr && issuer && tile < p.ntiles) {
        fence_proxy_async();                              // the buffer's previous readers are behind a named barrier
        mbar_expect_tx(&res_bar[n & 1], (uint32_t)p.BR * (uint32_t)p.F_out * 128u);
        tma_load_3d(res_buf + (n & 1) * RES_BYTES, &tmRes, &res_bar[n & 1], 0, 0, tile * p.BR);
      }
    };
    if (grp < ng) res_issue(blockIdx.x + grp * gridDim.x, 0);
    int nloc = 0;
    for (int it = grp, tile = blockIdx.x + grp * gridDim.x; grp < ng && tile < p.ntiles; tile += ng * gridDim.x, it += ng, ++nloc) {
      const int a = it & 1, pa = (it >> 1) & 1;
      const bool has_res = p.res != nullptr && fr < p.F_out && t < p.BR && tile * p.BR + t < p.rows;
      res_issue(tile + ng * gridDim.x, nloc + 1);
      mbar_wait(&tfull[a], pa);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * 32, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      if (p.res != nullptr) {
        mbar_wait(&res_bar[nloc & 1], (nloc >> 1) & 1);
        if (has_res) {
          const int pos = t * p.F_out + fr;               // row of the [BR * F_out][128 B] swizzled block
          const float4* rr = reinterpret_cast<const float4*>(res_buf + (nloc & 1) * RES_BYTES) + pos * 8;
#pragma unroll
          for (int c = 0; c < 8; ++c) { const float4 x = rr[c ^ (pos & 7)]; v[4 * c] += x.x; v[4 * c + 1] += x.y; v[4 * c + 2] += x.z; v[4 * c + 3] += x.w; }
        }
      }
      const int row = min(tile * p.BR + t, p.rows - 1);
      const bool live = p.row_clip[row] >= 0;
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = live ? fmaxf(v[c] + bias[c], 0.f) : 0.f;
      // the group's previous store has read the staging buffer (its issuer waited for that before arriving here)
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 8; ++c) so[c ^ (i & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      fence_proxy_async();
      if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
      if (issuer) {
        tma_store_3d(&tmOut, stage_out, 0, 0, tile * p.BR);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    pdl_trigger();                  // the next convolution may load its weights and set up
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// [prows][F][32] fp32 viewed as {32, P, F/P, prows}; box {32, 1, pitch, nrows}
static CUtensorMap make_map(const float* base, int prows, int F, int P, int pitch, int nrows) {
  CUtensorMap m;
  memset(&m, 0, sizeof m);
  cuuint64_t dims[4] = {32, (cuuint64_t)P, (cuuint64_t)(F / P), (cuuint64_t)prows};
  cuuint64_t strides[3] = {128, (cuuint64_t)128 * P, (cuuint64_t)128 * F};
  cuuint32_t box[4] = {32, 1, (cuuint32_t)pitch, (cuuint32_t)nrows};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  CUresult r = fn ? fn(&m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                  : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled(fcm2) failed (%d) prows=%d F=%d P=%d pitch=%d nrows=%d\n", (int)r, prows, F, P, pitch, nrows);
  return m;
}

// [rows][F][32] fp32 output / residual tensor as {32, F, rows}; box {32, pitch, BR}: a store clips f >= F and rows past the end
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
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled(fcm out) failed (%d) rows=%d F=%d pitch=%d BR=%d\n", (int)r, rows, F, pitch, BR);
  return m;
}

}  // namespace fcm

// in: [rows][F_in][32] with one pad row in front (in points at row 0, the map starts one row earlier); prows = rows of the
// buffer including both pad rows; sf = frequency stride; sc: optional 1x1 stride-2 shortcut source [rows][2*F_out][32].
void run_fcm_conv_tc(cbx_ctx* c, cudaStream_t st, const CUtensorMap& tmW, const float* bias, const float* in, int F_in, int F_out, int sf,
                     const float* sc, int F_sc, const float* res, float* out, const int32_t* row_clip, int rows, int prows, double flops, const char* tag, bool pdl) {
  using namespace fcm;
  Params p{};
  const int pitch = sf == 1 ? F_out + 2 : F_out + 1;
  const int BR = 128 / pitch;
  p.pitch = pitch; p.BR = BR; p.F_out = F_out; p.rows = rows; p.row_base = 1; p.ntiles = (rows + BR - 1) / BR;
  p.bias = bias; p.res = res; p.out = out; p.row_clip = row_clip;
  auto plane_bytes = [&](int nrows) { return (uint32_t)(nrows * pitch * 128); };
  auto align1k = [](uint32_t x) { return (x + 1023u) & ~1023u; };
  CUtensorMap tm[3];
  uint32_t off = 0;
  int np = 0;
  if (sf == 1) {
    p.plane[np] = Plane{0, -1, -1, BR + 2, plane_bytes(BR + 2), off};
    tm[np] = make_map(in - (size_t)F_in * kFcmC, prows, F_in, 1, pitch, BR + 2);
    off += align1k(p.plane[np].bytes); ++np;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) p.tap[kh * 3 + kw] = Tap{0, kw * pitch + kh};
  } else {
    p.plane[np] = Plane{0, 0, -1, BR + 2, plane_bytes(BR + 2), off};          // even input frequencies 2f
    tm[np] = make_map(in - (size_t)F_in * kFcmC, prows, F_in, 2, pitch, BR + 2);
    off += align1k(p.plane[np].bytes); ++np;
    p.plane[np] = Plane{1, -1, -1, BR + 2, plane_bytes(BR + 2), off};         // odd input frequencies 2f-1 (index f-1)
    tm[np] = tm[0];
    off += align1k(p.plane[np].bytes); ++np;
    for (int kw = 0; kw < 3; ++kw) {
      p.tap[0 * 3 + kw] = Tap{1, kw * pitch};          // kh = 0: f_in = 2f - 1
      p.tap[1 * 3 + kw] = Tap{0, kw * pitch};          // kh = 1: f_in = 2f
      p.tap[2 * 3 + kw] = Tap{1, kw * pitch + 1};      // kh = 2: f_in = 2f + 1
    }
  }
  p.ntaps = 9;
  if (sc) {
    p.plane[np] = Plane{0, 0, 0, BR, plane_bytes(BR), off};
    tm[np] = make_map(sc - (size_t)F_sc * kFcmC, prows, F_sc, 2, pitch, BR);
    off += align1k(p.plane[np].bytes);
    p.tap[9] = Tap{np, 0};
    p.ntaps = 10; ++np;
  }
  CUtensorMap tmOut = make_map_out(out, rows, F_out, pitch, BR);
  CUtensorMap tmRes = res ? make_map_out(res, rows, F_out, F_out, BR) : tmOut;      // box {32, F_out, BR}: positions t * F_out + f
  for (int q = np; q < 3; ++q) tm[q] = tm[0];
  p.nplanes = np;
  p.stage_bytes = off;
  constexpr int kSmemMax = 227 * 1024;
  p.nstages = STAGES; p.ngroups = 2;
  auto need = [&]() { return W_BYTES + p.ngroups * (OUT_BYTES / 2) + p.nstages * (int)off + 2048 + (res ? 4 * RES_BYTES : 0) + 1024 + 256; };
  if (need() > kSmemMax) p.ngroups = 1;             // the stride-2 convs of layer 1: two 26 KB parity planes per stage
  if (need() > kSmemMax) p.nstages = 2;
  const int smem = need();
  ensure_max_smem(fcm_conv_kernel, kSmemMax);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  const int grid = p.ntiles < nsm ? p.ntiles : nsm;
  Scope scp(c->launches, st, tag, flops, 128.0 * rows * (F_in + F_out + (sc ? F_out : 0) + (res ? F_out : 0)));   // in + out (+ shortcut / residual)
  tc::launch_pdl(fcm_conv_kernel, dim3(grid), dim3(320), smem, st, pdl, tm[0], tm[1], tm[2], tmW, tmOut, tmRes, p);
}

}  // namespace cbx

// CAM layer "linear_local" (xvector.py:203-219): Conv1d(128 -> 32, k = 3, dilation d, zero padding d) on the bottleneck
// output u, multiplied by the context gate, appended to the dense block's concat buffer -- one persistent tcgen05 kernel.
//
// A CTA loops over 128-frame tiles.  Per tile the halo block of u (128 + 2d frames x 128 channels) is loaded ONCE as four
// 128B-swizzled K-block planes; the three taps are row-shifted descriptor views of those planes (shift = tap * d frames), so u
// crosses L2 -> SM once instead of three times.  Guard rows between clips are zero in u and rows outside the tensor are TMA
// zero fill: that is the convolution's zero padding.  Weights ([32][3*128]) stay resident; the [128 x 32] accumulator is
// double-buffered in TMEM; the epilogue multiplies by gate[segment(row)] and leaves through a staging tile and one TMA store
// into columns [cin, cin + 32) of the concat buffer.
#include "cbx_internal.h"
#include "tc.cuh"
#include "epi.cuh"

namespace cbx {
namespace lconv {

using namespace tc;

constexpr int STAGES = 2;
constexpr int PLANE_BYTES = 17 * 1024;            // (128 + 2*2) rows x 128 B rounded up to the 1 KB swizzle atom
constexpr int STAGE_BYTES = 4 * PLANE_BYTES;
constexpr int W_BYTES = 12 * 4096;                // [tap][k block][32 x 128 B]
constexpr int OUT_BYTES = 2 * 128 * 128;          // one staging tile per epilogue group
constexpr int SMEM_BYTES = W_BYTES + OUT_BYTES + STAGES * STAGE_BYTES + 1024 + 256;

struct Params {
  int M, dil, col0, ntiles;
  const float* gate; const int32_t* row_seg;
  uint16_t* shadow; int ldh;            // bf16 copy of the concatenation buffer (option cat_bf16), or null
};

// U16 (bf16 mode): u and the weights are bf16 -- a 128-byte row holds 64 channels, so the halo block is two planes instead of four and
// a tap is 8 tcgen05.mma kind::f16 (K = 16) instead of 16 kind::tf32 (K = 8): half the MMAs and half the A-operand bytes per tile.
template <bool U16>
__global__ void __launch_bounds__(320, 1)
local_conv_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                  const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sOut = smem + W_BYTES;
  uint8_t* sIn = sOut + OUT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sIn + STAGES * STAGE_BYTES);
  uint64_t* full = bars;              // [STAGES]
  uint64_t* empty = bars + STAGES;    // [STAGES]
  uint64_t* tfull = empty + STAGES;   // [2]
  uint64_t* tempty = tfull + 2;       // [2]
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t plane_tx = (uint32_t)(128 + 2 * p.dil) * 128u;

  if (warp == 0) {
    if (lane == 0) {
      constexpr int NPL = U16 ? 2 : 4;          // planes (K blocks of one 128-byte row) per halo block
      constexpr int KCOLS = U16 ? 64 : 32;      // channels per plane
      mbar_expect_tx(wfull, 3 * NPL * 4096);
      for (int kb = 0; kb < 3 * NPL; ++kb) tma_load_2d(sW + kb * 4096, &tmW, wfull, kb * KCOLS, 0);
      pdl_wait();                   // u comes from the bottleneck GEMM, two kernels back (the weights above do not)
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], NPL * plane_tx);
        for (int kb = 0; kb < NPL; ++kb)
          tma_load_2d(sIn + s * STAGE_BYTES + kb * PLANE_BYTES, &tmU, &full[s], kb * KCOLS, tile * 128 - p.dil);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = U16 ? make_idesc_bf16(128, 32) : make_idesc_tf32(128, 32);
    constexpr int NPL = U16 ? 2 : 4;
    const uint64_t dhi = make_desc_sw128(0);
    const uint32_t w16 = smem_u32(sW) >> 4;
    const uint32_t tap16 = (uint32_t)p.dil * 8u;             // one tap = dil rows of 128 B, in 16-byte units
    mbar_wait(wfull, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      const int a = it & 1, pa = (it >> 1) & 1;
      mbar_wait(&tempty[a], pa ^ 1);
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t d = tmem_base + a * 32;
      const uint32_t in16 = smem_u32(sIn + s * STAGE_BYTES) >> 4;
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 3; ++tap)
#pragma unroll
          for (int kb = 0; kb < NPL; ++kb) {
            const uint64_t ad = dhi | (uint64_t)((in16 + kb * (PLANE_BYTES >> 4) + tap * tap16) & 0x3FFFu);
            const uint64_t bd = dhi | (uint64_t)((w16 + (tap * NPL + kb) * 256) & 0x3FFFu);
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // four instructions of 32 K bytes per 128-byte row either way
              if constexpr (U16) umma_bf16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (tap | kb | k) != 0);
              else umma_tf32(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (tap | kb | k) != 0);
            }
          }
        umma_commit(&empty[s]);
        umma_commit(&tfull[a]);
      }
      __syncwarp();
    }
  } else {
    // two epilogue groups of four warps on alternate tiles (group = accumulator): the epilogue is a latency-bound chain
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int i = q * 32 + lane;
    uint8_t* const stage_out = sOut + grp * (128 * 128);
    float4* so = reinterpret_cast<float4*>(stage_out) + i * 8;
    const bool issuer = (warp == 2 || warp == 6) && lane == 0;
    pdl_wait();                     // the gate comes from the kernel in front
    for (int it = grp, tile = blockIdx.x + grp * gridDim.x; tile < p.ntiles; tile += 2 * gridDim.x, it += 2) {
      const int a = grp, pa = (it >> 1) & 1;
      mbar_wait(&tfull[a], pa);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * 32, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      const int m = tile * 128 + i;
      const int seg = m < p.M ? p.row_seg[m] : -1;
      const float4* g = reinterpret_cast<const float4*>(p.gate + (size_t)max(seg, 0) * kGrowth);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 gg = __ldg(g + c);
        v[4 * c] = seg >= 0 ? v[4 * c] * gg.x : 0.f;
        v[4 * c + 1] = seg >= 0 ? v[4 * c + 1] * gg.y : 0.f;
        v[4 * c + 2] = seg >= 0 ? v[4 * c + 2] * gg.z : 0.f;
        v[4 * c + 3] = seg >= 0 ? v[4 * c + 3] * gg.w : 0.f;
      }
      if (p.shadow && m < p.M) tc::store32_bf16(p.shadow + (size_t)m * p.ldh + p.col0, v);
      // the group's previous store has read the staging tile
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 8; ++c) so[c ^ (i & 7)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      fence_proxy_async();
      if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
      if (issuer) {
        tma_store_2d(&tmOut, stage_out, p.col0, tile * 128);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    pdl_trigger();                  // the last tile's store is on its way: the next layer's GEMM may start setting up
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

}  // namespace lconv

// u [M][128] -> cat[:, col0 : col0 + 32] = conv_k3_dil(u) * gate[segment]; tmU: {128 cols, M rows} box {32, 128 + 2 dil} (tf32),
// tmOut: {ld cols, M rows} box {32, 128} (fp32) over the concat buffer
void run_local_conv_tc(cbx_ctx* c, cudaStream_t st, const CUtensorMap& tmU, const CUtensorMap& tmW, const CUtensorMap& tmOut, int M, int dil,
                       int col0, const float* gate, const int32_t* row_seg, bool pdl, uint16_t* shadow, int ldh, bool u16) {
  using namespace lconv;
  if (M <= 0) return;
  ensure_max_smem(local_conv_kernel<false>, SMEM_BYTES);
  ensure_max_smem(local_conv_kernel<true>, SMEM_BYTES);
  Params p{M, dil, col0, (M + 127) / 128, gate, row_seg, shadow, ldh};
  const int grid = p.ntiles < tc::sm_count() ? p.ntiles : tc::sm_count();
  Scope sc(c->launches, st, "dense_local_gemm", 2.0 * M * kGrowth * 3 * kBnC, (u16 ? 2.0 : 4.0) * M * kBnC + 4.0 * M * kGrowth);
  if (u16) tc::launch_pdl(local_conv_kernel<true>, dim3(grid), dim3(320), SMEM_BYTES, st, pdl, tmU, tmW, tmOut, p);
  else tc::launch_pdl(local_conv_kernel<false>, dim3(grid), dim3(320), SMEM_BYTES, st, pdl, tmU, tmW, tmOut, p);
}

}  // namespace cbx

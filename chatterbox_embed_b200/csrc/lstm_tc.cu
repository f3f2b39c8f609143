// Recurrent step of the VoiceEncoder LSTM (voice_encoder.py:152; nn.LSTM(40,256,3), gate order i,f,g,o) as a persistent
// thread-block-cluster kernel on the sm_100a tensor cores.
//
//   gates_t[1024] = xw_t[1024] (input projection + biases, precomputed by one dense GEMM)  +  W_hh[1024x256] . h_{t-1}
//
// One cluster of 8 CTAs advances a tile of 2 x 112 partials through all 160 steps of one layer.
//   * CTA j owns hidden units [32j, 32j+32) and all four of their gates: 128 gate rows of W_hh (row 4u+g), K = 256.
//     That slice (128 x 256 tf32 = 128 KB) is loaded ONCE into TMEM (256 of the 512 columns) and is the A operand of
//     every tcgen05.mma of the kernel (A-from-TMEM form), so the weights never touch shared memory or L2 again.
//   * h_{t-1} of the tile's partials is the B operand: [112 partials x 256 units] tf32, K-major, 128B-swizzled, in shared
//     memory (112 KB per sub-tile, two sub-tiles A/B ping-pong so the tensor pipe works on one while the other is in its
//     gate math).  D = gates^T [128 gate rows x 112 partials] fp32 accumulates in TMEM (112 columns per sub-tile).
//   * Gate math: thread = one gate row (TMEM lane), registers = partials.  Each thread applies its own non-linearity
//     (sigmoid, or tanh for g), then a 4x4 transpose over the four lanes of a unit (warp shuffles) hands every lane all four
//     gates of one partial in four; c_t stays in registers (fp32), h_t = o * tanh(c_t).
//   * Exchange: each CTA writes its [112 x 32] slice of h_t (rounded to tf32) into its own copy of the B tile and pushes
//     it to the 7 peers with cp.async.bulk over distributed shared memory, completing on the peers' "full" mbarriers.
//     "free" mbarriers (one remote arrive per CTA per step) tell the senders that every peer's MMA has finished reading
//     h_{t-1} before it is overwritten.
#include <type_traits>

#include "cbx_internal.h"
#include "tc.cuh"

namespace cbx {
namespace lstm {

constexpr int NSUB = 112;                  // partials per sub-tile (UMMA N)
constexpr int TILE = 2 * NSUB;             // partials per cluster
constexpr int CL = 8;                      // CTAs per cluster
constexpr int UNITS = kVeHidden / CL;      // 32 hidden units per CTA
constexpr int KB_BYTES = NSUB * 128;       // one K block (32 units) of the B tile: NSUB rows x 128 B
constexpr int H_BYTES = CL * KB_BYTES;     // 114688
constexpr int CH = 16;                     // partials per gate-math chunk (one tcgen05.ld x16)
static_assert(NSUB % CH == 0 && NSUB % 16 == 0, "UMMA N and the chunking");
constexpr int SMEM_BYTES = 2 * H_BYTES + 1024 /*alignment*/ + 2 * NSUB * 4 /*row bases*/ + 128 /*barriers*/;
constexpr int THREADS = 288;               // warp 0: MMA issuer; warps 1-4: gate math A; warps 5-8: gate math B
constexpr uint32_t COL_W = 0, COL_D = 256; // TMEM columns: W slice [0,256), D_A [256,256+NSUB), D_B after it

using namespace tc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
// push a local shared-memory region into a peer CTA, completing (bytes) on the peer's mbarrier
__device__ __forceinline__ void dsmem_push(uint32_t dst_cluster_addr, uint32_t src_addr, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(src_addr), "r"(bytes), "r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// sigmoid(x) for m = 1, tanh(x) = 2 sigmoid(2x) - 1 for m = 2 (branch-free: the gate type is a per-thread constant)
__device__ __forceinline__ float gate_act(float x, float m, float neg_m_log2e, float one_minus_m) {
  const float e = ex2_approx(x * neg_m_log2e);
  return fmaf(m, rcp_approx(1.f + e), one_minus_m);
}
__device__ __forceinline__ float tanh_acc(float x) {
  const float e = ex2_approx(x * (-2.f * 1.4426950408889634f));
  return fmaf(2.f, rcp_approx(1.f + e), -1.f);
}

#ifdef CBX_DEV_TOOLS   // v1 of the recurrence (DSMEM pushes), kept for tools/ comparisons only
struct Params {
  const float* xw;            // [rows][1024], columns permuted: col = 128 j + 4 u + g  <->  gate g of unit 32 j + u
  const int32_t* slot_row;    // layer 0: first xw row of each partial slot (mel row); nullptr: row = slot * 160
  const float* whh;           // [1024][256] tf32-rounded, rows permuted like the xw columns
  float* hseq;                // [n_slots * 160][256] or nullptr
  float* hlast;               // [n_slots][256] or nullptr (h of the last step)
  int n_slots;
  long long* trace;           // [160][2][8] clock64 stamps of CTA 0 (dbg & 64)
  int dbg;                    // timing experiments only (results are wrong when non-zero): 1 no pushes, 2 no hseq stores, 4 no xw loads, 8 no free handshake
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) lstm_rec_tc_kernel(Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // 1024-aligned, still a shared-space pointer
  uint8_t* hbuf = smem;                                                   // [2][8 K blocks][NSUB rows x 128 B]
  int32_t* rowbase = reinterpret_cast<int32_t*>(smem + 2 * H_BYTES);      // [2][NSUB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rowbase + 2 * NSUB);
  uint64_t* full = bars;          // [2] h_t of the sub-tile is complete in this CTA (1 local arrive + 7 pushes)
  uint64_t* freeb = bars + 2;     // [2] all 8 CTAs' MMAs have finished reading h_{t-1}
  uint64_t* accum = bars + 4;     // [2] this CTA's gates for step t are in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t j = cluster_rank();
  const int tile = blockIdx.x / CL;
  const int q_tile = tile * TILE;
  const bool tr = (p.dbg & 64) && blockIdx.x == 0 && p.trace;

  if (threadIdx.x == 0) {
    for (int x = 0; x < 2; ++x) { mbar_init(&full[x], 1); mbar_init(&freeb[x], CL); mbar_init(&accum[x], 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // h_{-1} = 0; row bases of the tile's partial slots
  for (int i = threadIdx.x; i < 2 * H_BYTES / 16; i += THREADS) reinterpret_cast<float4*>(hbuf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < TILE; i += THREADS) {
    const int q = min(q_tile + i, p.n_slots - 1);
    rowbase[i] = p.slot_row ? p.slot_row[q] : q * kVePartial;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 1) {
    // W_hh slice -> TMEM: gate row r = lane of TMEM; group A loads K columns [0,128), group B [128,256)
    const int wg = (warp - 1) >> 2, qd = warp & 3;
    const int r = qd * 32 + lane;
    const float* wrow = p.whh + ((size_t)j * 128 + r) * kVeHidden + wg * 128;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + c) + i);
        v[4 * i] = w.x; v[4 * i + 1] = w.y; v[4 * i + 2] = w.z; v[4 * i + 3] = w.w;
      }
      tmem_st32(tmem_base + ((uint32_t)(qd * 32) << 16) + COL_W + wg * 128 + c, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();            // every CTA's barriers are initialised and its h tiles zeroed before any push can land
  tc_fence_after();

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, NSUB);
      for (int t = 0; t < kVePartial; ++t) {
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
          if (t > 0) mbar_wait(&full[x], (t - 1) & 1);
          tc_fence_after();
          if (tr) p.trace[(t * 2 + x) * 8 + 0] = clock64();
          const uint32_t hb = smem_u32(hbuf + x * H_BYTES);
#pragma unroll
          for (int kb = 0; kb < CL; ++kb) {
            const uint64_t bd = make_desc_sw128(hb + kb * KB_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_ts(tmem_base + COL_D + x * NSUB, tmem_base + COL_W + kb * 32 + k * 8, bd + (uint64_t)(k * 32 >> 4), idesc, (kb | k) != 0);
          }
          umma_commit(&accum[x]);
          if (tr) p.trace[(t * 2 + x) * 8 + 1] = clock64();
        }
      }
    }
  } else {
    // ===================== gate math + exchange =====================
    const int x = (warp - 1) >> 2;                 // sub-tile of this warp group
    const int qd = warp & 3;                       // TMEM lane quadrant
    const int r = qd * 32 + lane;                  // gate row: unit u = r / 4, gate g = r % 4
    const int u = r >> 2, g = r & 3;
    const bool P0 = g & 1, P1 = g & 2;
    const float m = g == 2 ? 2.f : 1.f;
    const float neg_m_log2e = -m * 1.4426950408889634f, one_minus_m = 1.f - m;
    const int gt = (threadIdx.x - 32) & 127;       // thread index within the warp group
    const int q0 = q_tile + x * NSUB;
    uint8_t* htile = hbuf + x * H_BYTES + j * KB_BYTES;          // this CTA's K block of the B tile
    const uint32_t off0 = g * 128 + (((u >> 2) ^ g) << 4) + (u & 3) * 4;                  // rows n with n % 8 == g
    const uint32_t off1 = (4 + g) * 128 + (((u >> 2) ^ (4 + g)) << 4) + (u & 3) * 4;      // rows n with n % 8 == 4 + g
    const float* xcol = p.xw + j * 128 + r;
    const int32_t* rb = rowbase + x * NSUB;
    const uint32_t dtm = tmem_base + ((uint32_t)(qd * 32) << 16) + COL_D + x * NSUB;
    float cst[NSUB / 4];
#pragma unroll
    for (int i = 0; i < NSUB / 4; ++i) cst[i] = 0.f;
    // peers' addresses of this sub-tile's barriers / of my K block (lanes 0..7 of the group's first warp push)
    uint32_t peer_free = 0, peer_full = 0, peer_tile = 0;
    if (gt < CL) {
      peer_free = mapa(smem_u32(&freeb[x]), gt);
      peer_full = mapa(smem_u32(&full[x]), gt);
      peer_tile = mapa(smem_u32(htile), gt);
    }

    for (int t = 0; t < kVePartial; ++t) {
      // input projections of this step (issued before the accumulator wait: the loads fly while the MMA runs)
      const bool trt = tr && gt == 0;
      if (trt) p.trace[(t * 2 + x) * 8 + 2] = clock64();
      float xv[NSUB];
#pragma unroll
      for (int n = 0; n < NSUB; ++n) xv[n] = (p.dbg & 4) ? 0.f : __ldg(xcol + (size_t)(rb[n] + t) * kVeGates);

      mbar_wait(&accum[x], t & 1);
      tc_fence_after();
      if (trt) p.trace[(t * 2 + x) * 8 + 3] = clock64();
      if (!(p.dbg & 8)) {
        if (gt < CL) remote_arrive(peer_free);       // this CTA's MMA no longer reads h_{t-1} of sub-tile x
        mbar_wait_cluster(&freeb[x], t & 1);         // ... and neither does anybody else's: the tile may be overwritten
      }

      if (trt) p.trace[(t * 2 + x) * 8 + 4] = clock64();
      const bool last = t == kVePartial - 1;
#pragma unroll
      for (int c = 0; c < NSUB / CH; ++c) {
        float v[CH];
        tmem_ld16(dtm + c * CH, v);
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = gate_act(v[i] + xv[c * CH + i], m, neg_m_log2e, one_minus_m);
#pragma unroll
        for (int grp = 0; grp < CH / 4; ++grp) {
          const float a0 = v[4 * grp], a1 = v[4 * grp + 1], a2 = v[4 * grp + 2], a3 = v[4 * grp + 3];
          // 4x4 transpose over the lanes of one unit: lane g ends up with gates g, g^1, g^2, g^3 of partial 4 grp + g
          const float k0 = P0 ? a1 : a0, s0 = P0 ? a0 : a1;
          const float k1 = P0 ? a3 : a2, s1 = P0 ? a2 : a3;
          const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
          const float own = P1 ? k1 : k0, sown = P1 ? k0 : k1;
          const float par = P1 ? r1 : r0, spar = P1 ? r0 : r1;
          const float own2 = __shfl_xor_sync(0xffffffffu, sown, 2), par2 = __shfl_xor_sync(0xffffffffu, spar, 2);
          // own: gate g, par: gate g^1, own2: gate g^2, par2: gate g^3
          const float e0 = P0 ? par : own, o0 = P0 ? own : par;       // even / odd gate with bit 1 == P1
          const float e1 = P0 ? par2 : own2, o1 = P0 ? own2 : par2;   // even / odd gate with bit 1 != P1
          const float gi = P1 ? e1 : e0, gg = P1 ? e0 : e1, gf = P1 ? o1 : o0, go = P1 ? o0 : o1;
          const int ci = c * (CH / 4) + grp;
          const float cn = fmaf(gf, cst[ci], gi * gg);
          cst[ci] = cn;
          const float h = go * tanh_acc(cn);
          const int nb = c * CH + (grp & ~1) * 4;                  // first row of the 8-row swizzle atom
          if (!last)
            *reinterpret_cast<float*>(htile + nb * 128 + ((grp & 1) ? off1 : off0)) = to_tf32(h);
          const int q = q0 + c * CH + grp * 4 + g;
          if (q < p.n_slots) {
            if (p.hseq && !(p.dbg & 2)) p.hseq[((size_t)q * kVePartial + t) * kVeHidden + j * UNITS + u] = h;
            if (last && p.hlast) p.hlast[(size_t)q * kVeHidden + j * UNITS + u] = h;
          }
        }
      }
      if (trt) p.trace[(t * 2 + x) * 8 + 5] = clock64();
      if (!last) {
        fence_proxy_async();                       // generic-proxy writes of the slice -> visible to the async proxy
        tc_fence_before();
        if (x == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (trt) p.trace[(t * 2 + x) * 8 + 6] = clock64();
        if (gt < CL) {
          if ((uint32_t)gt == j) mbar_expect_tx(&full[x], (p.dbg & 1) ? 0 : (CL - 1) * KB_BYTES);    // local slice present + 7 pushes expected
          else if (!(p.dbg & 1)) dsmem_push(peer_tile, smem_u32(htile), KB_BYTES, peer_full);
        }
        if (trt) p.trace[(t * 2 + x) * 8 + 7] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();            // nobody leaves while a peer may still signal one of its barriers
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


#endif  // CBX_DEV_TOOLS

// =====================================================================================================================
// v2 of the recurrence.  Same decomposition (8-CTA cluster, W_hh slice in TMEM as the A operand, two ping-pong sub-tiles
// of 112 partials), but
//   * the exchange of h_t goes through L2: every CTA stores its [112 x 32] slice into the layer's hseq buffer (which the
//     next layer's input-projection GEMM needs anyway), then ONE thread issues a multicast TMA load of that slice into the
//     B tiles of all 8 CTAs.  (The DSMEM pushes of v1 ran at the ~20 B/clk/SM DSMEM port rate and cost 4-6 k cycles a step.)
//   * xw (layers 1, 2) and hseq use a tiled time-major row order, row(q, t) = ((tile*160 + t)*2 + x)*112 + n for partial
//     slot q = tile*224 + x*112 + n, so that everything a (tile, t, x) half-step touches is one contiguous block and all
//     the addressing in the loop is immediate offsets from two running pointers;
//   * the gate math of a half-step is spread over all 8 gate warps (two per TMEM lane quadrant, 56 partials each) and the
//     xw loads of the NEXT half-step are issued chunk by chunk while the current one is being computed.
// Gate warps per TMEM lane quadrant (NGW): 2 -> 8 gate warps, 56 partials and 8-column chunks per thread (round 1); 4 -> 16 gate
// warps (four per SM sub-partition), 28 partials and 4-column chunks per thread: the same instructions in total, but twice the
// warps to hide the TMEM-load / MUFU / shuffle latencies of the dependent chain behind (the gate math is latency bound).

// streaming 4-byte load that does not allocate in L1: the unified L1 / shared-memory array is busy feeding the tensor core
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// the same for an input projection stored as bf16 (bf16 mode): 2-byte load, the value is the upper half of the fp32 pattern
// (returned RAW, zero-extended: the shift that turns it into a float is done where the value is consumed, a half-step later -- a
// conversion placed at the load would make the thread wait for every prefetched load at once: measured 3.5 -> 5.7 ms per step)
__device__ __forceinline__ uint32_t ldg_stream_bf16_raw(const uint16_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg_stream_f32_raw(const float* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

struct Params2 {
  const void* xw;             // fp32 or bf16 (kXw16); layer 0: [mel rows][1024] gathered through slot_row; else tiled time-major [tiles*160*224][1024]
  const int32_t* slot_row;    // layer 0 only
  const float* whh;
  float* hseq;                // tiled time-major [tiles*160*224][256] (tf32-rounded h), also the exchange medium
  float* hlast;               // [n_slots][256] or nullptr
  int n_slots;
  long long* trace;           // [160][2][8] clock64 stamps of CTA 0 when non-null (tests/tools/lstm_trace.py)
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// (Loading a half-step's 28 accumulator columns with back-to-back tcgen05.ld and ONE wait was measured slower than the per-chunk
// loads: 3.25 -> 3.53 ms for the three layers -- the gate phase is bound by its instruction count, not by TMEM latency.)
constexpr int threads2(int ngw) { return 32 * (2 + 4 * ngw); }      // warp 0: MMA issuer, warps 1 .. 4 NGW: gate warps, last warp: exchange
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 4 reciprocals for one MUFU: r = 1/(d0 d1 d2 d3); the d's are 1 + 2^x with x clamped to 30, so the product stays finite
__device__ __forceinline__ void rcp4(float d0, float d1, float d2, float d3, float& i0, float& i1, float& i2, float& i3) {
  const float p01 = d0 * d1, p23 = d2 * d3;
  const float r = rcp_approx(p01 * p23);
  const float r01 = r * p23, r23 = r * p01;
  i0 = r01 * d1; i1 = r01 * d0; i2 = r23 * d3; i3 = r23 * d2;
}

template <bool kLayer0, bool kXw16, int NGW>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(threads2(NGW), 1)
lstm_rec_tc2_kernel(const __grid_constant__ CUtensorMap tmH, Params2 p) {
  constexpr int THREADS2 = threads2(NGW);
  constexpr int HALF = NSUB / NGW;           // partials per gate warp and half-step
  constexpr int CH2 = NGW == 2 ? 8 : 4;      // partials per tcgen05.ld
  constexpr int NGRP = CH2 / 4;              // groups of four partials per chunk (one 4x4 transpose each)
  constexpr int XWARP = 1 + 4 * NGW;         // the exchange warp
  static_assert(HALF % CH2 == 0 && NSUB % NGW == 0, "chunking");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* hbuf = smem;                                                   // [2][8 K blocks][NSUB rows x 128 B]
  int32_t* rowbase = reinterpret_cast<int32_t*>(smem + 2 * H_BYTES);      // [2][NSUB] (layer 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(rowbase + 2 * NSUB);
  uint64_t* full = bars;          // [2] 1 local arrive (expect 8 slices) + 8 multicast TMA loads
  uint64_t* freeb = bars + 2;     // [2] all 8 CTAs' MMAs have finished reading h_{t-1}
  uint64_t* accum = bars + 4;     // [2]
  uint64_t* ready = bars + 6;     // [2] this CTA's slice of h_t is stored and fenced (256 gate-thread arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t j = cluster_rank();
  const int tile = blockIdx.x / CL;
  const int q_tile = tile * TILE;
  const bool tr = p.trace != nullptr && blockIdx.x == 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmH);
    for (int x = 0; x < 2; ++x) { mbar_init(&full[x], 1); mbar_init(&freeb[x], CL); mbar_init(&accum[x], 1); mbar_init(&ready[x], 128 * NGW); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 2 * H_BYTES / 16; i += THREADS2) reinterpret_cast<float4*>(hbuf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kLayer0)
    for (int i = threadIdx.x; i < TILE; i += THREADS2) rowbase[i] = p.slot_row[min(q_tile + i, p.n_slots - 1)];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 1 && warp <= 8) {
    const int wg = (warp - 1) >> 2, qd = warp & 3;
    const int r = qd * 32 + lane;
    const float* wrow = p.whh + ((size_t)j * 128 + r) * kVeHidden + wg * 128;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + c) + i);
        v[4 * i] = w.x; v[4 * i + 1] = w.y; v[4 * i + 2] = w.z; v[4 * i + 3] = w.w;
      }
      tmem_st32(tmem_base + ((uint32_t)(qd * 32) << 16) + COL_W + wg * 128 + c, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();
  tc_fence_after();

  const size_t blk0 = (size_t)tile * kVePartial * 2 * NSUB;                      // first tiled row of the tile
  if (warp == 0) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues
    constexpr uint32_t idesc = make_idesc_tf32(128, NSUB);
    for (int t = 0; t < kVePartial; ++t) {
#pragma unroll 1
      for (int x = 0; x < 2; ++x) {
        if (t > 0) mbar_wait(&full[x], (t - 1) & 1);
        tc_fence_after();
        const uint32_t hb = smem_u32(hbuf + x * H_BYTES);
        if (elect_one()) {
          if (tr) p.trace[(t * 2 + x) * 8 + 0] = clock64();
#pragma unroll
          for (int kb = 0; kb < CL; ++kb) {
            const uint64_t bd = make_desc_sw128(hb + kb * KB_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_ts(tmem_base + COL_D + x * NSUB, tmem_base + COL_W + kb * 32 + k * 8, bd + (uint64_t)(k * 32 >> 4), idesc, (kb | k) != 0);
          }
          umma_commit(&accum[x]);
          if (tr) p.trace[(t * 2 + x) * 8 + 1] = clock64();
        }
        __syncwarp();
      }
    }
  } else if (warp == XWARP) {
    // ===================== exchange: slice stored -> peers' MMAs done with h_{t-1} -> multicast TMA load into all 8 B tiles
    if (lane == 0) {
      for (int t = 0; t < kVePartial - 1; ++t) {
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
          const int hs = 2 * t + x;
          mbar_wait(&ready[x], t & 1);
          // ONE gpu-scope fence per half-step, here: the gate threads' stores of h_t were ordered before their (release.cta)
          // arrivals on `ready`, which this thread has just acquired, and fences are cumulative -- so this fence makes all of
          // them visible at L2, where the TMA load below reads them.  (Round 1 had every gate thread execute __threadfence():
          // MEMBAR.SC + ERRBAR + CCTL.IVALL in 512 threads was 16 % of the kernel's warp-stall samples, ncu source page.)
          __threadfence();
          asm volatile("fence.proxy.async.global;" ::: "memory");
          if (tr) p.trace[(t * 2 + x) * 8 + 5] = clock64();
          mbar_wait_cluster(&freeb[x], t & 1);
          if (tr) p.trace[(t * 2 + x) * 8 + 6] = clock64();
          mbar_expect_tx(&full[x], CL * KB_BYTES);
          tma_load_2d_mc(hbuf + x * H_BYTES + j * KB_BYTES, &tmH, &full[x], j * UNITS, (int)(blk0 + (size_t)hs * NSUB), (uint16_t)0xff);
          if (tr) p.trace[(t * 2 + x) * 8 + 7] = clock64();
        }
      }
    }
  } else {
    // ===================== gate warps: quadrant qd, column half hf; all 8 warps work on sub-tile A, then on sub-tile B
    const int qd = warp & 3, hf = (warp - 1) >> 2;
    const int r = qd * 32 + lane;
    const int u = r >> 2, g = r & 3;
    const bool P0 = g & 1, P1 = g & 2;
    const float m = g == 2 ? 2.f : 1.f;
    const float neg_m_log2e = -m * 1.4426950408889634f, one_minus_m = 1.f - m;
    const int gt = threadIdx.x - 32;               // 0 .. 128 NGW - 1
    const int n0 = hf * HALF;                      // first partial (within a sub-tile) of this warp
    using XwT = typename std::conditional<kXw16, uint16_t, float>::type;
    const XwT* xq = static_cast<const XwT*>(p.xw) + (kLayer0 ? (size_t)0 : (blk0 + n0) * kVeGates) + j * 128 + r;
    auto ldx = [](const XwT* q) -> uint32_t {
      if constexpr (kXw16) return ldg_stream_bf16_raw(q); else return ldg_stream_f32_raw(q);
    };
    auto xval = [](uint32_t raw) -> float {
      if constexpr (kXw16) return __uint_as_float(raw << 16); else return __uint_as_float(raw);
    };
    float* hq = p.hseq + (blk0 + n0 + g) * kVeHidden + j * UNITS + u;
    const uint32_t dcol = tmem_base + ((uint32_t)(qd * 32) << 16) + COL_D + n0;
    float cst[2][HALF / 4];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int i = 0; i < HALF / 4; ++i) cst[x][i] = 0.f;
    uint32_t peer_free[2] = {0, 0};
    if (gt < CL) { peer_free[0] = mapa(smem_u32(&freeb[0]), gt); peer_free[1] = mapa(smem_u32(&freeb[1]), gt); }

    auto xw_load = [&](int hs, int n) -> uint32_t {     // input projection of partial n0 + n for half-step hs = 2 t + x
      if (kLayer0) {
        const int x = hs & 1, t = hs >> 1;
        return ldx(xq + (size_t)(rowbase[x * NSUB + n0 + n] + t) * kVeGates);
      }
      return ldx(xq + ((size_t)hs * NSUB + n) * kVeGates);
    };
    // activations of one 8-column chunk: own gate of 8 partials; two shared reciprocals
    auto activate = [&](int x, int c, const uint32_t* xin, float* a) {
      float v[CH2];
      if constexpr (CH2 == 8) tmem_ld8(dcol + x * NSUB + c * CH2, v); else tmem_ld4(dcol + x * NSUB + c * CH2, v);
#ifndef CBX_LSTM_EXP_RCP_ACT
      // MUFU.TANH: sigmoid(s) = 0.5 tanh(0.5 s) + 0.5.  Measured against the oracle the embedding error is the same as with
      // the ex2 + rcp formulation below (W1 1.5e-5, W2 3.6e-4: the TF32 products dominate), at 3 instructions per gate value.
      const float ta = g == 2 ? 1.f : 0.5f, tcn = g == 2 ? 0.f : 0.5f;
#pragma unroll
      for (int i = 0; i < CH2; ++i) {
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"((v[i] + xval(xin[c * CH2 + i])) * ta));
        a[i] = fmaf(ta, th, tcn);
      }
#else
      float d[CH2];
#pragma unroll
      for (int i = 0; i < CH2; ++i) d[i] = 1.f + ex2_approx(fminf((v[i] + xval(xin[c * CH2 + i])) * neg_m_log2e, 30.f));
      float inv[CH2];
#pragma unroll
      for (int i = 0; i < CH2; i += 4) rcp4(d[i], d[i + 1], d[i + 2], d[i + 3], inv[i], inv[i + 1], inv[i + 2], inv[i + 3]);
#pragma unroll
      for (int i = 0; i < CH2; ++i) a[i] = fmaf(m, inv[i], one_minus_m);
#endif
    };
    uint32_t xv[HALF];
#pragma unroll
    for (int n = 0; n < HALF; ++n) xv[n] = xw_load(0, n);

    for (int t = 0; t < kVePartial; ++t) {
      const bool last = t == kVePartial - 1;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int hs = 2 * t + x;
        const bool trt = tr && gt == 0;
        if (trt) p.trace[(t * 2 + x) * 8 + 2] = clock64();
        mbar_wait(&accum[x], t & 1);
        tc_fence_after();
        if (trt) p.trace[(t * 2 + x) * 8 + 3] = clock64();
        if (gt < CL) remote_arrive(peer_free[x]);
        float* hrow = hq + (size_t)hs * NSUB * kVeHidden;
        // software pipeline over the 7 chunks: the activations (MUFU) of chunk c+1 are issued alongside the transpose and
        // cell update (ALU / shuffle) of chunk c
        float act[2][CH2];
        activate(x, 0, xv, act[0]);
#pragma unroll
        for (int c = 0; c < HALF / CH2; ++c) {
          if (c + 1 < HALF / CH2) activate(x, c + 1, xv, act[(c + 1) & 1]);
          // the xw registers of chunk c have been consumed: they take the loads of the next half-step
          if (hs + 1 < 2 * kVePartial) {
#pragma unroll
            for (int i = 0; i < CH2; ++i) xv[c * CH2 + i] = xw_load(hs + 1, c * CH2 + i);
          }
          const float* v = act[c & 1];
          float cn[NGRP], gout[NGRP];
#pragma unroll
          for (int grp = 0; grp < NGRP; ++grp) {
            const float a0 = v[4 * grp], a1 = v[4 * grp + 1], a2 = v[4 * grp + 2], a3 = v[4 * grp + 3];
            const float k0 = P0 ? a1 : a0, s0 = P0 ? a0 : a1;
            const float k1 = P0 ? a3 : a2, s1 = P0 ? a2 : a3;
            const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
            const float own = P1 ? k1 : k0, sown = P1 ? k0 : k1;
            const float par = P1 ? r1 : r0, spar = P1 ? r0 : r1;
            const float own2 = __shfl_xor_sync(0xffffffffu, sown, 2), par2 = __shfl_xor_sync(0xffffffffu, spar, 2);
            const float e0 = P0 ? par : own, o0 = P0 ? own : par;
            const float e1 = P0 ? par2 : own2, o1 = P0 ? own2 : par2;
            const float gi = P1 ? e1 : e0, gg = P1 ? e0 : e1, gf = P1 ? o1 : o0;
            gout[grp] = P1 ? o0 : o1;
            const int ci = c * NGRP + grp;
            cn[grp] = fmaf(gf, cst[x][ci], gi * gg);
            cst[x][ci] = cn[grp];
          }
          // tanh(c) of the chunk's cells, h = o * tanh(c), stored (tf32-rounded) into the layer's hseq buffer = the exchange medium
          float hh[NGRP];
#ifndef CBX_LSTM_EXP_RCP_ACT
#pragma unroll
          for (int grp = 0; grp < NGRP; ++grp) {
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(cn[grp]));
            hh[grp] = gout[grp] * th;
          }
#else
#pragma unroll
          for (int grp = 0; grp < NGRP; ++grp) {
            const float dd = 1.f + ex2_approx(fminf(cn[grp] * (-2.f * 1.4426950408889634f), 60.f));
            hh[grp] = gout[grp] * fmaf(2.f, rcp_approx(dd), -1.f);
          }
#endif
#pragma unroll
          for (int grp = 0; grp < NGRP; ++grp) {
            hrow[(size_t)(c * CH2 + 4 * grp) * kVeHidden] = to_tf32(hh[grp]);           // partial n0 + c * CH2 + 4 grp + g
            if (last && p.hlast) {
              const int q = q_tile + x * NSUB + n0 + c * CH2 + 4 * grp + g;
              if (q < p.n_slots) p.hlast[(size_t)q * kVeHidden + j * UNITS + u] = hh[grp];
            }
          }
        }
        // publish this half-step's slice of h_t at once: the exchange (fence, wait for the peers, multicast TMA) and the MMA of
        // the next step of this sub-tile then run while the gate warps are in the other sub-tile's math
        if (hs + 2 < 2 * kVePartial) {
          asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy GLOBAL stores ordered before the exchange warp's TMA read
          mbar_arrive(&ready[x]);
        }
        if (trt) p.trace[(t * 2 + x) * 8 + 4] = clock64();
        tc_fence_before();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lstm

#ifdef CBX_DEV_TOOLS
void run_lstm_rec_tc(cbx_ctx* c, const float* xw, const int32_t* slot_row, const float* whh_perm, float* hseq, float* hlast,
                     int n_slots, cudaStream_t st) {
  if (n_slots <= 0) return;
  ensure_max_smem(lstm::lstm_rec_tc_kernel, lstm::SMEM_BYTES);
  lstm::Params p{xw, slot_row, whh_perm, hseq, hlast, n_slots, (long long*)c->lstm_trace, (int)c->lstm_dbg};
  const int tiles = (n_slots + lstm::TILE - 1) / lstm::TILE;
  Scope sc(c->launches, st, "lstm_rec_tc_kernel", 2.0 * n_slots * kVePartial * kVeHidden * kVeGates, 4.0 * n_slots * kVePartial * (kVeGates + kVeHidden));
  lstm::lstm_rec_tc_kernel<<<tiles * lstm::CL, lstm::THREADS, lstm::SMEM_BYTES, st>>>(p);
}

}  // namespace cbx

// Diagnostic: how many 8-CTA clusters of the recurrence kernel can be co-resident on this device.
extern "C" int cbx_lstm_max_clusters(cbx_ctx* c) {
  using namespace cbx;
  if (!c) return -1;
  DeviceGuard dev_guard(c->device);
  cudaFuncSetAttribute(lstm::lstm_rec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lstm::SMEM_BYTES);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(lstm::CL * 64); cfg.blockDim = dim3(lstm::THREADS); cfg.dynamicSmemBytes = lstm::SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = lstm::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, lstm::lstm_rec_tc_kernel, &cfg);
  if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return -2; }
  return n;
}
#else
}  // namespace cbx
#endif  // CBX_DEV_TOOLS

namespace cbx {
// v2 entry point.  hseq / xw rows are in the tiled time-major order (see lstm_rec_tc2_kernel); both must hold
// lstm_padded_slots(n_slots) * 160 rows.
int lstm_padded_slots(int n_slots) { return (n_slots + lstm::TILE - 1) / lstm::TILE * lstm::TILE; }

void run_lstm_rec_tc2(cbx_ctx* c, const void* xw, bool xw_bf16, const int32_t* slot_row, const float* whh_perm, float* hseq, float* hlast,
                      int n_slots, cudaStream_t st) {
  if (n_slots <= 0) return;
  const int tiles = (n_slots + lstm::TILE - 1) / lstm::TILE;
  const int64_t rows = (int64_t)tiles * lstm::TILE * kVePartial;
  CUtensorMap tmH = tc::make_map_2d(hseq, rows, kVeHidden, kVeHidden, lstm::NSUB, false);
  lstm::Params2 p{xw, slot_row, whh_perm, hseq, hlast, n_slots, (long long*)c->lstm_trace};
  Scope sc(c->launches, st, "lstm_rec_tc_kernel", 2.0 * n_slots * kVePartial * kVeHidden * kVeGates,
           (double)n_slots * kVePartial * ((xw_bf16 ? 2.0 : 4.0) * kVeGates + 4.0 * kVeHidden));
  auto go = [&](auto kern, int ngw) {
    ensure_max_smem(kern, lstm::SMEM_BYTES);
    kern<<<tiles * lstm::CL, lstm::threads2(ngw), lstm::SMEM_BYTES, st>>>(tmH, p);
  };
  if (c->lstm_gate_warps == 4) {
    if (slot_row) { if (xw_bf16) go(lstm::lstm_rec_tc2_kernel<true, true, 4>, 4); else go(lstm::lstm_rec_tc2_kernel<true, false, 4>, 4); }
    else { if (xw_bf16) go(lstm::lstm_rec_tc2_kernel<false, true, 4>, 4); else go(lstm::lstm_rec_tc2_kernel<false, false, 4>, 4); }
  } else {
    if (slot_row) { if (xw_bf16) go(lstm::lstm_rec_tc2_kernel<true, true, 2>, 2); else go(lstm::lstm_rec_tc2_kernel<true, false, 2>, 2); }
    else { if (xw_bf16) go(lstm::lstm_rec_tc2_kernel<false, true, 2>, 2); else go(lstm::lstm_rec_tc2_kernel<false, false, 2>, 2); }
  }
}
}  // namespace cbx

// sm_100a tensor-core GEMM engine: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory -> tcgen05.mma
// kind::tf32 with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
//   C[m][n] = epi( sum_k pro(A[m][k]) * W[n][k] )     A, W both K-major (row-major activations, PyTorch weights)
//
// One 128 x BN output tile per CTA.  Warp roles: warp 0 = TMA producer (one elected lane), warp 1 = TMEM
// allocator + MMA issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quadrant = warp % 4), and when the op
// has an A-operand prologue (BatchNorm + ReLU before the conv, xvector.py:266-271) warps 6..9 transform each landed A
// stage in place before it is handed to the MMA warp.  Pipelines: full/empty mbarriers per smem stage (TMA <-> MMA,
// with an extra "ready" barrier when the prologue warps sit in between) and one accumulator barrier (MMA -> epilogue).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cbx_internal.h"

namespace cbx {
namespace tc {

constexpr int BM = 128;       // UMMA M
constexpr int BK = 32;        // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;     // tf32: 32 bytes of K per instruction

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// streaming 16-byte global load that does not allocate in L1 (the unified L1 / shared-memory array is the scarce resource of
// the producer-fed GEMMs: tensor-core operand reads, stage stores and TMA writes all go through it)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// one lane of a fully converged warp (the warp-uniform tcgen05 / TMA instructions are issued from inside `if (elect_one())`
// with the whole warp running the surrounding loop: that keeps the issue path free of divergence bookkeeping)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// shared -> global tensor store (bulk async-group completion); elements of the box outside the tensor are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with bf16 operands (kind::f16, fp32 accumulation): 16 K elements = 32 bytes per instruction
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier when all MMAs issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- programmatic dependent launch
// The dense-layer chain (bottleneck GEMM -> CAM gate -> local conv, 52 times) is ~150 short dependent kernels: each one is
// launched with programmatic stream serialization, does its set-up (barriers, TMEM, weight loads) while the previous kernel
// drains, and calls pdl_wait() before it touches anything the previous kernels produced (or overwrites anything they read).
// A kernel launched without the attribute sees both instructions as no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 32 lanes x 32 columns: thread i of the warp gets columns [c, c+32) of TMEM lane (quadrant*32 + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float to_tf32(float x) {   // round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (stride byte offset), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version, bits [46,48)
  d |= (uint64_t)2 << 61;                               // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// kind::tf32 instruction descriptor: D=f32, A=B=tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- K-block -> A tile coordinates
// The K loop runs over nkb blocks of 32 columns.  Block kb belongs to "tap" kb / cpb (cpb = column blocks per tap) and
// reads A columns col0[tap] + (kb % cpb)*32 of rows m0 + shift[tap]: a time-shifted read is how the convolutions'
// taps are expressed; rows outside the tensor come back as zeros from TMA, rows between clips are zero guard rows.
struct TapMap {
  int cpb;
  int shift[5];
  int col0[5];
};
__host__ inline TapMap plain_map(int K) { TapMap t{}; t.cpb = (K + BK - 1) / BK; return t; }

// A-operand prologue: BatchNorm scale/shift + ReLU + tf32 rounding, applied in place on the landed smem stage
struct NoPrologue { static constexpr bool kOn = false; };
struct BnReluPrologue {
  static constexpr bool kOn = true;
  const float* a; const float* b;     // per input channel (column of A)
};

constexpr int smem_bytes(int BN, int stages) { return stages * (BM * BK * 4 + BN * BK * 4) + 1024 + 256; }

template <int BN, int STAGES, class Pro, class Epi>
__global__ void __launch_bounds__(Pro::kOn ? 320 : 192)
tgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int nkb, TapMap tap, Pro pro, Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* ready = empty + STAGES;          // only used with a prologue
  uint64_t* accum = ready + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&ready[s], 128);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES, ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        const int t = kb / tap.cpb, cb = kb - t * tap.cpb;
        tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], tap.col0[t] + cb * BK, m0 + tap.shift[t]);
        tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, ph = (kb / STAGES) & 1;
      mbar_wait(Pro::kOn ? &ready[s] : &full[s], ph);
      tc_fence_after();
      const uint64_t ad = make_desc_sw128(smem_u32(sA + s * A_BYTES));
      const uint64_t bd = make_desc_sw128(smem_u32(sB + s * B_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_tf32(tmem_base, ad + (uint64_t)(k * UMMA_K * 4 >> 4), bd + (uint64_t)(k * UMMA_K * 4 >> 4), idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
        if (kb == nkb - 1) umma_commit(accum);
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ===== epilogue: TMEM -> registers -> global =====
    mbar_wait(accum, 0);
    tc_fence_after();
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      epi(row, n0 + c, v);
    }
    tc_fence_before();
  } else {
    // ===== A-operand prologue warps (only instantiated for Pro::kOn) =====
    if constexpr (Pro::kOn) {
      const int t = threadIdx.x - 192;         // 0..127
      const int chunk = t & 7, r0 = t >> 3;    // 16-byte chunk within the 128-byte row, first row
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES, ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const int tp = kb / tap.cpb, cb = kb - tp * tap.cpb;
        const int kcol = tap.col0[tp] + cb * BK;
        float4* base = reinterpret_cast<float4*>(sA + s * A_BYTES);
#pragma unroll
        for (int i = 0; i < BM / 16; ++i) {
          const int r = r0 + i * 16;
          const int logical = chunk ^ (r & 7);                 // 128B swizzle: physical chunk = logical ^ (row % 8)
          const int k = kcol + logical * 4;
          const float4 sc = __ldg(reinterpret_cast<const float4*>(pro.a + k));
          const float4 sh = __ldg(reinterpret_cast<const float4*>(pro.b + k));
          float4 x = base[r * 8 + chunk];
          x.x = to_tf32(fmaxf(fmaf(x.x, sc.x, sh.x), 0.f));
          x.y = to_tf32(fmaxf(fmaf(x.y, sc.y, sh.y), 0.f));
          x.z = to_tf32(fmaxf(fmaf(x.z, sc.z, sh.z), 0.f));
          x.w = to_tf32(fmaxf(fmaf(x.w, sc.w, sh.w), 0.f));
          base[r * 8 + chunk] = x;
        }
        fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core's async proxy
        mbar_arrive(&ready[s]);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- pre-activation GEMM (BN + ReLU on the A operand)
//   C[m][n] = epi( sum_k relu(a[k] * X[m][k] + b[k]) * W[n][k] )
// The D-TDNN layers are pre-activation (BN -> ReLU -> conv, xvector.py:266-271), every layer with its OWN BatchNorm over the
// shared concat buffer, so the normalisation cannot be folded into weights and must happen while the A operand is staged.
// Here producer warps do the staging themselves: coalesced 16-byte global loads of X (many in flight), BN + ReLU + tf32
// rounding in registers, one store into the 128B-swizzled K-major stage.  Two producer groups of 4 warps alternate K blocks,
// so one group's load latency hides behind the other's arithmetic.  W tiles come by TMA; MMA / epilogue as in tgemm_kernel.
// Warps: 0 = TMA (W), 1 = MMA issuer, 2..5 = producer group 0, 6..9 = producer group 1; all eight run the epilogue.
constexpr int PG = 2;     // producer groups

// timeline probe (tools/gemm_trace.py): when set, every CTA records {smid, t_entry, t_setup, t_first_full, t_accum, t_end}
static __device__ unsigned long long* g_gemm_trace = nullptr;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// XM = 1: X is the bf16 copy of the buffer (option cat_bf16 = 1): half the bytes and half the load instructions per K block; BN,
// ReLU and the tf32 stage are unchanged (the MMA stays kind::tf32, the weights are not rounded any further).
// XM = 2 (cat_bf16 = 2): bf16 OPERANDS as well -- a K block is 64 channels per 128-byte row, the stage holds bf16, W comes
// from its bf16 copy (tmB: box {64, BN}), kind::f16 MMAs: half the stage stores and half the operand bytes per channel.
// K need not be a multiple of 64: the tail block's missing channels are zeros on both sides (TMA fill / masked loads).
// does the epilogue functor write the tile itself (then the kernel's staging + TMA store of C is skipped)?
template <class E> __device__ __forceinline__ auto epi_skips_c(const E& e, int) -> decltype(e.skip_c()) { return e.skip_c(); }
template <class E> __device__ __forceinline__ bool epi_skips_c(const E&, long) { return false; }

template <int BN, int STAGES, class Epi, int XM>
__global__ void __launch_bounds__(320, 2)
tgemm_bnrelu_kernel(const void* __restrict__ Xv, int lda, int M, int K, const float* __restrict__ bn_a, const float* __restrict__ bn_b,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, int nkb, Epi epi,
                    const __grid_constant__ CUtensorMap tmX, int pf_stride) {
  constexpr bool XH = XM == 1;
  constexpr int KB = XM == 2 ? 64 : BK;          // channels per K block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bfull = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* afull = bfull + STAGES;          // 128 producer arrivals
  uint64_t* empty = afull + STAGES;          // MMA commit: both the A and the B half of the stage are free
  uint64_t* accum = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  // BN scale / shift of all K columns, staged once (static data: loaded before the dependency wait): the producers then
  // need no global loads besides X, whose requests already fill the SM's load queue
  float* s_bn = reinterpret_cast<float*>(smem + STAGES * (A_BYTES + B_BYTES) + 256);      // [2][nkb * 32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;     // n fastest: the CTAs that share an X row tile run together (L2 reuse)
  // TF32 path: the producers put the loads of their FIRST K block in flight before the set-up (BN vector staging, barrier init,
  // TMEM allocation, the CTA-wide sync), so that ~1 us of every CTA's 3.4 - 4.2 us "until the first MMA can issue" overlaps it
  float4 pre[BM / 16];
  if constexpr (XM == 0) {
    if (warp >= 2) {
      const int g0 = (warp - 2) >> 2, t0 = (threadIdx.x - 64) & 127;
      const float* xp0 = reinterpret_cast<const float*>(Xv) + (size_t)(m0 + (t0 >> 3)) * lda + (t0 & 7) * 4 + g0 * BK;
      pdl_wait();               // X belongs to the kernels before this one
      if (g0 < nkb) {
#pragma unroll
        for (int i = 0; i < BM / 16; ++i)
          pre[i] = (m0 + (t0 >> 3) + i * 16 < M) ? ldg_stream(reinterpret_cast<const float4*>(xp0 + (size_t)i * 16 * lda)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  {
    const int kq = nkb * (KB / 4);                          // float4s per vector (the tail block's missing channels: zeros)
    float4* d = reinterpret_cast<float4*>(s_bn);
    for (int i = threadIdx.x; i < 2 * kq; i += 320) {
      const int j = i < kq ? i : i - kq;
      d[i] = 4 * j < K ? __ldg(reinterpret_cast<const float4*>(i < kq ? bn_a : bn_b) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  unsigned long long* tr = g_gemm_trace ? g_gemm_trace + 16 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (tr && threadIdx.x == 0) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); tr[0] = smid; tr[1] = gtime(); }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&afull[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tr && threadIdx.x == 0) tr[2] = gtime();

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES, ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&bfull[s], B_BYTES);
        tma_load_2d(sB + s * B_BYTES, &tmB, &bfull[s], kb * KB, n0);
      }
      // L2 prefetch for the CTA that will take over this slot: the first four K blocks of the row tile `pf_stride` (= CTA slots of the
      // GPU) further on, issued when this CTA's K loop is nearly done -- the successor's first loads then hit L2 instead of waiting
      // for DRAM behind everyone else's stream
      if (pf_stride > 0 && (int)(blockIdx.y + pf_stride) * BM < M)
        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmX)), "r"(0), "r"((int)(blockIdx.y + pf_stride) * BM) : "memory");
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = XM == 2 ? make_idesc_bf16(BM, BN) : make_idesc_tf32(BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, ph = (kb / STAGES) & 1;
      mbar_wait(&afull[s], ph);
      mbar_wait(&bfull[s], ph);
      tc_fence_after();
      if (tr && kb == 0 && lane == 0) tr[3] = gtime();
      const uint64_t ad = make_desc_sw128(smem_u32(sA + s * A_BYTES));
      const uint64_t bd = make_desc_sw128(smem_u32(sB + s * B_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {            // four instructions of 32 K bytes either way
          if constexpr (XM == 2) umma_bf16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          else umma_tf32(tmem_base, ad + (uint64_t)(k * UMMA_K * 4 >> 4), bd + (uint64_t)(k * UMMA_K * 4 >> 4), idesc, (kb | k) != 0);
        }
        umma_commit(&empty[s]);
        if (kb == nkb - 1) umma_commit(accum);
      }
      __syncwarp();
    }
  } else {
    // ===== producers (warps 2..9, two groups of 4 alternating K blocks): thread = (16-byte chunk of the 128-byte row, row
    // r0 + 16 i).  Two register buffers used in turn (the loop is unrolled by two so that no buffer is ever copied): the
    // loads of the group's NEXT K block are in flight while the current one is transformed, four K blocks per CTA towards
    // HBM.  (fence.proxy.async.shared::cta does not wait for outstanding global loads -- measured, tools/fencebench.cu.)
    // Warps 2..5 run the epilogue once their K blocks are done.
    const int g = (warp - 2) >> 2;
    const int t = (threadIdx.x - 64) & 127;
    if constexpr (XM == 2) {
      // bf16 X, bf16 stage: a K block is 128 bytes of a row = eight 16-byte chunks of 8 channels; thread = (chunk, row r0 + 16 i)
      const int chunk = t & 7, r0 = t >> 3;
      const uint16_t* xp = reinterpret_cast<const uint16_t*>(Xv) + (size_t)(m0 + r0) * lda + chunk * 8;
      uint4 xa[BM / 16], xb[BM / 16];
      const float4* sc4 = reinterpret_cast<const float4*>(s_bn) + 2 * chunk;
      const float4* sh4 = sc4 + nkb * (KB / 4);
      auto load = [&](uint4* dst, int kb) {
        const int kcol = kb * KB;
        const bool in_k = kcol + chunk * 8 < K;
#pragma unroll
        for (int i = 0; i < BM / 16; ++i)
          dst[i] = (in_k && m0 + r0 + i * 16 < M) ? ldg_stream(reinterpret_cast<const uint4*>(xp + (size_t)i * 16 * lda + kcol)) : make_uint4(0u, 0u, 0u, 0u);
      };
      auto bnrelu2 = [](uint32_t w, float s0, float h0, float s1, float h1) {
        const float a = fmaxf(fmaf(__uint_as_float(w << 16), s0, h0), 0.f), b = fmaxf(fmaf(__uint_as_float(w & 0xffff0000u), s1, h1), 0.f);
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        return r;
      };
      auto produce = [&](const uint4* x, int kb) {
        const int s = kb % STAGES, ph = (kb / STAGES) & 1;
        const float4 sc0 = sc4[kb * (KB / 4)], sc1 = sc4[kb * (KB / 4) + 1], sh0 = sh4[kb * (KB / 4)], sh1 = sh4[kb * (KB / 4) + 1];
        mbar_wait(&empty[s], ph ^ 1);
        uint4* base = reinterpret_cast<uint4*>(sA + s * A_BYTES);
#pragma unroll
        for (int i = 0; i < BM / 16; ++i) {
          const int r = r0 + i * 16;
          const uint4 q = x[i];
          base[r * 8 + (chunk ^ (r & 7))] = make_uint4(bnrelu2(q.x, sc0.x, sh0.x, sc0.y, sh0.y), bnrelu2(q.y, sc0.z, sh0.z, sc0.w, sh0.w),
                                                       bnrelu2(q.z, sc1.x, sh1.x, sc1.y, sh1.y), bnrelu2(q.w, sc1.z, sh1.z, sc1.w, sh1.w));
        }
        fence_proxy_async();
        mbar_arrive(&afull[s]);
      };
      pdl_wait();
      if (g < nkb) load(xa, g);
      for (int kb = g; kb < nkb; kb += 2 * PG) {
        if (kb + PG < nkb) load(xb, kb + PG);
        produce(xa, kb);
        if (kb + PG < nkb) {
          if (kb + 2 * PG < nkb) load(xa, kb + 2 * PG);
          produce(xb, kb + PG);
        }
      }
    } else if constexpr (XH) {
      // bf16 X: a K block is 64 bytes of a row = four 16-byte chunks of 8 channels; thread = (chunk, row r0 + 32 i)
      const int chunk = t & 3, r0 = t >> 2;
      const uint16_t* xp = reinterpret_cast<const uint16_t*>(Xv) + (size_t)(m0 + r0) * lda + chunk * 8;
      uint4 xa[BM / 32], xb[BM / 32];
      const float4* sc4 = reinterpret_cast<const float4*>(s_bn) + 2 * chunk;
      const float4* sh4 = sc4 + nkb * (BK / 4);
      auto load = [&](uint4* dst, int kb) {
        const int kcol = kb * BK;
#pragma unroll
        for (int i = 0; i < BM / 32; ++i)
          dst[i] = (m0 + r0 + i * 32 < M) ? ldg_stream(reinterpret_cast<const uint4*>(xp + (size_t)i * 32 * lda + kcol)) : make_uint4(0u, 0u, 0u, 0u);
      };
      auto produce = [&](const uint4* x, int kb) {
        const int s = kb % STAGES, ph = (kb / STAGES) & 1;
        const float4 sc0 = sc4[kb * (BK / 4)], sc1 = sc4[kb * (BK / 4) + 1], sh0 = sh4[kb * (BK / 4)], sh1 = sh4[kb * (BK / 4) + 1];
        mbar_wait(&empty[s], ph ^ 1);
        float4* base = reinterpret_cast<float4*>(sA + s * A_BYTES);
#pragma unroll
        for (int i = 0; i < BM / 32; ++i) {
          const int r = r0 + i * 32;
          const uint4 q = x[i];                 // element 2j in the low half of word j, 2j + 1 in the high half
          float4 y0, y1;
          y0.x = to_tf32(fmaxf(fmaf(__uint_as_float(q.x << 16), sc0.x, sh0.x), 0.f));
          y0.y = to_tf32(fmaxf(fmaf(__uint_as_float(q.x & 0xffff0000u), sc0.y, sh0.y), 0.f));
          y0.z = to_tf32(fmaxf(fmaf(__uint_as_float(q.y << 16), sc0.z, sh0.z), 0.f));
          y0.w = to_tf32(fmaxf(fmaf(__uint_as_float(q.y & 0xffff0000u), sc0.w, sh0.w), 0.f));
          y1.x = to_tf32(fmaxf(fmaf(__uint_as_float(q.z << 16), sc1.x, sh1.x), 0.f));
          y1.y = to_tf32(fmaxf(fmaf(__uint_as_float(q.z & 0xffff0000u), sc1.y, sh1.y), 0.f));
          y1.z = to_tf32(fmaxf(fmaf(__uint_as_float(q.w << 16), sc1.z, sh1.z), 0.f));
          y1.w = to_tf32(fmaxf(fmaf(__uint_as_float(q.w & 0xffff0000u), sc1.w, sh1.w), 0.f));
          base[r * 8 + ((2 * chunk) ^ (r & 7))] = y0;
          base[r * 8 + ((2 * chunk + 1) ^ (r & 7))] = y1;
        }
        fence_proxy_async();
        mbar_arrive(&afull[s]);
      };
      pdl_wait();
      if (g < nkb) load(xa, g);
      for (int kb = g; kb < nkb; kb += 2 * PG) {
        if (kb + PG < nkb) load(xb, kb + PG);
        produce(xa, kb);
        if (kb + PG < nkb) {
          if (kb + 2 * PG < nkb) load(xa, kb + 2 * PG);
          produce(xb, kb + PG);
        }
      }
    } else {
    const float* X = reinterpret_cast<const float*>(Xv);
    const int chunk = t & 7, r0 = t >> 3;
    const float* xp = X + (size_t)(m0 + r0) * lda + chunk * 4;
    float4 xa[BM / 16], xb[BM / 16];                // 8 rows of X each
    const float4* sc4 = reinterpret_cast<const float4*>(s_bn) + chunk;
    const float4* sh4 = sc4 + nkb * (BK / 4);
    auto load = [&](float4* dst, int kb) {
      const int kcol = kb * BK;
#pragma unroll
      for (int i = 0; i < BM / 16; ++i)
        dst[i] = (m0 + r0 + i * 16 < M) ? ldg_stream(reinterpret_cast<const float4*>(xp + (size_t)i * 16 * lda + kcol)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto produce = [&](const float4* x, int kb) {
      const int s = kb % STAGES, ph = (kb / STAGES) & 1;
      const float4 sc = sc4[kb * (BK / 4)], sh = sh4[kb * (BK / 4)];
      mbar_wait(&empty[s], ph ^ 1);
      float4* base = reinterpret_cast<float4*>(sA + s * A_BYTES);
#pragma unroll
      for (int i = 0; i < BM / 16; ++i) {
        const int r = r0 + i * 16;
        float4 y;
        y.x = to_tf32(fmaxf(fmaf(x[i].x, sc.x, sh.x), 0.f));
        y.y = to_tf32(fmaxf(fmaf(x[i].y, sc.y, sh.y), 0.f));
        y.z = to_tf32(fmaxf(fmaf(x[i].z, sc.z, sh.z), 0.f));
        y.w = to_tf32(fmaxf(fmaf(x[i].w, sc.w, sh.w), 0.f));
        base[r * 8 + (chunk ^ (r & 7))] = y;
      }
      fence_proxy_async();
      mbar_arrive(&afull[s]);
    };
    // (griddepcontrol.wait was executed before the early loads at the top: X, the segment sums and C belong to the kernels before
    // this one; the weights do not)
    if (g < nkb) {
#pragma unroll
      for (int i = 0; i < BM / 16; ++i) xa[i] = pre[i];
    }
    for (int kb = g; kb < nkb; kb += 2 * PG) {
      if (kb + PG < nkb) load(xb, kb + PG);
      produce(xa, kb);
      if (kb + PG < nkb) {
        if (kb + 2 * PG < nkb) load(xa, kb + 2 * PG);
        produce(xb, kb + PG);
      }
    }
    }
    pdl_trigger();              // every CTA has its loads behind it: the next kernel may start setting up
    {
      // ===== epilogue, all eight producer warps: group g takes the 32-column chunks c = g, g + 2, ... (a warp may only read
      // TMEM lanes 32 (warp % 4) ..., which both groups cover).  The C tile leaves through the idle A / B stages and TMA
      // tensor stores: one 128 x 32 box per chunk.
      if (tr && threadIdx.x == 64) tr[4] = gtime();
      mbar_wait(accum, 0);
      tc_fence_after();
      if (tr && threadIdx.x == 64) tr[5] = gtime();
      const int q = warp & 3;
      const int row = m0 + q * 32 + lane;
      const int i = q * 32 + lane;
      const bool issuer = (warp == 2 || warp == 6) && lane == 0;
      // staging: the idle A and B stages are one contiguous region, cut into 16 KB buffers; chunk ci uses buffer ci % NBUF
      // (NBUF even, so a buffer is only ever reused by the group that used it before)
      constexpr int NBUF = (STAGES * (A_BYTES + B_BYTES) / (BM * 128)) & ~1;
      static_assert(NBUF >= 2, "staging needs two buffers");
#pragma unroll 1
      for (int c = g * 32; c < BN; c += 32 * PG) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        if (tr && threadIdx.x == 64 && c < 128) tr[8 + (c >> 6) * 4] = gtime();
        epi(row, n0 + c, v);                               // transforms v in place (epi.out == nullptr: no store)
        if (tr && threadIdx.x == 64 && c < 128) tr[9 + (c >> 6) * 4] = gtime();
        if (epi_skips_c(epi, 0)) continue;                 // the functor wrote the tile (bf16 u): no fp32 C store
        const int ci = c >> 5;
        uint8_t* stage = smem + (ci % NBUF) * (BM * 128);
        if (ci >= NBUF) {                                  // buffer reuse: the store that last used it must have read it
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
        }
        float4* so = reinterpret_cast<float4*>(stage) + i * 8;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) so[jj ^ (i & 7)] = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
        fence_proxy_async();
        if (g == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
        if (tr && threadIdx.x == 64 && c < 128) tr[10 + (c >> 6) * 4] = gtime();
        if (issuer) {
          tma_store_2d(&tmC, stage, n0 + c, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (tr && threadIdx.x == 64 && c < 128) tr[11 + (c >> 6) * 4] = gtime();
      }
      if (tr && threadIdx.x == 64) tr[6] = gtime();
      if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tr && threadIdx.x == 0) tr[7] = gtime();
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));      // upper half <- first source, lower half <- second
  return r;
}

// ---------------------------------------------------------------- persistent GEMM
//   C[m][n] = sum_k A[m][k] * W[n][k] + bias[n],   one CTA per SM looping over 128 x BN output tiles (n fastest).
// The shared-memory stage ring and the barriers' phases run on across tiles, and the fp32 accumulator is double-buffered
// in TMEM (2 x BN columns), so the epilogue of tile i overlaps the K loop of tile i+1 and the per-CTA set-up (barrier init,
// TMEM allocation, descriptor prefetch) is paid once per SM instead of once per tile.  A and W tiles come by TMA (tf32
// rounding of A in the tensor map); warps 0 = TMA, 1 = MMA, 2..5 = epilogue.  The C tile leaves through two 16 KB staging
// buffers and TMA tensor stores (one 128 x 32 box per accumulator chunk) instead of per-thread 16-byte stores scattered over
// 32 rows.  (A second epilogue group, as in the convolution kernels, made this one slower: it is bound by its 4 KB per row
// of C writes, and the extra staging buffers cost a pipeline stage.)
// OUT16: C is stored as bf16 (the bf16 mode's LSTM input projections: the kernel is bound by its C writes, 4 KB -> 2 KB per row);
// one staging tile then carries 64 columns (128 bytes of bf16 per row) and tmC is a bf16 tensor map with a {64, 128} box.
template <int BN, int STAGES, bool OUT16 = false>
__global__ void __launch_bounds__(192, 1)
pgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
             const float* __restrict__ bias, int n_tiles_n, int n_tiles, int nkb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
  static_assert(2 * BN <= 512, "two accumulators must fit TMEM");
  constexpr int TMEM_COLS = 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bfull = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));   // TMA bytes (A and B)
  uint64_t* empty = bfull + STAGES;          // MMA commit
  uint64_t* tfull = empty + STAGES;          // [2] accumulator ready
  uint64_t* tempty = tfull + 2;              // [2] accumulator drained (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* sC = smem + STAGES * (A_BYTES + B_BYTES) + 1024;    // [2][128 rows x 128 B] C staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&bfull[s], A_BYTES + B_BYTES);
          tma_load_2d(sA + s * A_BYTES, &tmA, &bfull[s], kb * BK, m0);
          tma_load_2d(sB + s * B_BYTES, &tmB, &bfull[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    int it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int a = ti & 1, pa = (ti >> 1) & 1;
      mbar_wait(&tempty[a], pa ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + a * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&bfull[s], ph);
        tc_fence_after();
        const uint64_t ad = make_desc_sw128(smem_u32(sA + s * A_BYTES));
        const uint64_t bd = make_desc_sw128(smem_u32(sB + s * B_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_tf32(d, ad + (uint64_t)(k * UMMA_K * 4 >> 4), bd + (uint64_t)(k * UMMA_K * 4 >> 4), idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
          if (kb == nkb - 1) umma_commit(&tfull[a]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int i = q * 32 + lane;
    int ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
      const int a = ti & 1, pa = (ti >> 1) & 1;
      mbar_wait(&tfull[a], pa);
      tc_fence_after();
      if constexpr (OUT16) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          float v[64];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c), v);
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c + 32), v + 32);
          if (c + 64 >= BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);
          }
          const int buf = (c >> 6) & 1;
          const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c);
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bb = __ldg(b4 + j);
            pk[2 * j] = pack_bf16x2(v[4 * j] + bb.x, v[4 * j + 1] + bb.y);
            pk[2 * j + 1] = pack_bf16x2(v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w);
          }
          if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
          uint4* so = reinterpret_cast<uint4*>(sC + buf * (BM * 128)) + i * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) so[j ^ (i & 7)] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (warp == 2 && lane == 0) {
            tma_store_2d(&tmC, sC + buf * (BM * 128), n0 + c, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c), v);
        if (c + 32 >= BN) {                    // all of this warp's accumulator columns are in registers: release the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[a]);
        }
        const int buf = (c >> 5) & 1;
        {
          const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c);      // eight LSU requests instead of 32
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 bb = __ldg(b4 + j); v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last used `buf` has read it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float4* so = reinterpret_cast<float4*>(sC + buf * (BM * 128)) + i * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) so[j ^ (i & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          tma_store_2d(&tmC, sC + buf * (BM * 128), n0 + c, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      }
    }
    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

// 2-D fp32 row-major [rows][cols] with leading dimension ld (floats); box = 32 columns x box_rows rows, 128B swizzle
CUtensorMap make_map_2d(const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool round_tf32);
CUtensorMap make_map_2d_plain(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows, bool bf16 = false);   // no swizzle (L2 prefetch)
CUtensorMap make_map_2d_bf16(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

template <int BN, int STAGES, class Pro, class Epi>
inline void tgemm(Launches& L, cudaStream_t st, const char* tag, const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                  const TapMap& tap, int ntaps, Pro pro, Epi epi) {
  if (M <= 0 || N <= 0) return;
  auto kern = tgemm_kernel<BN, STAGES, Pro, Epi>;
  constexpr int SMEM = smem_bytes(BN, STAGES);
  ensure_max_smem(kern, SMEM);
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  const int nkb = tap.cpb * ntaps;
  Scope sc(L, st, tag, 2.0 * M * N * K, 4.0 * ((double)M * (K / ntaps) + (double)M * N));   // algorithmic bytes: A once + C
  kern<<<grid, Pro::kOn ? 320 : 192, SMEM, st>>>(tmA, tmB, nkb, tap, pro, epi);
}

// persistent GEMM launcher: C[M][N] (row-major, leading dimension ldc) = A . W^T + bias, C written with TMA tensor stores
int sm_count();
template <int BN, int STAGES, bool OUT16 = false>
inline void pgemm_bias_tma(Launches& L, cudaStream_t st, const char* tag, const CUtensorMap& tmA, const CUtensorMap& tmB, void* C, int ldc,
                           const float* bias, int M, int N, int K) {
  if (M <= 0 || N <= 0) return;
  auto kern = pgemm_kernel<BN, STAGES, OUT16>;
  constexpr int SMEM = smem_bytes(BN, STAGES) + 1024 + 2 * BM * 128;
  ensure_max_smem(kern, SMEM);
  const int tn = (N + BN - 1) / BN, tiles = ((M + BM - 1) / BM) * tn;
  CUtensorMap tmC = OUT16 ? make_map_2d_bf16(C, M, N, ldc, BM) : make_map_2d(static_cast<float*>(C), M, N, ldc, BM, false);
  Scope sc(L, st, tag, 2.0 * M * N * K, 4.0 * (double)M * K + (OUT16 ? 2.0 : 4.0) * (double)M * N);
  kern<<<tiles < sm_count() ? tiles : sm_count(), 192, SMEM, st>>>(tmA, tmB, tmC, bias, tn, tiles, (K + BK - 1) / BK);
}

template <int BN, int STAGES, class Epi, int XM = 0>
inline void tgemm_bnrelu(Launches& L, cudaStream_t st, const char* tag, const void* X, int lda, const float* bn_a, const float* bn_b,
                         const CUtensorMap& tmB, float* C, int ldc, int M, int N, int K, Epi epi, bool pdl = false) {
  if (M <= 0 || N <= 0) return;
  auto kern = tgemm_bnrelu_kernel<BN, STAGES, Epi, XM>;
  constexpr int SMEM = smem_bytes(BN, STAGES) + 2 * 1024 * 4;      // + BN scale / shift of up to 1024 columns
  if (K > 1024) { fprintf(stderr, "libcbx: tgemm_bnrelu supports K <= 1024\n"); return; }
  ensure_max_smem(kern, SMEM);
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  Scope sc(L, st, tag, 2.0 * M * N * K, (XM ? 2.0 : 4.0) * (double)M * K + 4.0 * (double)M * N);
  CUtensorMap tmC = make_map_2d(C, M, N, ldc, BM, false);
  constexpr int KB = XM == 2 ? 64 : BK;
  // option bn_prefetch: a plain tensor map of X, box = 512 bytes of a row x 128 rows, for the successor-tile L2 prefetch; the
  // successor of a CTA in its slot is the row tile (CTA slots of the GPU) / (column tiles) further on (2 CTAs per SM)
  CUtensorMap tmX = tmC;
  int pf_stride = 0;
  if (L.bn_prefetch > 0 && K >= 128) {
    tmX = make_map_2d_plain(X, M, K, lda, XM ? 256 : 128, BM, XM != 0);          // 512 bytes of every row either way: four (bf16 X: eight / four) K blocks
    pf_stride = (2 * sm_count() + (int)grid.x - 1) / (int)grid.x;
  }
  launch_pdl(kern, grid, dim3(320), SMEM, st, pdl, X, lda, M, K, bn_a, bn_b, tmB, tmC, (K + KB - 1) / KB, epi, tmX, pf_stride);
}

}  // namespace tc
}  // namespace cbx


// Host side of the tcgen05 GEMM engine: tensor-map encoding (driver entry point fetched through the runtime, so the
// library has no link-time dependency on libcuda) and the raw GEMM entry points used by the kernel unit tests.
#include <vector>

#include "tc.cuh"
#include "epi.cuh"

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>

namespace cbx {

void ensure_max_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;          // (kernel, device) -> bytes granted
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  int& have = done[{kernel, dev}];
  if (have < bytes) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); have = bytes; }
}

namespace tc {

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

CUtensorMap make_map_2d(const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool round_tf32) {
  CUtensorMap m;
  std::memset(&m, 0, sizeof m);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn fn = encode_fn();
  CUresult r = fn ? fn(&m, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                  : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld\n", (int)r,
                                 (long long)rows, (long long)cols, (long long)ld);
  return m;
}

CUtensorMap make_map_2d_plain(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows, bool bf16) {
  CUtensorMap m;
  std::memset(&m, 0, sizeof m);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (bf16 ? 2 : sizeof(float))};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn fn = encode_fn();
  CUresult r = fn ? fn(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                  : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled (plain) failed (%d) rows=%lld cols=%lld ld=%lld\n", (int)r,
                                 (long long)rows, (long long)cols, (long long)ld);
  return m;
}

// bf16 [rows][cols] (leading dimension ld elements): box {64 columns = 128 bytes, box_rows}, 128-byte swizzle
CUtensorMap make_map_2d_bf16(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  CUtensorMap m;
  std::memset(&m, 0, sizeof m);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn fn = encode_fn();
  CUresult r = fn ? fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                  : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled (bf16) failed (%d) rows=%lld cols=%lld ld=%lld\n", (int)r,
                                 (long long)rows, (long long)cols, (long long)ld);
  return m;
}

}  // namespace tc
}  // namespace cbx

#ifdef CBX_DEV_TOOLS   // kernel unit-test / timing entry points for the scripts under tools/: NOT part of the product library (build.py, CBX_DEV_TOOLS=1)
using namespace cbx;

// same functor under a type of this translation unit: the kernel template is then instantiated HERE (next to the trace pointer
// this file sets) instead of being merged with xv.cu's instantiation
namespace { struct EpiSegsumProbe : tc::EpiBiasReluMaskSegsum { }; }

extern "C" int cbx_test_tgemm(cbx_ctx* c, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                              int M, int N, int K, const float* bias, const float* pro_a, const float* pro_b, int variant,
                              int shift1, void* stream) {
  if (!c) return CBX_ERR_ARG;
  DeviceGuard dev_guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!tc::encode_fn()) { c->err = "cuTensorMapEncodeTiled unavailable"; return CBX_ERR_CUDA; }
  tc::TapMap tap = tc::plain_map(K);
  int ntaps = 1;
  if (variant == 3) {           // two taps: rows m and m+shift1, K columns each (W is [N][2K])
    ntaps = 2; tap.shift[1] = shift1;
  }
  const bool pro = variant == 1;
  CUtensorMap tmA = tc::make_map_2d(A, M, K, lda, tc::BM, !pro);
  tc::EpiBias epi{C, (int)ldc, bias, M};
  if (variant == 0 || variant == 3) {
    CUtensorMap tmB = tc::make_map_2d(W, N, (int64_t)K * ntaps, ldw, 128, true);
    tc::tgemm<128, 3>(c->launches, st, "test_tgemm", tmA, tmB, M, N, K * ntaps, tap, ntaps, tc::NoPrologue{}, epi);
  } else if (variant == 1) {
    CUtensorMap tmB = tc::make_map_2d(W, N, K, ldw, 128, true);
    tc::tgemm<128, 3>(c->launches, st, "test_tgemm_pro", tmA, tmB, M, N, K, tap, 1, tc::BnReluPrologue{pro_a, pro_b}, epi);
  } else if (variant == 2) {
    CUtensorMap tmB = tc::make_map_2d(W, N, K, ldw, 32, true);
    tc::tgemm<32, 4>(c->launches, st, "test_tgemm_n32", tmA, tmB, M, N, K, tap, 1, tc::NoPrologue{}, epi);
  } else if (variant == 6) {    // set / clear the timeline probe of the pre-activation GEMM: C = trace buffer (16 x u64 per CTA) or NULL
    unsigned long long* p = reinterpret_cast<unsigned long long*>(C);
    CBX_CUDA_OK(c, cudaMemcpyToSymbol(tc::g_gemm_trace, &p, sizeof(p)));
  } else if (variant == 7) {    // the bottleneck layer's production epilogue (bias + ReLU + mask + segment sums, TMA-stored C)
    static int32_t* row_seg = nullptr; static unsigned long long* seg_sum = nullptr; static int cap = 0;
    if (cap < M) {
      cudaFree(row_seg); cudaFree(seg_sum);
      std::vector<int32_t> h(M);
      for (int i = 0; i < M; ++i) h[i] = (i % 501 == 500) ? -1 : (i / 501) * 5 + (i % 501) / 100;     // 500-frame clips, one guard row, 100-frame segments
      CBX_CUDA_OK(c, cudaMalloc((void**)&row_seg, sizeof(int32_t) * M));
      CBX_CUDA_OK(c, cudaMalloc((void**)&seg_sum, sizeof(unsigned long long) * 128 * (size_t)(M / 100 + 8)));
      CBX_CUDA_OK(c, cudaMemcpy(row_seg, h.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice));
      cap = M;
    }
    CUtensorMap tmB = tc::make_map_2d(W, N, K, ldw, 128, true);
    tc::tgemm_bnrelu<128, 2>(c->launches, st, "test_tgemm_bnrelu", A, (int)lda, pro_a, pro_b, tmB, C, (int)ldc, M, N, K,
                             EpiSegsumProbe{{nullptr, (int)ldc, bias, row_seg, seg_sum, M}});
  } else if (variant == 4) {    // pre-activation GEMM (register producers), single CTA
    CUtensorMap tmB = tc::make_map_2d(W, N, K, ldw, 128, true);
    tc::tgemm_bnrelu<128, 2>(c->launches, st, "test_tgemm_bnrelu", A, (int)lda, pro_a, pro_b, tmB, C, (int)ldc, M, N, K, epi);
  } else {
    c->err = "bad variant"; return CBX_ERR_ARG;
  }
  CBX_CUDA_OK(c, cudaGetLastError());
  return CBX_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Unit test: A operand descriptors that start at an arbitrary 128-byte row of a swizzled tile (the implicit-GEMM convs
// read their taps as row-shifted views of ONE halo tile in shared memory).  C[128][32] = A[shift .. shift+128) . W^T
namespace cbx { namespace tc {
__global__ void __launch_bounds__(128) shift_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                         float* C, int shift, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                  // [256 rows][128 B]
  uint8_t* sW = smem + 256 * 128;      // [32 rows][128 B]
  uint64_t* full = reinterpret_cast<uint64_t*>(sW + 32 * 128);
  uint64_t* accum = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(accum, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, 256 * 128 + 32 * 128);
    tma_load_2d(sA, &tmA, full, 0, 0);
    tma_load_2d(sA + 128 * 128, &tmA, full, 0, 128);
    tma_load_2d(sW, &tmW, full, 0, 0);
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + shift * 128;
    uint64_t ad = make_desc_sw128(a_addr);
    if (use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bd = make_desc_sw128(smem_u32(sW));
    constexpr uint32_t idesc = make_idesc_tf32(128, 32);
    for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, ad + (uint64_t)(k * 32 >> 4), bd + (uint64_t)(k * 32 >> 4), idesc, k != 0);
    umma_commit(accum);
  }
  mbar_wait(accum, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  for (int i = 0; i < 32; ++i) C[(size_t)(warp * 32 + lane) * 32 + i] = v[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}
}}  // namespace cbx::tc

extern "C" int cbx_test_shift_gemm(cbx_ctx* c, const float* A /*[256][32]*/, const float* W /*[32][32]*/, float* C /*[128][32]*/,
                                   int shift, int use_base_offset) {
  if (!c) return CBX_ERR_ARG;
  DeviceGuard dev_guard(c->device);
  CUtensorMap tmA = tc::make_map_2d(A, 256, 32, 32, 128, true);
  CUtensorMap tmW = tc::make_map_2d(W, 32, 32, 32, 32, true);
  const int smem = 256 * 128 + 32 * 128 + 1024 + 64;
  tc::shift_gemm_kernel<<<1, 128, smem>>>(tmA, tmW, C, shift, use_base_offset);
  CBX_CUDA_OK(c, cudaDeviceSynchronize());
  return CBX_OK;
}
#endif  // CBX_DEV_TOOLS

// Host side of the tcgen05 GEMM engine: tensor-map encoding (driver entry point fetched through the runtime, so the
// library has no link-time dependency on libcuda) and the raw GEMM entry points used by the kernel unit tests.
#include "tc.cuh"
#include "epi.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>

namespace cbx {
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

CUtensorMap make_map_fcm(const float* base, int rows, int F, int P, int F_out, int BR) {
  CUtensorMap m;
  std::memset(&m, 0, sizeof m);
  cuuint64_t dims[4] = {32, (cuuint64_t)P, (cuuint64_t)(F / P), (cuuint64_t)rows};
  cuuint64_t strides[3] = {128, (cuuint64_t)128 * P, (cuuint64_t)128 * F};
  cuuint32_t box[4] = {32, 1, (cuuint32_t)F_out, (cuuint32_t)BR};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  EncodeTiledFn fn = encode_fn();
  CUresult r = fn ? fn(&m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                  : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) fprintf(stderr, "libcbx: cuTensorMapEncodeTiled(fcm) failed (%d) rows=%d F=%d P=%d\n", (int)r, rows, F, P);
  return m;
}

}  // namespace tc
}  // namespace cbx

using namespace cbx;

extern "C" int cbx_test_tgemm(cbx_ctx* c, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                              int M, int N, int K, const float* bias, const float* pro_a, const float* pro_b, int variant,
                              int shift1, void* stream) {
  if (!c) return CBX_ERR_ARG;
  cudaSetDevice(c->device);
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
  } else {
    c->err = "bad variant"; return CBX_ERR_ARG;
  }
  CBX_CUDA_OK(c, cudaGetLastError());
  return CBX_OK;
}

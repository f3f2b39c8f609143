// Latency of fence.proxy.async (+ variants) after shared-memory stores, with and without global loads in flight.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float4* g, float* out, long long* cyc, int mode) {
  extern __shared__ float4 sm[];
  const int t = threadIdx.x;
  float4 acc = make_float4(0, 0, 0, 0);
  long long tot = 0;
  for (int it = 0; it < 64; ++it) {
    float4 ld = make_float4(0, 0, 0, 0);
    if (mode & 1) ld = __ldg(g + ((size_t)blockIdx.x * 4096 + it * 64 + t) * 97 % (1 << 22));   // a global load in flight
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[(t + i * blockDim.x) & 2047] = make_float4(it, t, i, 1);
    long long t0 = clock64();
    if (mode & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (mode & 4) asm volatile("fence.proxy.async;" ::: "memory");
    if (mode & 8) __threadfence();
    if (mode & 16) __threadfence_block();
    long long t1 = clock64();
    tot += t1 - t0;
    acc.x += ld.x + sm[(t * 7 + it) & 2047].x;
  }
  out[blockIdx.x * blockDim.x + t] = acc.x;
  if (t == 0 && blockIdx.x == 0) *cyc = tot / 64;
}
int main() {
  float4* g; float* o; long long* c;
  cudaMalloc(&g, sizeof(float4) << 22); cudaMemset(g, 0, sizeof(float4) << 22); cudaMalloc(&o, 1 << 22); cudaMalloc(&c, 8);
  const char* names[] = {"(nothing)", "fence.proxy.async.shared::cta", "fence.proxy.async", "__threadfence", "__threadfence_block"};
  const int modes[] = {0, 2, 4, 8, 16};
  for (int ldm = 0; ldm < 2; ++ldm)
    for (int i = 0; i < 5; ++i) {
      k<<<148, 128, 32768>>>(g, o, c, modes[i] | ldm); cudaDeviceSynchronize();
      k<<<148, 128, 32768>>>(g, o, c, modes[i] | ldm); cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("%-32s %s: %lld cycles\n", names[i], ldm ? "with a global load in flight" : "no loads in flight         ", h);
    }
  return 0;
}

// Calibrate clock64() against CUDA-event time (is the SM counter the SM clock?).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long ticks, long long* out) {
  long long t0 = clock64();
  unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  while (clock64() - t0 < ticks) {}
  unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = clock64() - t0; out[1] = (long long)(g1 - g0); }
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 4; ++rep) {
    for (int grid : {1, 148}) {
      cudaEventRecord(e0); spin<<<grid, 128>>>(200000000LL, d); cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("grid %3d: %lld ticks in %.3f ms (events) / %.3f ms (globaltimer) -> %.1f MHz\n", grid, h[0], ms, h[1] / 1e6, h[0] / (ms * 1e3));
    }
  }
  return 0;
}

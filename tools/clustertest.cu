// Where do the two CTAs of a (2,1,1) cluster land?  nvcc -arch=sm_100a -o tools/bin/clustertest tools/clustertest.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 2) kx(int* o) {
  extern __shared__ char sm[];
  unsigned smid, rank;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) o[blockIdx.x] = (int)(smid | (rank << 16));
  // stay resident a while so that co-residency is exercised
  long long t0 = clock64(); while (clock64() - t0 < 200000) {}
}
int main() {
  const int n = 2000;
  int* d; cudaMalloc(&d, 4 * n); cudaMemset(d, 0xff, 4 * n);
  cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  kx<<<dim3(n), 320, 100 * 1024>>>(d);
  printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  static int h[n]; cudaMemcpy(h, d, 4 * n, cudaMemcpyDeviceToHost);
  int bad_pair = 0, odd_leader = 0, maxsm = 0;
  for (int i = 0; i < n; i += 2) {
    int s0 = h[i] & 0xffff, s1 = h[i + 1] & 0xffff, r0 = h[i] >> 16, r1 = h[i + 1] >> 16;
    if (r0 != 0 || r1 != 1) printf("rank order unexpected at %d: %d %d\n", i, r0, r1);
    if ((s0 ^ 1) != s1) ++bad_pair;
    if (s0 & 1) ++odd_leader;
    if (s0 > maxsm) maxsm = s0; if (s1 > maxsm) maxsm = s1;
  }
  printf("clusters %d: peer != smid^1 in %d, leader on odd smid in %d, max smid %d\n", n / 2, bad_pair, odd_leader, maxsm);
  for (int i = 0; i < 12; i += 2) printf("  cluster %d: smid %d / %d\n", i / 2, h[i] & 0xffff, h[i + 1] & 0xffff);
}

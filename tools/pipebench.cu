// Per-SMSP throughput of MUFU.EX2 / MUFU.RCP / SHFL / FSEL / FFMA on this GPU (cycles per warp instruction).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0.5f + 0.001f * (threadIdx.x + i);
  const bool p = threadIdx.x & 1;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
      if (OP == 3) asm volatile("{.reg .pred q; setp.ne.s32 q, %2, 0; selp.f32 %0, %0, %1, q;}" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "r"((int)p));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 5) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 1 << 22); cudaMalloc(&c, 8);
  const char* names[] = {"MUFU.EX2", "MUFU.RCP", "SHFL.BFLY", "FSEL", "FFMA", "MUFU.TANH"};
  for (int warps : {4, 8, 16}) {
    for (int op = 0; op < 6; ++op) {
      const int iters = 2000;
      auto launch = [&](int o_) {
        switch (o_) { case 0: k<0><<<148, warps * 32>>>(o, c, iters); break; case 1: k<1><<<148, warps * 32>>>(o, c, iters); break;
          case 2: k<2><<<148, warps * 32>>>(o, c, iters); break; case 3: k<3><<<148, warps * 32>>>(o, c, iters); break;
          case 4: k<4><<<148, warps * 32>>>(o, c, iters); break; case 5: k<5><<<148, warps * 32>>>(o, c, iters); break; }
      };
      launch(op); cudaDeviceSynchronize(); launch(op); cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      const double per_smsp_instr = (double)iters * 8 * (warps / 4.0);
      printf("warps/SM %2d  %-10s %6.2f cycles per warp-instruction per SMSP\n", warps, names[op], h / per_smsp_instr);
    }
  }
  return 0;
}

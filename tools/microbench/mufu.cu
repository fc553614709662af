// Microbenchmark: throughput of MUFU.EX2 / F2FP / FFMA / FMNMX per SM sub-partition on this GPU (cycles per warp instruction),
// as a function of warps per scheduler.   nvcc -arch=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 15])); acc ^= r; }
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + 1) & 15]), "f"(x[(i + 2) & 15]));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[(i + 3) & 15]));
      if (OP == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); if (i & 1) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[i - 1])); acc ^= r; } }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[] = {"MUFU.EX2", "F2FP.BF16x2", "FFMA", "FMNMX", "EX2 + F2FP per 2"};
  for (int op = 0; op < 5; ++op)
    for (int warps = 4; warps <= 16; warps *= 2) {     // 1, 2, 4 warps per scheduler
      const int iters = 256;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
        if (op == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
        if (op == 2) k<2><<<1, warps * 32>>>(out, cyc, iters);
        if (op == 3) k<3><<<1, warps * 32>>>(out, cyc, iters);
        if (op == 4) k<4><<<1, warps * 32>>>(out, cyc, iters);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double per = (double)h / (iters * 16.0 * (warps / 4));
      printf("%-18s warps/scheduler %d : %.2f cycles per warp-instruction per scheduler (%s)\n", names[op], warps / 4, per, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}

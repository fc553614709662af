// Microbenchmark: tensor-memory read throughput (tcgen05.ld 32x32b.x32 = 4 KB per warp instruction) on this GPU, per SM, as a
// function of warps per CTA, CTAs per SM and loads per tcgen05.wait::ld.   nvcc -arch=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
      "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
    : "r"(taddr) : "memory");
}
template <int PER_WAIT>
__global__ void k(long long* cyc, uint32_t* sink, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const uint32_t base = tptr + (((warp & 3) * 32u) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[PER_WAIT][32];
#pragma unroll
    for (int p = 0; p < PER_WAIT; ++p) ld32(base + ((it * PER_WAIT + p) & 3) * 32 + (warp >> 2) * 128 % 256, v[p]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int p = 0; p < PER_WAIT; ++p) acc ^= v[p][it & 31];
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tptr) : "memory");
}
int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 8); cudaMalloc(&sink, 1 << 22);
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  for (int per = 1; per <= 4; per *= 4)
    for (int warps = 4; warps <= 8; warps *= 2)
      for (int ctas = 1; ctas <= 2; ++ctas) {
        const int iters = 2048;
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          if (per == 1) k<1><<<nsm * ctas, warps * 32>>>(cyc, sink, iters); else k<4><<<nsm * ctas, warps * 32>>>(cyc, sink, iters);
          cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        }
        const double loads = (double)iters * per * warps * ctas;          // 4 KB each, per SM
        printf("loads/wait %d  warps/CTA %d  CTAs/SM %d : %.1f cycles per x32 load per warp, %.0f B/cycle/SM  (%s)\n", per, warps, ctas,
               (double)h / (iters * per), loads * 4096.0 / h, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}

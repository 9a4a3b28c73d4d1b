// Micro-benchmark: how fast can the epilogue warps of one CTA drain TMEM?  (tcgen05.ld 32x32b, x16 / x32 / x64, 4..28 warps,
// one or two loads in flight per warp.)  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o ldtm_rate ldtm_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int X>
__device__ __forceinline__ void ld(uint32_t addr, int (&r)[X]);
template <>
__device__ __forceinline__ void ld<16>(uint32_t addr, int (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t addr, int (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
}

template <int X, int INFLIGHT>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* clocks, int* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  int acc = 0;
  int r0[X], r1[X];
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)(((warp >> 2) * X * 2 + i * 64) & 511) & ~(uint32_t)(X - 1);
    ld<X>(base + (col & (512 - X)), r0);
    if (INFLIGHT == 2) ld<X>(base + ((col + X) & (512 - X)), r1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < X; ++j) acc ^= r0[j];
    if (INFLIGHT == 2) {
#pragma unroll
      for (int j = 0; j < X; ++j) acc ^= r1[j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int X, int INFLIGHT>
void run(int warps) {
  long long* d;
  int* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  k<X, INFLIGHT><<<148, warps * 32>>>(iters, d, sink);
  k<X, INFLIGHT><<<148, warps * 32>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * warps * INFLIGHT * X * 32 * 4;
  printf("x%-2d inflight %d warps %2d: %7.1f B/clk/SM  (%lld clk, %s)\n", X, INFLIGHT, warps, bytes / (double)h[0], h[0], cudaGetErrorString(e));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16, 24, 28}) { run<16, 1>(w); run<16, 2>(w); run<32, 1>(w); }
  return 0;
}

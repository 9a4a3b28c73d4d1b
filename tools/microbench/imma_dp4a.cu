// Microbenchmark: legacy mma.sync int8 (IMMA m16n8k32) and dp4a issue rates on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void imma_kernel(int iters, int* out) {
  int d[8][4] = {};
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(d[j][0]), "+r"(d[j][1]), "+r"(d[j][2]), "+r"(d[j][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  int s = 0;
  for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dp4a_kernel(int iters, int* out) {
  int d[16] = {};
  unsigned a = threadIdx.x * 7 + 1, b = threadIdx.x * 13 + 5;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(d[j]) : "r"(a + j), "r"(b));
  }
  int s = 0;
  for (int j = 0; j < 16; ++j) s += d[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void imad_kernel(int iters, int* out) {
  int d[16] = {};
  int a = threadIdx.x * 7 + 1, b = threadIdx.x * 13 + 5;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(d[j]) : "r"(a + j), "r"(b));
  }
  int s = 0;
  for (int j = 0; j < 16; ++j) s += d[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int* out;
  cudaMalloc(&out, sms * 8 * 1024 * sizeof(int));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 20000;
    float ms;
    imma_kernel<<<sms, warps * 32>>>(100, out);
    cudaEventRecord(e0); imma_kernel<<<sms, warps * 32>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double imma = (double)iters * 8 * warps * sms;
    printf("warps/SM %2d  IMMA m16n8k32: %.3f ms  %.1f TOPS  (%.2f IMMA/clk/SM @%d MHz nominal)\n", warps, ms,
           imma * 4096 * 2 / ms / 1e9, imma / sms / (ms * 1e-3 * clk * 1e3), clk / 1000);
    dp4a_kernel<<<sms, warps * 32>>>(100, out);
    cudaEventRecord(e0); dp4a_kernel<<<sms, warps * 32>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)iters * 16 * warps * 32 * sms;
    printf("             dp4a        : %.3f ms  %.1f G dp4a/s  (%.1f lanes/clk/SM)\n", ms, n / ms / 1e6,
           n / sms / (ms * 1e-3 * clk * 1e3));
    imad_kernel<<<sms, warps * 32>>>(100, out);
    cudaEventRecord(e0); imad_kernel<<<sms, warps * 32>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("             imad        : %.3f ms  (%.1f lanes/clk/SM)\n", ms, n / sms / (ms * 1e-3 * clk * 1e3));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

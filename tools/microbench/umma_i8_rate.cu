// Microbenchmark: issue rate of tcgen05.mma kind::i8 (and kind::f8f6f4 for comparison) from fixed shared-memory
// operands, one CTA per SM, no TMA / epilogue: the achievable int8 tensor peak of this B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int KIND>  // 0 = i8, 1 = f8f6f4
__global__ void __launch_bounds__(128, 1) umma_kernel(int iters, int N, int b_mn_major, unsigned long long* cycles, int random_data) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(raw)[i] = random_data ? h : 0x01010101u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 32) {
    // instruction descriptor: D = S32 (i8) / F32 (f8), A K-major, B K- or MN-major, M = 128
    uint32_t idesc = ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | ((uint32_t)b_mn_major << 16);
    if (KIND == 0) idesc |= (2u << 4) | (1u << 7) | (0u << 10);
    else idesc |= (1u << 4);  // F32 accumulate, E4M3 x E4M3
    const uint32_t sa = base, sb = base + 16384;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = desc(sa + kk * 32, 0, 1024);
        const uint64_t bd = b_mn_major ? desc(sb + kk * 4096, 16384, 1024) : desc(sb + kk * 32, 0, 1024);
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem + (uint32_t)((it & 1) * 256)), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem + (uint32_t)((it & 1) * 256)), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const unsigned long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const size_t smem = 64 * 1024 + 1024;
  cudaFuncSetAttribute(umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rnd = 0; rnd < 2; ++rnd)
  for (int kind = 0; kind < 2; ++kind) {
    for (int N : {128, 256}) {
      for (int mn = 0; mn < 2; ++mn) {
        if (kind == 1 && mn == 1) continue;
        float ms;
        unsigned long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          if (kind == 0) umma_kernel<0><<<sms, 128, smem>>>(iters, N, mn, cyc, rnd);
          else umma_kernel<1><<<sms, 128, smem>>>(iters, N, mn, cyc, rnd);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double macs = (double)iters * 4 * 128.0 * N * 32 * sms;
        printf("%s %s M=128 N=%3d B %s-major: %.3f ms  %.0f TOPS  %.1f clk per MMA  (%s)\n", rnd ? "random-data" : "const-data ", kind ? "f8f6f4" : "i8    ", N,
               mn ? "MN" : "K ", ms, 2 * macs / ms / 1e9, (double)h / (iters * 4.0), cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}

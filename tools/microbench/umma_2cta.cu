// Bring-up: tcgen05.mma.cta_group::2 kind::i8, M = 256 (128 rows per CTA), N = 256 (128 columns of B per CTA),
// A K-major SW128, B MN-major SW128 (the GEMM's operand layouts).  Checks the accumulators of both CTAs against a
// host reference and times back-to-back issue.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spin = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (++spin > (1u << 22)) __trap();
  }
}
__host__ __device__ inline int a_val(int m, int k) { return ((m * 7 + k * 3) % 11) - 5; }
__host__ __device__ inline int b_val(int k, int n) { return ((k * 5 + n * 2) % 13) - 6; }

// out: [2 CTAs][128 rows][256 cols] int32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int* out, int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;      // [A 16K][B 16K]
  uint8_t* gen = raw + (base - smem_u32(raw));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A: this CTA's 128 rows (global rows rank*128 ..), K = 128 bytes, K-major SW128
  for (int i = threadIdx.x; i < 128 * 128; i += 128) {
    const int r = i / 128, kb = i % 128;
    gen[r * 128 + ((((kb >> 4) ^ (r & 7)) << 4) | (kb & 15))] = (uint8_t)(int8_t)a_val(rank * 128 + r, kb);
  }
  // B: this CTA's 128 columns (global n = rank*128 ..), 128 k rows, MN-major SW128: (k, n) at k*128 + swz(n)
  for (int i = threadIdx.x; i < 128 * 128; i += 128) {
    const int kk = i / 128, n = i % 128;
    gen[16384 + kk * 128 + ((((n >> 4) ^ (kk & 7)) << 4) | (n & 15))] = (uint8_t)(int8_t)b_val(kk, rank * 128 + n);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (rank == 0 && warp == 0) {
    // whole warp, one elected lane issues (uniform descriptors)
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    // D = S32, A signed K-major, B signed MN-major, N = 256, M = 256
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t hi_a = desc(0, 0, 1024), hi_b = desc(0, 16384, 1024);
    const uint32_t a_lo = (base & 0x3FFFF) >> 4, b_lo = ((base + 16384) & 0x3FFFF) >> 4;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(hi_a | (uint64_t)(a_lo + kk * 2)), "l"(hi_b | (uint64_t)(b_lo + kk * 256)), "r"(idesc),
                         "r"((it | kk) ? 1u : 0u) : "memory");
      }
      __syncwarp();
    }
    if (leader)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    wait(smem_u32(&bar), 0u);
    const unsigned long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  // every warp of both CTAs: wait for the accumulators, then dump this CTA's 128 lanes x 256 columns
  wait(smem_u32(&bar), 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (blockIdx.x < 2) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 256; c0 += 16) {
      int v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + row) * 256 + c0 + j] = v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int* out; cudaMalloc(&out, 2 * 128 * 256 * 4); cudaMemset(out, 0xff, 2 * 128 * 256 * 4);
  unsigned long long* cyc; cudaMalloc(&cyc, 8);
  const size_t smem = 2 * 16384 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // correctness: one pair, one pass over K = 128
  k<<<2, 128, smem>>>(out, 1, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<int> h(2 * 128 * 256);
  cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 256; ++n) {
      int ref = 0;
      for (int kk = 0; kk < 128; ++kk) ref += a_val(m, kk) * b_val(kk, n);
      if (h[(size_t)m * 256 + n] != ref) {
        if (bad < 5) printf("mismatch m=%d n=%d got %d want %d\n", m, n, h[(size_t)m * 256 + n], ref);
        ++bad;
      }
    }
  printf("2-CTA MMA M=256 N=256 K=128: %ld / %d mismatches\n", bad, 256 * 256);
  // rate: all pairs
  const int iters = 4000;
  k<<<(sms / 2) * 2, 128, smem>>>(out, iters, cyc);
  e = cudaDeviceSynchronize();
  unsigned long long c = 0;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("rate: %.1f clk per MMA (M=256 N=256 K=32 across the pair) (%s)\n", (double)c / (iters * 4.0), cudaGetErrorString(e));
  return 0;
}

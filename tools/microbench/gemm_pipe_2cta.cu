// Microbenchmark: the bare TMA -> tcgen05.mma pipeline with cta_group::2.  Each CTA of a pair keeps its own 128 rows
// of weights (4 slabs) resident and streams ITS half (128 of 256 columns) of every activation slab; the leader
// issues M=256 x N=256 MMAs.  Per SM the tensor work per slab equals the 1-CTA MT=2 N=128 schedule (512 clk),
// but the shared-memory traffic drops from 80 KB (64 read + 16 written) to 48 KB per slab.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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
    if (++spin > (1u << 24)) __trap();
  }
}
constexpr int kMaxStages = 10;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __grid_constant__ CUtensorMap map, int slabs, int S, int who, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;   // [A: 4 x 16K][B ring: S x 16K]
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages], fin;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t a_bytes = 4u * 16384u;
  if (threadIdx.x == 32) {
    for (int s = 0; s < kMaxStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_u32(&full[s])));    // both producers arrive (leader's copy is used)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));   // multicast commit
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fin)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < (a_bytes + S * 16384u) / 4; i += 128) {
    uint32_t h = i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    *reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)) + 4 * i) = h;
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
  if (warp == 0 && rank == 0 && (who & 1)) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t hi_a = desc(0, 0, 1024), hi_b = desc(0, 16384, 1024);
    const uint32_t a_lo0 = (base & 0x3FFFF) >> 4, b_lo0 = ((base + a_bytes) & 0x3FFFF) >> 4;
    int stage = 0; uint32_t phase = 0, b_lo = b_lo0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < slabs; ++it) {
      if (who & 2) {
        wait(smem_u32(&full[stage]), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t a_lo = a_lo0 + (uint32_t)(it & 3) * 1024u;
      const uint32_t d = tmem + (uint32_t)(((it >> 2) & 1) * 256);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(hi_a | (uint64_t)(a_lo + kk * 2)), "l"(hi_b | (uint64_t)(b_lo + kk * 256)), "r"(idesc),
                         "r"(((it & 3) | kk) ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&empty[stage])), "h"((uint16_t)3) : "memory");
      }
      b_lo += 1024u;
      if (++stage == S) { stage = 0; phase ^= 1u; b_lo = b_lo0; }
    }
    if (leader)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&fin)), "h"((uint16_t)3) : "memory");
    wait(smem_u32(&fin), 0u);
    const unsigned long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  } else if (warp == 1 && (who & 2)) {
    // producer in BOTH CTAs: own half of every slab into own smem, completion on the LEADER's full barrier
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    int stage = 0; uint32_t phase = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < slabs; ++it) {
      if (who & 1) wait(smem_u32(&empty[stage]), phase ^ 1u);
      if (leader) {
        uint32_t lbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lbar) : "r"(smem_u32(&full[stage])), "r"(0u));
        asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(lbar), "r"(16384u) : "memory");
        const int n = (blockIdx.x >> 1) * 7 + (it >> 2);
        asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(base + a_bytes + stage * 16384), "l"(&map), "r"(lbar), "r"(((n * 2 + (int)rank) % 6) * 128),
                       "r"((it & 3) * 128), "r"((n / 3) % 64) : "memory");
      }
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    const unsigned long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cycles[1] = t1 - t0;
  }
  // the peer must not exit (its smem / TMEM are in use) before the leader is done
  if (who & 1) wait(smem_u32(&fin), 0u);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int B = 64, C = 512, T = 800, ld = 800;
  uint8_t* x; cudaMalloc(&x, (size_t)B * C * ld); cudaMemset(x, 3, (size_t)B * C * ld);
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * C};
  cuuint32_t box[3] = {128, 128, 1}, es[3] = {1, 1, 1};
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  unsigned long long* cyc; cudaMalloc(&cyc, 32);
  const size_t smem = 14 * 16384 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int slabs = 4000;
  struct Cfg { int S, who; } cfgs[] = {{6, 1}, {6, 3}, {10, 3}, {3, 3}};
  for (const Cfg& c : cfgs) {
    cudaMemset(cyc, 0, 32);
    k<<<(sms / 2) * 2, 128, smem>>>(map, slabs, c.S, c.who, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[4];
    cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("S=%2d mma=%d tma=%d: MMA warp %.0f clk/slab, producer %.0f clk/slab (%s)\n", c.S, c.who & 1, (c.who >> 1) & 1,
           (double)h[0] / slabs, (double)h[1] / slabs, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}

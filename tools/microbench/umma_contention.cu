// Microbenchmark: does tcgen05.mma (i8, M=128 N=128 K=32, SS mode) slow down when (a) TMA keeps writing other
// shared-memory stages and / or (b) other warps drain TMEM with tcgen05.ld at the same time?
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
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// warps: 0 = alloc + MMA issuer (lane 0), 1 = TMA producer (lane 0) + consumer arrive, 4..19 = TMEM drain
__global__ void __launch_bounds__(640, 1) k(const __grid_constant__ CUtensorMap map, int iters, int do_tma, int do_drain,
                                            volatile int* stop_flag_unused, unsigned long long* cycles, int* sink) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;   // [A 16K][B 16K][4 TMA stages x 16K]
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar, tfull[4];
  __shared__ volatile int done_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tfull[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    done_flag = 0;
  }
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 640) reinterpret_cast<uint32_t*>(raw)[i] = 0x01010101u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (1u << 16) | (2u << 4) | (1u << 7);
    const uint32_t sa = base, sb = base + 16384;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = desc(sa + kk * 32, 0, 1024), bd = desc(sb + kk * 4096, 16384, 1024);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem + (uint32_t)((it & 1) * 128)), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    wait(smem_u32(&bar), 0u);
    const unsigned long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
    done_flag = 1;
  } else if (threadIdx.x == 32 && do_tma) {
    // stream 16 KB boxes into 4 stages until the MMA loop is done (self-consumed)
    int stage = 0; uint32_t phase = 0; long long n = 0;
    while (!done_flag) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tfull[stage])), "r"(16384u) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(base + 32768 + stage * 16384), "l"(&map), "r"(smem_u32(&tfull[stage])), "r"((int)((n * 128) % 768)),
                     "r"((int)((n % 4) * 128)), "r"((int)((blockIdx.x + n) % 64)) : "memory");
      wait(smem_u32(&tfull[stage]), phase);   // one box in flight per stage visit: ~4 in flight overall is not needed here
      if (++stage == 4) { stage = 0; phase ^= 1u; }
      ++n;
    }
    if (blockIdx.x == 0) cycles[1] = (unsigned long long)n;
  } else if (warp >= 4 && do_drain) {
    int acc = 0;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)(((warp - 4) >> 2) * 32);
    long long n = 0;
    while (!done_flag) {
      int v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += v[i];
      ++n;
    }
    if (acc == 0x7fffffff) sink[0] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 128) cycles[2] = (unsigned long long)n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int B = 64, C = 512, T = 800, ld = 800;
  uint8_t* x; cudaMalloc(&x, (size_t)B * C * ld); cudaMemset(x, 1, (size_t)B * C * ld);
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * C};
  cuuint32_t box[3] = {128, 128, 1}, es[3] = {1, 1, 1};
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  unsigned long long* cyc; cudaMalloc(&cyc, 32);
  int* sink; cudaMalloc(&sink, 4);
  const size_t smem = 6 * 16384 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 5000;
  for (int mode = 0; mode < 4; ++mode) {
    cudaMemset(cyc, 0, 32);
    k<<<sms, 640, smem>>>(map, iters, mode & 1, (mode >> 1) & 1, nullptr, cyc, sink);
    cudaDeviceSynchronize();
    unsigned long long h[4];
    cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("TMA=%d drain=%d: %.1f clk per MMA; %llu boxes loaded, %llu tcgen05.ld.x16 per drain warp (%s)\n", mode & 1, (mode >> 1) & 1,
           (double)h[0] / (iters * 4.0), h[1], h[2], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

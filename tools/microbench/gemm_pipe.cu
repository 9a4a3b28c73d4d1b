// Microbenchmark: the bare TMA -> tcgen05.mma pipeline of a weight-stationary int8 GEMM, without any epilogue.
// A (weights, MT x 4 slabs of 128x128 B) is resident in shared memory; B slabs (128 k-rows x 128 t-bytes, MN-major)
// stream through an S-stage ring with full/empty mbarriers exactly as the product kernel does.  Reports the
// average time per slab seen by the MMA issuer so that the dependency chain itself can be timed in isolation.
//   knobs: stages S, MT (M tiles per slab: MMAs per slab = 4*MT), dep (1 = MMA waits for full / producer waits
//   for empty, 0 = free running), who (bit0 = run MMA, bit1 = run TMA)
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
constexpr int kMaxStages = 8;

template <int kMT, int kN>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int slabs, int S, int MT, int dep, int who,
                                            int ctrl_warp_hi, unsigned long long* cycles, int a_fixed, int b_fixed, int acc_always, int N) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;   // [A: MT*4 x 16K][B ring: S x 16K]
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages], fin;
  const int warp = threadIdx.x >> 5;
  const uint32_t a_bytes = (uint32_t)MT * 4u * 16384u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < kMaxStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fin)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < (a_bytes + S * 16384u) / 4; i += 128) {
    uint32_t h = i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    *reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)) + 4 * i) = h;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int mma_thread = ctrl_warp_hi ? 96 : 0, tma_thread = ctrl_warp_hi ? 64 : 32;
  const int uwarp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp index as a uniform value
  if (ctrl_warp_hi == 2 && uwarp == 3) {
    if (who & 1) {
      // whole warp walks the loop (uniform control flow, descriptors in uniform registers); one elected lane issues
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      const uint32_t idesc = ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (1u << 16) | (2u << 4) | (1u << 7);
      const uint64_t desc_hi = ((uint64_t)((1024u >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t bdesc_hi = desc_hi | ((uint64_t)((16384u >> 4) & 0x3FFF) << 16);
      int stage = 0; uint32_t phase = 0;
      const unsigned long long t0 = clock64();
      const uint32_t a_lo0 = (base & 0x3FFFF) >> 4, b_lo0 = ((base + a_bytes) & 0x3FFFF) >> 4;
      constexpr uint32_t kBStep = (4096u * (kN / 128)) >> 4, kStageLo = (16384u * (kN / 128)) >> 4;
      uint32_t b_lo = b_lo0;
      const bool chain = dep && (who & 2);
      for (int it = 0; it < slabs; ++it) {
        if (chain) {
          wait(smem_u32(&full[stage]), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t ks = (uint32_t)(it & 3);
        const uint32_t dbase = tmem + (uint32_t)(((it >> 2) & 1) * 256);
        if (leader) {
#pragma unroll
          for (int ms = 0; ms < kMT; ++ms) {
            const uint32_t a_lo = a_lo0 + (uint32_t)(ms * 4) * 1024u + ks * 1024u;
            const uint32_t d = kN == 256 ? dbase : dbase + (uint32_t)(ms * 128);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = desc_hi | (uint64_t)(a_lo + kk * 2), bd = bdesc_hi | (uint64_t)(b_lo + kk * kBStep);
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                           ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"((ks | kk) ? 1u : 0u) : "memory");
            }
          }
          if (dep) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
        }
        b_lo += kStageLo;
        if (++stage == S) { stage = 0; phase ^= 1u; b_lo = b_lo0; }
      }
      if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&fin)) : "memory");
      wait(smem_u32(&fin), 0u);
      const unsigned long long t1 = clock64();
      if (blockIdx.x == 0 && leader) cycles[0] = t1 - t0;
    }
  } else if (ctrl_warp_hi != 2 && (int)threadIdx.x == mma_thread && (who & 1)) {
    const uint32_t idesc = ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (1u << 16) | (2u << 4) | (1u << 7);
    int stage = 0; uint32_t phase = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < slabs; ++it) {
      if (dep && (who & 2)) {
        wait(smem_u32(&full[stage]), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t sb = base + a_bytes + (b_fixed ? 0u : (uint32_t)stage * 16384u * (N / 128));
      for (int ms = 0; ms < MT; ++ms) {
        const uint32_t sa = base + (uint32_t)(ms * 4 + (a_fixed ? 0 : (it & 3))) * 16384u;
        const uint32_t d = N == 256 ? tmem + (uint32_t)(((it >> 2) & 1) * 256) : tmem + (uint32_t)(((it >> 2) & 1) * 256 + ms * 128);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ad = desc(sa + kk * 32, 0, 1024), bd = desc(sb + kk * 4096 * (N / 128), 16384, 1024);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"((acc_always || ((it & 3) | kk)) ? 1u : 0u) : "memory");
        }
      }
      if (dep) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&fin)) : "memory");
    wait(smem_u32(&fin), 0u);
    const unsigned long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  } else if ((int)threadIdx.x == tma_thread && (who & 2)) {
    int stage = 0; uint32_t phase = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < slabs; ++it) {
      if (it >= S) {
        if (dep && (who & 1)) wait(smem_u32(&empty[stage]), phase ^ 1u);
        else wait(smem_u32(&full[stage]), phase ^ 1u);   // self-paced: previous load into this stage has landed
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stage])), "r"(16384u) : "memory");
      const int n = blockIdx.x * 7 + (it >> 2);
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(base + a_bytes + stage * 16384), "l"(&map), "r"(smem_u32(&full[stage])), "r"((n % 6) * 128),
                     "r"((it & 3) * 128), "r"((n / 6) % 64) : "memory");
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    // wait for the last S loads
    for (int i = 0; i < S && i < slabs; ++i) {
      const int it = slabs - 1 - i;
      if (!(dep && (who & 1))) wait(smem_u32(&full[it % S]), (uint32_t)((it / S) & 1));
    }
    const unsigned long long t1 = clock64();
    if (blockIdx.x == 0) cycles[1] = t1 - t0;
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
  uint8_t* x; cudaMalloc(&x, (size_t)B * C * ld); cudaMemset(x, 3, (size_t)B * C * ld);
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * C};
  cuuint32_t box[3] = {128, 128, 1}, es[3] = {1, 1, 1};
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  unsigned long long* cyc; cudaMalloc(&cyc, 32);
  const size_t smem = 14 * 16384 + 1024;
  cudaFuncSetAttribute(k<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int slabs = 4000;
  struct Cfg { int S, MT, dep, who, hi, af, bf, acc, N; } cfgs[] = {
      {6, 2, 0, 1, 0, 0, 0, 0, 128}, {6, 2, 1, 3, 0, 0, 0, 0, 128},
      {6, 2, 0, 1, 2, 0, 0, 0, 128}, {6, 1, 0, 1, 2, 0, 0, 0, 128}, {3, 1, 0, 1, 2, 0, 0, 0, 256}, {6, 2, 1, 3, 2, 0, 0, 0, 128},
      {6, 1, 1, 3, 2, 0, 0, 0, 128}, {3, 1, 1, 3, 2, 0, 0, 0, 256}, {3, 2, 1, 3, 2, 0, 0, 0, 128}, {2, 2, 1, 3, 2, 0, 0, 0, 128}};
  for (const Cfg& c : cfgs) {
    cudaMemset(cyc, 0, 32);
    if (c.N == 256) k<1, 256><<<sms, 128, smem>>>(map, slabs, c.S, c.MT, c.dep, c.who, c.hi, cyc, c.af, c.bf, c.acc, c.N);
    else if (c.MT == 2) k<2, 128><<<sms, 128, smem>>>(map, slabs, c.S, c.MT, c.dep, c.who, c.hi, cyc, c.af, c.bf, c.acc, c.N);
    else k<1, 128><<<sms, 128, smem>>>(map, slabs, c.S, c.MT, c.dep, c.who, c.hi, cyc, c.af, c.bf, c.acc, c.N);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[4];
    cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("N=%d Afixed=%d Bfixed=%d acc1=%d S=%d MT=%d dep=%d mma=%d tma=%d hi=%d: MMA thread %.0f clk/slab (%.1f clk/MMA), TMA thread %.0f clk/slab (%s)\n", c.N, c.af, c.bf, c.acc, c.S, c.MT,
           c.dep, c.who & 1, (c.who >> 1) & 1, c.hi, (double)h[0] / slabs, (double)h[0] / slabs / (4.0 * c.MT), (double)h[1] / slabs,
           cudaGetErrorString(e));
  }
  return 0;
}

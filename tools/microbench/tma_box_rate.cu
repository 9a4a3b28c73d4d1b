// Microbenchmark: TMA tiled-load delivery rate per SM for the GEMM's activation boxes (128 rows x 128 bytes,
// SWIZZLE_128B) as a function of row pitch, pipeline depth and sharing between CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int stages, int boxes_per_cta,
                                                    int n_tiles, int share, int C) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[12], empty[12];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int cblocks = C / 128;
  if (threadIdx.x == 0) {          // producer
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < boxes_per_cta; ++i) {
      const int tile = ((blockIdx.x / share) + (i / cblocks) * (gridDim.x / share)) % n_tiles;  // (b, t-tile)
      const int nt = tile % 7, b = tile / 7, cb = i % cblocks;
      wait(smem_u32(&empty[stage]), phase ^ 1u);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stage])), "r"(16384u) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(base + stage * 16384), "l"(&map), "r"(smem_u32(&full[stage])), "r"(nt * 128), "r"(cb * 128), "r"(b) : "memory");
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {  // consumer: frees the stage immediately
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < boxes_per_cta; ++i) {
      wait(smem_u32(&full[stage]), phase);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int B = 64, C = 512, T = 800;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384 + 1024);
  for (int ld : {800, 1024}) {
    uint8_t* x; cudaMalloc(&x, (size_t)B * C * ld); cudaMemset(x, 1, (size_t)B * C * ld);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * C};
    cuuint32_t box[3] = {128, 128, 1}, es[3] = {1, 1, 1};
    enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int share : {1, 2, 4}) {
      for (int stages : {2, 4, 6, 10}) {
        const int boxes = 96;  // per CTA
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          tma_kernel<<<sms, 64, stages * 16384 + 1024>>>(map, stages, boxes, B * 7, share, C);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms, e0, e1);
        }
        const double bytes = (double)sms * boxes * 16384;
        printf("ld=%4d share=%d stages=%2d: %7.1f us  %6.2f TB/s delivered  %5.1f GB/s/SM  %.2f us/box/SM  (%s)\n", ld, share,
               stages, ms * 1e3, bytes / ms / 1e9, bytes / sms / ms / 1e6, ms * 1e3 / boxes, cudaGetErrorString(cudaGetLastError()));
      }
    }
    cudaFree(x);
  }
  return 0;
}

// Microbenchmark: TMA delivery rate per SM as a function of WHERE the rows of a box live.
//   mode 0: box [128 B x 128 channels x 1 utterance]  (rows 800 B apart: the GEMM's activation box)
//   mode 1: box [128 B x 1 channel x 64 utterances]    (rows C*ld = 400 KB apart: the depthwise kernel's rows)
//   mode 2: as 1 with the utterance stride padded by `bpad` bytes (breaks the power-of-two-ish stride)
//   mode 3: box [128 B x 5 time blocks (stride 160 B, overlapping) x 1 channel x 25 utterances] = 125 rows
// All data is L2 resident (26 MB tensor re-read many times).
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

__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int mode, int stages, int boxes_per_cta,
                                                    uint32_t box_bytes) {
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
  if (threadIdx.x == 0) {          // producer
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < boxes_per_cta; ++i) {
      const int u = blockIdx.x * boxes_per_cta + i;
      wait(smem_u32(&empty[stage]), phase ^ 1u);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stage])), "r"(box_bytes) : "memory");
      const uint32_t dst = base + stage * 16384, bar = smem_u32(&full[stage]);
      if (mode == 0) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(bar), "r"((u % 6) * 128), "r"(((u / 6) % 4) * 128), "r"((u / 24) % 64) : "memory");
      } else if (mode == 1 || mode == 2) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(bar), "r"((u % 6) * 128), "r"((u / 6) % 512), "r"(0) : "memory");
      } else {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(bar), "r"((u & 1) * 16), "r"(0), "r"((u / 2) % 512), "r"(((u / 1024) % 2) * 25) : "memory");
      }
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
  const int B = 64, C = 512, T = 800, ld = 800;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384 + 1024);
  struct Cfg { int mode; long long bpad; } cfgs[] = {{0, 0}, {1, 0}, {2, 128}, {2, 800}, {2, 4096 + 128}, {3, 0}, {3, 800}};
  for (const Cfg& c : cfgs) {
    const size_t bstride = (size_t)C * ld + c.bpad;
    uint8_t* x; cudaMalloc(&x, (size_t)B * bstride + 4096); cudaMemset(x, 1, (size_t)B * bstride + 4096);
    CUtensorMap map;
    CUresult r;
    uint32_t box_bytes;
    if (c.mode <= 2) {
      cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)bstride};
      cuuint32_t box0[3] = {128, 128, 1}, box1[3] = {128, 1, 64}, es[3] = {1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, x, dims, strides, c.mode == 0 ? box0 : box1, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      box_bytes = c.mode == 0 ? 16384 : 8192;
    } else {
      cuuint64_t dims[4] = {(cuuint64_t)T, 5, (cuuint64_t)C, (cuuint64_t)B};
      cuuint64_t strides[3] = {160, (cuuint64_t)ld, (cuuint64_t)bstride};
      cuuint32_t box[4] = {128, 5, 1, 25}, es[4] = {1, 1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      box_bytes = 125 * 128;
    }
    if (r != CUDA_SUCCESS) { printf("mode %d: encode failed %d\n", c.mode, (int)r); continue; }
    for (int stages : {4, 8}) {
      const int boxes = 96;
      float ms = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        tma_kernel<<<sms, 64, stages * 16384 + 1024>>>(map, c.mode, stages, boxes, box_bytes);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      const double bytes = (double)sms * boxes * box_bytes;
      printf("mode %d bpad=%5lld stages=%d: %7.1f us  %6.2f TB/s  %5.1f GB/s/SM  %.0f clk per 128-B row  (%s)\n", c.mode, c.bpad, stages,
             ms * 1e3, bytes / ms / 1e9, bytes / sms / ms / 1e6, ms * 1e-3 * 1.965e9 / (boxes * (box_bytes / 128.0)),
             cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(x);
  }
  return 0;
}

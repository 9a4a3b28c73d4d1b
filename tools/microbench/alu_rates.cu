// Microbenchmark: issue rate (thread-instructions per clock per SM) of the integer / float instructions a
// requantisation epilogue can be built from.  8 independent dependency chains per thread, 1024 threads per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 512

template <int OP>
__device__ __forceinline__ void step(int& v, int a, int b) {
  if (OP == 0) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
  if (OP == 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
  if (OP == 2) asm volatile("shr.s32 %0, %0, %1;" : "+r"(v) : "r"(b & 1));
  if (OP == 3) asm volatile("max.s32 %0, %0, %1;" : "+r"(v) : "r"(a));
  if (OP == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(v) : "r"(a));
  if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b));
  if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(v) : "r"(a));
  if (OP == 7) asm volatile("{\n\t.reg .f32 f;\n\tmov.b32 f, %0;\n\tfma.rn.f32 f, f, %1, %2;\n\tmov.b32 %0, f;\n\t}" : "+r"(v) : "f"(1.0001f), "f"(0.5f));
  if (OP == 8) asm volatile("{\n\t.reg .f32 f;\n\tmov.b32 f, %0;\n\tadd.rn.f32 f, f, %1;\n\tmov.b32 %0, f;\n\t}" : "+r"(v) : "f"(0.5f));
  if (OP == 9) asm volatile("{\n\t.reg .f32 f;\n\tmov.b32 f, %0;\n\tmax.f32 f, f, %1;\n\tmov.b32 %0, f;\n\t}" : "+r"(v) : "f"(0.5f));
  if (OP == 10) asm volatile("{\n\t.reg .f32 f;\n\tcvt.rn.f32.s32 f, %0;\n\tmov.b32 %0, f;\n\t}" : "+r"(v));
  if (OP == 11) asm volatile("{\n\t.reg .f32 f;\n\tmov.b32 f, %0;\n\tcvt.rni.s32.f32 %0, f;\n\t}" : "+r"(v));
  if (OP == 12) asm volatile("cvt.pack.sat.s8.s32.b32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
  if (OP == 13) asm volatile("{\n\t.reg .b32 t;\n\tmax.s32 t, %0, %1;\n\tmin.s32 %0, t, %2;\n\t}" : "+r"(v) : "r"(a), "r"(b));   // clamp
  if (OP == 14) asm volatile("vmax2.s32.s32.s32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
  if (OP == 15) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(v) : "r"(a));
  if (OP == 16) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b & 7));
  if (OP == 17) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(v) : "r"(a));
  if (OP == 18) asm volatile("{\n\t.reg .b64 w;\n\tmul.wide.s32 w, %0, %1;\n\tshr.s64 w, w, %2;\n\tcvt.u32.u64 %0, w;\n\t}" : "+r"(v) : "r"(a), "r"(40));
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(int a, int b, int* out, unsigned long long* cyc) {
  int v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = threadIdx.x * 7 + i;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) step<OP>(v[i], a, b);
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s ^= v[i];
  if (s == 0x12345678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int* out, unsigned long long* cyc, int sms) {
  k<OP><<<sms, 1024>>>(12345, 3, out, cyc);
  cudaDeviceSynchronize();
  unsigned long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %6.1f thread-instr / clk / SM\n", name, 1024.0 * CHAINS * ITERS / (double)h);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int* out; unsigned long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  run<0>("mad.hi.s32 (IMAD.HI)", out, cyc, sms);
  run<17>("mul.hi.s32", out, cyc, sms);
  run<1>("mad.lo.s32 (IMAD)", out, cyc, sms);
  run<2>("shr.s32 (SHF)", out, cyc, sms);
  run<16>("shf.r.clamp", out, cyc, sms);
  run<3>("max.s32 (VIMNMX)", out, cyc, sms);
  run<13>("max+min clamp (2 instr?)", out, cyc, sms);
  run<15>("max.s16x2", out, cyc, sms);
  run<14>("vmax2 (emulated?)", out, cyc, sms);
  run<4>("add.s32 (IADD3)", out, cyc, sms);
  run<5>("lop3", out, cyc, sms);
  run<6>("prmt", out, cyc, sms);
  run<12>("cvt.pack.sat.s8.s32 (I2IP)", out, cyc, sms);
  run<7>("fma.f32", out, cyc, sms);
  run<8>("add.f32", out, cyc, sms);
  run<9>("max.f32 (FMNMX)", out, cyc, sms);
  run<10>("cvt.f32.s32 (I2F)", out, cyc, sms);
  run<11>("cvt.rni.s32.f32 (F2I)", out, cyc, sms);
  run<18>("mul.wide + shr.s64 (3 ops)", out, cyc, sms);
  return 0;
}

// Shared pieces of the IMMA depthwise engines (dwconv.cu: one row at a time; dwconv_rt.cu: row tapes).
#pragma once
#include "qasr_common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kThreads = 32 * kWarps;
constexpr int kTile = 128;   // outputs per warp tile
constexpr int kBufs = 3;     // input rows in flight per warp (bulk async copies, 2 ahead)
constexpr uint32_t kSpinLimit = 1u << 24;

struct DwArgs {
  const uint8_t* x;
  const int8_t* w;
  void* out;
  int x_ld, out_ld, B, C, T_in, T_out, k, stride, dil, pad;
  int d;          // (-pad) mod 8
  int bchunk;     // utterances per warp work item
  int nbg;        // ceil(B / bchunk)
  int halo;       // zero bytes in front of the row inside a staging buffer (multiple of 16, >= pad + d)
  int copy_bytes; // T_in rounded up to 16: bytes of one bulk copy
  int buf_bytes;  // one staging buffer (multiple of 128)
  int period;     // row-tape kernel: bytes from one utterance row to the next inside a tape (gap of zeros + copy_bytes)
  unsigned long long* stamp;   // in-graph timing slot (qasr_profile_begin) or nullptr
  EpiArgs ep;
};

template <bool XS>
__device__ __forceinline__ void imma_16832(int (&acc)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (XS)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// first MMA of a chain: D = A x B + C with every C register = c (the initial accumulator value), D written to
// fresh registers, so the loop needs no moves to re-initialise its accumulators
template <bool XS>
__device__ __forceinline__ void imma_16832_init(int (&acc)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int c) {
  if (XS)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c));
  else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c));
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > kSpinLimit) __trap();
  }
}
// one input row global -> shared through the bulk-copy engine, completion on an mbarrier
__device__ __forceinline__ void row_copy_async(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

inline int dw_misalign(int pad) { return ((-pad) % 8 + 8) % 8; }

inline int dw_ksteps(int k, int stride, int dilation, int pad) {
  const int d = dw_misalign(pad);
  return (15 * stride + d + (k - 1) * dilation + 1 + 31) / 32;
}

}  // namespace

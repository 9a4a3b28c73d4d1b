// K1-tape: depthwise 1-D convolution on 8-bit codes on the tcgen05 tensor path WITHOUT any operand expansion.
// Same arithmetic as dwconv.cu (F.conv1d(x_int, w_int, groups=C) of QuantConv1d.int_conv,
// nemo/quantization/utils/quant_modules.py:301-305, in true integers) and the same fused epilogue.
//
// The trick is the shared-memory matrix descriptor.  In the K-major, non-swizzled canonical layout a core matrix is
// 8 rows x 16 bytes with the rows 16 bytes apart, LBO = distance between the two 16-byte K chunks of one MMA and
// SBO = distance between 8-row groups.  With LBO = 16 and SBO = 128 the element A[r][kk] of an M = 128, K = 32
// operand sits at byte 16 r + kk of a plain byte string: row r of the operand IS the window of the signal that
// starts 16 bytes after row r-1's.  So the utterance rows of one channel, copied back to back into shared memory (a
// "tape", each row preceded by a gap of zeros >= the conv padding), are directly the A operand of a banded product
//     D[r][n] = sum_kk tape[16 r + kk] * Bm[n][kk],   Bm[n][kk] = w[(kk - n - d) / dil]      (16 x K, per channel)
// with D[r][n] = out[16 r + n]: 128 lanes x 16 columns of TMEM = 2048 consecutive output frames per MMA group, and
// every TMEM lane owns 16 CONSECUTIVE frames = one 16-byte store.  No Toeplitz expansion of the activations, no
// per-window TMA rows (the input arrives as one bulk copy per utterance row), K = 16 + (k-1) dil + d bytes of window
// per row (k = 33: two MMAs of K = 32).  The 16 x K weight matrices are packed once per layer (qasr_pack_dw_weights).
//
// Work item = (channel, chunk of `nb` utterances); a CTA walks a contiguous range of items, double-buffering tapes.
// Warp roles (608 threads, 1 CTA / SM): warps 0..15 epilogue (quarter = warp & 3 is the TMEM lane quarter, warp >> 2
// the unit of the round), warp 16 producer (bulk copies), warp 17 MMA issuer, warp 18 TMEM allocator.
// A "round" = 4 units (one per epilogue group) = 64 TMEM columns; 8 rounds of accumulators are in flight.
#include <cuda.h>

#include "qasr_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace qasr_tc;

constexpr int kEpiWarps = 16;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kAllocWarp = kEpiWarps + 2;
constexpr int kThreads = 32 * (kEpiWarps + 3);
constexpr int kMaxBufs = 4;
constexpr int kSlots = 8;                 // accumulator rounds in flight: 8 x 64 columns = the 512 TMEM columns
constexpr int kMaxKS = 8;
constexpr uint32_t kBmatBytes = 512;      // one K step of the 16 x K weight matrix: 2 n-groups x 2 k-chunks x 128 B
constexpr uint32_t kSmemBudget = 227 * 1024;

struct TapeArgs {
  const uint8_t* x;
  const uint8_t* bmat;   // [C][KS][512]
  void* out;
  int x_ld, out_ld, B, C, T_in, T_out;
  int KS;            // K steps of 32
  int G;             // zero gap in front of every utterance row (multiple of 16, >= pad)
  int P;             // tape bytes per utterance: G + copy_bytes
  int rpu;           // tape rows (16 bytes) per utterance: P / 16
  int rows_valid;    // rows of an utterance that hold outputs: ceil(T_out / 16)
  int copy_bytes;    // T_in rounded up to 16
  int nb, nbc, items;
  int nbuf;
  uint32_t tape_bytes;   // one tape buffer
  float inv_rpu;
  uint32_t idesc;
  unsigned long long* stamp;
  EpiArgs ep;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}
// K-major operand without swizzle (layout type 0): start address, LBO, SBO in 16-byte units, descriptor version 1
__device__ __forceinline__ uint64_t desc_noswizzle(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(kThreads, 1) dwconv_tape_kernel(const TapeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bmat_base = smem_base + (uint32_t)a.nbuf * a.tape_bytes;
  const uint32_t bars = bmat_base + (uint32_t)a.nbuf * kMaxKS * kBmatBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxBufs + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kMaxBufs + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kMaxBufs + kSlots + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxBufs + 2 * kSlots);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int i0 = (int)((long long)blockIdx.x * a.items / gridDim.x);
  const int i1 = (int)((long long)(blockIdx.x + 1) * a.items / gridDim.x);

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kMaxBufs; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < kSlots; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kAllocWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the tapes start as zeros: the gaps between the utterance rows (= the conv padding) are never written again
  {
    uint4* z = reinterpret_cast<uint4*>(smem_gen);
    const int n16 = (int)(((uint32_t)a.nbuf * a.tape_bytes) >> 4);
    for (int i = threadIdx.x; i < n16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros before async-proxy copies / MMA reads
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nbc = a.nbc, nb = a.nb, rpu = a.rpu;

  if (warp == kProducerWarp) {
    // ===================== producer: one bulk copy per utterance row + the channel's weight matrix =====================
    pdl_wait();   // the activation rows come from the previous kernel
    if (lane == 0) qasr_stamp_begin(a.stamp);
    int buf = 0;
    uint32_t phase = 0;
    int c = i0 / nbc, bc = i0 - c * nbc;
    for (int it = i0; it < i1; ++it) {
      const int b0 = bc * nb, nrows = min(nb, a.B - b0);
      mbar_wait(empty_bar(buf), phase ^ 1u);
      if (lane == 0) {
        mbar_expect_tx(full_bar(buf), (uint32_t)(nrows * a.copy_bytes) + (uint32_t)a.KS * kBmatBytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(bmat_base + (uint32_t)buf * kMaxKS * kBmatBytes), "l"(a.bmat + (size_t)c * a.KS * kBmatBytes),
                       "r"((uint32_t)a.KS * kBmatBytes), "r"(full_bar(buf)) : "memory");
      }
      __syncwarp();
      const uint32_t dst0 = smem_base + (uint32_t)buf * a.tape_bytes + (uint32_t)a.G;
      const uint8_t* src0 = a.x + ((size_t)b0 * a.C + c) * a.x_ld;
      for (int r = lane; r < nrows; r += 32)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst0 + (uint32_t)(r * a.P)), "l"(src0 + (size_t)r * a.C * a.x_ld), "r"((uint32_t)a.copy_bytes),
                       "r"(full_bar(buf)) : "memory");
      if (++buf == a.nbuf) { buf = 0; phase ^= 1u; }
      if (++bc == nbc) { bc = 0; ++c; }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc = a.idesc;
    const int KS = a.KS;
    int buf = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
    int bc = i0 % nbc;
    for (int it = i0; it < i1; ++it) {
      const int nrows = min(nb, a.B - bc * nb);
      const int units = (nrows * rpu + 127) >> 7;
      mbar_wait(full_bar(buf), phase);
      tc_fence_after();
      const uint32_t tape = smem_base + (uint32_t)buf * a.tape_bytes;
      const uint32_t bm = bmat_base + (uint32_t)buf * kMaxKS * kBmatBytes;
      for (int u = 0; u < units; u += 4) {
        mbar_wait(tempty_bar(slot), sphase ^ 1u);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (u + g < units) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 64 + g * 16);
              const uint32_t a_addr = tape + (uint32_t)(u + g) * 2048u;
              for (int ks = 0; ks < KS; ++ks)
                tc_mma_i8(d_tmem, desc_noswizzle(a_addr + 32u * ks, 16u, 128u),
                          desc_noswizzle(bm + (uint32_t)ks * kBmatBytes, 128u, 256u), idesc, ks != 0 ? 1u : 0u);
            }
          }
          tc_commit(tfull_bar(slot));
        }
        __syncwarp();
        if (++slot == kSlots) { slot = 0; sphase ^= 1u; }
      }
      if (leader) tc_commit(empty_bar(buf));   // the tape is free once every MMA that reads it has retired
      __syncwarp();
      if (++buf == a.nbuf) { buf = 0; phase ^= 1u; }
      if (++bc == nbc) bc = 0;
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue: TMEM -> codes -> one 16-byte store per lane =====================
    const int quarter = warp & 3, grp = warp >> 2;
    const EpiConsts kc = make_epi_consts(a.ep, false);
    const bool masked = a.ep.flags & QASR_F_MASK;
    const bool literal_all = a.ep.flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64);
    const bool dyn = a.ep.flags & QASR_F_DEVICE_TABLES;
    pdl_wait();   // out / out_lens (and device-written tables) belong to the stream order
    int slot = 0, rot = 0;   // rot: which position of a round this group takes; advances so units go round-robin over groups
    uint32_t sphase = 0;
    int c = i0 / nbc, bc = i0 - c * nbc;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int it = i0; it < i1; ++it) {
      const int b0 = bc * nb, nrows = min(nb, a.B - b0);
      const int rows_total = nrows * rpu, units = (rows_total + 127) >> 7;
      const ChanRegs c0 = load_chan0(a.ep, c);
      // 0 = float fast path, 1 = integer fast path, 2 = literal recipe: one channel per item, so warp-uniform
      const int mode = c0.r != 0.0f ? 0 : ((literal_all && !(dyn && !c0.literal)) ? 2 : 1);
      uint8_t* const out_c = reinterpret_cast<uint8_t*>(a.out) + (size_t)c * a.out_ld;
      for (int u = 0; u < units; u += 4) {
        mbar_wait(tfull_bar(slot), sphase);
        tc_fence_after();
        int acc[16];
        const int pos = (grp + rot) & 3;
        const bool have = u + pos < units;
        if (have) {
          tmem_ld16(lane_addr + (uint32_t)(slot * 64 + pos * 16), acc);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive_warp(tempty_bar(slot));
        if (++slot == kSlots) { slot = 0; sphase ^= 1u; }
        if (!have) continue;
        const int R = (u + pos) * 128 + quarter * 32 + lane;      // tape row of this lane
        const int bl = __float2int_rz(((float)R + 0.5f) * a.inv_rpu);
        const int i = R - bl * rpu;
        if (R >= rows_total || i >= a.rows_valid) continue;       // a row of the gap: no output frames
        const int b = b0 + bl, t0 = 16 * i;
        const int len = masked ? min(a.ep.out_lens[b], a.T_out) : a.T_out;
        uint32_t w[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          int q[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int v = acc[4 * q4 + j];
            q[j] = mode == 0 ? (int)epilogue_f32_bits(kc, v, c0)
                             : (mode == 1 ? epilogue_fast<false, false, false>(a.ep, kc, v, c0, 0, c0, 0)
                                          : epilogue_exact(a.ep, kc, v, c0, false, 0, c0, 0));
          }
          w[q4] = pack4(q[0], q[1], q[2], q[3]);
        }
        if (t0 + 16 > len) {   // length mask / row tail: zero the bytes at t >= len
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int nvalid = len - (t0 + 4 * q4);
            w[q4] &= nvalid >= 4 ? 0xffffffffu : (nvalid <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nvalid))));
          }
        }
        *reinterpret_cast<uint4*>(out_c + (size_t)b * a.C * a.out_ld + t0) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      rot = (rot - units) & 3;   // the next item's first unit goes to the group after the one that took this item's last
      if (++bc == nbc) { bc = 0; ++c; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) qasr_stamp_end(a.stamp);
  if (warp == kAllocWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// [C][k] int8 taps -> [C][KS][512 B]: the 16 x 32 K-major, non-swizzled weight matrix of every K step,
// byte (n, kk) of step ks at  ks * 512 + (n / 8) * 256 + ((kk % 32) / 16) * 128 + (n % 8) * 16 + kk % 16
__global__ void pack_tape_kernel(const int8_t* __restrict__ w, uint8_t* __restrict__ bm, int C, int k, int dil, int d, int KS) {
  const long long total = (long long)C * KS * 512;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx & 511), ks = (int)((idx >> 9) % KS), c = (int)(idx / ((long long)KS * 512));
    const int n = (o >> 8) * 8 + ((o >> 4) & 7);
    const int kk = 32 * ks + ((o >> 7) & 1) * 16 + (o & 15);
    const int u = kk - n - d;
    int8_t v = 0;
    if (u >= 0 && u % dil == 0 && u / dil < k) v = w[(size_t)c * k + u / dil];
    bm[idx] = (uint8_t)v;
  }
}

inline int round_up_i(int a, int b) { return (a + b - 1) / b * b; }

struct TapeGeom {
  int G, d, KS;
};
inline TapeGeom tape_geom(int k, int dil, int pad) {
  TapeGeom g;
  g.G = round_up_i(pad > 0 ? pad : 1, 16);
  g.d = g.G - pad;
  g.KS = (16 + (k - 1) * dil + g.d + 31) / 32;
  return g;
}

}  // namespace

bool qasr_dwconv_tape_supported(int k, int stride, int dilation, int pad) {
  if (stride != 1 || dilation < 1 || pad < 0 || k < 1) return false;
  return tape_geom(k, dilation, pad).KS <= kMaxKS;
}

int64_t qasr_dwconv_tape_packed_bytes(int C, int k, int stride, int dilation, int pad) {
  if (!qasr_dwconv_tape_supported(k, stride, dilation, pad)) return 0;
  return (int64_t)C * tape_geom(k, dilation, pad).KS * kBmatBytes;
}

int qasr_dwconv_tape_pack(const int8_t* w, uint8_t* bm, int C, int k, int dilation, int pad, cudaStream_t st) {
  const TapeGeom g = tape_geom(k, dilation, pad);
  const long long total = (long long)C * g.KS * 512;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  pack_tape_kernel<<<blocks, 256, 0, st>>>(w, bm, C, k, dilation, g.d, g.KS);
  return qasr_check_launch("pack_tape_kernel");
}

int qasr_dwconv_tape_launch(const void* x, int x_ld, const uint8_t* bmat, void* out, int out_ld, int B, int C, int T_in,
                            int T_out, int k, int dilation, int pad, const qasr_epilogue* epi, void* stream) {
  const TapeGeom g = tape_geom(k, dilation, pad);
  TapeArgs a;
  memset(&a, 0, sizeof(a));
  a.x = static_cast<const uint8_t*>(x);
  a.bmat = bmat;
  a.out = out;
  a.x_ld = x_ld; a.out_ld = out_ld; a.B = B; a.C = C; a.T_in = T_in; a.T_out = T_out;
  a.KS = g.KS; a.G = g.G;
  a.copy_bytes = round_up_i(T_in, 16);
  a.P = a.G + a.copy_bytes;
  a.rpu = a.P / 16;
  a.rows_valid = (T_out + 15) / 16;
  a.inv_rpu = 1.0f / (float)a.rpu;
  // utterances per work item: tapes of <= ~24 KB (a few units each), and enough items to balance 148 CTAs
  const int sms = qasr_sm_count();
  int nb = 24 * 1024 / a.P;
  if (nb < 1) nb = 1;
  if (nb > B) nb = B;
  while (nb > 4 && (long long)C * ((B + nb - 1) / nb) < 8LL * sms) nb = (nb + 1) / 2;
  nb = (B + (B + nb - 1) / nb - 1) / ((B + nb - 1) / nb);   // same number of chunks, evenly sized
  if (qasr_debug_get(15) > 0) nb = (int)qasr_debug_get(15) < B ? (int)qasr_debug_get(15) : B;   // bring-up knob 15: utterances per item
  a.nb = nb;
  a.nbc = (B + nb - 1) / nb;
  a.items = C * a.nbc;
  const int units_max = (nb * a.rpu + 127) / 128;
  a.tape_bytes = (uint32_t)round_up_i(units_max * 2048 + 32 * a.KS + 16, 128);
  const uint32_t fixed = 128 /*align*/ + 8u * (2 * kMaxBufs + 2 * kSlots + 2);
  int nbuf = (int)((kSmemBudget - fixed) / (a.tape_bytes + kMaxKS * kBmatBytes));
  if (nbuf > kMaxBufs) nbuf = kMaxBufs;
  if (nbuf < 2) {
    qasr_set_error("dwconv(tape): rows of %d frames do not fit two tape buffers in shared memory", T_in);
    return QASR_ERR_UNSUPPORTED;
  }
  a.nbuf = nbuf;
  const uint32_t a_signed = (epi->flags & QASR_F_B_SIGNED) ? 1u : 0u;
  // instruction descriptor: D = S32, A = activations (un)signed 8-bit K-major, B = weights signed 8-bit K-major, N = 16, M = 128
  a.idesc = (2u << 4) | (a_signed << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(16 >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  a.ep = make_epi_args(epi);
  a.stamp = qasr_next_stamp();
  const size_t smem = fixed + (size_t)nbuf * (a.tape_bytes + kMaxKS * kBmatBytes);
  int grid = sms < a.items ? sms : a.items;
  static bool attr_set[64] = {};
  int dev = 0;
  QASR_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    QASR_CHECK_CUDA(cudaFuncSetAttribute(dwconv_tape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    attr_set[dev] = true;
  }
  QASR_CHECK_CUDA(qasr_launch_pdl(dwconv_tape_kernel, dim3(grid), dim3(kThreads), smem, (cudaStream_t)stream, a));
  return qasr_check_launch("dwconv_tape_kernel");
}

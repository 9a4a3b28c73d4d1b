// K1: depthwise (time-channel-separable) 1-D convolution on 8-bit activation codes with the fused
// requantisation epilogue.  Restates F.conv1d(x_int.double(), w_int.double(), groups=C) of the
// reference's QuantConv1d.int_conv (nemo/quantization/utils/quant_modules.py:301-305) in true
// integers.
//
// The op moves 2 bytes per output and needs k = 33..87 MACs per output: on the CUDA-core int pipe
// (dp4a measured at 64 lanes/clk/SM on B200 = 18.5 T dp4a/s) it is ALU-bound at 2-4x the HBM time.
// It is therefore computed as a banded (Toeplitz) matrix product on the warp-level int8 tensor path:
//
//   one warp tile = 128 consecutive output frames of one (utterance, channel) row, viewed as
//       D[16 x 8],  D[i][j] = out[t0 + 16 j + i]
//       A[16 x K],  A[i][kk] = w[(kk - S i - d) / dil]      (banded; depends on the channel only)
//       B[K  x 8],  B[kk][j] = x[P + 16 S j + kk],  P = S t0 - pad - d   (4-byte aligned)
//   K = 32 * KS covers 15 S + d + (k-1) dil + 1 taps; d = (-pad) mod 8 absorbs the misalignment.
//   The K index of the MMA is a PERMUTATION of the window bytes, chosen so that the 2 KS words a lane needs are
//   contiguous: k' = 32 ks + 16 h + 4 tg + e  <->  window byte v(k') = 4 (2 KS tg + 2 ks + h) + e; A is packed with
//   the same permutation, so the sum over k is unchanged and every K step costs one 8-byte LDS instead of two
//   4-byte ones.
//   * A fragments (KS x 4 registers per lane) are packed once per layer by qasr_pack_dw_weights and
//     loaded with KS*4 coalesced words per warp, then reused for every tile of that channel
//     (several utterances per warp).
//   * Input rows are staged in shared memory by the bulk-copy engine (cp.async.bulk, one copy per
//     row, completion on an mbarrier), kBufs rows per warp in flight, with zero halos on both sides
//     of the row standing in for the conv padding: the HBM read is one coalesced stream per row and
//     is issued two rows ahead of its use.
//   * B fragments are plain aligned 64-bit words of the staged row: lane (g, tg) needs the word pairs
//     (P/4 + 4 S g + 2 KS tg + 2 ks, +1), no bounds checks.
//   * D (int32, 4 per lane) goes through the fused epilogue, the 4 result bytes are transposed
//     through a 128-byte per-warp shared-memory line so the warp stores 128 contiguous bytes.
// mma.sync m16n8k32 (IMMA) was measured at 0.48 instr/clk/SM (1.14 POPS) on B200: the MMA part of
// a 512-channel layer at batch 64 x 800 frames is ~4.6 us, below its 8 us HBM time.
#include <type_traits>

#include "qasr_common.cuh"
#include "dw_common.cuh"

#ifndef QASR_TAPE_DEFAULT
#define QASR_TAPE_DEFAULT 0   // engine choice without knob 8: 1 = tape engine wherever it applies
#endif

namespace {

// EK: 0 = REQUANT all-integer fast path, 1 = REQUANT literal path, 2 = DEQUANT_F32 / RAW_I32 / STATS,
//     3 = REQUANT float fast path (QASR_F_FAST_F32), 4 = REQUANT with device-written tables (QASR_F_DEVICE_TABLES):
//     integer fast path for the channels qasr_dyn_finalize proved, literal path for the others
// PROF: in-graph timing stamps (qasr_profile_begin).  A separate instantiation: the two extra 64-bit values cost the
// product kernel 4-6 registers and with them a resident CTA per SM (measured: +3 % step time).
template <int KS, int S, bool XS, int EK, bool PROF>
#ifndef QASR_DW_MIN_BLOCKS
#define QASR_DW_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(kThreads, QASR_DW_MIN_BLOCKS) dwconv_mma_kernel(const DwArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t stage[kWarps][128];  // four 128-byte transpose lines per warp (two per loop iteration, alternating)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  pdl_trigger();
  const int item = blockIdx.x * kWarps + warp;
  if (item >= a.C * a.nbg) return;  // whole warps exit together; only warp-level sync below
  const int c = item / a.nbg, bg = item - c * a.nbg;
  const int b_begin = bg * a.bchunk;
  const int nrows = min(a.B, b_begin + a.bchunk) - b_begin;

  // ---- per-warp staging ring: kBufs x [halo zeros | row | zeros], one mbarrier each
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * kBufs;
  uint8_t* bufs = smem + 128 + (size_t)warp * kBufs * a.buf_bytes;
  for (int i = lane; i < kBufs * a.buf_bytes / 16; i += 32)
    reinterpret_cast<uint4*>(bufs)[i] = make_uint4(0, 0, 0, 0);
  if (lane == 0) {
    for (int i = 0; i < kBufs; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + i)));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // zero fill (generic proxy) before the async copies
  __syncwarp();
  const uint8_t* xbase = a.x + ((size_t)b_begin * a.C + c) * a.x_ld;
  const size_t xstep = (size_t)a.C * a.x_ld;  // next utterance, same channel
  // ---- banded A fragments of this channel: pre-packed per lane by qasr_pack_dw_weights (constants of the plan:
  //      loaded before the dependency wait so they overlap the previous kernel's tail)
  uint32_t af[KS][4];
  {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a.w) + (size_t)c * KS * 128 + lane;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int r = 0; r < 4; ++r) af[ks][r] = __ldg(wp + (ks * 4 + r) * 32);
    }
  }
  const ChanRegs c0 = load_chan0(a.ep, c);
  const bool f32_path = c0.r != 0.0f;   // warp-uniform: one channel per work item
  const EpiConsts kc = make_epi_consts(a.ep, false);
  pdl_wait();   // activations (and the output buffer) belong to the stream order; everything above is set-up
  if (lane == 0) {
    // thousands of warps on ONE address at the same moment cost a stamped launch microseconds: the first CTA on data is
    // one of the first the grid dispatched (see dwconv_rt.cu)
    if (PROF && warp == 0 && blockIdx.x < 296) qasr_stamp_begin(a.stamp);
    for (int r = 0; r < min(nrows, kBufs - 1); ++r)
      row_copy_async(smem_addr(bufs + r * a.buf_bytes + a.halo), xbase + r * xstep, a.copy_bytes,
                     smem_addr(bars + r));
  }

  // offsets of this lane's 4 accumulators inside the 128-frame tile: D[row][col] -> 16 col + row
  const int o0 = 32 * tg + g, o1 = o0 + 16, o2 = o0 + 8, o3 = o0 + 24;
  uint8_t* sb = reinterpret_cast<uint8_t*>(stage[warp]);
  // word offset of this lane's first B word relative to a tile: window start = halo - pad - d (bytes)
  const int lane_w = ((a.halo - a.pad - a.d) >> 2) + 4 * S * g + 2 * KS * tg;   // even: 8-byte aligned pairs
  const int out_words = ((a.T_out + 15) >> 4) << 2;  // stores cover T_out rounded up to 16 (zero filled)

  // per-row bookkeeping is carried incrementally (no modulo / 64-bit multiplies on the row path)
  const bool masked = a.ep.flags & QASR_F_MASK;
  const int T_out = a.T_out, buf_bytes = a.buf_bytes;
  const uint32_t copy_bytes = (uint32_t)a.copy_bytes;
  const size_t out_row_step = (size_t)a.C * a.out_ld * (a.ep.mode == QASR_EPI_REQUANT_I8 ? 1 : 4);
  uint8_t* out_row = reinterpret_cast<uint8_t*>(a.out) + ((size_t)b_begin * a.C + c) * a.out_ld * (a.ep.mode == QASR_EPI_REQUANT_I8 ? 1 : 4);
  const uint8_t* x_next = xbase + (size_t)(kBufs - 1) * xstep;       // row r + kBufs - 1
  int buf = 0, nb = kBufs - 1;
  uint32_t par = 0;
  long long st_lo = 0x7fffffffffffffffLL, st_hi = -0x7fffffffffffffffLL - 1;   // QASR_EPI_STATS: this lane's extremes
  // ... kept in 32 bits while acc + bias cannot leave int32 (|acc| < 2^22 for every depthwise window this kernel takes):
  // two VIMNMX per counted output instead of 64-bit compares and selects
  const bool st_narrow = c0.bias > -(1 << 30) && c0.bias < (1 << 30);
  int st_lo32 = 2147483647, st_hi32 = -2147483647 - 1;
  for (int r = 0; r < nrows; ++r, out_row += out_row_step, x_next += xstep) {
    // keep kBufs-1 rows in flight: the buffer being refilled was consumed in iteration r-1
    if (lane == 0 && r + kBufs - 1 < nrows) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      row_copy_async(smem_addr(bufs + nb * buf_bytes + a.halo), x_next, copy_bytes, smem_addr(bars + nb));
    }
    bar_wait(smem_addr(bars + buf), par);
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(bufs + buf * buf_bytes) + lane_w;
    if (++nb == kBufs) nb = 0;
    if (++buf == kBufs) { buf = 0; par ^= 1u; }
    const int b = b_begin + r;
    const int len = masked ? min(a.ep.out_lens[b], T_out) : T_out;
    // the tile loop exists twice, float and integer epilogue, chosen once per row (warp-uniform)
    auto tiles = [&](auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;   // 0 = integer fast path, 1 = float fast path, 2 = literal recipe
    constexpr bool F32 = MODE == 1;
    // float path: the MMA's initial accumulator is 0x4B400000 + bias, so the finished accumulator's BITS are the
    // float 1.5 * 2^23 + (acc + bias): the bias add and the int -> float conversion cost no instruction
    const int acc_init = F32 ? kF32MagicBits + c0.bias : 0;
    // stores of whole tiles below `full_end` need neither the length mask nor the row-end test (warp-uniform split)
    const int full_end = min(len, out_words * 4);
    uint32_t* const orow_w = reinterpret_cast<uint32_t*>(out_row) + lane;
    auto code = [&](int acc_v) -> uint8_t {
      if (F32) return (uint8_t)epilogue_f32_bits_biased(kc, acc_v, c0);
      return (uint8_t)(MODE == 0 ? epilogue_fast<false, false, false>(a.ep, kc, acc_v, c0, 0, c0, 0)
                                 : epilogue_exact(a.ep, kc, acc_v, c0, false, 0, c0, 0));
    };
    // finished accumulators of the tile at t0 -> codes -> (transpose through the warp's staging line `sl`) -> store
    auto finish_tile = [&](int t0, const int (&acc)[4], int sl) {
      if (EK != 2) {
        uint8_t* sbl = sb + 128 * sl;
        sbl[o0] = code(acc[0]);
        sbl[o1] = code(acc[1]);
        sbl[o2] = code(acc[2]);
        sbl[o3] = code(acc[3]);
        __syncwarp();
        // no second barrier: the next iteration writes the OTHER pair of lines, and the barrier of that iteration
        // orders these reads before the writes that reuse this line two iterations later
        uint32_t word = stage[warp][32 * sl + lane];
        if (t0 + kTile <= full_end) {
          orow_w[t0 >> 2] = word;
        } else {   // last tile(s) of the row: length mask (zero the bytes at t >= len) and row end
          const int tl = t0 + 4 * lane;
          if (tl + 4 > len) {
            const int nvalid = len - tl;
            word &= nvalid <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nvalid)));
          }
          if ((tl >> 2) < out_words) orow_w[t0 >> 2] = word;
        }
      } else if (a.ep.mode == QASR_EPI_STATS) {
        // statistics pass of dynamic quantisation: extremes of acc + bias over the counted frames of this channel
        const int offs[4] = {o0, o1, o2, o3};
        if (st_narrow) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool counted = t0 + offs[i] < len;
            const int z = acc[i] + c0.bias;
            st_lo32 = min(st_lo32, counted ? z : 2147483647);
            st_hi32 = max(st_hi32, counted ? z : -2147483647 - 1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (t0 + offs[i] < len) {
              const long long z = qasr_acc_bias64(acc[i], c0.bias);
              st_lo = z < st_lo ? z : st_lo;
              st_hi = z > st_hi ? z : st_hi;
            }
          }
        }
      } else if (a.ep.mode == QASR_EPI_DEQUANT_F32) {
        float* orow = reinterpret_cast<float*>(out_row);
        if (t0 + o0 < a.T_out) orow[t0 + o0] = epilogue_dequant(a.ep, acc[0], c0);
        if (t0 + o1 < a.T_out) orow[t0 + o1] = epilogue_dequant(a.ep, acc[1], c0);
        if (t0 + o2 < a.T_out) orow[t0 + o2] = epilogue_dequant(a.ep, acc[2], c0);
        if (t0 + o3 < a.T_out) orow[t0 + o3] = epilogue_dequant(a.ep, acc[3], c0);
      } else {
        int32_t* orow = reinterpret_cast<int32_t*>(out_row);
        if (t0 + o0 < a.T_out) orow[t0 + o0] = acc[0] + c0.bias;
        if (t0 + o1 < a.T_out) orow[t0 + o1] = acc[1] + c0.bias;
        if (t0 + o2 < a.T_out) orow[t0 + o2] = acc[2] + c0.bias;
        if (t0 + o3 < a.T_out) orow[t0 + o3] = acc[3] + c0.bias;
      }
    };
    // two tiles per iteration: their dependent MMA chains interleave (no issue slots lost to the MMA latency) and
    // their epilogues give the scheduler independent work
    auto mma_pair = [&](const uint32_t* tw0, const uint32_t* tw1, int (&acc0)[4], int (&acc1)[4]) {
      const uint2* p0 = reinterpret_cast<const uint2*>(tw0);
      const uint2* p1 = reinterpret_cast<const uint2*>(tw1);
      uint2 b0 = p0[0], b1 = p1[0];
      imma_16832_init<XS>(acc0, af[0], b0.x, b0.y, acc_init);
      imma_16832_init<XS>(acc1, af[0], b1.x, b1.y, acc_init);
#pragma unroll
      for (int ks = 1; ks < KS; ++ks) {
        b0 = p0[ks];
        b1 = p1[ks];
        imma_16832<XS>(acc0, af[ks], b0.x, b0.y);
        imma_16832<XS>(acc1, af[ks], b1.x, b1.y);
      }
    };
    int t0 = 0, sp = 0;   // sp: which pair of staging lines this iteration uses
    if (EK != 2) {
      // pairs of whole tiles below `full_end`: no mask, no row-end test, pointers advance by constants and the
      // staging lines are compile-time constants (the loop body is two pairs: lines 0-1, then lines 2-3)
      const uint32_t* tw = xw;
      uint32_t* ow = orow_w;
      constexpr int kPairIn = (2 * kTile * S) >> 2, kPairOut = (2 * kTile) >> 2;
      auto pair_fast = [&](auto line_tag, const uint32_t* twp, uint32_t* owp) {
        constexpr int LINE = decltype(line_tag)::value;
        int acc0[4], acc1[4];
        mma_pair(twp, twp + ((kTile * S) >> 2), acc0, acc1);
        uint8_t* const l0 = sb + 128 * LINE;
        l0[o0] = code(acc0[0]); l0[o1] = code(acc0[1]); l0[o2] = code(acc0[2]); l0[o3] = code(acc0[3]);
        l0[128 + o0] = code(acc1[0]); l0[128 + o1] = code(acc1[1]); l0[128 + o2] = code(acc1[2]); l0[128 + o3] = code(acc1[3]);
        __syncwarp();
        // no second barrier: the next pair writes the OTHER two lines (see finish_tile)
        const uint32_t w0 = stage[warp][32 * LINE + lane], w1 = stage[warp][32 * LINE + 32 + lane];
        owp[0] = w0;
        owp[kTile >> 2] = w1;
      };
      for (; t0 + 4 * kTile <= full_end; t0 += 4 * kTile, tw += 2 * kPairIn, ow += 2 * kPairOut) {
        pair_fast(std::integral_constant<int, 0>{}, tw, ow);
        pair_fast(std::integral_constant<int, 2>{}, tw + kPairIn, ow + kPairOut);
      }
      if (t0 + 2 * kTile <= full_end) {
        pair_fast(std::integral_constant<int, 0>{}, tw, ow);
        t0 += 2 * kTile;
        sp = 2;
      }
    }
    for (; t0 + kTile < a.T_out; t0 += 2 * kTile, sp ^= 2) {
      const uint32_t* tw0 = xw + ((t0 * S) >> 2);
      int acc0[4], acc1[4];
      mma_pair(tw0, tw0 + ((kTile * S) >> 2), acc0, acc1);
      finish_tile(t0, acc0, sp);
      finish_tile(t0 + kTile, acc1, sp + 1);
    }
    if (t0 < a.T_out) {   // odd tile count: the last one alone
      const uint32_t* tw = xw + ((t0 * S) >> 2);
      const uint2* pw2 = reinterpret_cast<const uint2*>(tw);
      int acc[4];
      uint2 bw = pw2[0];
      imma_16832_init<XS>(acc, af[0], bw.x, bw.y, acc_init);
#pragma unroll
      for (int ks = 1; ks < KS; ++ks) {
        bw = pw2[ks];
        imma_16832<XS>(acc, af[ks], bw.x, bw.y);
      }
      finish_tile(t0, acc, sp);
    }
    };
    if (EK == 3 || (EK == 0 && f32_path)) tiles(std::integral_constant<int, 1>{});
    else if (EK == 0 || (EK == 4 && !c0.literal)) tiles(std::integral_constant<int, 0>{});
    else tiles(std::integral_constant<int, 2>{});
    __syncwarp();  // every lane is done reading this buffer before lane 0 refills it
  }
  if (PROF && lane == 0) qasr_stamp_end(a.stamp);
  if (EK == 2 && a.ep.mode == QASR_EPI_STATS) {
    if (st_narrow && st_hi32 >= st_lo32) { st_lo = st_lo32; st_hi = st_hi32; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long l2 = __shfl_xor_sync(0xffffffffu, st_lo, o), h2 = __shfl_xor_sync(0xffffffffu, st_hi, o);
      st_lo = l2 < st_lo ? l2 : st_lo;
      st_hi = h2 > st_hi ? h2 : st_hi;
    }
    if (lane == 0 && st_hi >= st_lo) {
      qasr_stats* st = reinterpret_cast<qasr_stats*>(a.out) + c;
      atomicMin(reinterpret_cast<long long*>(&st->lo), st_lo);
      atomicMax(reinterpret_cast<long long*>(&st->hi), st_hi);
    }
  }
}

template <int KS, int S>
int launch(const DwArgs& a, cudaStream_t stream) {
  const int items = a.C * a.nbg;
  const int grid = (items + kWarps - 1) / kWarps;
  const size_t smem = 128 + (size_t)kWarps * kBufs * a.buf_bytes;
  const bool xs = a.ep.flags & QASR_F_B_SIGNED;
  const bool fast = !(a.ep.flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64 | QASR_F_STAGE2)) &&
                    a.ep.prev == nullptr;
  int ek = a.ep.mode != QASR_EPI_REQUANT_I8 ? 2 : ((a.ep.flags & QASR_F_FAST_F32) ? 3 : (fast ? 0 : 1));
  if (ek == 1 && (a.ep.flags & QASR_F_DEVICE_TABLES) && !(a.ep.flags & QASR_F_STAGE2) && a.ep.prev == nullptr) ek = 4;
  // stamped instantiations exist for the kernels of the static plan (EK 0 and 3); other launches carry no stamp
  const bool prof = a.stamp != nullptr && (ek == 0 || ek == 3);
#define QASR_DW_GO(XS_, EK_)                                                                         \
  do {                                                                                               \
    auto kern = (prof && (EK_ == 0 || EK_ == 3)) ? dwconv_mma_kernel<KS, S, XS_, EK_, (EK_ == 0 || EK_ == 3)>  \
                                                 : dwconv_mma_kernel<KS, S, XS_, EK_, false>;        \
    static size_t attr_smem[2] = {0, 0};                                                             \
    if (smem > 48 * 1024 && smem > attr_smem[prof]) {                                                \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr_smem[prof] = smem;                                                                        \
    }                                                                                                \
    QASR_CHECK_CUDA(qasr_launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, a));             \
  } while (0)
  if (xs) {
    if (ek == 0) QASR_DW_GO(true, 0); else if (ek == 1) QASR_DW_GO(true, 1); else if (ek == 2) QASR_DW_GO(true, 2); else if (ek == 3) QASR_DW_GO(true, 3); else QASR_DW_GO(true, 4);
  } else {
    if (ek == 0) QASR_DW_GO(false, 0); else if (ek == 1) QASR_DW_GO(false, 1); else if (ek == 2) QASR_DW_GO(false, 2); else if (ek == 3) QASR_DW_GO(false, 3); else QASR_DW_GO(false, 4);
  }
#undef QASR_DW_GO
  return qasr_check_launch("dwconv_mma_kernel");
}

// [C][k] int8 taps -> [C][KS][4 regs][32 lanes] A-fragment words of the banded matrix
__global__ void pack_dw_kernel(const int8_t* __restrict__ w, uint32_t* __restrict__ wp, int C, int k, int S,
                               int dil, int d, int KS) {
  const int total = C * KS * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, r = (i >> 5) & 3, ks = (i >> 7) % KS, c = i / (KS * 128);
    const int g = lane >> 2, tg = lane & 3;
    const int row = g + 8 * (r & 1);
    // MMA k index 32 ks + 16 (r >> 1) + 4 tg + bb <-> window byte 4 (2 KS tg + 2 ks + (r >> 1)) + bb (see the header)
    const int v0 = 4 * (2 * KS * tg + 2 * ks + (r >> 1));
    uint32_t word = 0;
    for (int bb = 0; bb < 4; ++bb) {
      const int u = v0 + bb - S * row - d;
      int v = 0;
      if (u >= 0) {
        const int tap = u / dil;
        if (tap * dil == u && tap < k) v = w[(size_t)c * k + tap];
      }
      word |= (uint32_t)(v & 0xff) << (8 * bb);
    }
    wp[i] = word;
  }
}

}  // namespace

// tensor-core variant (dwconv_tc.cu)
bool qasr_dwconv_tc_supported(int B, int C, int T_in, int T_out, int k, int stride, int dilation, int pad);
int qasr_dwconv_tc_launch(const void* x, int x_ld, const int8_t* taps, void* out, int out_ld, int B, int C, int T_in,
                          int T_out, int k, int stride, int dilation, int pad, const qasr_epilogue* epi, void* stream);

// tape engine (dwconv_tape.cu): tcgen05 on the utterance rows themselves through overlapping no-swizzle descriptors
bool qasr_dwconv_tape_supported(int k, int stride, int dilation, int pad);
int64_t qasr_dwconv_tape_packed_bytes(int C, int k, int stride, int dilation, int pad);
int qasr_dwconv_tape_pack(const int8_t* w, uint8_t* bm, int C, int k, int dilation, int pad, cudaStream_t st);
int qasr_dwconv_tape_launch(const void* x, int x_ld, const uint8_t* bmat, void* out, int out_ld, int B, int C, int T_in,
                            int T_out, int k, int dilation, int pad, const qasr_epilogue* epi, void* stream);

// row-tape walk of the IMMA engine (dwconv_rt.cu): stride 1, requantising fast epilogues
bool qasr_dwconv_rt_supported(int T_in, int k, int dilation, int pad);
int qasr_dwconv_rt_launch(const void* x, int x_ld, const int8_t* w, void* out, int out_ld, int B, int C, int T_in, int T_out,
                          int k, int dilation, int pad, int ek, const qasr_epilogue* epi, void* stream);

// packed layout: [IMMA A-fragment table (absent when the tap window exceeds 6 x 32)] [raw taps [C][k], for the
// tcgen05 kernel which builds its banded matrices on the fly] [16 x K weight matrices of the tape engine]
static int64_t dw_table_bytes(int C, int k, int stride, int dilation, int pad) {
  const int ks = dw_ksteps(k, stride, dilation, pad);
  return ks > 6 ? 0 : (int64_t)C * ks * 128 * 4;
}

extern "C" int64_t qasr_dw_packed_bytes(int32_t C, int32_t k, int32_t stride, int32_t dilation, int32_t pad) {
  if (C <= 0 || k <= 0 || (stride != 1 && stride != 2) || dilation < 1 || pad < 0) return -1;
  const int64_t table = dw_table_bytes(C, k, stride, dilation, pad);
  if (table == 0 && !qasr_dwconv_tc_supported(1, C, 1, 1, k, stride, dilation, pad)) return -1;
  return table + (((int64_t)C * k + 15) / 16) * 16 + qasr_dwconv_tape_packed_bytes(C, k, stride, dilation, pad);
}

extern "C" int qasr_pack_dw_weights(const int8_t* w, void* wp, int32_t C, int32_t k, int32_t stride,
                                    int32_t dilation, int32_t pad, void* stream) {
  QASR_CHECK_ARG(w && wp && qasr_aligned16(wp), "pack_dw_weights: null / unaligned pointer");
  QASR_CHECK_ARG(qasr_dw_packed_bytes(C, k, stride, dilation, pad) > 0,
                 "pack_dw_weights: unsupported shape (k=%d stride=%d dilation=%d pad=%d)", k, stride, dilation, pad);
  const int ks = dw_ksteps(k, stride, dilation, pad);
  const int64_t table = dw_table_bytes(C, k, stride, dilation, pad);
  QASR_CHECK_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(wp) + table, w, (size_t)C * k, cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream));
  if (qasr_dwconv_tape_supported(k, stride, dilation, pad)) {
    int rc = qasr_dwconv_tape_pack(w, static_cast<uint8_t*>(wp) + table + (((int64_t)C * k + 15) / 16) * 16, C, k, dilation, pad,
                                   (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (table == 0) return QASR_OK;
  const int total = C * ks * 128;
  pack_dw_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w, static_cast<uint32_t*>(wp), C, k, stride, dilation, dw_misalign(pad), ks);
  return qasr_check_launch("pack_dw_kernel");
}

extern "C" int qasr_dwconv1d_i8(const void* x, int32_t x_ld, const int8_t* w, void* out, int32_t out_ld,
                                int32_t B, int32_t C, int32_t T_in, int32_t T_out, int32_t k, int32_t stride,
                                int32_t dilation, int32_t pad, const qasr_epilogue* epi, void* stream) {
  QASR_CHECK_ARG(x && w && out && epi, "dwconv: null pointer");
  QASR_CHECK_ARG(qasr_aligned16(w), "dwconv: packed weights (qasr_pack_dw_weights) must be 16-byte aligned");
  QASR_CHECK_ARG(B > 0 && C > 0 && T_in > 0 && T_out > 0 && k > 0, "dwconv: bad shape");
  QASR_CHECK_ARG(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2 (got %d)", stride);
  QASR_CHECK_ARG(dilation >= 1 && pad >= 0, "dwconv: bad dilation/pad");
  QASR_CHECK_ARG((T_in + 2 * pad - dilation * (k - 1) - 1) / stride + 1 == T_out,
                 "dwconv: T_out %d inconsistent with T_in %d k %d stride %d dil %d pad %d", T_out, T_in, k,
                 stride, dilation, pad);
  QASR_CHECK_ARG(x_ld % 16 == 0 && x_ld >= T_in && qasr_aligned16(x), "dwconv: x pitch/alignment");
  QASR_CHECK_ARG(out_ld >= ((T_out + 15) / 16) * 16 && qasr_aligned16(out),
                 "dwconv: out pitch must cover T_out rounded up to 16");
  QASR_CHECK_ARG(epi->mode != QASR_EPI_REQUANT_I8 || out_ld % 16 == 0, "dwconv: out pitch % 16");
  int rc = qasr_validate_epilogue(epi, false);
  if (rc) return rc;
  qasr_set_plain_launch((epi->flags & QASR_F_DEVICE_TABLES) != 0);

  // engine choice (tools/dw_bench.py, batch 64 x 800 frames x 512 channels, float epilogue): IMMA 16.6 / 18.3 / 30.9 us
  // for k = 33 / 75 / 87-dilated against 21.8 / 23.7 / 27.5 us on tcgen05 with TMA-fetched windows (its input
  // delivery costs 6.5 clk per 128-byte row per SM at 1.6x window overlap = 11 us, which the 20 epilogue warps'
  // drain does not yet hide).  The IMMA kernel runs wherever its 6 x 32-tap window reaches, the tcgen05 kernel
  // covers the longer windows.  Bring-up knob 8 forces one engine (0 = IMMA, 1 = tcgen05).
  const int64_t table_bytes = dw_table_bytes(C, k, stride, dilation, pad);
  const bool tc_ok = qasr_dwconv_tc_supported(B, C, T_in, T_out, k, stride, dilation, pad);
  const long long knob = qasr_debug_get(8);
  bool use_tc = false;
  if (knob == 0 && table_bytes > 0) use_tc = false;
  if (knob == 1 && tc_ok) use_tc = true;
  if (table_bytes == 0) use_tc = true;   // only the tensor-core kernel covers this tap window
  // 6 K steps (k = 87 dilated): the lanes' 8-byte fragment loads are 12 words apart and collide on banks
  // (30.9 us against 27.5 us on tcgen05, 512 channels, batch 64 x 800 frames)
  // ... for layers of at least ~14 M outputs: the tcgen05 kernel pays 11 us per launch before its lower marginal cost shows
  // (BASELINE config 1, batch 8 x 250 frames: 18 us against 4 us on the IMMA engine)
  if (knob < 0 && tc_ok && dw_ksteps(k, stride, dilation, pad) >= 6 && (long long)B * C * T_out >= 14000000LL) use_tc = true;
  // tape engine (tcgen05 on the rows themselves): requantising launches with stride 1.  Bring-up knob 8 = 2 forces it,
  // 0 / 1 force the other engines.
  const bool tape_ok = epi->mode == QASR_EPI_REQUANT_I8 && qasr_dwconv_tape_supported(k, stride, dilation, pad) &&
                       !(epi->flags & QASR_F_STAGE2) && epi->prev == nullptr && out_ld % 16 == 0;
  const bool use_tape = tape_ok && (knob == 2 || (knob < 0 && QASR_TAPE_DEFAULT));
  if (use_tape)
    return qasr_dwconv_tape_launch(x, x_ld, reinterpret_cast<const uint8_t*>(w) + table_bytes + (((int64_t)C * k + 15) / 16) * 16,
                                   out, out_ld, B, C, T_in, T_out, k, dilation, pad, epi, stream);
  if (epi->mode == QASR_EPI_STATS) {   // the statistics epilogue lives in the IMMA engine
    if (table_bytes == 0) {
      qasr_set_error("dwconv: STATS pass needs the IMMA engine (tap window of k=%d dilation=%d exceeds 6 x 32)", k, dilation);
      return QASR_ERR_UNSUPPORTED;
    }
    use_tc = false;
  }
  if (use_tc) {
    if (!tc_ok) {
      qasr_set_error("dwconv: tap span (k=%d, dilation=%d, stride=%d) exceeds both kernels' windows", k, dilation, stride);
      return QASR_ERR_UNSUPPORTED;
    }
    return qasr_dwconv_tc_launch(x, x_ld, w + table_bytes, out, out_ld, B, C, T_in, T_out, k, stride, dilation, pad,
                                 epi, stream);
  }

  // stride-1 launches on the integer / float fast epilogues walk row tapes (dwconv_rt.cu: -25 % warp instructions per
  // row of 800 frames); bring-up knob 16 = 0 keeps them on the row-at-a-time kernel below
  if (stride == 1 && epi->mode == QASR_EPI_REQUANT_I8 && qasr_debug_get(16) != 0 &&
      qasr_dwconv_rt_supported(T_in, k, dilation, pad)) {
    const bool fast = !(epi->flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64 | QASR_F_STAGE2)) &&
                      epi->prev == nullptr;
    int ek = (epi->flags & QASR_F_FAST_F32) ? 3 : (fast ? 0 : 1);
    // device-written tables (dynamic quantisation): per-channel choice between the integer fast path and the literal recipe
    if (ek == 1 && (epi->flags & QASR_F_DEVICE_TABLES) && !(epi->flags & QASR_F_STAGE2) && epi->prev == nullptr) ek = 4;
    if (ek != 1)
      return qasr_dwconv_rt_launch(x, x_ld, w, out, out_ld, B, C, T_in, T_out, k, dilation, pad, ek, epi, stream);
  }

  DwArgs a;
  memset(&a, 0, sizeof(a));
  a.x = static_cast<const uint8_t*>(x);
  a.w = w;
  a.out = out;
  a.x_ld = x_ld; a.out_ld = out_ld; a.B = B; a.C = C; a.T_in = T_in; a.T_out = T_out;
  a.k = k; a.stride = stride; a.dil = dilation; a.pad = pad;
  a.d = dw_misalign(pad);
  const int span = (k - 1) * dilation + 1;
  const int ksteps = (15 * stride + a.d + span + 31) / 32;
  if (ksteps > 6) {
    qasr_set_error("dwconv: tap span %d (k=%d, dilation=%d, stride=%d) exceeds the supported 6 x 32 window",
                   span, k, dilation, stride);
    return QASR_ERR_UNSUPPORTED;
  }
  // staging buffer: [halo | row (copy_bytes) | zeros up to the furthest word a tile can touch]
  a.halo = ((pad + a.d + 15) / 16) * 16;
  a.copy_bytes = ((T_in + 15) / 16) * 16;
  const int n_tiles = (T_out + kTile - 1) / kTile;
  const int reach = a.halo - pad - a.d + n_tiles * kTile * stride + 32 * ksteps + 16;  // bytes read at most
  int buf = a.halo + a.copy_bytes;
  if (buf < reach) buf = reach;
  a.buf_bytes = ((buf + 127) / 128) * 128;
  const size_t smem = 128 + (size_t)kWarps * kBufs * a.buf_bytes;
  if (smem > 200 * 1024) {
    qasr_set_error("dwconv: rows of %d frames need %zu bytes of staging shared memory (limit 200 KiB)", T_in, smem);
    return QASR_ERR_UNSUPPORTED;
  }
  // utterances per warp: amortise the A-fragment set-up while keeping the SM's 64 warp slots filled (the tile loop
  // is a serial LDS -> IMMA -> epilogue chain per warp; tools/dw_bench.py occupancy: 32 -> 64 warps/SM = -15 %)
  const long long per_sm = qasr_debug_get(10) > 0 ? qasr_debug_get(10) : 16LL * kWarps;   // knob 10: warps per SM wanted
  const long long want_items = per_sm * qasr_sm_count();
  int bchunk = (int)(((long long)B * C + want_items - 1) / want_items);
  if (bchunk < 1) bchunk = 1;
  if (bchunk > 16) bchunk = 16;
  if (bchunk > B) bchunk = B;
  a.bchunk = bchunk;
  a.nbg = (B + bchunk - 1) / bchunk;
  a.ep = make_epi_args(epi);
  a.stamp = qasr_next_stamp();
  cudaStream_t st = static_cast<cudaStream_t>(stream);

#define QASR_DW_CASE(KS_)                                      \
  if (ksteps == KS_) return stride == 1 ? launch<KS_, 1>(a, st) : launch<KS_, 2>(a, st);
  QASR_DW_CASE(1)
  QASR_DW_CASE(2)
  QASR_DW_CASE(3)
  QASR_DW_CASE(4)
  QASR_DW_CASE(5)
  QASR_DW_CASE(6)
#undef QASR_DW_CASE
  return QASR_ERR_UNSUPPORTED;
}

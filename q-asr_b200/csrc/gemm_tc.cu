// K2: 1-D convolution (k >= 1, groups == 1) on 8-bit codes as an int8 tensor-core GEMM /
// implicit GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma kind::i8 -> TMEM -> fused
// requantisation epilogue.  Restates the F.conv1d(x_int.double(), w_int.double(), bias) of
// QuantConv1d.int_conv (nemo/quantization/utils/quant_modules.py:301-305) for the pointwise
// 1x1 convs, the residual 1x1 convs, Jasper's dense convs and the decoder projection.
//
// GEMM view (per utterance b):   D[C_out, T_out] = sum_tap  W_tap[C_out, C_in] * X[C_in, T_out + shift(tap)]
//   A operand = packed weights [tap][C_out_pad][C_in_pad], K-major, 128B swizzle
//   B operand = activation codes in the reference's own (B, C, T) layout: time (N) contiguous,
//               i.e. MN-major, 128B swizzle; one TMA box = 128 channels x 128 frames.  The conv
//               taps are K-loop iterations that shift the TMA time coordinate (negative / past-the-
//               end coordinates are zero-filled by TMA = the conv's zero padding).
//   D         = int32 in TMEM, 128 lanes (output channels) x 128 columns (frames) per tile;
//               two tile sets are double buffered so the epilogue of tile i overlaps the MMAs of
//               tile i+1.  An optional second operand pair (the block's residual 1x1 conv,
//               jasper.py:667-672) accumulates into a second TMEM tile and both meet in the
//               res_act epilogue (jasper.py:681) without an int32 round trip through HBM.
// Weight-stationary schedule: a CTA owns one group of MT (1 or 2) 128-row M tiles for its whole life
// and walks over (utterance, time-tile) pairs.  When the packed weights of that group fit in shared
// memory (all 1x1 layers of QuartzNet / the decoder) they are loaded ONCE and stay resident; only the
// 16 KB activation boxes stream through the pipeline, and with MT = 2 every box feeds two M tiles.
// This halves / quarters the L2->SM traffic of a naive 128x128 tiling (measured 471 MB per 78 MB of
// algorithmic bytes on a 512-channel res_act layer before this change), which was what bound the
// kernel.  Larger K (weights not resident) falls back to streaming A with B.
// Warp roles (640 threads, 1 CTA / SM, persistent over tiles):
//   warps 0..15: epilogue (TMEM -> regs -> HBM), four per scheduler so TMEM-load / store latencies of one
//   warp hide behind the others;  warp 16: TMA producer;  warp 17: MMA issuer;  warp 18: TMEM allocator.
//   The control warps sit at the highest warp ids because the scheduler favours them (measured).
// Issue discipline (measured with tools/microbench/gemm_pipe.cu): a single thread that builds the shared-memory
// descriptors in ordinary registers inside a divergent `if (lane == 0)` pays an ELECT loop + R2UR per
// tcgen05.mma and issues one MMA per ~145 clk against the tensor pipe's 64 clk.  Here the WHOLE producer / MMA
// warp walks the loops with warp-uniform values (warp index broadcast by shuffle, descriptors advanced by adding
// to their low word), so descriptors live in uniform registers and the MMAs issue back to back (64.4 clk).
#include <cuda.h>

#include <type_traits>

#include "qasr_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace qasr_tc;

constexpr int BM = 128, BN = 128, BK = 128;
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 16;
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;
constexpr uint32_t kABytes = BM * BK, kBBytes = BK * BN;
constexpr uint32_t kSmemBudget = 227 * 1024;
constexpr uint32_t kSmemSlack = 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t kTmemCols = 512;

struct GemmSeg {
  int taps, cblocks, stride, dil, pad, c_out_pad;
  int slab0;       // index of this segment's first slab in the resident A region
  uint32_t idesc;
};

struct GemmArgs {
  void* out;
  int out_ld, B, C_out, T_out;
  int nseg, m_tiles, n_tiles;
  int MT;            // M tiles per CTA (1 or 2)
  int m_groups;      // m_tiles / MT
  int n_total;       // B * n_tiles
  int n_div, n_split; // tile walk: indices below n_split are the FULL time tiles (b = i / n_div, nt = i % n_div), the rest the
                     // utterances' partial last tiles (b = i - n_split, nt = n_div): the cheap tiles fill the last round
  int n_step;        // CTAs per M group (= stride of the n walk)
  int a_resident;    // weights of the CTA's M group stay in shared memory
  int a_slabs;       // 16 KB weight slabs per M tile (all segments)
  int stages;        // pipeline depth
  int bn;            // CTA-pair kernel: frames per tile across the pair (256, or 128 with two operands)
  uint32_t stage_bytes;
  GemmSeg seg[2];
  uint32_t b_lbo, b_sbo, a_sbo;  // smem descriptor strides (bytes)
  uint32_t b_layout;             // smem descriptor layout type of the activation operand: 2 = 128-byte swizzle, 4 = 64-byte swizzle
  uint32_t b_kstep;              // bytes between two K steps (32 channels) of the activation operand inside a stage
  int dbg_align, dbg_nonneg;     // bring-up knobs (qasr_debug_set 2 / 3)
  int dbg_skip;                  // qasr_debug_set 6: 1 = epilogue does nothing, 2 = no MMA issue, 4 = no stores
  unsigned long long* stamp;     // in-graph timing slot (qasr_profile_begin) or nullptr
  EpiArgs ep;
};

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
// EPI: bit 0 = all-integer fast epilogue, bits 1..3 (fast path only) = second conv operand,
// previous-res_act addend, second requantisation stage.  EPI == 0 is the literal (runtime-flag) path.
constexpr int kEpiFast = 1, kEpiDual = 2, kEpiPrev = 4, kEpiStage2 = 8, kEpiF32 = 16;
// device-written tables (QASR_F_DEVICE_TABLES, dynamic quantisation): the integer fast path runs where qasr_dyn_finalize
// proved it (per channel), the literal recipe elsewhere, decided per warp (its 32 rows never change)
constexpr int kEpiDyn = 32;
// QASR_F_RANGE_CHECK: integer fast path with the per-element |acc + bias| < 2^22 test (literal recipe beyond)
constexpr int kEpiRange = 64;
// QASR_F_CLAMP16: the float path clamps on packed 16-bit lanes (epilogue_f32_pack16)
constexpr int kEpiC16 = 128;
// QASR_F_STAGE2_16: second requantisation stage in 16-bit fixed point (epilogue_fast_s16x4)
constexpr int kEpiS16 = 256;

constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kAllocWarp = kEpiWarps + 2;

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}

// 32 accumulators of one output row (32 consecutive frames) -> codes -> one 32-byte (or two 16-byte) store.
// Shared by the 1-CTA and the CTA-pair kernels.
template <int EPI, bool F32>
__device__ __forceinline__ void finish32(const GemmArgs& g, const EpiConsts& kc, const ChanRegs& c0, const ChanRegs& c1,
                                         bool row_ok, int cc, int b, int tv, int len, bool wide_ok, const int (&a0)[32],
                                         const int (&a1)[32]) {
    const size_t row = (size_t)b * g.C_out + cc;
    if (EPI != 0 || g.ep.mode == QASR_EPI_REQUANT_I8) {
      uint8_t* optr = reinterpret_cast<uint8_t*>(g.out) + row * g.out_ld + tv;
      uint32_t pw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if ((EPI == 0 || (EPI & kEpiPrev)) && g.ep.prev != nullptr && row_ok) {
        const int8_t* pp = g.ep.prev + row * g.ep.prev_ld + tv;
        if (tv < g.T_out) { const uint4 v = *reinterpret_cast<const uint4*>(pp); pw[0] = v.x; pw[1] = v.y; pw[2] = v.z; pw[3] = v.w; }
        if (tv + 16 < g.T_out) { const uint4 v = *reinterpret_cast<const uint4*>(pp + 16); pw[4] = v.x; pw[5] = v.y; pw[6] = v.z; pw[7] = v.w; }
      }
      uint32_t packed[8];
#pragma unroll
      for (int w4 = 0; w4 < 8; ++w4) {
        if (F32 && (EPI & kEpiC16)) {
          packed[w4] = epilogue_f32_pack16(kc, a0[4 * w4], a0[4 * w4 + 1], a0[4 * w4 + 2], a0[4 * w4 + 3], c0);
        } else if ((EPI & kEpiFast) && (EPI & kEpiS16)) {
          packed[w4] = epilogue_fast_s16x4<(EPI & kEpiDual) != 0, (EPI & kEpiPrev) != 0, (EPI & kEpiRange) != 0>(
              g.ep, kc, &a0[4 * w4], c0, &a1[(EPI & kEpiDual) ? 4 * w4 : 0], c1, pw[w4]);
        } else {
        int q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * w4 + i;
          if (F32) {
            q[i] = (int)epilogue_f32_bits(kc, a0[j], c0);
          } else if (EPI & kEpiFast) {
            const int prev = (EPI & kEpiPrev) ? (int)(int8_t)(pw[w4] >> (8 * i)) : 0;
            q[i] = epilogue_fast<(EPI & kEpiDual) != 0, (EPI & kEpiPrev) != 0, (EPI & kEpiStage2) != 0, (EPI & kEpiRange) != 0>(
                g.ep, kc, a0[j], c0, (EPI & kEpiDual) ? a1[j] : 0, c1, prev);
          } else {
            const int prev = (int)(int8_t)((pw[w4] >> (8 * i)) & 0xff);
            q[i] = epilogue_exact(g.ep, kc, a0[j], c0, g.nseg > 1, g.nseg > 1 ? a1[j] : 0, c1, prev);
          }
        }
        packed[w4] = pack4(q[0], q[1], q[2], q[3]);
        }
      }
      if (tv + 32 > len) {  // length mask / row tail: zero the bytes at t >= len
#pragma unroll
        for (int w4 = 0; w4 < 8; ++w4) {
          const int nvalid = len - (tv + 4 * w4);
          packed[w4] &= nvalid >= 4 ? 0xffffffffu : (nvalid <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nvalid))));
        }
      }
      if (row_ok && !(g.dbg_skip & 4)) {
        if (wide_ok && tv + 16 < g.T_out) {
          // one 32-byte store = one full L2 sector per lane (the lanes of a warp are 32 different rows)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(optr), "r"(packed[0]), "r"(packed[1]),
                       "r"(packed[2]), "r"(packed[3]), "r"(packed[4]), "r"(packed[5]), "r"(packed[6]), "r"(packed[7]) : "memory");
        } else {
          if (tv < g.T_out) *reinterpret_cast<uint4*>(optr) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          if (tv + 16 < g.T_out) *reinterpret_cast<uint4*>(optr + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
    } else if (g.ep.mode == QASR_EPI_DEQUANT_F32) {
      float* orow = reinterpret_cast<float*>(g.out) + row * g.out_ld;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (row_ok && tv + j < g.T_out) orow[tv + j] = epilogue_dequant(g.ep, a0[j], c0);
    } else {
      int32_t* orow = reinterpret_cast<int32_t*>(g.out) + row * g.out_ld;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (row_ok && tv + j < g.T_out) orow[tv + j] = a0[j] + c0.bias;
    }
}


// Statistics pass of dynamic quantisation (QASR_EPI_STATS): this lane's channel, the counted frames of its 32 columns,
// folded into the lane's running extremes (flushed with one pair of atomics when the CTA is done).  One operand: int64
// extremes of acc + bias; two operands: order-preserving keys of f32(acc1+b1)*s1 + f32(acc0+b0)*s0 (quant_modules.py:108).
__device__ __forceinline__ void stats32(bool dual, const ChanRegs& c0, const ChanRegs& c1, int tv, int len, const int (&a0)[32],
                                        const int (&a1)[32], long long& lo, long long& hi) {
  if (!dual) {
    if (c0.bias > -(1 << 30) && c0.bias < (1 << 30)) {     // |acc| < 2^30 always: 32-bit arithmetic is exact
      int l = 2147483647, h = -2147483647 - 1;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (tv + j < len) {
          const int z = a0[j] + c0.bias;
          l = min(l, z);
          h = max(h, z);
        }
      }
      if (h >= l) {
        lo = (long long)l < lo ? (long long)l : lo;
        hi = (long long)h > hi ? (long long)h : hi;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (tv + j < len) {
          const long long z = qasr_acc_bias64(a0[j], c0.bias);
          lo = z < lo ? z : lo;
          hi = z > hi ? z : hi;
        }
      }
    }
  } else {
    unsigned l = (unsigned)lo, h = (unsigned)hi;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (tv + j < len) {
        const float y0 = __fmul_rn((float)qasr_acc_bias64(a0[j], c0.bias), c0.s);
        const float y1 = __fmul_rn((float)qasr_acc_bias64(a1[j], c1.bias), c1.s);
        const unsigned key = qasr_f32_key(__fadd_rn(y1, y0));     // identity + x
        l = min(l, key);
        h = max(h, key);
      }
    }
    lo = (long long)l;
    hi = (long long)h;
  }
}
__device__ __forceinline__ void stats_flush(bool dual, qasr_stats* st, long long lo, long long hi) {
  if (!dual) {
    if (hi >= lo) {
      atomicMin(reinterpret_cast<long long*>(&st->lo), lo);
      atomicMax(reinterpret_cast<long long*>(&st->hi), hi);
    }
  } else if ((unsigned)hi >= (unsigned)lo) {
    atomicMin(reinterpret_cast<unsigned*>(&st->lo), (unsigned)lo);
    atomicMax(reinterpret_cast<unsigned*>(&st->hi), (unsigned)hi);
  }
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_region = smem_base;                                       // resident weights (or empty)
  const uint32_t a_region_bytes = g.a_resident ? (uint32_t)(g.MT * g.a_slabs) * kABytes : 0u;
  const uint32_t stage_base = smem_base + a_region_bytes;
  const uint32_t bars = stage_base + (uint32_t)g.stages * g.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + 2 + s); };
  const uint32_t afull_bar = bars + 8u * (2 * kMaxStages + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int mgroup = blockIdx.x % g.m_groups;   // fixed for the CTA's life: weight-stationary
  const int n_first = blockIdx.x / g.m_groups;

  if (warp == kProducerWarp && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
    if (g.nseg > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    }
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
    mbar_init(afull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kAllocWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == kProducerWarp) {
    // ===================== TMA producer (whole warp walks, one elected lane issues) =====================
    const bool leader = elect_one();
    if (g.a_resident && leader) {  // the CTA's weights: once
      mbar_expect_tx(afull_bar, a_region_bytes);
      for (int ms = 0; ms < g.MT; ++ms) {
        for (int sg = 0; sg < g.nseg; ++sg) {
          const GemmSeg& s = g.seg[sg];
          const CUtensorMap* mA = sg == 0 ? &tmA0 : &tmA1;
          for (int tap = 0; tap < s.taps; ++tap)
            for (int cb = 0; cb < s.cblocks; ++cb)
              tma_load_2d(a_region + (uint32_t)(ms * g.a_slabs + s.slab0 + tap * s.cblocks + cb) * kABytes, mA,
                          cb * BK, tap * s.c_out_pad + (mgroup * g.MT + ms) * BM, afull_bar);
        }
      }
    }
    __syncwarp();
    pdl_wait();   // weights above are constants of the plan; activations below come from the previous kernel
    if (leader) qasr_stamp_begin(g.stamp);
    const int n_total = g.n_total, n_step = g.n_step, nseg = g.nseg, stages = g.stages;
    const int a_res = g.a_resident, MT = g.MT;
    const uint32_t stage_bytes = g.stage_bytes;
    int stage = 0;
    uint32_t phase = 0;
    // (time tile, utterance) of the walk carried incrementally: no division on the per-round path
    const int n_div = g.n_div, n_split = g.n_split;
    const int step_nt = n_step % n_div, step_b = n_step / n_div;
    int wnt = n_first % n_div, wb = n_first / n_div;
    const int seg_taps[2] = {g.seg[0].taps, g.seg[1].taps}, seg_cblocks[2] = {g.seg[0].cblocks, g.seg[1].cblocks};
    const int seg_tstep[2] = {BN * g.seg[0].stride, BN * g.seg[1].stride}, seg_dil[2] = {g.seg[0].dil, g.seg[1].dil};
    const int seg_pad[2] = {g.seg[0].pad, g.seg[1].pad};
    for (int nidx = n_first; nidx < n_total; nidx += n_step, wnt += step_nt, wb += step_b) {
      if (wnt >= n_div) { wnt -= n_div; ++wb; }
      const int nt = nidx >= n_split ? n_div : wnt, b = nidx >= n_split ? nidx - n_split : wb;
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        if (sg >= nseg) break;
        const GemmSeg& s = g.seg[sg];
        const CUtensorMap* mA = sg == 0 ? &tmA0 : &tmA1;
        const CUtensorMap* mB = sg == 0 ? &tmB0 : &tmB1;
        const int taps = seg_taps[sg], cblocks = seg_cblocks[sg];
        for (int tap = 0; tap < taps; ++tap) {
          const int tcoord = nt * seg_tstep[sg] + tap * seg_dil[sg] - seg_pad[sg];
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (leader) {
              uint32_t dst = stage_base + (uint32_t)stage * stage_bytes;
              mbar_expect_tx(full_bar(stage), stage_bytes);
              if (!a_res) {
                for (int ms = 0; ms < MT; ++ms, dst += kABytes)
                  tma_load_2d(dst, mA, cb * BK, tap * s.c_out_pad + (mgroup * MT + ms) * BM, full_bar(stage));
              }
              tma_load_3d(dst, mB, tcoord, cb * BK, b, full_bar(stage));
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (whole warp walks, one elected lane issues) =====================
    const bool leader = elect_one();
    const int n_total = g.n_total, n_step = g.n_step, nseg = g.nseg, stages = g.stages;
    const int a_res = g.a_resident, MT = g.MT, a_slabs = g.a_slabs, skip_mma = g.dbg_skip & 2;
    // descriptors: constant high words, low word = (smem address & 0x3FFFF) >> 4 advanced by plain adds
    const uint64_t a_hi = make_smem_desc(0, 0, g.a_sbo), b_hi = make_smem_desc(0, g.b_lbo, g.b_sbo);
    const uint32_t a_region_lo = (a_region & 0x3FFFFu) >> 4, stage_lo0 = (stage_base & 0x3FFFFu) >> 4;
    const uint32_t stage_step_lo = g.stage_bytes >> 4;
    const uint32_t b_in_stage_lo = a_res ? 0u : ((uint32_t)MT * kABytes) >> 4;
    constexpr uint32_t kSlabLo = kABytes >> 4, kKStepA = 32u >> 4, kKStepB = (32u * BN) >> 4;
    const uint32_t acc_per_stage = (uint32_t)(MT * nseg * BN);
    int stage = 0;
    uint32_t phase = 0, stage_lo = stage_lo0;
    int as = 0;
    uint32_t aphase = 0;
    if (a_res) {
      mbar_wait(afull_bar, 0u);
      tc_fence_after();
    }
    const int n_div = g.n_div, n_split = g.n_split, step_nt = n_step % n_div, T_out_m = g.T_out;
    int wnt = n_first % n_div;
    // per-segment constants live in registers: the kernel-parameter bank costs a dependent load per use
    const int seg_slabs[2] = {g.seg[0].taps * g.seg[0].cblocks, g.seg[1].taps * g.seg[1].cblocks};
    const uint32_t seg_idesc[2] = {g.seg[0].idesc & ~(0x3Fu << 17), g.seg[1].idesc & ~(0x3Fu << 17)};
    const uint32_t seg_a_lo[2] = {a_region_lo + (uint32_t)g.seg[0].slab0 * kSlabLo, a_region_lo + (uint32_t)g.seg[1].slab0 * kSlabLo};
    for (int nidx = n_first; nidx < n_total; nidx += n_step, wnt += step_nt) {
      if (wnt >= n_div) wnt -= n_div;
      const int nt_mma = nidx >= n_split ? n_div : wnt;
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_round = tmem_base + (uint32_t)as * acc_per_stage;
      // the last time tile of an utterance is partial: issue only the frames it has (N rounded up to 16; the
      // B descriptor walks the same 128-column box, the untouched accumulator columns are never stored)
      const int cols = min(BN, ((T_out_m - nt_mma * BN + 15) >> 4) << 4);
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        if (sg >= nseg) break;
        const int slabs = seg_slabs[sg];
        const uint32_t idesc = seg_idesc[sg] | ((uint32_t)(cols >> 3) << 17);
        uint32_t a_res_lo = seg_a_lo[sg];
        for (int it = 0; it < slabs; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int ms = 0; ms < 2; ++ms) {
              if (ms < MT) {
                const uint32_t a_lo = a_res ? a_res_lo + (uint32_t)(ms * a_slabs) * kSlabLo : stage_lo + (uint32_t)ms * kSlabLo;
                const uint32_t b_lo = stage_lo + b_in_stage_lo;
                const uint32_t d_tmem = d_round + (uint32_t)((ms * nseg + sg) * BN);
#pragma unroll
                for (int kk = 0; kk < BK / 32; ++kk) {
                  if (!skip_mma)
                    tc_mma_i8(d_tmem, a_hi | (uint64_t)(a_lo + kk * kKStepA), b_hi | (uint64_t)(b_lo + kk * kKStepB), idesc,
                              (it | kk) != 0 ? 1u : 0u);
                }
              }
            }
            tc_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
          }
          a_res_lo += kSlabLo;
          stage_lo += stage_step_lo;
          if (++stage == stages) { stage = 0; phase ^= 1u; stage_lo = stage_lo0; }
        }
      }
      if (leader) tc_commit(tfull_bar(as));  // accumulators of this round are complete
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int cgroup = warp >> 2;             // which 32 of the tile's 128 columns
    const EpiConsts kc = make_epi_consts(g.ep, g.nseg > 1);
    // per-channel parameters: the CTA's M group never changes
    ChanRegs c0[2], c1[2];
    bool row_ok[2];
    int cc[2];
#pragma unroll
    for (int ms = 0; ms < 2; ++ms) {
      const int c = (mgroup * g.MT + ms) * BM + quarter * 32 + lane;
      row_ok[ms] = ms < g.MT && c < g.C_out;
      cc[ms] = row_ok[ms] ? c : 0;
      c0[ms] = load_chan0(g.ep, cc[ms]);
      c1[ms] = g.nseg > 1 ? load_chan(g.ep.ch1, cc[ms]) : c0[ms];
    }
    // float epilogue per warp: all 32 channels of this warp's rows proved (the rows never change)
    const bool warp_f32[2] = {__all_sync(0xffffffffu, c0[0].r != 0.0f) != 0, __all_sync(0xffffffffu, c0[1].r != 0.0f) != 0};
    const bool warp_lit[2] = {__any_sync(0xffffffffu, c0[0].literal || c1[0].literal) != 0,
                              __any_sync(0xffffffffu, c0[1].literal || c1[1].literal) != 0};
    const bool stats_mode = EPI == 0 && g.ep.mode == QASR_EPI_STATS, stats_dual = g.nseg > 1;
    long long st_lo[2], st_hi[2];
#pragma unroll
    for (int ms = 0; ms < 2; ++ms) {
      st_lo[ms] = stats_dual ? 0xffffffffLL : 0x7fffffffffffffffLL;
      st_hi[ms] = stats_dual ? 0LL : -0x7fffffffffffffffLL - 1;
    }
    const uint32_t acc_per_stage = (uint32_t)(g.MT * g.nseg * BN);
    int as = 0;
    uint32_t aphase = 0;
    pdl_wait();   // out / prev / out_lens belong to the stream order
    const int MT = g.MT, nseg = g.nseg, T_out = g.T_out;
    const bool masked = g.ep.flags & QASR_F_MASK;
    const bool wide_ok = (reinterpret_cast<uintptr_t>(g.out) & 31) == 0 && (g.out_ld & 31) == 0;   // 32-byte aligned rows
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cgroup * 32);

    // 32 accumulators (one row, 32 consecutive frames) -> codes -> two 16-byte stores
    auto finish_t = [&](auto f32_tag, int ms, int b, int tv, int len, const int (&a0)[32], const int (&a1)[32]) {
      finish32<EPI, decltype(f32_tag)::value>(g, kc, c0[ms], c1[ms], row_ok[ms], cc[ms], b, tv, len, wide_ok, a0, a1);
    };
    // float or integer epilogue: a per-warp, per-M-tile constant (the warp's 32 rows never change)
    auto finish = [&](int ms, int b, int tv, int len, const int (&a0)[32], const int (&a1)[32]) {
      if (tv >= T_out) return;   // warp-uniform: this column group lies past the end of the utterance
      if (stats_mode) { stats32(stats_dual, c0[ms], c1[ms], tv, len, a0, a1, st_lo[ms], st_hi[ms]); return; }
      if ((EPI & kEpiDyn) && warp_lit[ms]) { finish32<0, false>(g, kc, c0[ms], c1[ms], row_ok[ms], cc[ms], b, tv, len, wide_ok, a0, a1); return; }
      if ((EPI & kEpiF32) || ((EPI & ~kEpiC16) == kEpiFast && warp_f32[ms])) finish_t(std::true_type{}, ms, b, tv, len, a0, a1);
      else finish_t(std::false_type{}, ms, b, tv, len, a0, a1);
    };

    const int n_total = g.n_total, n_step = g.n_step, n_div = g.n_div, n_split = g.n_split;
    const int step_nt = n_step % n_div, step_b = n_step / n_div;
    int wnt = n_first % n_div, wb = n_first / n_div;
    for (int nidx = n_first; nidx < n_total; nidx += n_step, wnt += step_nt, wb += step_b) {
      if (wnt >= n_div) { wnt -= n_div; ++wb; }
      const int nt = nidx >= n_split ? n_div : wnt, b = nidx >= n_split ? nidx - n_split : wb;
      const int len = masked ? min(g.ep.out_lens[b], T_out) : T_out;
      const int tv = nt * BN + cgroup * 32;   // first frame of this thread's 32 columns
      const uint32_t taddr = lane_base + (uint32_t)as * acc_per_stage;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (g.dbg_skip & 1) {
        // (bring-up: accumulators are dropped)
        tc_fence_before();
        mbar_arrive_warp(tempty_bar(as));
      } else if (!(EPI & kEpiDual) && nseg == 1 && MT == 2) {
        // both M tiles' accumulators leave TMEM first: the buffer goes back to the MMA warp before the math
        int a0[32], a1[32];
        tmem_ld32(taddr, a0);
        tmem_ld32(taddr + BN, a1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(tempty_bar(as));
        finish(0, b, tv, len, a0, a0);
        finish(1, b, tv, len, a1, a1);
      } else {
#pragma unroll
        for (int ms = 0; ms < 2; ++ms) {
          if (ms >= MT) break;
          int a0[32], a1[32];
          // the fast variants know at compile time whether there is a second operand (a runtime choice between the
          // two register arrays costs 32 moves per chunk)
          const bool two = EPI != 0 ? (EPI & kEpiDual) != 0 : nseg > 1;
          tmem_ld32(taddr + (uint32_t)(ms * nseg * BN), a0);
          if (two) tmem_ld32(taddr + (uint32_t)(ms * nseg * BN + BN), a1);
          tmem_ld_wait();
          if (ms == MT - 1) {
            tc_fence_before();
            mbar_arrive_warp(tempty_bar(as));
          }
          if (EPI != 0 && (EPI & kEpiDual)) finish(ms, b, tv, len, a0, a1);
          else if (EPI != 0) finish(ms, b, tv, len, a0, a0);
          else finish(ms, b, tv, len, a0, nseg > 1 ? a1 : a0);
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (stats_mode) {
#pragma unroll
      for (int ms = 0; ms < 2; ++ms)
        if (row_ok[ms]) stats_flush(stats_dual, reinterpret_cast<qasr_stats*>(g.out) + cc[ms], st_lo[ms], st_hi[ms]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) qasr_stamp_end(g.stamp);
  if (warp == kAllocWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// CTA-pair kernel (tcgen05 cta_group::2)
// ------------------------------------------------------------------------------------------
// Two CTAs of a cluster (the two SMs of a TPC) run ONE MMA stream of M = 256: each keeps the weights of its own
// 128-row M tile resident and streams its own HALF of every activation slab (bn / 2 of the tile's bn frames), the
// leader issues tcgen05.mma.cta_group::2 and both drain their own 128 accumulator rows.  Why: a 128x128x32 SS
// MMA reads 8 KB of shared memory per 64 clk = the SM's whole 128 B/clk, so in the 1-CTA kernel the TMA writes
// and the operand reads time-share the port (measured 602-640 clk per 16 KB slab, loads and MMAs adding up
// instead of overlapping); the pair reads 8 KB per 128 clk per SM and the bare pipeline runs at the tensor
// pipe's 512 clk per slab with TMA fully hidden (tools/microbench/gemm_pipe_2cta.cu, umma_2cta.cu).
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t leader_addr(uint32_t local) {   // same shared-memory offset in cluster rank 0
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(0u));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// relaxed: the hand-off orders TMEM reads (tcgen05.wait::ld + tcgen05 fence), not global memory; a .release at
// cluster scope would first drain the previous round's output stores (~1 us per round per warp, measured)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on the same barrier offset in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_region = smem_base;                                  // this CTA's M tile: a_slabs x 16 KB, resident
  const uint32_t a_region_bytes = (uint32_t)g.a_slabs * kABytes;
  const uint32_t stage_base = smem_base + a_region_bytes;               // stages x 16 KB: this CTA's half of a slab
  const uint32_t bars = stage_base + (uint32_t)g.stages * g.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };                 // leader's copy collects both CTAs' loads
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + s); };       // up to 4 accumulator buffers
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + 4 + s); };   // leader's copy, both CTAs arrive
  const uint32_t afull_bar = bars + 8u * (2 * kMaxStages + 8);
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 9);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int pair = blockIdx.x >> 1;
  const int mpair = pair % g.m_groups;          // which pair of M tiles: fixed for the pair's life
  const int n_first = pair / g.m_groups;
  const int mt = mpair * 2 + (int)rank;         // this CTA's M tile
  const int bn = g.bn;

  if (warp == kProducerWarp && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
    if (g.nseg > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    }
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * kEpiWarps); }
    mbar_init(afull_bar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kAllocWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_total = g.n_total, n_step = g.n_step, nseg = g.nseg, stages = g.stages, T_out = g.T_out;
  const int n_div = g.n_div, n_split = g.n_split;
  const int step_nt = n_step % n_div, step_b = n_step / n_div;
  const int nbuf = min(4, (int)kTmemCols / (nseg * bn));   // accumulator buffers that fit the 512 TMEM columns

  if (warp == kProducerWarp) {
    // ===================== TMA producer (both CTAs: own weights once, own half of every activation slab) =====================
    const bool leader = elect_one();
    if (leader) {
      const uint32_t lafull = leader_addr(afull_bar);
      mbar_expect_tx_cluster(lafull, a_region_bytes);
      for (int sg = 0; sg < nseg; ++sg) {
        const GemmSeg& s = g.seg[sg];
        const CUtensorMap* mA = sg == 0 ? &tmA0 : &tmA1;
        for (int tap = 0; tap < s.taps; ++tap)
          for (int cb = 0; cb < s.cblocks; ++cb)
            tma_load_2d_pair(a_region + (uint32_t)(s.slab0 + tap * s.cblocks + cb) * kABytes, mA, cb * BK,
                             tap * s.c_out_pad + mt * BM, lafull);
      }
    }
    __syncwarp();
    pdl_wait();   // weights above are constants of the plan; activations below come from the previous kernel
    if (leader) qasr_stamp_begin(g.stamp);
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage_bytes = g.stage_bytes;   // 16 KB (128-frame boxes) or 8 KB (64-frame boxes, bn = 128)
    int wnt = n_first % n_div, wb = n_first / n_div;
    const int seg_taps[2] = {g.seg[0].taps, g.seg[1].taps}, seg_cblocks[2] = {g.seg[0].cblocks, g.seg[1].cblocks};
    const int seg_dil[2] = {g.seg[0].dil, g.seg[1].dil}, seg_pad[2] = {g.seg[0].pad, g.seg[1].pad};
    for (int nidx = n_first; nidx < n_total; nidx += n_step, wnt += step_nt, wb += step_b) {
      if (wnt >= n_div) { wnt -= n_div; ++wb; }
      const int nt = nidx >= n_split ? n_div : wnt, b = nidx >= n_split ? nidx - n_split : wb;
      // frames of this tile, rounded up to 32: the pair splits them in halves of a multiple of 16 (TMA alignment)
      const int cols = min(bn, ((T_out - nt * bn + 31) >> 5) << 5);
      const int t_half = nt * bn + (int)rank * (cols >> 1);
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        if (sg >= nseg) break;
        const CUtensorMap* mB = sg == 0 ? &tmB0 : &tmB1;
        const int taps = seg_taps[sg], cblocks = seg_cblocks[sg];
        for (int tap = 0; tap < taps; ++tap) {
          const int tcoord = t_half + tap * seg_dil[sg] - seg_pad[sg];
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (leader) {
              const uint32_t lfull = leader_addr(full_bar(stage));
              mbar_expect_tx_cluster(lfull, stage_bytes);
              tma_load_3d_pair(stage_base + (uint32_t)stage * stage_bytes, mB, tcoord, cb * BK, b, lfull);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp && rank == 0) {
    // ===================== MMA issuer (leader CTA only) =====================
    const bool leader = elect_one();
    const int skip_mma = g.dbg_skip & 2;
    const uint64_t a_hi = make_smem_desc(0, 0, g.a_sbo);
    const uint64_t b_hi = (make_smem_desc(0, g.b_lbo, g.b_sbo) & ~((uint64_t)7 << 61)) | ((uint64_t)g.b_layout << 61);
    const uint32_t a_region_lo = (a_region & 0x3FFFFu) >> 4, stage_lo0 = (stage_base & 0x3FFFFu) >> 4;
    constexpr uint32_t kSlabLo = kABytes >> 4, kKStepA = 32u >> 4;
    const uint32_t kStageLo = g.stage_bytes >> 4, kKStepB = g.b_kstep >> 4;
    const uint32_t acc_per_stage = (uint32_t)(nseg * bn);
    int stage = 0;
    uint32_t phase = 0, stage_lo = stage_lo0;
    int as = 0;
    uint32_t aphase = 0;
    mbar_wait(afull_bar, 0u);   // both CTAs' weights have landed
    tc_fence_after();
    int wnt = n_first % n_div;
    const int seg_slabs[2] = {g.seg[0].taps * g.seg[0].cblocks, g.seg[1].taps * g.seg[1].cblocks};
    // instruction descriptor: the operand's own formats, N = this tile's frames, M = 256 across the pair
    const uint32_t fmt_mask = ~((0x3Fu << 17) | (0x1Fu << 24));
    const uint32_t seg_idesc[2] = {(g.seg[0].idesc & fmt_mask) | ((uint32_t)(256 >> 4) << 24),
                                   (g.seg[1].idesc & fmt_mask) | ((uint32_t)(256 >> 4) << 24)};
    const uint32_t seg_a_lo[2] = {a_region_lo + (uint32_t)g.seg[0].slab0 * kSlabLo, a_region_lo + (uint32_t)g.seg[1].slab0 * kSlabLo};
    for (int nidx = n_first; nidx < n_total; nidx += n_step, wnt += step_nt) {
      if (wnt >= n_div) wnt -= n_div;
      const int nt = nidx >= n_split ? n_div : wnt;
      const int cols = min(bn, ((T_out - nt * bn + 31) >> 5) << 5);
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_round = tmem_base + (uint32_t)as * acc_per_stage;
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        if (sg >= nseg) break;
        const int slabs = seg_slabs[sg];
        const uint32_t idesc = seg_idesc[sg] | ((uint32_t)(cols >> 3) << 17);
        const uint32_t d_tmem = d_round + (uint32_t)(sg * bn);
        uint32_t a_lo = seg_a_lo[sg];
        for (int it = 0; it < slabs; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < BK / 32; ++kk) {
              if (!skip_mma)
                tc_mma_i8_pair(d_tmem, a_hi | (uint64_t)(a_lo + kk * kKStepA), b_hi | (uint64_t)(stage_lo + kk * kKStepB), idesc,
                               (it | kk) != 0 ? 1u : 0u);
            }
            tc_commit_pair(empty_bar(stage));   // frees the slot in both CTAs once these MMAs have read it
          }
          a_lo += kSlabLo;
          stage_lo += kStageLo;
          if (++stage == stages) { stage = 0; phase ^= 1u; stage_lo = stage_lo0; }
        }
      }
      if (leader) tc_commit_pair(tfull_bar(as));   // both CTAs' accumulator halves are complete
      if (++as == nbuf) { as = 0; aphase ^= 1u; }
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (each CTA drains its own 128 rows x bn columns) =====================
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int cgroup = warp >> 2;             // which quarter of the tile's columns
    const EpiConsts kc = make_epi_consts(g.ep, nseg > 1);
    const int c = mt * BM + quarter * 32 + lane;
    const bool row_ok = c < g.C_out;
    const int cc = row_ok ? c : 0;
    const ChanRegs c0 = load_chan0(g.ep, cc);
    const ChanRegs c1 = nseg > 1 ? load_chan(g.ep.ch1, cc) : c0;
    const bool warp_f32 = __all_sync(0xffffffffu, c0.r != 0.0f) != 0;
    const bool warp_lit = __any_sync(0xffffffffu, c0.literal || c1.literal) != 0;
    const bool stats_mode = EPI == 0 && g.ep.mode == QASR_EPI_STATS, stats_dual = nseg > 1;
    long long st_lo = stats_dual ? 0xffffffffLL : 0x7fffffffffffffffLL, st_hi = stats_dual ? 0LL : -0x7fffffffffffffffLL - 1;
    const uint32_t acc_per_stage = (uint32_t)(nseg * bn);
    const uint32_t ltempty0 = leader_addr(tempty_bar(0));   // the barriers are 8 bytes apart in the leader too
    int as = 0;
    uint32_t aphase = 0;
    pdl_wait();   // out / prev / out_lens belong to the stream order
    const bool masked = g.ep.flags & QASR_F_MASK;
    const bool wide_ok = (reinterpret_cast<uintptr_t>(g.out) & 31) == 0 && (g.out_ld & 31) == 0;
    const int chunks = bn >> 7;               // 32-column chunks per warp: 2 for bn = 256, 1 for bn = 128
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cgroup * chunks * 32);
    auto finish = [&](int b, int tv, int len, const int (&a0)[32], const int (&a1)[32]) {
      if (tv >= T_out) return;   // warp-uniform: past the end of the utterance
      if (stats_mode) { stats32(stats_dual, c0, c1, tv, len, a0, a1, st_lo, st_hi); return; }
      if ((EPI & kEpiDyn) && warp_lit) { finish32<0, false>(g, kc, c0, c1, row_ok, cc, b, tv, len, wide_ok, a0, a1); return; }
      if ((EPI & kEpiF32) || ((EPI & ~kEpiC16) == kEpiFast && warp_f32)) finish32<EPI, true>(g, kc, c0, c1, row_ok, cc, b, tv, len, wide_ok, a0, a1);
      else finish32<EPI, false>(g, kc, c0, c1, row_ok, cc, b, tv, len, wide_ok, a0, a1);
    };
    int wnt = n_first % n_div, wb = n_first / n_div;
    for (int nidx = n_first; nidx < n_total; nidx += n_step, wnt += step_nt, wb += step_b) {
      if (wnt >= n_div) { wnt -= n_div; ++wb; }
      const int nt = nidx >= n_split ? n_div : wnt, b = nidx >= n_split ? nidx - n_split : wb;
      const int len = masked ? min(g.ep.out_lens[b], T_out) : T_out;
      const int tv = nt * bn + cgroup * chunks * 32;   // first frame of this warp's columns
      const uint32_t taddr = lane_base + (uint32_t)as * acc_per_stage;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (g.dbg_skip & 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ltempty0 + 8u * as);
      } else if (nseg == 1 && chunks == 2) {
        int a0[32], a1[32];
        tmem_ld32(taddr, a0);
        tmem_ld32(taddr + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ltempty0 + 8u * as);   // the buffer goes back to the leader before the math
        finish(b, tv, len, a0, a0);
        finish(b, tv + 32, len, a1, a1);
      } else {
        for (int h = 0; h < chunks; ++h) {
          int a0[32], a1[32];
          const bool two = EPI != 0 ? (EPI & kEpiDual) != 0 : nseg > 1;
          tmem_ld32(taddr + (uint32_t)(h * 32), a0);
          if (two) tmem_ld32(taddr + (uint32_t)(bn + h * 32), a1);
          tmem_ld_wait();
          if (h == chunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(ltempty0 + 8u * as);
          }
          if (EPI != 0 && (EPI & kEpiDual)) finish(b, tv + h * 32, len, a0, a1);
          else if (EPI != 0) finish(b, tv + h * 32, len, a0, a0);
          else finish(b, tv + h * 32, len, a0, nseg > 1 ? a1 : a0);
        }
      }
      if (++as == nbuf) { as = 0; aphase ^= 1u; }
    }
    if (stats_mode && row_ok) stats_flush(stats_dual, reinterpret_cast<qasr_stats*>(g.out) + cc, st_lo, st_hi);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's shared memory and TMEM stay alive until the leader's last MMA has retired
  if (threadIdx.x == 0) qasr_stamp_end(g.stamp);
  if (warp == kAllocWarp) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// Order of the (utterance, time tile) walk: all FULL tiles first, the utterances' partial last tiles after them.  With 64
// utterances x 7 tiles of 128 frames (800 frames: 6 full + one of 32) the 448 tiles leave 4 for a last round on 148 / 74 /
// 37 CTAs per M group; in utterance-major order those were full tiles (4 CTAs x 512 frames against 384 for the rest), now
// they are 32-frame ones.  Bring-up knob 18 = 0 keeps the utterance-major order.
static void set_tile_walk(GemmArgs& g, int bn) {
  const int n_full = g.T_out / bn;
  if (n_full >= 1 && n_full < g.n_tiles && qasr_debug_get(18) != 0) {
    g.n_div = n_full;
    g.n_split = n_full * g.B;
  } else {
    g.n_div = g.n_tiles;
    g.n_split = g.n_total;
  }
}

// ------------------------------------------------------------------------------------------
// weight packing: [C_out][C_in][k] -> [k][C_out_pad][C_in_pad], zero filled
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const int8_t* __restrict__ w, int8_t* __restrict__ wp, int C_out,
                                    int C_in, int k, int C_out_pad, int C_in_pad) {
  const long long total = (long long)k * C_out_pad * C_in_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C_in_pad);
    const long long r = i / C_in_pad;
    const int co = (int)(r % C_out_pad);
    const int tap = (int)(r / C_out_pad);
    int8_t v = 0;
    if (ci < C_in && co < C_out) v = w[((size_t)co * C_in + ci) * k + tap];
    wp[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------
int make_map_weights(CUtensorMap* map, const int8_t* wp, int rows, int c_in_pad) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { qasr_set_error("cuTensorMapEncodeTiled entry point not available"); return QASR_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)c_in_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)c_in_pad};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(wp), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { qasr_set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return QASR_ERR_CUDA; }
  return QASR_OK;
}

// frames = 128: the [128 frames][128 channels] box of the 128-byte swizzle; frames = 64: the half box a CTA of a pair
// streams when the pair's tile is 128 frames wide (64-byte rows, 64-byte swizzle)
int make_map_acts(CUtensorMap* map, const void* x, int T, int C, int B, int ld, int frames = BN) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { qasr_set_error("cuTensorMapEncodeTiled entry point not available"); return QASR_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * (cuuint64_t)C};
  cuuint32_t box[3] = {(cuuint32_t)frames, (cuuint32_t)BK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, frames == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { qasr_set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return QASR_ERR_CUDA; }
  return QASR_OK;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }


}  // namespace

extern "C" int64_t qasr_packed_weight_bytes(int32_t C_out, int32_t C_in, int32_t k) {
  if (C_out <= 0 || C_in <= 0 || k <= 0) return -1;
  return (int64_t)k * round_up(C_out, BM) * round_up(C_in, BK);
}

extern "C" int qasr_pack_conv_weights(const int8_t* w, int8_t* wp, int32_t C_out, int32_t C_in, int32_t k,
                                      void* stream) {
  QASR_CHECK_ARG(w && wp && C_out > 0 && C_in > 0 && k > 0, "pack_conv_weights: bad arguments");
  const long long total = qasr_packed_weight_bytes(C_out, C_in, k);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, wp, C_out, C_in, k, round_up(C_out, BM),
                                                               round_up(C_in, BK));
  return qasr_check_launch("pack_weights_kernel");
}

extern "C" int qasr_conv1d_gemm_i8(const qasr_gemm_operand* op0, const qasr_gemm_operand* op1, void* out,
                                   int32_t out_ld, int32_t B, int32_t C_out, int32_t T_out,
                                   const qasr_epilogue* epi, void* stream) {
  QASR_CHECK_ARG(op0 && out && epi, "conv1d_gemm: null pointer");
  QASR_CHECK_ARG(B > 0 && C_out > 0 && T_out > 0, "conv1d_gemm: bad shape");
  QASR_CHECK_ARG(qasr_aligned16(out) && out_ld >= round_up(T_out, 16), "conv1d_gemm: out pitch/alignment");
  QASR_CHECK_ARG(epi->mode != QASR_EPI_REQUANT_I8 || out_ld % 16 == 0, "conv1d_gemm: out pitch % 16");
  int rc = qasr_validate_epilogue(epi, true);
  if (rc) return rc;
  QASR_CHECK_ARG((op1 != nullptr) == (epi->ch1 != nullptr), "conv1d_gemm: second operand and epi->ch1 must come together");
  QASR_CHECK_ARG(!op1 || epi->mode == QASR_EPI_REQUANT_I8 || epi->mode == QASR_EPI_STATS,
                 "conv1d_gemm: second operand needs REQUANT or STATS mode");
  qasr_set_plain_launch((epi->flags & QASR_F_DEVICE_TABLES) != 0);

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  const qasr_gemm_operand* ops[2] = {op0, op1};
  g.nseg = op1 ? 2 : 1;
  for (int i = 0; i < g.nseg; ++i) {
    const qasr_gemm_operand* o = ops[i];
    QASR_CHECK_ARG(o->x && o->wp && o->C_in > 0 && o->T_in > 0 && o->k > 0, "conv1d_gemm: operand %d bad", i);
    if (o->stride != 1) {
      qasr_set_error("conv1d_gemm: stride %d not supported by the (B,C,T) tensor-core path", o->stride);
      return QASR_ERR_UNSUPPORTED;
    }
    QASR_CHECK_ARG(o->dilation >= 1 && o->pad >= 0, "conv1d_gemm: bad dilation/pad");
    // TMA needs a 16-byte aligned box start along the contiguous (time) axis (measured on B200: an
    // unaligned inner coordinate raises an illegal-instruction fault), so in the (B,C,T) layout only
    // taps whose shift is a multiple of 16 frames can be fetched: in practice the 1x1 convs.
    for (int tap = 0; tap < o->k; ++tap) {
      if ((tap * o->dilation - o->pad) % 16 != 0) {
        qasr_set_error("conv1d_gemm: k=%d dilation=%d pad=%d needs unaligned time shifts; the (B,C,T) "
                       "tensor-core path handles 1x1 convolutions only", o->k, o->dilation, o->pad);
        return QASR_ERR_UNSUPPORTED;
      }
    }
    QASR_CHECK_ARG((o->T_in + 2 * o->pad - o->dilation * (o->k - 1) - 1) / o->stride + 1 == T_out,
                   "conv1d_gemm: operand %d: T_out %d inconsistent with T_in %d", i, T_out, o->T_in);
    QASR_CHECK_ARG(o->x_ld % 16 == 0 && o->x_ld >= o->T_in && qasr_aligned16(o->x) && qasr_aligned16(o->wp),
                   "conv1d_gemm: operand %d pitch/alignment", i);
    GemmSeg& s = g.seg[i];
    s.taps = o->k;
    s.cblocks = (o->C_in + BK - 1) / BK;
    s.stride = o->stride;
    s.dil = o->dilation;
    s.pad = o->pad;
    s.c_out_pad = round_up(C_out, BM);
    s.slab0 = g.a_slabs;
    g.a_slabs += s.taps * s.cblocks;
    const uint32_t b_signed = (epi->flags & (i == 0 ? QASR_F_B_SIGNED : QASR_F_B1_SIGNED)) ? 1u : 0u;
    // instruction descriptor: D=S32, A=signed 8-bit K-major, B=(un)signed 8-bit MN-major, N=128, M=128
    s.idesc = (2u << 4) | (1u << 7) | (b_signed << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
              ((uint32_t)(BM >> 4) << 24);
    rc = make_map_weights(&maps[2 * i], o->wp, o->k * s.c_out_pad, round_up(o->C_in, BK));
    if (rc) return rc;
    rc = make_map_acts(&maps[2 * i + 1], o->x, o->T_in, o->C_in, B, o->x_ld);
    if (rc) return rc;
  }
  if (g.nseg == 1) { maps[2] = maps[0]; maps[3] = maps[1]; }
  g.out = out;
  g.out_ld = out_ld;
  g.B = B;
  g.C_out = C_out;
  g.T_out = T_out;
  g.m_tiles = (C_out + BM - 1) / BM;
  g.n_tiles = (T_out + BN - 1) / BN;
  g.n_total = g.n_tiles * B;
  set_tile_walk(g, BN);
  // M tiles per CTA: two when both accumulator sets still double-buffer in TMEM (512 columns)
  g.MT = (g.nseg == 1 && g.m_tiles % 2 == 0) ? 2 : 1;
  // less than one wave of (M tile, time tile) pairs (BASELINE config 1: batch 8 x 250 frames): one M tile per CTA doubles
  // the CTAs at work and halves the weights each loads before its only tile (256 x 256: 3.8 -> 3.0 us)
  const bool sub_wave = (long long)g.n_total * g.m_tiles <= qasr_sm_count();
  if (sub_wave) g.MT = 1;
  const uint32_t budget = kSmemBudget - kSmemSlack;
  auto a_bytes = [&](int mt) { return (uint32_t)(mt * g.a_slabs) * kABytes; };
  if (g.MT == 2 && a_bytes(2) + 4 * kBBytes > budget && a_bytes(1) + 4 * kBBytes <= budget) g.MT = 1;
  g.a_resident = a_bytes(g.MT) + 4 * kBBytes <= budget ? 1 : 0;
  if (qasr_debug_get(4) == 1) g.a_resident = 0;         // bring-up knob: force the streaming schedule
  if (qasr_debug_get(5) >= 1 && qasr_debug_get(5) <= 2 && g.m_tiles % (int)qasr_debug_get(5) == 0 && g.nseg * (int)qasr_debug_get(5) <= 2)
    g.MT = (int)qasr_debug_get(5);
  g.m_groups = g.m_tiles / g.MT;
  g.stage_bytes = g.a_resident ? kBBytes : (uint32_t)g.MT * kABytes + kBBytes;
  const uint32_t avail = budget - (g.a_resident ? a_bytes(g.MT) : 0u);
  g.stages = (int)(avail / g.stage_bytes);
  if (g.stages > kMaxStages) g.stages = kMaxStages;
  QASR_CHECK_ARG(g.stages >= 2, "conv1d_gemm: internal: pipeline does not fit shared memory");
  const uint32_t smem_bytes = (g.a_resident ? a_bytes(g.MT) : 0u) + (uint32_t)g.stages * g.stage_bytes + kSmemSlack;
  g.a_sbo = 1024;
  g.b_sbo = qasr_debug_get(1) >= 0 ? (uint32_t)qasr_debug_get(1) : 1024;
  g.b_lbo = qasr_debug_get(0) >= 0 ? (uint32_t)qasr_debug_get(0) : (uint32_t)(BK * 128);
  g.b_layout = 2;
  g.b_kstep = 32u * BN;
  g.ep = make_epi_args(epi);
  g.dbg_align = qasr_debug_get(2) > 0 ? (int)qasr_debug_get(2) : 0;
  g.dbg_nonneg = qasr_debug_get(3) > 0 ? 1 : 0;
  g.dbg_skip = qasr_debug_get(6) > 0 ? (int)qasr_debug_get(6) : 0;
  g.stamp = qasr_next_stamp();

  const bool fast = !(epi->flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64)) &&
                    epi->mode == QASR_EPI_REQUANT_I8;
  int variant = 0;
  if (fast)
    variant = kEpiFast | (op1 ? kEpiDual : 0) | (epi->prev ? kEpiPrev : 0) | ((epi->flags & QASR_F_STAGE2) ? kEpiStage2 : 0);
  if (fast && (epi->flags & QASR_F_RANGE_CHECK)) variant |= kEpiRange;
  if (fast && (epi->flags & QASR_F_FAST_F32)) variant = kEpiFast | kEpiF32;
  if ((variant == kEpiFast || variant == (kEpiFast | kEpiF32)) && (epi->flags & QASR_F_CLAMP16)) variant |= kEpiC16;
  if ((variant & ~kEpiRange) == (kEpiFast | kEpiDual | kEpiStage2) && (epi->flags & QASR_F_STAGE2_16)) variant |= kEpiS16;
  // dynamic quantisation: the tables come from the device, the literal flags stay set for the channels whose fast path
  // qasr_dyn_finalize could not prove; everything else runs the integer fast path
  if ((epi->flags & QASR_F_DEVICE_TABLES) && epi->mode == QASR_EPI_REQUANT_I8 && !epi->prev && !(epi->flags & QASR_F_STAGE2))
    variant = kEpiFast | kEpiDyn | (op1 ? kEpiDual : 0);
  // CTA-pair kernel (tcgen05 cta_group::2): an even number of M tiles keep their weights resident one tile per CTA;
  // halves the shared-memory operand reads per MMA so TMA writes and MMA reads stop time-sharing the port.
  // Bit-exact on every GEMM test.  Its loads + MMAs run faster than the 1-CTA kernel's (11.4 us against 13.2 us,
  // 512x512 at batch 64 x 800 frames), but both kernels end up bound by the 16 epilogue warps' issue slots and the
  // pair pays ~0.3 us more hand-off per round, so it is chosen where it measured faster (tools/gemm_bench.py, float
  // epilogue): single-operand layers with K >= 512 and 4 M tiles (15.6 us against 16.9 us), or K >= 1024 (41.6 us
  // against 56.8 us, where one CTA cannot keep two M tiles resident); 256-channel and two-operand layers stay on the
  // 1-CTA kernel (8.5 / 12.6 / 36.1 us against 9.7 / 14.1 / 35.6 us).  Bring-up knob 12 forces either (0 / 1).
  const bool pair_auto = !sub_wave && g.nseg == 1 && g.m_tiles >= 4 && (g.a_slabs >= 8 || (g.a_slabs >= 4 && g.m_tiles == 4));
  bool pair = (qasr_debug_get(12) == 1 || (qasr_debug_get(12) != 0 && pair_auto)) && g.m_tiles % 2 == 0 &&
              a_bytes(1) + 4 * kBBytes <= budget && qasr_sm_count() >= 2;
  for (int i = 0; i < g.nseg; ++i) pair = pair && g.seg[i].stride == 1;
  if (pair) {
    g.MT = 1;
    g.a_resident = 1;
    g.bn = g.nseg == 1 ? 256 : 128;   // two buffers of (operands x bn) accumulator columns fill the 512 TMEM columns
    if (qasr_debug_get(13) == 128 || qasr_debug_get(13) == 256) g.bn = (int)qasr_debug_get(13);   // bring-up: 4 buffers / 1 buffer
    g.m_groups = g.m_tiles / 2;
    g.n_tiles = (T_out + g.bn - 1) / g.bn;
    g.n_total = g.n_tiles * B;
    set_tile_walk(g, g.bn);
    g.stage_bytes = kBBytes;
    // 128-frame tiles: each CTA streams 64 frames of every slab as a [64 frames][128 channels] box in the 64-byte
    // swizzle (8 KB per stage, nothing fetched twice).  Bring-up knob 14 = 0 keeps the full 128-frame boxes.
    if (g.bn == 128 && qasr_debug_get(14) != 0) {
      for (int i = 0; i < g.nseg; ++i) {
        rc = make_map_acts(&maps[2 * i + 1], ops[i]->x, ops[i]->T_in, ops[i]->C_in, B, ops[i]->x_ld, 64);
        if (rc) return rc;
      }
      if (g.nseg == 1) maps[3] = maps[1];
      g.stage_bytes = kBBytes / 2;
      g.b_layout = 4;
      g.b_kstep = 32u * 64u;
      g.b_sbo = qasr_debug_get(1) >= 0 ? (uint32_t)qasr_debug_get(1) : 512;
      g.b_lbo = qasr_debug_get(0) >= 0 ? (uint32_t)qasr_debug_get(0) : (uint32_t)(BK * 64);
    }
    g.stages = (int)((budget - a_bytes(1)) / g.stage_bytes);
    if (g.stages > kMaxStages) g.stages = kMaxStages;
    const uint32_t pair_smem = a_bytes(1) + (uint32_t)g.stages * g.stage_bytes + kSmemSlack;
    int pairs_per_group = (qasr_sm_count() / 2) / g.m_groups;
    if (pairs_per_group < 1) pairs_per_group = 1;
    if (pairs_per_group > g.n_total) pairs_per_group = g.n_total;
    g.n_step = pairs_per_group;
    const int pair_grid = 2 * pairs_per_group * g.m_groups;
    cudaStream_t pst = (cudaStream_t)stream;
    static bool pair_attr_set[512][64] = {};
    int pdev = 0;
    QASR_CHECK_CUDA(cudaGetDevice(&pdev));
    const bool pneed = pdev >= 0 && pdev < 64 && !pair_attr_set[variant][pdev];
#define QASR_PAIR_CASE(V_)                                                                                          \
  case V_:                                                                                                          \
    if (pneed)                                                                                                      \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(gemm_i8_pair_kernel<V_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)kSmemBudget));                                                      \
    QASR_CHECK_CUDA(qasr_launch_pdl_pair(gemm_i8_pair_kernel<V_>, dim3(pair_grid), dim3(kGemmThreads), pair_smem, pst, \
                                         maps[0], maps[1], maps[2], maps[3], g));                                   \
    break;
    switch (variant) {
      QASR_PAIR_CASE(0)
      QASR_PAIR_CASE(1)
      QASR_PAIR_CASE(3)
      QASR_PAIR_CASE(5)
      QASR_PAIR_CASE(7)
      QASR_PAIR_CASE(9)
      QASR_PAIR_CASE(11)
      QASR_PAIR_CASE(13)
      QASR_PAIR_CASE(15)
      QASR_PAIR_CASE(17)
      QASR_PAIR_CASE(33)
      QASR_PAIR_CASE(35)
      QASR_PAIR_CASE(65)
      QASR_PAIR_CASE(67)
      QASR_PAIR_CASE(69)
      QASR_PAIR_CASE(71)
      QASR_PAIR_CASE(73)
      QASR_PAIR_CASE(75)
      QASR_PAIR_CASE(77)
      QASR_PAIR_CASE(79)
      QASR_PAIR_CASE(129)
      QASR_PAIR_CASE(145)
      QASR_PAIR_CASE(267)
      QASR_PAIR_CASE(331)
      default:
        qasr_set_error("conv1d_gemm: internal: no pair kernel variant %d", variant);
        return QASR_ERR_INVALID;
    }
#undef QASR_PAIR_CASE
    if (pneed) pair_attr_set[variant][pdev] = true;
    return qasr_check_launch("gemm_i8_pair_kernel");
  }
  int per_group = qasr_sm_count() / g.m_groups;
  if (per_group < 1) per_group = 1;
  if (per_group > g.n_total) per_group = g.n_total;
  g.n_step = per_group;
  const int grid = per_group * g.m_groups;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[512][64] = {};
  int dev = 0;
  QASR_CHECK_CUDA(cudaGetDevice(&dev));
  const bool need_attr = dev >= 0 && dev < 64 && !attr_set[variant][dev];
#define QASR_GEMM_CASE(V_)                                                                                     \
  case V_:                                                                                                     \
    if (need_attr)                                                                                             \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<V_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)kSmemBudget));                                                 \
    QASR_CHECK_CUDA(qasr_launch_pdl(gemm_i8_kernel<V_>, dim3(grid), dim3(kGemmThreads), smem_bytes, st, maps[0],  \
                                    maps[1], maps[2], maps[3], g));                                            \
    break;
  switch (variant) {
    QASR_GEMM_CASE(0)
    QASR_GEMM_CASE(1)
    QASR_GEMM_CASE(3)
    QASR_GEMM_CASE(5)
    QASR_GEMM_CASE(7)
    QASR_GEMM_CASE(9)
    QASR_GEMM_CASE(11)
    QASR_GEMM_CASE(13)
    QASR_GEMM_CASE(15)
    QASR_GEMM_CASE(17)
    QASR_GEMM_CASE(33)
    QASR_GEMM_CASE(35)
    QASR_GEMM_CASE(65)
    QASR_GEMM_CASE(67)
    QASR_GEMM_CASE(69)
    QASR_GEMM_CASE(71)
    QASR_GEMM_CASE(73)
    QASR_GEMM_CASE(75)
    QASR_GEMM_CASE(77)
    QASR_GEMM_CASE(79)
    QASR_GEMM_CASE(129)
    QASR_GEMM_CASE(145)
    QASR_GEMM_CASE(267)
    QASR_GEMM_CASE(331)
    default:
      qasr_set_error("conv1d_gemm: internal: no kernel variant %d", variant);
      return QASR_ERR_INVALID;
  }
#undef QASR_GEMM_CASE
  if (need_attr) attr_set[variant][dev] = true;
  return qasr_check_launch("gemm_i8_kernel");
}

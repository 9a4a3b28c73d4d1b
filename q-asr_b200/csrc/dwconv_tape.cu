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
// Work item = (channel, chunk of `nb` utterances, at most 16 units of 128 tape rows); a CTA walks a contiguous range of
// items, up to four tapes in flight.
// Warp roles (608 threads, 1 CTA / SM): warps 0..15 epilogue (quarter = warp & 3 is the TMEM lane quarter, warp >> 2
// the group that takes every fourth unit), warp 16 producer, warp 17 MMA issuer, warp 18 TMEM allocator.
// Input delivery, measured per item of 22 rows x 800 bytes: one bulk copy per row 1.0-1.3 us (~47 clk per copy on the
// SM's copy engine), LDG / STS pairs 3.3 us (a full L2 latency per batch of 8), 16-byte cp.async 1.8-2.2 us; hence ONE
// TMA tensor copy per item through a 4-D view of the tensor.
// An item is ONE accumulator round (256 TMEM columns, two rounds in flight): every warp waits once per item for the
// accumulators and hands the slot back once.  With a round per 4 units (the first version) every one of the 16 epilogue
// warps walked the chain wait -> tcgen05.ld -> arrive ~31 times per CTA at ~400 clk each whatever the number of slots:
// 19.8 us on 512 channels x B64 x T800 (k = 33), 16.9 us with a round per item.  tcgen05.ld.32x32b.x16 itself drains
// 426 B / clk / SM (tools/ubench/ldtm_rate.cu), far from binding.  Every epilogue warp keeps the tcgen05.ld of its next unit
// in flight while it requantises and stores the current one (two register buffers), across item boundaries.
#include <cuda.h>

#include "qasr_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace qasr_tc;

constexpr int kEpiWarps = 16;
constexpr int kProdWarps = 1;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + kProdWarps, kAllocWarp = kMmaWarp + 1;
constexpr int kThreads = 32 * (kEpiWarps + kProdWarps + 2);
constexpr int kMaxBufs = 4;
constexpr int kSlots = 2;                 // accumulator rounds in flight: 2 x 256 columns = the 512 TMEM columns
constexpr int kRound = 16;                // units per round = per work item (the host sizes items to <= 16 units): ONE
                                          // full / empty handshake per item and warp instead of one per 4 units
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
  int inner;         // TMA view of a row: [T / inner][inner] (inner divides G and copy_bytes)
  uint32_t tape_bytes;   // one tape buffer
  float inv_rpu;
  uint32_t idesc;
  int dbg_skip;      // qasr_debug_set 6: 1 = epilogue drains TMEM only, 2 = no MMA issue, 4 = no input copies
  unsigned long long* dbg_trace;   // qasr_debug_set 7: device buffer, role timeline of CTA 0 (4 x 256 slots)
  unsigned long long* stamp;
  EpiArgs ep;
};

// The producer runs at most kMaxBufs items ahead of the MMA issuer, which runs at most kSlots rounds (>= 1 round per item)
// ahead of the epilogue: 16 slots never wrap onto a record that is still to be read.
constexpr int kChanRing = 16;
struct alignas(16) ChanSlot {
  int4 rec;
  float r;
  int pad[3];
};

#define TAPE_TRACE(role, n) do { if (a.dbg_trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (n) < 256) { a.dbg_trace[(role) * 256 + (n)] = qasr_globaltimer(); } } while (0)

__device__ __forceinline__ void mbar_poll(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();   // a wedged pipeline traps instead of hanging the GPU
  }
}
// suspending wait without a time hint: the warp sleeps in hardware (no issue slots taken from the computing warps of its
// scheduler, which a polling loop of 16 waiting warps does) and is woken by the phase change or a short system time-out
__device__ __forceinline__ void mbar_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
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

// 16 accumulators of one lane (16 consecutive frames of one channel) -> 16 codes.  MODE: 0 = float fast path,
// 1 = integer fast path, 2 = literal recipe (one channel per work item: the choice is warp-uniform)
template <int MODE>
__device__ __forceinline__ void tape_codes(const EpiArgs& ep, const EpiConsts& kc, const ChanRegs& c0, const int (&acc)[16],
                                           uint32_t (&w)[4]) {
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    int q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = acc[4 * q4 + j];
      q[j] = MODE == 0 ? (int)epilogue_f32_bits(kc, v, c0)
                       : (MODE == 1 ? epilogue_fast<false, false, false>(ep, kc, v, c0, 0, c0, 0)
                                    : epilogue_exact(ep, kc, v, c0, false, 0, c0, 0));
    }
    w[q4] = pack4(q[0], q[1], q[2], q[3]);
  }
}

// LIT: the epilogue carries the literal (double-precision) recipe as well -- only layers with a channel that needs it
template <bool LIT>
__global__ void __launch_bounds__(kThreads, 1) dwconv_tape_kernel(const __grid_constant__ CUtensorMap tmX, const TapeArgs a) {
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
  // channel records of the items in flight (written by the producer, read by the epilogue warps): {int4 record, float r}
  ChanSlot* const chan_ring = reinterpret_cast<ChanSlot*>(smem_gen + (bars + 8u * (2 * kMaxBufs + 2 * kSlots + 2) - smem_base));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int i0 = (int)((long long)blockIdx.x * a.items / gridDim.x);
  const int i1 = (int)((long long)(blockIdx.x + 1) * a.items / gridDim.x);
  if (threadIdx.x == 0) TAPE_TRACE(3, 0);

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
  if (threadIdx.x == 0) TAPE_TRACE(3, 1);

  if (warp == kProducerWarp) {
    // ===================== producer: ONE tensor copy per item =====================
    // The (B, C, T) tensor is described to TMA as [B][C][T / inner][inner]; a box [nb][1][rpu * 16 / inner][inner] that starts
    // G / inner rows BEFORE the utterance lands in shared memory as nb x P contiguous bytes = the tape, with the gap in
    // front of every utterance (and whatever lies past the end of a row or of the batch) zero-filled by TMA.
    pdl_wait();   // the activation rows come from the previous kernel
    if (lane == 0) {
      qasr_stamp_begin(a.stamp);
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    }
    int buf = 0;
    uint32_t phase = 0;
    int c = i0 / nbc, bc = i0 - c * nbc;
    for (int it = i0; it < i1; ++it) {
      mbar_wait(empty_bar(buf), phase ^ 1u);
      TAPE_TRACE(0, 2 * (it - i0));
      if (lane == 0) {
        // the item's channel record, with the float multiplier of the fast path derived here (off the epilogue's path)
        const int4 rec = __ldg(reinterpret_cast<const int4*>(a.ep.ch0) + c);
        ChanSlot& cs = chan_ring[(it - i0) & (kChanRing - 1)];
        cs.rec = rec;
        cs.r = (a.ep.flags & QASR_F_FAST_F32) ? __double2float_rn(dyad_mult(rec.z, rec.w & 0x7fffffff))
                                              : (a.ep.rf != nullptr ? __ldg(a.ep.rf + c) : 0.0f);
        mbar_expect_tx(full_bar(buf), (uint32_t)(nb * a.P) + (uint32_t)a.KS * kBmatBytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(bmat_base + (uint32_t)buf * kMaxKS * kBmatBytes), "l"(a.bmat + (size_t)c * a.KS * kBmatBytes),
                       "r"((uint32_t)a.KS * kBmatBytes), "r"(full_bar(buf)) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_base + (uint32_t)buf * a.tape_bytes), "l"(&tmX), "r"(full_bar(buf)), "r"(0), "r"(-(a.G / a.inner)), "r"(c),
              "r"(bc * nb) : "memory");
      }
      __syncwarp();
      TAPE_TRACE(0, 2 * (it - i0) + 1);
      if (++buf == a.nbuf) { buf = 0; phase ^= 1u; }
      if (++bc == nbc) { bc = 0; ++c; }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc = a.idesc;
    const int KS = a.KS;
    const uint64_t a_hi = desc_noswizzle(0, 16u, 128u), b_hi = desc_noswizzle(0, 128u, 256u);
    int buf = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
    int bc = i0 % nbc;
    for (int it = i0; it < i1; ++it) {
      const int nrows = min(nb, a.B - bc * nb);
      const int units = (nrows * rpu + 127) >> 7;
      TAPE_TRACE(1, 3 * (it - i0));
      mbar_wait(full_bar(buf), phase);
      TAPE_TRACE(1, 3 * (it - i0) + 1);
      tc_fence_after();
      const uint32_t tape = smem_base + (uint32_t)buf * a.tape_bytes;
      const uint32_t bm = bmat_base + (uint32_t)buf * kMaxKS * kBmatBytes;
      for (int u = 0; u < units; u += kRound) {
        mbar_poll(tempty_bar(slot), sphase ^ 1u);
        tc_fence_after();
        // descriptors as warp-uniform values (low word = address >> 4, advanced by plain adds): they live in uniform
        // registers and the elected lane issues back to back (see gemm_tc.cu)
        const uint32_t a_lo0 = ((tape & 0x3FFFFu) >> 4) + (uint32_t)u * 128u, b_lo0 = (bm & 0x3FFFFu) >> 4;
        const uint32_t d0 = tmem_base + (uint32_t)(slot * kRound * 16);
        const int ng = min(kRound, units - u);
        if (leader && !(a.dbg_skip & 2)) {
          // K step outermost: consecutive MMAs write different accumulators (a dependent accumulate waits for the
          // whole pipeline latency of its predecessor, which at N = 16 is several times the MMA's own slot)
          for (int ks = 0; ks < KS; ++ks) {
            for (int g = 0; g < ng; ++g)
              tc_mma_i8(d0 + (uint32_t)(g * 16), a_hi | (uint64_t)(a_lo0 + (uint32_t)(g * 128 + ks * 2)),
                        b_hi | (uint64_t)(b_lo0 + (uint32_t)(ks * 32)), idesc, ks != 0 ? 1u : 0u);
          }
        }
        if (leader) tc_commit(tfull_bar(slot));
        __syncwarp();
        if (++slot == kSlots) { slot = 0; sphase ^= 1u; }
      }
      if (leader) tc_commit(empty_bar(buf));   // the tape is free once every MMA that reads it has retired
      __syncwarp();
      TAPE_TRACE(1, 3 * (it - i0) + 2);
      if (++buf == a.nbuf) { buf = 0; phase ^= 1u; }
      if (++bc == nbc) bc = 0;
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue: TMEM -> codes -> one 16-byte store per lane =====================
    // Every warp keeps the TMEM load of its NEXT unit in flight while it requantises and stores the current one (two
    // register buffers, alternating), also across item boundaries.  What a unit needs later is reduced to a 4-register
    // job at issue time (where its lane stores, how many of its 16 frames are valid, which channel record it uses).
    const int quarter = warp & 3, grp = warp >> 2;
    const EpiConsts kc = make_epi_consts(a.ep, false);
    const bool masked = a.ep.flags & QASR_F_MASK;
    const bool literal_all = a.ep.flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64);
    const bool dyn = a.ep.flags & QASR_F_DEVICE_TABLES;
    pdl_wait();   // out / out_lens (and device-written tables) belong to the stream order
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int lane_row = quarter * 32 + lane;
    uint8_t* const out = reinterpret_cast<uint8_t*>(a.out);
    constexpr int kNoStore = INT_MIN;

    struct Job {
      uint8_t* ptr;   // this lane's 16 output bytes
      int nvalid;     // frames of them below the utterance's length (<= 0: all zero); kNoStore: a row of the gap
      int ring;       // slot of the channel record
    };
    // Walk over the CTA's items.  An item is ONE round: the MMA issuer commits all its units (<= 16, 16 TMEM columns each)
    // at once; group g drains units pos, pos + 4, pos + 8, pos + 12 with pos = (g + rot) & 3, rot carrying the deal over the
    // item boundary so that the groups stay evenly loaded.  Every warp waits once per item for the accumulators and hands
    // the slot back once, after the last of its loads has landed.
    int it = i0, c = i0 / nbc, bc = i0 - c * nbc;
    int unit = 0, units = 0, rows_total = 0, rot = 0;
    uint32_t slot = 0, sphase = 0;
    bool open = false;   // the current item's accumulators have been waited for
    // after the previous load has landed: moves to this warp's next unit (retiring items on the way), starts its TMEM load
    // and fills in its job; false when the CTA's items are exhausted
    auto next = [&](int (&acc)[16], Job& j) -> bool {
      for (;;) {
        if (it >= i1) return false;
        if (!open) {
          rows_total = min(nb, a.B - bc * nb) * rpu;
          units = (rows_total + 127) >> 7;
          unit = (grp + rot) & 3;
          mbar_sleep(tfull_bar(slot), sphase);
          tc_fence_after();
          open = true;
        }
        if (unit < units) break;
        // this warp is through with the item: hand the slot back, step to the next item
        tc_fence_before();
        mbar_arrive_warp(tempty_bar(slot));
        if (++slot == kSlots) { slot = 0; sphase ^= 1u; }
        rot = (rot - units) & 3;
        ++it;
        if (++bc == nbc) { bc = 0; ++c; }
        open = false;
      }
      tmem_ld16(lane_addr + (uint32_t)(slot * (kRound * 16) + unit * 16), acc);
      const int R = unit * 128 + lane_row;                        // tape row of this lane
      const int bl = __float2int_rz(((float)R + 0.5f) * a.inv_rpu);
      const int i = R - bl * rpu;
      const int b = bc * nb + bl;
      const bool valid = R < rows_total && i < a.rows_valid;       // else: a row of the gap, no output frames
      int len = a.T_out;
      if (masked && valid) len = min(__ldg(a.ep.out_lens + b), len);
      j.ptr = out + ((size_t)b * a.C + c) * a.out_ld + 16 * i;
      j.nvalid = valid ? len - 16 * i : kNoStore;
      j.ring = (it - i0) & (kChanRing - 1);
      unit += 4;
      return true;
    };
    // 16 accumulators -> 16 codes, length mask, one 16-byte store
    auto finish = [&](const int (&acc)[16], const Job& j) {
      if (j.nvalid == kNoStore || (a.dbg_skip & 1)) return;
      // the producer wrote the record before the tape became full, the MMA issuer saw the tape full before it committed
      // the round this unit came from
      const ChanSlot& cs = chan_ring[j.ring];
      ChanRegs c0;
      c0.r = cs.r;
      c0.bias = cs.rec.y;
      uint32_t w[4];
      if (c0.r != 0.0f) {                                  // float fast path (one channel per unit: warp-uniform)
        tape_codes<0>(a.ep, kc, c0, acc, w);
      } else {
        const int4 rec = cs.rec;
        c0.s = __int_as_float(rec.x);
        c0.m = rec.z;
        c0.e = rec.w & 0x7fffffff;
        c0.literal = rec.w < 0;
        if (LIT && literal_all && !(dyn && !c0.literal)) {  // literal recipe
          c0.mult = dyad_mult(c0.m, c0.e);
          tape_codes<2>(a.ep, kc, c0, acc, w);
        } else {                                           // integer fast path
          c0.sh = max(c0.e - 32, 0);
          c0.add = (long long)c0.bias * (long long)c0.m + ((c0.e >= 1 && c0.e <= 62) ? (1ll << (c0.e - 1)) : 0ll);
          tape_codes<1>(a.ep, kc, c0, acc, w);
        }
      }
      if (j.nvalid < 16) {   // length mask / row tail: zero the bytes at t >= len
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int nv = j.nvalid - 4 * q4;
          w[q4] &= nv >= 4 ? 0xffffffffu : (nv <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nv))));
        }
      }
      *reinterpret_cast<uint4*>(j.ptr) = make_uint4(w[0], w[1], w[2], w[3]);
    };

    // two register buffers: the TMEM load of the next unit is in flight while the current one is requantised and stored
    {
      int accA[16], accB[16];
      Job jA, jB;
      bool hA = next(accA, jA);
      while (hA) {
        tmem_ld_wait();
        const bool hB = next(accB, jB);
        finish(accA, jA);
        if (!hB) break;
        tmem_ld_wait();
        hA = next(accA, jA);
        finish(accB, jB);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TAPE_TRACE(3, 2);
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
  g.G = round_up_i(pad > 0 ? pad : 1, 32);   // multiple of 32: rows whose length is one too can travel as 32-byte TMA rows
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
  // utterances per work item: at most 16 units (one accumulator round of 256 TMEM columns), and enough items to balance
  // 148 CTAs
  const int sms = qasr_sm_count();
  const int nb_cap = (kRound * 128) / a.rpu;
  if (nb_cap < 1) {
    qasr_set_error("dwconv(tape): rows of %d frames exceed one accumulator round", T_in);
    return QASR_ERR_UNSUPPORTED;
  }
  int nb = nb_cap < B ? nb_cap : B;
  while (nb > 4 && (long long)C * ((B + nb - 1) / nb) < 8LL * sms) nb = (nb + 1) / 2;
  nb = (B + (B + nb - 1) / nb - 1) / ((B + nb - 1) / nb);   // same number of chunks, evenly sized
  if (qasr_debug_get(15) > 0) nb = (int)qasr_debug_get(15) < B ? (int)qasr_debug_get(15) : B;   // bring-up knob 15: utterances per item
  if (nb > nb_cap) nb = nb_cap;
  a.nb = nb;
  a.nbc = (B + nb - 1) / nb;
  a.items = C * a.nbc;
  const int units_max = (nb * a.rpu + 127) / 128;
  a.tape_bytes = (uint32_t)round_up_i(units_max * 2048 + 32 * a.KS + 16, 128);
  const uint32_t fixed = 128 /*align*/ + 8u * (2 * kMaxBufs + 2 * kSlots + 2) + kChanRing * (uint32_t)sizeof(ChanSlot);
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
  // row view for TMA: the largest inner length that divides both the gap and the row (16 always does)
  int inner = 16;
  for (int cand = 256; cand > 16; cand -= 16)
    if (a.G % cand == 0 && a.copy_bytes % cand == 0) { inner = cand; break; }
  if (qasr_debug_get(14) >= 16 && a.G % (int)qasr_debug_get(14) == 0 && a.copy_bytes % (int)qasr_debug_get(14) == 0)
    inner = (int)qasr_debug_get(14);   // bring-up: force the inner length
  a.inner = inner;
  CUtensorMap map;
  {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { qasr_set_error("cuTensorMapEncodeTiled entry point not available"); return QASR_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)(a.copy_bytes / inner), (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)inner, (cuuint64_t)x_ld, (cuuint64_t)x_ld * (cuuint64_t)C};
    cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)(a.P / inner), 1, (cuuint32_t)a.nb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { qasr_set_error("cuTensorMapEncodeTiled(tape) failed: %d", (int)r); return QASR_ERR_CUDA; }
  }
  a.ep = make_epi_args(epi);
  a.dbg_skip = qasr_debug_get(6) > 0 ? (int)qasr_debug_get(6) : 0;
  a.dbg_trace = qasr_debug_get(7) > 0 ? reinterpret_cast<unsigned long long*>(qasr_debug_get(7)) : nullptr;
  a.stamp = qasr_next_stamp();
  const size_t smem = fixed + (size_t)nbuf * (a.tape_bytes + kMaxKS * kBmatBytes);
  int grid = sms < a.items ? sms : a.items;
  static bool attr_set[64] = {};
  int dev = 0;
  QASR_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    QASR_CHECK_CUDA(cudaFuncSetAttribute(dwconv_tape_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    QASR_CHECK_CUDA(cudaFuncSetAttribute(dwconv_tape_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    attr_set[dev] = true;
  }
  const bool lit = a.ep.flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64);
  if (lit) {
    QASR_CHECK_CUDA(qasr_launch_pdl(dwconv_tape_kernel<true>, dim3(grid), dim3(kThreads), smem, (cudaStream_t)stream, map, a));
  } else {
    QASR_CHECK_CUDA(qasr_launch_pdl(dwconv_tape_kernel<false>, dim3(grid), dim3(kThreads), smem, (cudaStream_t)stream, map, a));
  }
  return qasr_check_launch("dwconv_tape_kernel");
}

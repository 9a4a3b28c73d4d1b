// K1-tc: depthwise 1-D convolution on 8-bit codes as a banded (Toeplitz) GEMM on the tcgen05 tensor path.
// Same arithmetic as dwconv.cu (F.conv1d(x_int, w_int, groups=C) of QuantConv1d.int_conv,
// nemo/quantization/utils/quant_modules.py:301-305, in true integers), different engine: the IMMA version
// is issue-bound at ~75 warp instructions per 128 outputs (29 us for a 512-channel layer at batch 64 x 800
// frames); tcgen05 has ~4x the int8 rate of IMMA and needs no per-warp fragment work, so even a Toeplitz
// product that is 13-35 % dense costs ~3 us of tensor time for the same layer.
//
// GEMM view for ONE channel c (the CTA keeps c fixed for as long as it can):
//   D[r, n] = sum_j A[r, j] * Tc[n, j]          r = (time block, utterance), n = frame inside the block
//   A[r, j] = x[b, c, S * N * tblk - pad - o + j]  the input window of that block: plain contiguous bytes of
//             the reference's (B, C, T) layout = K-major rows.  One TMA box = [128 B of time] x [1 channel] x
//             [bbox utterances]; window starts are rounded down to 16 bytes (o = (-pad) mod 16: TMA faults
//             on unaligned inner coordinates) and out-of-range frames are zero-filled by TMA = conv padding.
//   Tc[n, j] = w[c, (j - o - S n) / dil]           banded matrix of the channel's taps, built in shared
//             memory by the epilogue warps (only the band chunks are rewritten per channel, the zero chunks
//             are written once), K-major with the 128-byte swizzle, double buffered across channels.
//   D: int32 in TMEM, lane = r, column = n; each epilogue thread owns 32 consecutive frames of one
//      (utterance, channel) row -> requantise with CTA-uniform channel constants -> one 32-byte store.
// Warp roles (736 threads, 1 CTA / SM, static contiguous range of (channel, row tile) units per CTA):
//   warps 0..19 epilogue + Tc builder (quarter = warp & 3 is the TMEM lane quarter, warp >> 2 the 32-column
//   group), warp 20 TMA producer, warp 21 MMA issuer, warp 22 TMEM allocator.  Producer / MMA warps walk
//   their loops with warp-uniform values so descriptors stay in uniform registers (see gemm_tc.cu).
#include <cuda.h>

#include <type_traits>

#include "qasr_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace qasr_tc;

constexpr int kEpiWarps = 20;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kAllocWarp = kEpiWarps + 2;
constexpr int kThreads = 32 * (kEpiWarps + 3);
constexpr int kColGroups = kEpiWarps / 4;
constexpr int kMaxStages = 6;
constexpr int kMaxSlabs = 3;
constexpr uint32_t kSlabBytes = 128 * 128;   // one K slab of the A tile: 128 rows x 128 bytes
constexpr uint32_t kSmemBudget = 227 * 1024;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;         // TMEM columns between the two accumulator buffers

struct DwTcArgs {
  const int8_t* taps;   // [C][k] int8
  void* out;
  int out_ld, B, C, T_in, T_out, k, stride, dil, pad;
  int o;             // (-pad) mod 16: bytes the window start was rounded down by
  int span;          // (k - 1) * dil + 1
  int N;             // output frames per time block (multiple of 32)
  int slabs;         // 128-byte K slabs of the input window
  int bbox;          // utterances per TMA box (multiple of 8, divides 128)
  int tb_per_tile;   // time blocks per 128-row tile = 128 / bbox
  int n_tblk;        // ceil(T_out / N)
  int ntbg;          // ceil(n_tblk / tb_per_tile)
  int tiles_per_ch;  // ceil(B / bbox) * ntbg
  int units;         // C * tiles_per_ch
  int stages;
  int zl;            // length of one shifted copy of the zero-padded dilated taps
  int nb;            // band chunks (16 B) per Tc row, upper bound
  uint32_t idesc;
  uint32_t tmat_bytes, stage_bytes;
  // row-staged variant (ROWS): per unit, the piece of every utterance row that its time blocks need (one bulk copy
  // per row, two piece buffers in flight) lands in shared memory and the epilogue warps cut the overlapping
  // K-major windows out of it; bytes outside [0, copy_bytes) of a row are the conv padding and are cut as zeros
  const uint8_t* x;
  int x_ld;
  int copy_bytes;      // T_in rounded up to 16: valid bytes of one row
  int row_pitch;       // bytes of one row's piece: (tb_per_tile - 1) * N * stride + slabs * 128
  uint32_t rows_bytes; // one piece buffer: bbox * row_pitch (two of them follow the A stages)
  unsigned long long* dbg_trace;   // qasr_debug_set 7: device buffer, role timeline of CTA 0 (3 x 256 slots)
  int dbg_skip;      // qasr_debug_set 6: 1 = epilogue does nothing, 2 = no MMA issue, 4 = no stores, 8 = no Tc rebuild
  unsigned long long* stamp;   // in-graph timing slot (qasr_profile_begin) or nullptr
  EpiArgs ep;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

#define DWTC_TRACE(role, n) do { if (a.dbg_trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (n) < 256) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.dbg_trace[(role) * 256 + (n)] = t_; } } while (0)

// EK: 0 = REQUANT all-integer fast path, 1 = REQUANT literal path, 2 = DEQUANT_F32 / RAW_I32,
//     3 = REQUANT float fast path (QASR_F_FAST_F32)
// ROWS: false = input windows fetched as TMA boxes (any window, <= kMaxSlabs slabs), true = row-staged variant
template <int EK, bool ROWS>
__global__ void __launch_bounds__(kThreads, 1) dwconv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const DwTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));   // generic pointer to smem_base
  const uint32_t tmat_base = smem_base;                                      // 2 x Tc
  const uint32_t stage_base = tmat_base + 2u * a.tmat_bytes;                 // stages x A tile
  const uint32_t rows_off = 2u * a.tmat_bytes + (uint32_t)a.stages * a.stage_bytes;   // ROWS: staged input rows
  const uint32_t z_off = rows_off + (ROWS ? 2u * a.rows_bytes : 0u);
  const uint32_t bars = smem_base + z_off + 4u * (uint32_t)a.zl;             // zl is a multiple of 16
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + 2 + s); };
  auto mfull_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + 4 + s); };
  auto pfull_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + 6 + s); };
  auto pempty_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + 8 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 10);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int u0 = (int)((long long)blockIdx.x * a.units / gridDim.x);
  const int u1 = (int)((long long)(blockIdx.x + 1) * a.units / gridDim.x);

  if (warp == kProducerWarp && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full_bar(s), ROWS ? kEpiWarps : 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(pfull_bar(s), 1);
      mbar_init(pempty_bar(s), kEpiWarps);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps);
      mbar_init(mfull_bar(s), kEpiWarps);
    }
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

  if (warp == kProducerWarp && ROWS) {
    // ===================== row producer: per unit, one bulk copy per utterance row =====================
    const int tiles_per_ch = a.tiles_per_ch, ntbg = a.ntbg, bbox = a.bbox;
    const int piece_step = a.tb_per_tile * a.N * a.stride, piece_len = a.row_pitch, tshift = a.pad + a.o;
    pdl_wait();                // the activation rows come from the previous kernel
    if (lane == 0) qasr_stamp_begin(a.stamp);
    int c = u0 / tiles_per_ch, bc, tbg;
    {
      const int r = u0 - c * tiles_per_ch;
      bc = r / ntbg;
      tbg = r - bc * ntbg;
    }
    uint32_t phase = 0;
    int pb = 0;
    for (int u = u0; u < u1; ++u) {
      const int b0 = bc * bbox;
      const int nrows = min(bbox, a.B - b0);
      // the piece covers row bytes [t_p, t_p + piece_len); only its part inside [0, copy_bytes) exists in memory
      const int t_p = tbg * piece_step - tshift;
      const int src0 = max(t_p, 0);
      const int len = min(t_p + piece_len, a.copy_bytes) - src0;   // multiples of 16 throughout
      mbar_wait(pempty_bar(pb), phase ^ 1u);
      if (lane == 0) mbar_expect_tx(pfull_bar(pb), (len > 0 && !(a.dbg_skip & 32)) ? (uint32_t)(nrows * len) : 0u);
      __syncwarp();
      if (len > 0 && !(a.dbg_skip & 32)) {
        const uint32_t dst0 = smem_base + rows_off + (uint32_t)pb * a.rows_bytes + (uint32_t)(src0 - t_p);
        const uint8_t* srow = a.x + ((size_t)b0 * a.C + c) * a.x_ld + src0;
        for (int r = lane; r < nrows; r += 32)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst0 + (uint32_t)(r * piece_len)), "l"(srow + (size_t)r * a.C * a.x_ld), "r"((uint32_t)len),
                         "r"(pfull_bar(pb)) : "memory");
      }
      pb ^= 1;
      phase ^= (uint32_t)(pb == 0);
      if (++tbg == ntbg) { tbg = 0; if (++bc * ntbg == tiles_per_ch) { bc = 0; ++c; } }
    }
  } else if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    pdl_wait();   // the activation rows come from the previous kernel
    if (leader) qasr_stamp_begin(a.stamp);
    const int tiles_per_ch = a.tiles_per_ch, ntbg = a.ntbg, tbpt = a.tb_per_tile, n_tblk = a.n_tblk;
    const int slabs = a.slabs, bbox = a.bbox, stages = a.stages;
    const int tstep = a.N * a.stride, tshift = a.pad + a.o;
    const uint32_t box_bytes = (uint32_t)bbox * 128u;
    int stage = 0;
    uint32_t phase = 0;
    for (int u = u0; u < u1; ++u) {
      const int c = u / tiles_per_ch, r = u - c * tiles_per_ch;
      const int bc = r / ntbg, tbg = r - bc * ntbg;
      const int nbox = min(tbpt, n_tblk - tbg * tbpt);
      mbar_wait(empty_bar(stage), phase ^ 1u);
      if (leader) {
        const uint32_t dst = stage_base + (uint32_t)stage * a.stage_bytes;
        mbar_expect_tx(full_bar(stage), (uint32_t)(nbox * slabs) * box_bytes);
        for (int tb = 0; tb < nbox; ++tb) {
          const int t_start = (tbg * tbpt + tb) * tstep - tshift;
          for (int sl = 0; sl < slabs; ++sl)
            tma_load_3d(dst + (uint32_t)sl * kSlabBytes + (uint32_t)tb * box_bytes, &tmX, t_start + sl * 128, c,
                        bc * bbox, full_bar(stage));
        }
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const int tiles_per_ch = a.tiles_per_ch, slabs = a.slabs, stages = a.stages;
    const uint64_t desc_hi = make_smem_desc(0, 0, 1024);
    const uint32_t tmat_lo0 = (tmat_base & 0x3FFFFu) >> 4, tmat_step_lo = a.tmat_bytes >> 4;
    const uint32_t tslab_lo = ((uint32_t)a.N * 128u) >> 4;
    const uint32_t stage_lo0 = (stage_base & 0x3FFFFu) >> 4, stage_step_lo = a.stage_bytes >> 4;
    const uint32_t idesc = a.idesc;
    int stage = 0;
    uint32_t phase = 0, stage_lo = stage_lo0;
    int as = 0;
    uint32_t aphase = 0;
    int cur_c = -1, chan_idx = -1;
    uint32_t tmat_lo = tmat_lo0;
    for (int u = u0; u < u1; ++u) {
      const int c = u / tiles_per_ch;
      if (c != cur_c) {   // new channel: its banded matrix must have been built
        cur_c = c;
        ++chan_idx;
        mbar_wait(mfull_bar(chan_idx & 1), (uint32_t)((chan_idx >> 1) & 1));
        tmat_lo = tmat_lo0 + (uint32_t)(chan_idx & 1) * tmat_step_lo;
      }
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      DWTC_TRACE(1, 2 * (u - u0));
      if (leader) {
        const uint32_t d_tmem = tmem_base + (uint32_t)as * kAccStride;
        for (int sl = 0; sl < slabs; ++sl) {
          const uint32_t a_lo = stage_lo + (uint32_t)sl * (kSlabBytes >> 4);
          const uint32_t b_lo = tmat_lo + (uint32_t)sl * tslab_lo;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (!(a.dbg_skip & 2))
              tc_mma_i8(d_tmem, desc_hi | (uint64_t)(a_lo + kk * 2), desc_hi | (uint64_t)(b_lo + kk * 2), idesc,
                        (sl | kk) != 0 ? 1u : 0u);
        }
        if (!ROWS) tc_commit(empty_bar(stage));   // ROWS: the builders reuse a stage only after draining its tile
        tc_commit(tfull_bar(as));
      }
      DWTC_TRACE(1, 2 * (u - u0) + 1);
      stage_lo += stage_step_lo;
      if (++stage == stages) { stage = 0; phase ^= 1u; stage_lo = stage_lo0; }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp < kEpiWarps) {
    // ===================== Tc builder + epilogue =====================
    const int tid = threadIdx.x;   // 0 .. kEpiThreads-1
    const int quarter = warp & 3, cg = warp >> 2;
    const EpiConsts kc = make_epi_consts(a.ep, false);
    const int tiles_per_ch = a.tiles_per_ch, ntbg = a.ntbg, tbpt = a.tb_per_tile, bbox = a.bbox;
    const int N = a.N, S = a.stride, o = a.o, span = a.span, zl = a.zl, nb = a.nb, nq = a.slabs * 8;
    uint8_t* const zbuf = smem_gen + z_off;

    // zero both Tc buffers once: the band positions never move, so zero chunks stay zero for every channel
    for (uint32_t i = tid; i < 2u * a.tmat_bytes / 16u; i += kEpiThreads)
      reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0, 0, 0, 0);
    epi_bar_sync();

    auto build = [&](int c, int buf) {
      if ((a.dbg_skip & 8) && c != u0 / tiles_per_ch) { mbar_arrive_warp(mfull_bar(buf)); return; }
      // 4 byte-shifted copies of the zero-padded dilated taps: copy s, index m holds wd[m + s - 16]
      for (int idx = tid; idx < 4 * zl; idx += kEpiThreads) {
        const int s = idx / zl, m = idx - s * zl;
        const int i = m + s - 16;
        int8_t v = 0;
        if (i >= 0 && i < span) {
          const int tap = i / a.dil;
          if (tap * a.dil == i) v = a.taps[(size_t)c * a.k + tap];
        }
        reinterpret_cast<int8_t*>(zbuf)[idx] = v;
      }
      epi_bar_sync();
      uint8_t* const tm = smem_gen + (uint32_t)buf * a.tmat_bytes;
      for (int idx = tid; idx < N * nb; idx += kEpiThreads) {
        const int n = idx / nb, qi = idx - n * nb;
        const int first = o + S * n;                 // window byte of tap 0 for output n
        const int qg = (first >> 4) + qi;            // 16-byte chunk of the Tc row
        if (qg < nq && qg * 16 <= first + span - 1) {
          const int p = qg * 16 - first + 16;        // >= 1: index into the unshifted padded taps
          const int s = p & 3;
          const uint32_t* zw = reinterpret_cast<const uint32_t*>(zbuf + s * zl + (p - s));
          const uint4 v = make_uint4(zw[0], zw[1], zw[2], zw[3]);
          const int slab = qg >> 3, q = qg & 7;
          *reinterpret_cast<uint4*>(tm + (size_t)slab * (N * 128) + (n >> 3) * 1024 + (n & 7) * 128 + ((q ^ (n & 7)) << 4)) = v;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tcgen05 (async proxy) reads
      mbar_arrive_warp(mfull_bar(buf));
      epi_bar_sync();   // nobody still reads the tap copies when the next build overwrites them
    };

    // Per-unit control is kept free of divisions: a single warp runs dependent scalar code at ~6 clk per
    // instruction (measured: ~1 us per unit of pure index arithmetic and barrier traffic before this rewrite), so
    // the (channel, utterance chunk, time-block group) decomposition is carried incrementally and everything that
    // depends on the thread only is computed once.
    const int n_tblk = a.n_tblk;
    // ROWS: this thread's 16-byte chunks of a 128-row x (slabs x 128 B) window tile
    constexpr int kCut = 4;   // 128 rows x 16 chunks (2 slabs) / 640 threads
    int cut_src[kCut], cut_dst[kCut], cut_tb[kCut], cut_t[kCut];
#pragma unroll
    for (int i = 0; i < kCut; ++i) {
      const int idx = tid + i * kEpiThreads;
      const int rr = idx / nq, q = idx - rr * nq;
      const int tb = rr / bbox, bl = rr - tb * bbox;
      cut_tb[i] = idx < 128 * nq ? tb : (1 << 20);                      // beyond the tile: never valid
      cut_t[i] = tb * (N * S) + q * 16;
      cut_src[i] = bl * a.row_pitch + cut_t[i];
      cut_dst[i] = (q >> 3) * (int)kSlabBytes + rr * 128 + (((q & 7) ^ (rr & 7)) << 4);
    }
    // cut the windows of unit number i (time-block group tbg) out of piece buffer i & 1 into A stage i & 1
    const int piece_step = tbpt * N * S, tshift = a.pad + o, copy_bytes = a.copy_bytes;
    uint32_t piece_phase = 0;
    auto prebuild = [&](int i, int tbg) {
      const int pb = i & 1;
      mbar_wait(pfull_bar(pb), piece_phase);
      piece_phase ^= (uint32_t)pb;   // flips after buffer 1, i.e. once per pair of units
      uint8_t* const dst = smem_gen + 2u * a.tmat_bytes + (uint32_t)pb * a.stage_bytes;
      const uint8_t* const src = smem_gen + rows_off + (uint32_t)pb * a.rows_bytes;
      const int tb_left = n_tblk - tbg * tbpt;    // time blocks that exist in this group
      const int t_p = tbg * piece_step - tshift;  // row byte at piece offset 0
      const bool interior = t_p >= 0 && t_p + a.row_pitch <= copy_bytes;   // warp-uniform: no padding inside
      if (!(a.dbg_skip & 16)) {
#pragma unroll
        for (int j = 0; j < kCut; ++j) {
          if (cut_tb[j] < tb_left) {
            uint4 v = make_uint4(0, 0, 0, 0);
            const int t = t_p + cut_t[j];
            if (interior || (t >= 0 && t < copy_bytes)) v = *reinterpret_cast<const uint4*>(src + cut_src[j]);
            *reinterpret_cast<uint4*>(dst + cut_dst[j]) = v;
          }
        }
      }
      // generic reads of the piece are done and generic writes of the windows must reach the async proxy (tcgen05
      // reads them, the next bulk copy overwrites the piece): the warp's accesses are ordered before lane 0 by the
      // __syncwarp, lane 0 fences for the async proxy and arrives (one fence per warp, not per thread)
      __syncwarp();
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(full_bar(pb));
        mbar_arrive(pempty_bar(pb));
      }
    };

    // this thread's row of the 128-row tile
    const int rowi = quarter * 32 + lane;
    const int my_tb = rowi / bbox, my_bl = rowi - my_tb * bbox;
    const int t_store_end = ((a.T_out + 15) >> 4) << 4;   // stores cover T_out rounded up to 16 (zero filled)
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool masked = a.ep.flags & QASR_F_MASK;

    int chan_idx = 0;
    // unit u0 decomposed once; afterwards (c, bc, tbg) advance incrementally
    int c = u0 / tiles_per_ch, bc, tbg;
    {
      const int r = u0 - c * tiles_per_ch;
      bc = r / ntbg;
      tbg = r - bc * ntbg;
    }
    if (u0 < u1) build(c, 0);   // taps are constants of the plan: built before the dependency wait
    pdl_wait();
    if (ROWS && u0 < u1) prebuild(0, tbg);
    int as = 0;
    uint32_t aphase = 0;
    bool new_channel = true;
    ChanRegs c0 = load_chan0(a.ep, c);
    for (int u = u0; u < u1; ++u) {
      // next unit's coordinates
      int nc = c, nbc = bc, ntg = tbg + 1;
      if (ntg == ntbg) { ntg = 0; if (++nbc * ntbg == tiles_per_ch) { nbc = 0; ++nc; } }
      const bool has_next = u + 1 < u1;
      if (warp == 0) DWTC_TRACE(2, 4 * (u - u0));
      // the next unit's windows are cut before this unit is drained, so its MMAs overlap the drain (its piece was
      // requested two units ago; its A stage held unit u - 1, whose accumulator is already drained)
      if (ROWS && has_next) prebuild(u + 1 - u0, ntg);
      if (warp == 0) DWTC_TRACE(2, 4 * (u - u0) + 1);
      if (new_channel) {
        if (u != u0) { ++chan_idx; c0 = load_chan0(a.ep, c); }
        // the other buffer belonged to channel c-1, whose last accumulator we have already drained
        if ((c + 1) * tiles_per_ch < u1) build(c + 1, (chan_idx + 1) & 1);
      }
      const int b = bc * bbox + my_bl;
      const int tblk = tbg * tbpt + my_tb;
      const bool valid = tblk < n_tblk && b < a.B;
      const int bb = valid ? b : 0;
      const int len = masked ? min(a.ep.out_lens[bb], a.T_out) : a.T_out;
      const size_t row = (size_t)bb * a.C + c;

      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (warp == 0) DWTC_TRACE(2, 4 * (u - u0) + 2);
      auto drain = [&](auto f32_tag) {   // float or integer epilogue, chosen once per unit (one channel per unit)
      constexpr bool F32 = decltype(f32_tag)::value;
      for (int ck = cg; ck * 32 < N && !(a.dbg_skip & 1); ck += kColGroups) {
        int acc[32];
        tmem_ld32(lane_taddr + (uint32_t)as * kAccStride + (uint32_t)(ck * 32), acc);
        tmem_ld_wait();
        const int t0 = tblk * N + ck * 32;
        if (EK != 2) {
          uint32_t packed[8];
#pragma unroll
          for (int w4 = 0; w4 < 8; ++w4) {
            int q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              q[i] = F32 ? (int)epilogue_f32_bits(kc, acc[4 * w4 + i], c0)
                     : EK == 0 ? epilogue_fast<false, false, false>(a.ep, kc, acc[4 * w4 + i], c0, 0, c0, 0)
                               : epilogue_exact(a.ep, kc, acc[4 * w4 + i], c0, false, 0, c0, 0);
            packed[w4] = pack4(q[0], q[1], q[2], q[3]);
          }
          if (t0 + 32 > len) {   // length mask / row tail: zero the bytes at t >= len
#pragma unroll
            for (int w4 = 0; w4 < 8; ++w4) {
              const int nvalid = len - (t0 + 4 * w4);
              packed[w4] &= nvalid >= 4 ? 0xffffffffu : (nvalid <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nvalid))));
            }
          }
          uint8_t* orow = reinterpret_cast<uint8_t*>(a.out) + row * a.out_ld + t0;
          if (a.dbg_skip & 4) continue;
          if (valid && t0 < t_store_end) *reinterpret_cast<uint4*>(orow) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          if (valid && t0 + 16 < t_store_end)
            *reinterpret_cast<uint4*>(orow + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        } else if (a.ep.mode == QASR_EPI_DEQUANT_F32) {
          float* orow = reinterpret_cast<float*>(a.out) + row * a.out_ld;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (valid && t0 + j < a.T_out) orow[t0 + j] = epilogue_dequant(a.ep, acc[j], c0);
        } else {
          int32_t* orow = reinterpret_cast<int32_t*>(a.out) + row * a.out_ld;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (valid && t0 + j < a.T_out) orow[t0 + j] = acc[j] + c0.bias;
        }
      }
      };
      if (EK == 3 || (EK == 0 && c0.r != 0.0f)) drain(std::true_type{}); else drain(std::false_type{});
      tc_fence_before();
      mbar_arrive_warp(tempty_bar(as));
      as ^= 1;
      aphase ^= (uint32_t)(as == 0);
      if (warp == 0) DWTC_TRACE(2, 4 * (u - u0) + 3);
      new_channel = nc != c;
      c = nc; bc = nbc; tbg = ntg;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) qasr_stamp_end(a.stamp);
  if (warp == kAllocWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

int make_map_rows(CUtensorMap* map, const void* x, int T, int C, int B, int ld, int bbox) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { qasr_set_error("cuTensorMapEncodeTiled entry point not available"); return QASR_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * (cuuint64_t)C};
  cuuint32_t box[3] = {128, 1, (cuuint32_t)bbox};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { qasr_set_error("cuTensorMapEncodeTiled(dw rows) failed: %d", (int)r); return QASR_ERR_CUDA; }
  return QASR_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace

// Plans the tensor-core depthwise conv; returns false when the shape is outside what the kernel covers.
bool qasr_dwconv_tc_supported(int B, int C, int T_in, int T_out, int k, int stride, int dilation, int pad) {
  (void)C; (void)T_in;
  if (stride < 1 || stride > 2 || B < 1 || T_out < 1) return false;
  const int span = (k - 1) * dilation + 1, o = ((-pad) % 16 + 16) % 16;
  const int window = stride * 31 + span + o;   // smallest tile (N = 32)
  return ceil_div(window, 128) <= kMaxSlabs;
}

int qasr_dwconv_tc_launch(const void* x, int x_ld, const int8_t* taps, void* out, int out_ld, int B, int C, int T_in,
                          int T_out, int k, int stride, int dilation, int pad, const qasr_epilogue* epi, void* stream) {
  DwTcArgs a;
  memset(&a, 0, sizeof(a));
  a.taps = taps; a.out = out; a.out_ld = out_ld; a.B = B; a.C = C; a.T_in = T_in; a.T_out = T_out;
  a.k = k; a.stride = stride; a.dil = dilation; a.pad = pad;
  a.o = ((-pad) % 16 + 16) % 16;
  a.span = (k - 1) * dilation + 1;
  a.zl = ((a.span + 32 + 15) / 16) * 16;
  const uint32_t fixed = 1024 /*align*/ + 4u * (uint32_t)a.zl + 8u * (2 * kMaxStages + 12);
  // time-block length.  Row-staged variant (knob 11 = 1): the epilogue warps are the critical resource, so take the
  // N with the fewest drain rounds (time blocks x passes of the 5 column groups) among those with <= 2 slabs whose
  // buffers fit; TMA-window variant: least padded work (fetched window bytes x MMA depth).
  const bool want_rows = qasr_debug_get(11) == 1;
  const int bbox_plan = B >= 128 ? 128 : (B > 32 ? 64 : (B > 16 ? 32 : (B > 8 ? 16 : 8)));
  int bestN = 0, best_cost = 0;
  bool rows = false;
  if (want_rows) {
    for (int N : {160, 128, 96, 64, 32}) {
      const int slabs = ceil_div(stride * (N - 1) + a.span + a.o, 128);
      if (slabs > 2) continue;
      const int n_tblk = ceil_div(T_out, N);
      const int pitch = (128 / bbox_plan - 1) * N * stride + slabs * 128;   // one row's piece
      if (fixed + 2u * slabs * N * 128u + 2u * slabs * kSlabBytes + 2u * (uint32_t)bbox_plan * pitch > kSmemBudget) continue;
      const int cost = n_tblk * ceil_div(N / 32, kColGroups);
      if (bestN == 0 || cost < best_cost) { bestN = N; best_cost = cost; rows = true; }
    }
  }
  if (!rows) {
    for (int N : {160, 128, 96, 64, 32}) {
      const int slabs = ceil_div(stride * (N - 1) + a.span + a.o, 128);
      if (slabs > kMaxSlabs) continue;
      if (fixed + 2u * slabs * N * 128u + 2u * slabs * kSlabBytes > kSmemBudget) continue;
      const int cost = ceil_div(T_out, N) * ceil_div(N / 32, kColGroups) * slabs;   // ~ epilogue rounds x MMA depth
      if (bestN == 0 || cost < best_cost) { bestN = N; best_cost = cost; }
    }
  }
  if (bestN == 0) {
    qasr_set_error("dwconv(tc): tap span %d does not fit the %d x 128-byte window", a.span, kMaxSlabs);
    return QASR_ERR_UNSUPPORTED;
  }
  a.N = bestN;
  a.slabs = ceil_div(stride * (a.N - 1) + a.span + a.o, 128);
  a.nb = (a.span + 14) / 16 + 1;
  a.bbox = bbox_plan;
  a.tb_per_tile = 128 / a.bbox;
  a.n_tblk = ceil_div(T_out, a.N);
  a.ntbg = ceil_div(a.n_tblk, a.tb_per_tile);
  a.tiles_per_ch = ceil_div(B, a.bbox) * a.ntbg;
  a.units = C * a.tiles_per_ch;
  a.tmat_bytes = (uint32_t)a.slabs * a.N * 128u;
  a.stage_bytes = (uint32_t)a.slabs * kSlabBytes;
  // row-staged variant (bring-up knob 11 = 1): two piece buffers next to two A stages.  Bit-exact, but measured slower
  // than TMA windows (34.7 us against 22.2 us, k = 33, 512 channels, float epilogue): 64 bulk copies of ~416 B per
  // unit cost ~47 clk each on the SM's copy engine (16 us per layer), more than the 6.5 clk per 128-byte TMA row.
  a.x = static_cast<const uint8_t*>(x);
  a.x_ld = x_ld;
  a.copy_bytes = ((T_in + 15) / 16) * 16;
  a.row_pitch = (a.tb_per_tile - 1) * a.N * stride + a.slabs * 128;
  a.rows_bytes = (uint32_t)a.bbox * (uint32_t)a.row_pitch;
  if (rows) {
    a.stages = 2;
  } else {
    a.rows_bytes = 0;
    a.stages = (int)((kSmemBudget - fixed - 2u * a.tmat_bytes) / a.stage_bytes);
    if (a.stages > kMaxStages) a.stages = kMaxStages;
  }
  const uint32_t smem_bytes = fixed + 2u * a.tmat_bytes + (uint32_t)a.stages * a.stage_bytes + 2u * a.rows_bytes;
  const uint32_t a_signed = (epi->flags & QASR_F_B_SIGNED) ? 1u : 0u;   // the activation codes are operand A here
  // instruction descriptor: D=S32, A=(un)signed 8-bit K-major (codes), B=signed 8-bit K-major (Tc), N, M=128
  a.idesc = (2u << 4) | (a_signed << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(a.N >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  a.ep = make_epi_args(epi);
  a.stamp = qasr_next_stamp();
  a.dbg_skip = qasr_debug_get(6) > 0 ? (int)qasr_debug_get(6) : 0;
  a.dbg_trace = qasr_debug_get(7) > 0 ? reinterpret_cast<unsigned long long*>(qasr_debug_get(7)) : nullptr;
  CUtensorMap map;
  int rc = make_map_rows(&map, x, T_in, C, B, x_ld, a.bbox);
  if (rc) return rc;

  const bool fast = !(epi->flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64 | QASR_F_STAGE2)) &&
                    epi->prev == nullptr;
  const int ek = epi->mode != QASR_EPI_REQUANT_I8 ? 2 : ((epi->flags & QASR_F_FAST_F32) ? 3 : (fast ? 0 : 1));
  int grid = qasr_sm_count();
  if (grid > a.units) grid = a.units;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[2][4][64] = {};
  int dev = 0;
  QASR_CHECK_CUDA(cudaGetDevice(&dev));
  const bool need_attr = dev >= 0 && dev < 64 && !attr_set[rows][ek][dev];
#define QASR_DWTC_GO(EK_, ROWS_)                                                                                   \
  do {                                                                                                             \
    if (need_attr)                                                                                                 \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(dwconv_tc_kernel<EK_, ROWS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)kSmemBudget));                                                     \
    QASR_CHECK_CUDA(qasr_launch_pdl(dwconv_tc_kernel<EK_, ROWS_>, dim3(grid), dim3(kThreads), smem_bytes, st, map, a)); \
  } while (0)
#define QASR_DWTC_CASE(EK_) \
  case EK_:                 \
    if (rows) QASR_DWTC_GO(EK_, true); else QASR_DWTC_GO(EK_, false); \
    break;
  switch (ek) {
    QASR_DWTC_CASE(0)
    QASR_DWTC_CASE(1)
    QASR_DWTC_CASE(2)
    QASR_DWTC_CASE(3)
  }
#undef QASR_DWTC_CASE
#undef QASR_DWTC_GO
  if (need_attr) attr_set[rows][ek][dev] = true;
  return qasr_check_launch("dwconv_tc_kernel");
}

// K1-rt: the IMMA depthwise engine over ROW TAPES (stride 1, requantising fast epilogues).
// Same arithmetic and the same banded-Toeplitz warp MMA as dwconv.cu (F.conv1d(x_int, w_int, groups=C) of
// QuantConv1d.int_conv, nemo/quantization/utils/quant_modules.py:301-305, in true integers); what changes is the walk.
//
// dwconv.cu handles one utterance row at a time.  A row of 800 frames is 6.25 tiles of 128 outputs, so its branch-free
// 4-tile loop body runs ONCE per row and the rest of the row goes through the pair, the masked single tile and the row
// bookkeeping: 356 warp instructions per row where 206 would do at the loop body's rate (SASS of the k = 63 kernel).
// Here the R rows of a warp's work item (one channel, R utterances) are copied back to back into ONE staging buffer,
//     [halo | row 0 | gap | row 1 | gap | ... | row R-1 | zeros],        gap = zeros >= the conv padding,
// and the conv runs over that tape as if it were one signal: tile q covers tape outputs [q, q + 128) whatever row they
// belong to.  Outputs that fall into a gap are computed and dropped (4 % of the tape at 800 frames, against 12 % of
// padding in the seventh tile of every row before).  Tiles that lie inside one row's valid frames take the fast path
// (constant pointer steps, unmasked 128-byte stores); the one tile per row that straddles the row end, the gap and the
// first frames of the next row takes a generic path with per-lane (row, frame) resolution, length mask and zero fill.
// All rows are requested up front (one bulk copy each, one mbarrier each); the warp waits for a row when it first needs it.
#include <algorithm>
#include <type_traits>

#include "qasr_common.cuh"
#include "dw_common.cuh"

namespace {

constexpr int kMaxR = 8;   // rows per tape (mbarriers per warp)

// EK: 0 = integer fast path, float fast path for the channels of the f32_mult table; 3 = float fast path everywhere;
//     4 = device-written tables (QASR_F_DEVICE_TABLES, dynamic quantisation): integer fast path for the channels
//     qasr_dyn_finalize proved, literal recipe for the others
#ifndef QASR_RT_MIN_BLOCKS
#define QASR_RT_MIN_BLOCKS 1
#endif
template <int KS, bool XS, int EK, bool PROF>
__global__ void __launch_bounds__(kThreads, QASR_RT_MIN_BLOCKS) dwconv_rt_kernel(const DwArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(512) uint32_t stage[kWarps][128];   // four 128-byte transpose lines per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  pdl_trigger();
  const int item = blockIdx.x * kWarps + warp;
  if (item >= a.C * a.nbg) return;   // whole warps exit together; only warp-level sync below
  const int c = item / a.nbg, bg = item - c * a.nbg;
  const int b_begin = bg * a.bchunk;
  const int nrows = min(a.B, b_begin + a.bchunk) - b_begin;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * kMaxR;
  uint8_t* tape = smem + kWarps * kMaxR * 8 + (size_t)warp * a.buf_bytes;
  for (int i = lane; i < a.buf_bytes / 16; i += 32) reinterpret_cast<uint4*>(tape)[i] = make_uint4(0, 0, 0, 0);
  if (lane == 0) {
    for (int i = 0; i < nrows; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + i)));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zero fill (generic proxy) before the async copies
  __syncwarp();
  // banded A fragments of this channel, channel record: constants of the plan, loaded before the dependency wait
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
  const bool f32_path = EK == 3 || c0.r != 0.0f;   // warp-uniform: one channel per work item
  const EpiConsts kc = make_epi_consts(a.ep, false);
  const int P = a.period;
  pdl_wait();   // activations (and the output buffer) belong to the stream order; everything above is set-up
  if (lane == 0) {
    const unsigned long long t_begin = (PROF && a.stamp != nullptr) ? qasr_globaltimer() : ~0ull;
    const uint8_t* src = a.x + ((size_t)b_begin * a.C + c) * a.x_ld;
    const size_t xstep = (size_t)a.C * a.x_ld;
    for (int r = 0; r < nrows; ++r, src += xstep)
      row_copy_async(smem_addr(tape + a.halo + r * P), src, (uint32_t)a.copy_bytes, smem_addr(bars + r));
    // in-graph timing: 8192 warps hitting ONE address with atomics at the same moment cost the stamped kernel 3.6 us
    // (13.1 against 9.5 us on the 256-channel layers); the first CTA on data is one of the first the grid dispatched
    if (PROF && warp == 0 && blockIdx.x < 296) atomicMin(a.stamp, t_begin);
  }

  const int o0 = 32 * tg + g;   // this lane's accumulators inside a tile: o0, o0 + 16, o0 + 8, o0 + 24
  uint8_t* const sb = reinterpret_cast<uint8_t*>(stage[warp]) + o0;
  const uint32_t* const sw = stage[warp] + lane;
  const uint32_t* const xw = reinterpret_cast<const uint32_t*>(tape) + ((a.halo - a.pad - a.d) >> 2) + 4 * g + 2 * KS * tg;
  const bool masked = a.ep.flags & QASR_F_MASK;
  const int T_out = a.T_out;
  const int ow4 = ((T_out + 15) >> 4) << 4;   // stores cover T_out rounded up to 16 (zero filled)
  const size_t out_row_step = (size_t)a.C * a.out_ld;
  uint8_t* out_row = reinterpret_cast<uint8_t*>(a.out) + ((size_t)b_begin * a.C + c) * a.out_ld;

  auto walk = [&](auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;   // 0 = integer fast path, 1 = float fast path, 2 = literal recipe
    constexpr bool F32 = MODE == 1;
    // float path: the MMA's initial accumulator is 0x4B400000 + bias (see dwconv.cu)
    const int acc_init = F32 ? kF32MagicBits + c0.bias : 0;
    auto code = [&](int v) -> uint8_t {
      if (F32) return (uint8_t)epilogue_f32_bits_biased(kc, v, c0);
      if (MODE == 2) return (uint8_t)epilogue_exact(a.ep, kc, v, c0, false, 0, c0, 0);
      return (uint8_t)epilogue_fast<false, false, false>(a.ep, kc, v, c0, 0, c0, 0);
    };
    auto mma_one = [&](const uint32_t* tw0, int (&acc)[4]) {
      const uint2* p0 = reinterpret_cast<const uint2*>(tw0);
      uint2 b0 = p0[0];
      imma_16832_init<XS>(acc, af[0], b0.x, b0.y, acc_init);
#pragma unroll
      for (int ks = 1; ks < KS; ++ks) {
        b0 = p0[ks];
        imma_16832<XS>(acc, af[ks], b0.x, b0.y);
      }
    };
    auto mma_pair = [&](const uint32_t* tw0, int (&acc0)[4], int (&acc1)[4]) {
      const uint2* p0 = reinterpret_cast<const uint2*>(tw0);
      const uint2* p1 = reinterpret_cast<const uint2*>(tw0 + (kTile >> 2));
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
    // two whole tiles inside one row: codes -> transpose through lines LINE, LINE + 1 -> two unmasked 128-byte stores
    auto pair_fast = [&](auto line_tag, const uint32_t* twp, uint32_t* owp) {
      constexpr int LINE = decltype(line_tag)::value;
      int acc0[4], acc1[4];
      mma_pair(twp, acc0, acc1);
      uint8_t* const l0 = sb + 128 * LINE;
      l0[0] = code(acc0[0]); l0[16] = code(acc0[1]); l0[8] = code(acc0[2]); l0[24] = code(acc0[3]);
      l0[128] = code(acc1[0]); l0[144] = code(acc1[1]); l0[136] = code(acc1[2]); l0[152] = code(acc1[3]);
      __syncwarp();
      const uint32_t w0 = sw[32 * LINE], w1 = sw[32 * LINE + 32];
      owp[0] = w0;
      owp[kTile >> 2] = w1;
    };

    int q = 0;   // tape output position of the next tile (multiple of 128)
    int len_r = masked ? min(__ldg(a.ep.out_lens + b_begin), T_out) : T_out;
    bar_wait(smem_addr(bars), 0);
    for (int r = 0, base = 0; r < nrows; ++r, base += P, out_row += out_row_step) {
      const bool has_next = r + 1 < nrows;
      const int len_n = has_next ? (masked ? min(__ldg(a.ep.out_lens + b_begin + r + 1), T_out) : T_out) : 0;
      // ---- tiles inside this row's valid frames (base <= q < base + 128 here: the frames below q were stored by the
      //      previous row's last tile)
      const int fe = base + len_r;
      const uint32_t* tw = xw + (q >> 2);
      uint32_t* ow = reinterpret_cast<uint32_t*>(out_row) + ((q - base) >> 2) + lane;
      for (; q + 4 * kTile <= fe; q += 4 * kTile, tw += kTile, ow += kTile) {
        // lines 0-1, then 2-3: each pair's barrier orders the other pair's reads before the writes that reuse its lines
        pair_fast(std::integral_constant<int, 0>{}, tw, ow);
        pair_fast(std::integral_constant<int, 2>{}, tw + (kTile >> 1), ow + (kTile >> 1));
      }
      if (q + 2 * kTile <= fe) {
        pair_fast(std::integral_constant<int, 0>{}, tw, ow);
        __syncwarp();   // steps outside the 4-tile loop leave no reads of the staging lines pending
        q += 2 * kTile; tw += kTile >> 1; ow += kTile >> 1;
      }
      if (q + kTile <= fe) {
        int acc[4];
        mma_one(tw, acc);
        sb[0] = code(acc[0]); sb[16] = code(acc[1]); sb[8] = code(acc[2]); sb[24] = code(acc[3]);
        __syncwarp();
        ow[0] = sw[0];
        __syncwarp();
        q += kTile;
      }
      // ---- tiles that hold the row end, the gap and the first frames of the next row (P >= 128: at most two rows)
      if (has_next) bar_wait(smem_addr(bars + r + 1), 0);
      const int stop = has_next ? base + P : base + ow4;
      for (; q < stop; q += kTile) {
        int acc[4];
        mma_one(xw + (q >> 2), acc);
        sb[0] = code(acc[0]); sb[16] = code(acc[1]); sb[8] = code(acc[2]); sb[24] = code(acc[3]);
        __syncwarp();
        uint32_t word = sw[0];
        __syncwarp();
        const int p = q + 4 * lane - base;   // >= 0
        const bool nx = p >= P;
        const int f = nx ? p - P : p;
        const int nvalid = (nx ? len_n : len_r) - f;
        if (nvalid < 4) word &= nvalid <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nvalid)));
        uint8_t* dst = (nx ? out_row + out_row_step : out_row) + f;
        if (f < ow4 && (!nx || has_next)) *reinterpret_cast<uint32_t*>(dst) = word;
      }
      len_r = len_n;
    }
  };
  if (EK == 4) {
    if (c0.literal) walk(std::integral_constant<int, 2>{});
    else walk(std::integral_constant<int, 0>{});
  } else if (f32_path) {
    walk(std::integral_constant<int, 1>{});
  } else {
    walk(std::integral_constant<int, 0>{});
  }
  if (PROF && lane == 0) qasr_stamp_end(a.stamp);   // spread over the kernel's tail: no burst on the address
}

template <int KS>
int launch_rt(const DwArgs& a, int ek, cudaStream_t stream) {
  const int items = a.C * a.nbg;
  const int grid = (items + kWarps - 1) / kWarps;
  const size_t smem = kWarps * kMaxR * 8 + (size_t)kWarps * a.buf_bytes;
  const bool xs = a.ep.flags & QASR_F_B_SIGNED;
  const bool prof = a.stamp != nullptr;
#define QASR_RT_GO(XS_, EK_)                                                                             \
  do {                                                                                                   \
    auto kern = prof ? dwconv_rt_kernel<KS, XS_, EK_, true> : dwconv_rt_kernel<KS, XS_, EK_, false>;     \
    static size_t attr_smem[2] = {0, 0};                                                                 \
    if (smem > 48 * 1024 && smem > attr_smem[prof]) {                                                    \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr_smem[prof] = smem;                                                                            \
    }                                                                                                    \
    QASR_CHECK_CUDA(qasr_launch_pdl(kern, dim3(grid), dim3(kThreads), smem, stream, a));                 \
  } while (0)
  if (xs) {
    if (ek == 0) QASR_RT_GO(true, 0); else if (ek == 3) QASR_RT_GO(true, 3); else QASR_RT_GO(true, 4);
  } else {
    if (ek == 0) QASR_RT_GO(false, 0); else if (ek == 3) QASR_RT_GO(false, 3); else QASR_RT_GO(false, 4);
  }
#undef QASR_RT_GO
  return qasr_check_launch("dwconv_rt_kernel");
}

}  // namespace

// Row-tape geometry for a stride-1 launch; false when the shape stays on the row-at-a-time kernel.
bool qasr_dwconv_rt_supported(int T_in, int k, int dilation, int pad) {
  const int ksteps = dw_ksteps(k, 1, dilation, pad);
  if (ksteps > 6) return false;
  const int gap = ((std::max(pad, 1) + 15) / 16) * 16;
  return gap + ((T_in + 15) / 16) * 16 >= kTile;   // a tile touches at most two rows
}

int qasr_dwconv_rt_launch(const void* x, int x_ld, const int8_t* w, void* out, int out_ld, int B, int C, int T_in, int T_out,
                          int k, int dilation, int pad, int ek, const qasr_epilogue* epi, void* stream) {
  DwArgs a;
  memset(&a, 0, sizeof(a));
  a.x = static_cast<const uint8_t*>(x);
  a.w = w;
  a.out = out;
  a.x_ld = x_ld; a.out_ld = out_ld; a.B = B; a.C = C; a.T_in = T_in; a.T_out = T_out;
  a.k = k; a.stride = 1; a.dil = dilation; a.pad = pad;
  a.d = dw_misalign(pad);
  const int ksteps = dw_ksteps(k, 1, dilation, pad);
  a.halo = ((pad + a.d + 15) / 16) * 16;
  a.copy_bytes = ((T_in + 15) / 16) * 16;
  // right overhang of a row's last valid output = pad bytes past T_in (T_out = T_in + 2 pad - (k - 1) dil), left overhang of
  // the next row = pad: one gap of >= pad zeros serves both
  a.period = ((std::max(pad, 1) + 15) / 16) * 16 + a.copy_bytes;
  // rows per tape: the SM's 64 warp slots filled (see dwconv.cu), at most kMaxR
  const long long per_sm = qasr_debug_get(10) > 0 ? qasr_debug_get(10) : 16LL * kWarps;
  const long long want_items = per_sm * qasr_sm_count();
  int bchunk = (int)(((long long)B * C + want_items - 1) / want_items);
  if (bchunk < 1) bchunk = 1;
  if (qasr_debug_get(17) > 0) bchunk = (int)qasr_debug_get(17);   // bring-up knob 17: rows per tape
  if (bchunk > kMaxR) bchunk = kMaxR;
  if (bchunk > B) bchunk = B;
  a.bchunk = bchunk;
  a.nbg = (B + bchunk - 1) / bchunk;
  const int ow4 = ((T_out + 15) / 16) * 16;
  const int n_tiles = ((bchunk - 1) * a.period + ow4 + kTile - 1) / kTile;
  const int reach = a.halo - pad - a.d + n_tiles * kTile + 32 * ksteps + 16;   // bytes read at most
  int buf = a.halo + bchunk * a.period;
  if (buf < reach) buf = reach;
  a.buf_bytes = ((buf + 127) / 128) * 128;
  const size_t smem = kWarps * kMaxR * 8 + (size_t)kWarps * a.buf_bytes;
  if (smem > 200 * 1024) {
    qasr_set_error("dwconv(row tape): %d rows of %d frames need %zu bytes of staging shared memory", bchunk, T_in, smem);
    return QASR_ERR_UNSUPPORTED;
  }
  a.ep = make_epi_args(epi);
  a.stamp = qasr_next_stamp();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define QASR_RT_CASE(KS_) if (ksteps == KS_) return launch_rt<KS_>(a, ek, st);
  QASR_RT_CASE(1)
  QASR_RT_CASE(2)
  QASR_RT_CASE(3)
  QASR_RT_CASE(4)
  QASR_RT_CASE(5)
  QASR_RT_CASE(6)
#undef QASR_RT_CASE
  return QASR_ERR_UNSUPPORTED;
}

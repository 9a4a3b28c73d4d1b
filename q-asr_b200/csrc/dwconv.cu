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
//   K = 32 * KS covers 15 S + d + (k-1) dil + 1 taps; d = (-pad) mod 4 absorbs the misalignment.
//   * A fragments (KS x 4 registers) are built once per warp from the channel's k int8 taps and
//     reused for every tile of that channel (several utterances per warp).
//   * B fragments are plain aligned 32-bit words of the input row: lane L of the warp needs words
//     (P/4 + 4 S (L>>2) + (L&3) + 8 ks [+4]) -> for S = 1 the 32 lanes read 128 contiguous bytes,
//     fully coalesced; neighbouring k-steps overlap in L1.  No shared-memory staging of x at all.
//   * D (int32, 4 per lane) goes through the fused epilogue, the 4 result bytes are transposed
//     through a 128-byte per-warp shared-memory line so the warp stores 128 contiguous bytes.
// mma.sync m16n8k32 (IMMA) was measured at 0.48 instr/clk/SM (1.14 POPS) on B200: the MMA part of
// a 512-channel layer at batch 64 x 800 frames is ~4.6 us, below its 8 us HBM time.
#include "qasr_common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kThreads = 32 * kWarps;
constexpr int kTile = 128;  // outputs per warp tile

struct DwArgs {
  const uint8_t* x;
  const int8_t* w;
  void* out;
  int x_ld, out_ld, B, C, T_in, T_out, k, stride, dil, pad;
  int d;        // (-pad) mod 4
  int bchunk;   // utterances per warp work item
  int nbg;      // ceil(B / bchunk)
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

template <int KS, int S, bool XS, bool FAST>
__global__ void __launch_bounds__(kThreads) dwconv_mma_kernel(const DwArgs a) {
  __shared__ uint32_t stage[kWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int item = blockIdx.x * kWarps + warp;
  if (item >= a.C * a.nbg) return;  // whole warps exit together; only __syncwarp below
  const int c = item / a.nbg, bg = item - c * a.nbg;

  // ---- banded A fragments of this channel (registers, reused for every tile)
  uint32_t af[KS][4];
  {
    const int8_t* wr = a.w + (size_t)c * a.k;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = g + 8 * (r & 1);
        const int kk0 = 32 * ks + 16 * (r >> 1) + 4 * tg;
        uint32_t word = 0;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const int u = kk0 + bb - S * row - a.d;
          int v = 0;
          if (u >= 0) {
            const int tap = u / a.dil;
            if (tap * a.dil == u && tap < a.k) v = __ldg(wr + tap);
          }
          word |= (uint32_t)(v & 0xff) << (8 * bb);
        }
        af[ks][r] = word;
      }
    }
  }
  const ChanRegs c0 = load_chan(a.ep.ch0, c);
  const EpiConsts kc = make_epi_consts(a.ep, false);
  // offsets of this lane's 4 accumulators inside the 128-frame tile: D[row][col] -> 16 col + row
  const int o0 = 32 * tg + g, o1 = o0 + 16, o2 = o0 + 8, o3 = o0 + 24;
  const int woff = 4 * S * g + tg;  // word offset of b0 inside the tile window
  const int b_end = min(a.B, (bg + 1) * a.bchunk);
  const int nwords_in = (a.T_in + 3) >> 2;
  const int out_words = ((a.T_out + 15) >> 4) << 2;  // stores cover T_out rounded up to 16 (zero filled)

  for (int b = bg * a.bchunk; b < b_end; ++b) {
    const size_t row = (size_t)b * a.C + c;
    const uint32_t* xrow = reinterpret_cast<const uint32_t*>(a.x + row * a.x_ld);
    const int len = (a.ep.flags & QASR_F_MASK) ? min(a.ep.out_lens[b], a.T_out) : a.T_out;
    for (int t0 = 0; t0 < a.T_out; t0 += kTile) {
      const int pw = (t0 * S - a.pad - a.d) >> 2;  // exact: the numerator is a multiple of 4
      int acc[4] = {0, 0, 0, 0};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int w0 = pw + woff + 8 * ks, w1 = w0 + 4;
        const uint32_t b0 = (w0 >= 0 && w0 < nwords_in) ? __ldg(xrow + w0) : 0u;
        const uint32_t b1 = (w1 >= 0 && w1 < nwords_in) ? __ldg(xrow + w1) : 0u;
        imma_16832<XS>(acc, af[ks], b0, b1);
      }
      if (a.ep.mode == QASR_EPI_REQUANT_I8) {
        const int q0 = epilogue_requant<FAST>(a.ep, kc, acc[0], c0, false, 0, c0, 0, t0 + o0 < len);
        const int q1 = epilogue_requant<FAST>(a.ep, kc, acc[1], c0, false, 0, c0, 0, t0 + o1 < len);
        const int q2 = epilogue_requant<FAST>(a.ep, kc, acc[2], c0, false, 0, c0, 0, t0 + o2 < len);
        const int q3 = epilogue_requant<FAST>(a.ep, kc, acc[3], c0, false, 0, c0, 0, t0 + o3 < len);
        uint8_t* sb = reinterpret_cast<uint8_t*>(stage[warp]);
        sb[o0] = (uint8_t)q0;
        sb[o1] = (uint8_t)q1;
        sb[o2] = (uint8_t)q2;
        sb[o3] = (uint8_t)q3;
        __syncwarp();
        const uint32_t word = stage[warp][lane];
        __syncwarp();
        const int wi = (t0 >> 2) + lane;
        if (wi < out_words)
          reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(a.out) + row * a.out_ld)[wi] = word;
      } else if (a.ep.mode == QASR_EPI_DEQUANT_F32) {
        float* orow = reinterpret_cast<float*>(a.out) + row * a.out_ld;
        if (t0 + o0 < a.T_out) orow[t0 + o0] = epilogue_dequant(a.ep, acc[0], c0);
        if (t0 + o1 < a.T_out) orow[t0 + o1] = epilogue_dequant(a.ep, acc[1], c0);
        if (t0 + o2 < a.T_out) orow[t0 + o2] = epilogue_dequant(a.ep, acc[2], c0);
        if (t0 + o3 < a.T_out) orow[t0 + o3] = epilogue_dequant(a.ep, acc[3], c0);
      } else {
        int32_t* orow = reinterpret_cast<int32_t*>(a.out) + row * a.out_ld;
        if (t0 + o0 < a.T_out) orow[t0 + o0] = acc[0] + c0.bias;
        if (t0 + o1 < a.T_out) orow[t0 + o1] = acc[1] + c0.bias;
        if (t0 + o2 < a.T_out) orow[t0 + o2] = acc[2] + c0.bias;
        if (t0 + o3 < a.T_out) orow[t0 + o3] = acc[3] + c0.bias;
      }
    }
  }
}

template <int KS, int S>
int launch(const DwArgs& a, cudaStream_t stream) {
  const int items = a.C * a.nbg;
  const int grid = (items + kWarps - 1) / kWarps;
  const bool xs = a.ep.flags & QASR_F_B_SIGNED;
  const bool fast = !(a.ep.flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64));
  if (xs && fast) dwconv_mma_kernel<KS, S, true, true><<<grid, kThreads, 0, stream>>>(a);
  else if (xs) dwconv_mma_kernel<KS, S, true, false><<<grid, kThreads, 0, stream>>>(a);
  else if (fast) dwconv_mma_kernel<KS, S, false, true><<<grid, kThreads, 0, stream>>>(a);
  else dwconv_mma_kernel<KS, S, false, false><<<grid, kThreads, 0, stream>>>(a);
  return qasr_check_launch("dwconv_mma_kernel");
}

}  // namespace

extern "C" int qasr_dwconv1d_i8(const void* x, int32_t x_ld, const int8_t* w, void* out, int32_t out_ld,
                                int32_t B, int32_t C, int32_t T_in, int32_t T_out, int32_t k, int32_t stride,
                                int32_t dilation, int32_t pad, const qasr_epilogue* epi, void* stream) {
  QASR_CHECK_ARG(x && w && out && epi, "dwconv: null pointer");
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

  DwArgs a;
  a.x = static_cast<const uint8_t*>(x);
  a.w = w;
  a.out = out;
  a.x_ld = x_ld; a.out_ld = out_ld; a.B = B; a.C = C; a.T_in = T_in; a.T_out = T_out;
  a.k = k; a.stride = stride; a.dil = dilation; a.pad = pad;
  a.d = ((-pad) % 4 + 4) % 4;
  // utterances per warp: amortise the A-fragment set-up while keeping >= ~8 warps per SM in flight
  const long long want_items = 8LL * kWarps * qasr_sm_count();
  int bchunk = (int)(((long long)B * C + want_items - 1) / want_items);
  if (bchunk < 1) bchunk = 1;
  if (bchunk > 16) bchunk = 16;
  if (bchunk > B) bchunk = B;
  a.bchunk = bchunk;
  a.nbg = (B + bchunk - 1) / bchunk;
  a.ep = make_epi_args(epi);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int span = (k - 1) * dilation + 1;
  const int ksteps = (15 * stride + a.d + span + 31) / 32;
#define QASR_DW_CASE(KS_)                                      \
  if (ksteps == KS_) return stride == 1 ? launch<KS_, 1>(a, st) : launch<KS_, 2>(a, st);
  QASR_DW_CASE(1)
  QASR_DW_CASE(2)
  QASR_DW_CASE(3)
  QASR_DW_CASE(4)
  QASR_DW_CASE(5)
  QASR_DW_CASE(6)
#undef QASR_DW_CASE
  qasr_set_error("dwconv: tap span %d (k=%d, dilation=%d, stride=%d) exceeds the supported 6 x 32 window", span,
                 k, dilation, stride);
  return QASR_ERR_UNSUPPORTED;
}

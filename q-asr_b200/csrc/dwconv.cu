// K1: depthwise 1-D convolution on 8-bit activation codes with the fused requantisation
// epilogue.  Restates F.conv1d(x_int.double(), w_int.double(), groups=C) of the reference's
// QuantConv1d.int_conv (nemo/quantization/utils/quant_modules.py:301-305) in true integers.
//
// Layout / algorithm (HBM-bound by design, dp4a-issue-bound for the widest kernels):
//   * x is [B][C][x_ld] with time contiguous.  One thread produces one 16-byte output chunk
//     (16 consecutive frames of one channel); a warp therefore writes 512 contiguous bytes.
//   * Work items (channel, chunk) are flattened inside one utterance so that short rows
//     (T' = 250 ... 1008) still fill whole warps.
//   * The k taps of a channel are pre-arranged, per CTA in shared memory, as NW int32 words
//     for each of the 4 byte phases an output can have inside an aligned input word.
//     Output o then is  sum_n dp4a(xw[(o*stride >> 2) + n], W[n][(o*stride) & 3])  on ALIGNED
//     input words: no byte shuffles, int32 accumulation.  Dilation and the (-pad mod 4)
//     misalignment are folded into that table (zero bytes).
//   * The input window (15*stride/4 + NW + 1 words) lives in registers; loads are read-only
//     32-bit words, neighbouring threads overlap in L1.
#include "qasr_common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kChunk = 16;  // outputs per thread

struct DwArgs {
  const uint8_t* x;
  const int8_t* w;
  void* out;
  int x_ld, out_ld, B, C, T_in, T_out, k, stride, dil, pad;
  int nchunk;      // chunks per row
  int nw;          // tap words actually needed (<= NW)
  int d;           // (-pad) mod 4 folded into the tap table
  EpiArgs ep;
};

template <bool XS>
__device__ __forceinline__ int dp4a_mixed(uint32_t x, uint32_t w, int acc) {
  int d;
  if (XS) asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(w), "r"(acc));
  else    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(w), "r"(acc));
  return d;
}

// NW: tap words per phase (compile time, table zero padded), STRIDE: 1 or 2.
template <int NW, int STRIDE, bool XS, bool FAST>
__global__ void __launch_bounds__(kThreads) dwconv_kernel(const DwArgs a) {
  constexpr int WIN = ((kChunk - 1) * STRIDE) / 4 + NW;  // highest word index + 1
  extern __shared__ uint32_t wtab[];                      // [rows][NW][4]

  const int b = blockIdx.y;
  const int item0 = blockIdx.x * kThreads;
  const int items = a.C * a.nchunk;
  const int row0 = item0 / a.nchunk;
  const int row_last = min(items - 1, item0 + kThreads - 1) / a.nchunk;
  const int nrows = row_last - row0 + 1;

  // ---- tap table: word (n, r) byte bb <-> tap j with j*dil = 4n + bb - r - d
  for (int i = threadIdx.x; i < nrows * NW * 4; i += kThreads) {
    const int r = i & 3, n = (i >> 2) % NW, row = (i >> 2) / NW;
    const int8_t* wr = a.w + (size_t)(row0 + row) * a.k;
    uint32_t word = 0;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int u = 4 * n + bb - r - a.d;
      int v = 0;
      if (u >= 0 && (u % a.dil) == 0 && (u / a.dil) < a.k) v = wr[u / a.dil];
      word |= (uint32_t)(v & 0xff) << (8 * bb);
    }
    wtab[i] = word;
  }
  __syncthreads();

  const int item = item0 + threadIdx.x;
  if (item >= items) return;
  const int c = item / a.nchunk;
  const int chunk = item - c * a.nchunk;
  const int t0 = chunk * kChunk;

  // ---- input window: aligned words starting at position p0 (multiple of 4)
  const uint8_t* xrow = a.x + ((size_t)b * a.C + c) * a.x_ld;
  const int p0 = t0 * STRIDE - a.pad - a.d;
  uint32_t xw[WIN];
#pragma unroll
  for (int i = 0; i < WIN; ++i) {
    const int p = p0 + 4 * i;
    xw[i] = (p >= 0 && p < a.T_in) ? __ldg(reinterpret_cast<const uint32_t*>(xrow + p)) : 0u;
  }

  int acc[kChunk];
#pragma unroll
  for (int o = 0; o < kChunk; ++o) acc[o] = 0;
  const uint4* wt = reinterpret_cast<const uint4*>(wtab) + (size_t)(c - row0) * NW;
#pragma unroll
  for (int n = 0; n < NW; ++n) {
    const uint4 wv = wt[n];
    const uint32_t wph[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int o = 0; o < kChunk; ++o) {
      acc[o] = dp4a_mixed<XS>(xw[((o * STRIDE) >> 2) + n], wph[(o * STRIDE) & 3], acc[o]);
    }
  }

  // ---- epilogue
  const ChanRegs c0 = load_chan(a.ep.ch0, c);
  const int len = (a.ep.flags & QASR_F_MASK) ? a.ep.out_lens[b] : a.T_out;
  if (a.ep.mode == QASR_EPI_REQUANT_I8) {
    const double mult2 = dyad_mult(a.ep.m2, a.ep.e2);
    const long long half2 = (a.ep.e2 >= 1) ? (1ll << (a.ep.e2 - 1)) : 0;
    uint32_t packed[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      int q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + 4 * g + i;
        int prev = 0;
        if (a.ep.prev != nullptr && t < a.T_out)
          prev = a.ep.prev[((size_t)b * a.C + c) * a.ep.prev_ld + t];
        q[i] = epilogue_requant<FAST>(a.ep, acc[4 * g + i], c0, false, 0, c0, prev, mult2, half2,
                                      t < len && t < a.T_out);
      }
      packed[g] = pack4(q[0], q[1], q[2], q[3]);
    }
    uint8_t* orow = reinterpret_cast<uint8_t*>(a.out) + ((size_t)b * a.C + c) * a.out_ld;
    *reinterpret_cast<uint4*>(orow + t0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  } else if (a.ep.mode == QASR_EPI_DEQUANT_F32) {
    float* orow = reinterpret_cast<float*>(a.out) + ((size_t)b * a.C + c) * a.out_ld;
#pragma unroll
    for (int o = 0; o < kChunk; ++o)
      if (t0 + o < a.T_out) orow[t0 + o] = epilogue_dequant(a.ep, acc[o], c0);
  } else {
    int32_t* orow = reinterpret_cast<int32_t*>(a.out) + ((size_t)b * a.C + c) * a.out_ld;
#pragma unroll
    for (int o = 0; o < kChunk; ++o)
      if (t0 + o < a.T_out) orow[t0 + o] = acc[o] + c0.bias;
  }
}

template <int NW, int STRIDE>
int launch(const DwArgs& a, cudaStream_t stream) {
  const int items = a.C * a.nchunk;
  dim3 grid((items + kThreads - 1) / kThreads, a.B);
  const int max_rows = (kThreads + a.nchunk - 1) / a.nchunk + 1;
  const size_t smem = (size_t)max_rows * NW * 4 * sizeof(uint32_t);
  const bool xs = a.ep.flags & QASR_F_B_SIGNED;
  const bool fast = !(a.ep.flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64));
#define QASR_DW_GO(XS_, FAST_)                                                             \
  do {                                                                                     \
    auto kern = dwconv_kernel<NW, STRIDE, XS_, FAST_>;                                     \
    if (smem > 48 * 1024)                                                                  \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem));                                    \
    kern<<<grid, kThreads, smem, stream>>>(a);                                             \
  } while (0)
  if (xs && fast) QASR_DW_GO(true, true);
  else if (xs) QASR_DW_GO(true, false);
  else if (fast) QASR_DW_GO(false, true);
  else QASR_DW_GO(false, false);
#undef QASR_DW_GO
  return qasr_check_launch("dwconv_kernel");
}

}  // namespace

extern "C" int qasr_dwconv1d_i8(const void* x, int32_t x_ld, const int8_t* w, void* out,
                                int32_t out_ld, int32_t B, int32_t C, int32_t T_in, int32_t T_out,
                                int32_t k, int32_t stride, int32_t dilation, int32_t pad,
                                const qasr_epilogue* epi, void* stream) {
  QASR_CHECK_ARG(x && w && out && epi, "dwconv: null pointer");
  QASR_CHECK_ARG(B > 0 && C > 0 && T_in > 0 && T_out > 0 && k > 0, "dwconv: bad shape");
  QASR_CHECK_ARG(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2 (got %d)", stride);
  QASR_CHECK_ARG(dilation >= 1 && pad >= 0, "dwconv: bad dilation/pad");
  QASR_CHECK_ARG((T_in + 2 * pad - dilation * (k - 1) - 1) / stride + 1 == T_out,
                 "dwconv: T_out %d inconsistent with T_in %d k %d stride %d dil %d pad %d", T_out,
                 T_in, k, stride, dilation, pad);
  QASR_CHECK_ARG(x_ld % 16 == 0 && x_ld >= T_in && qasr_aligned16(x), "dwconv: x pitch/alignment");
  QASR_CHECK_ARG(out_ld >= ((T_out + 15) / 16) * 16 && qasr_aligned16(out),
                 "dwconv: out pitch must cover T_out rounded up to 16");
  QASR_CHECK_ARG(epi->mode != QASR_EPI_REQUANT_I8 || out_ld % 16 == 0, "dwconv: out pitch % 16");
  QASR_CHECK_ARG(B <= 65535, "dwconv: B too large for grid.y");
  int rc = qasr_validate_epilogue(epi, false);
  if (rc) return rc;

  DwArgs a;
  a.x = static_cast<const uint8_t*>(x);
  a.w = w;
  a.out = out;
  a.x_ld = x_ld; a.out_ld = out_ld; a.B = B; a.C = C; a.T_in = T_in; a.T_out = T_out;
  a.k = k; a.stride = stride; a.dil = dilation; a.pad = pad;
  a.nchunk = (T_out + kChunk - 1) / kChunk;
  a.d = ((-pad) % 4 + 4) % 4;
  const int span = (k - 1) * dilation + 1;
  a.nw = (a.d + 3 + span + 3) / 4;
  a.ep = make_epi_args(epi);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

#define QASR_DW_CASE(NW_)                                    \
  if (a.nw <= NW_)                                           \
    return stride == 1 ? launch<NW_, 1>(a, st) : launch<NW_, 2>(a, st);
  QASR_DW_CASE(4)
  QASR_DW_CASE(9)
  QASR_DW_CASE(11)
  QASR_DW_CASE(15)
  QASR_DW_CASE(17)
  QASR_DW_CASE(21)
  QASR_DW_CASE(32)
  QASR_DW_CASE(45)
#undef QASR_DW_CASE
  qasr_set_error("dwconv: kernel span %d (k=%d, dilation=%d) exceeds the supported 176 taps", span, k,
                 dilation);
  return QASR_ERR_UNSUPPORTED;
}

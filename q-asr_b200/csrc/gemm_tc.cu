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
// Warp roles (640 threads, 1 CTA / SM, persistent over tiles):
//   warp 0: TMA producer (one elected lane)      warp 1: MMA issuer (one elected lane)
//   warp 2: TMEM allocator                        warps 4..19: epilogue (TMEM -> regs -> HBM),
//   four per scheduler so TMEM-load / store latencies of one warp hide behind the others
#include <cuda.h>

#include "qasr_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace qasr_tc;

constexpr int BM = 128, BN = 128, BK = 128;
constexpr int kStages = 5;
constexpr int kEpiWarps = 16;
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;
constexpr uint32_t kABytes = BM * BK, kBBytes = BK * BN;
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t kTmemCols = 512;

struct GemmSeg {
  int taps, cblocks, stride, dil, pad, c_out_pad;
  uint32_t idesc;
};

struct GemmArgs {
  void* out;
  int out_ld, B, C_out, T_out;
  int nseg, m_tiles, n_tiles, total_tiles;
  GemmSeg seg[2];
  uint32_t b_lbo, b_sbo, a_sbo;  // smem descriptor strides (bytes)
  int dbg_align, dbg_nonneg;     // bring-up knobs (qasr_debug_set 2 / 3)
  EpiArgs ep;
};

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
// EPI: bit 0 = all-integer fast epilogue, bits 1..3 (fast path only) = second conv operand,
// previous-res_act addend, second requantisation stage.  EPI == 0 is the literal (runtime-flag) path.
constexpr int kEpiFast = 1, kEpiDual = 2, kEpiPrev = 4, kEpiStage2 = 8;

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
    if (g.nseg > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 32 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const int mt = tile % g.m_tiles;
        const int rest = tile / g.m_tiles;
        const int nt = rest % g.n_tiles, b = rest / g.n_tiles;
        for (int sg = 0; sg < g.nseg; ++sg) {
          const GemmSeg& s = g.seg[sg];
          const CUtensorMap* mA = sg == 0 ? &tmA0 : &tmA1;
          const CUtensorMap* mB = sg == 0 ? &tmB0 : &tmB1;
          for (int tap = 0; tap < s.taps; ++tap) {
            int tcoord = nt * BN * s.stride + tap * s.dil - s.pad;
            if (g.dbg_align > 1) tcoord = (tcoord / g.dbg_align) * g.dbg_align;
            if (g.dbg_nonneg && tcoord < 0) tcoord = 0;
            for (int cb = 0; cb < s.cblocks; ++cb) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
              mbar_expect_tx(full_bar(stage), kStageBytes);
              tma_load_2d(sa, mA, cb * BK, tap * s.c_out_pad + mt * BM, full_bar(stage));
              tma_load_3d(sb, mB, tcoord, cb * BK, b, full_bar(stage));
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        for (int sg = 0; sg < g.nseg; ++sg) {
          const GemmSeg& s = g.seg[sg];
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * 2 * BN + sg * BN);
          const int slabs = s.taps * s.cblocks;
          for (int it = 0; it < slabs; ++it) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
#pragma unroll
            for (int kk = 0; kk < BK / 32; ++kk) {
              const uint64_t adesc = make_smem_desc(sa + kk * 32, 0, g.a_sbo);
              const uint64_t bdesc = make_smem_desc(sb + kk * 32 * BN, g.b_lbo, g.b_sbo);
              tc_mma_i8(d_tmem, adesc, bdesc, s.idesc, (it | kk) != 0 ? 1u : 0u);
            }
            tc_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        tc_commit(tfull_bar(as));  // accumulators of this tile are complete
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int cgroup = (warp - 4) >> 2;       // which 32 of the tile's 128 columns
    const EpiConsts kc = make_epi_consts(g.ep, g.nseg > 1);
    int as = 0;
    uint32_t aphase = 0;
    int last_mt = -1;
    bool row_ok = false;
    int cc = 0;
    ChanRegs c0, c1;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const int mt = tile % g.m_tiles;
      const int rest = tile / g.m_tiles;
      const int nt = rest % g.n_tiles, b = rest / g.n_tiles;
      if (mt != last_mt) {  // per-channel parameters change only with the M tile
        const int c = mt * BM + quarter * 32 + lane;
        row_ok = c < g.C_out;
        cc = row_ok ? c : 0;
        c0 = load_chan(g.ep.ch0, cc);
        c1 = g.nseg > 1 ? load_chan(g.ep.ch1, cc) : c0;
        last_mt = mt;
      }
      const int len = (g.ep.flags & QASR_F_MASK) ? min(g.ep.out_lens[b], g.T_out) : g.T_out;
      const size_t row = (size_t)b * g.C_out + cc;

      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const int col0 = cgroup * 32 + ch * 16;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * 2 * BN + col0);
        int acc0[16], acc1[16];
        tmem_ld16(taddr, acc0);
        if (g.nseg > 1) tmem_ld16(taddr + BN, acc1);
        tmem_ld_wait();
        const int tv = nt * BN + col0;   // first frame of this 16-column vector
        if (g.ep.mode == QASR_EPI_REQUANT_I8) {
          uint8_t* orow = reinterpret_cast<uint8_t*>(g.out) + row * g.out_ld;
          uint4 pv = make_uint4(0, 0, 0, 0);
          if ((EPI == 0 || (EPI & kEpiPrev)) && g.ep.prev != nullptr && row_ok && tv < g.T_out)
            pv = *reinterpret_cast<const uint4*>(g.ep.prev + row * g.ep.prev_ld + tv);
          const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
          uint32_t packed[4];
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            int q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = 4 * w4 + i;
              if (EPI & kEpiFast) {
                const int prev = (EPI & kEpiPrev) ? (int)(int8_t)(pw[w4] >> (8 * i)) : 0;
                q[i] = epilogue_fast<(EPI & kEpiDual) != 0, (EPI & kEpiPrev) != 0, (EPI & kEpiStage2) != 0>(
                    g.ep, kc, acc0[j], c0, (EPI & kEpiDual) ? acc1[j] : 0, c1, prev);
              } else {
                const int prev = (int)(int8_t)((pw[w4] >> (8 * i)) & 0xff);
                q[i] = epilogue_exact(g.ep, kc, acc0[j], c0, g.nseg > 1, g.nseg > 1 ? acc1[j] : 0, c1, prev);
              }
            }
            packed[w4] = pack4(q[0], q[1], q[2], q[3]);
          }
          if (tv + 16 > len) {  // length mask / row tail: zero the bytes at t >= len
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) {
              const int nvalid = len - (tv + 4 * w4);
              const uint32_t m = nvalid >= 4 ? 0xffffffffu : (nvalid <= 0 ? 0u : (0xffffffffu >> (8 * (4 - nvalid))));
              packed[w4] &= m;
            }
          }
          if (row_ok && tv < g.T_out)
            *reinterpret_cast<uint4*>(orow + tv) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        } else if (g.ep.mode == QASR_EPI_DEQUANT_F32) {
          float* orow = reinterpret_cast<float*>(g.out) + row * g.out_ld;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (row_ok && tv + j < g.T_out) orow[tv + j] = epilogue_dequant(g.ep, acc0[j], c0);
        } else {
          int32_t* orow = reinterpret_cast<int32_t*>(g.out) + row * g.out_ld;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (row_ok && tv + j < g.T_out) orow[tv + j] = acc0[j] + c0.bias;
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
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

int make_map_acts(CUtensorMap* map, const void* x, int T, int C, int B, int ld) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { qasr_set_error("cuTensorMapEncodeTiled entry point not available"); return QASR_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * (cuuint64_t)C};
  cuuint32_t box[3] = {(cuuint32_t)BN, (cuuint32_t)BK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { qasr_set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return QASR_ERR_CUDA; }
  return QASR_OK;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// debug overrides of the B-operand descriptor strides (set through qasr_debug_set)
long long g_dbg[8] = {-1, -1, -1, -1, -1, -1, -1, -1};

}  // namespace

extern "C" int qasr_debug_set(int key, long long value) {
  if (key < 0 || key >= 8) return QASR_ERR_INVALID;
  g_dbg[key] = value;
  return QASR_OK;
}

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
  QASR_CHECK_ARG(!op1 || epi->mode == QASR_EPI_REQUANT_I8, "conv1d_gemm: second operand needs REQUANT mode");

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
  g.total_tiles = g.m_tiles * g.n_tiles * B;
  g.a_sbo = 1024;
  g.b_sbo = g_dbg[1] >= 0 ? (uint32_t)g_dbg[1] : 1024;
  g.b_lbo = g_dbg[0] >= 0 ? (uint32_t)g_dbg[0] : (uint32_t)(BK * 128);
  g.ep = make_epi_args(epi);
  g.dbg_align = g_dbg[2] > 0 ? (int)g_dbg[2] : 0;
  g.dbg_nonneg = g_dbg[3] > 0 ? 1 : 0;

  const bool fast = !(epi->flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64)) &&
                    epi->mode == QASR_EPI_REQUANT_I8;
  int variant = 0;
  if (fast)
    variant = kEpiFast | (op1 ? kEpiDual : 0) | (epi->prev ? kEpiPrev : 0) | ((epi->flags & QASR_F_STAGE2) ? kEpiStage2 : 0);
  int grid = qasr_sm_count();
  if (grid > g.total_tiles) grid = g.total_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[16][64] = {};
  int dev = 0;
  QASR_CHECK_CUDA(cudaGetDevice(&dev));
  const bool need_attr = dev >= 0 && dev < 64 && !attr_set[variant][dev];
#define QASR_GEMM_CASE(V_)                                                                                     \
  case V_:                                                                                                     \
    if (need_attr)                                                                                             \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<V_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)kSmemBytes));                                                  \
    gemm_i8_kernel<V_><<<grid, kGemmThreads, kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], g);         \
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
    default:
      qasr_set_error("conv1d_gemm: internal: no kernel variant %d", variant);
      return QASR_ERR_INVALID;
  }
#undef QASR_GEMM_CASE
  if (need_attr) attr_set[variant][dev] = true;
  return qasr_check_launch("gemm_i8_kernel");
}

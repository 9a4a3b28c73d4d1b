// K2b: dense 1-D convolution (any k, stride, dilation; groups == 1) on channels-last 8-bit codes
// [B][T][C] as an implicit int8 GEMM on tcgen05 -- Jasper's k = 11..29 convs, its strided head, and any
// 1x1 conv of a channels-last network.  Restates F.conv1d(x_int.double(), w_int.double(), bias) of
// QuantConv1d.int_conv (nemo/quantization/utils/quant_modules.py:301-305).
//
// Why a second layout: TMA needs a 16-byte aligned box start along the contiguous axis (measured: an
// unaligned inner coordinate faults), so in the reference's (B,C,T) layout the conv taps cannot be fetched
// as time-shifted boxes.  With channels contiguous the tap shift moves the NON-contiguous (time)
// coordinate, which may be any integer, negative or past the end (zero filled = the conv padding), and a
// strided conv is a TMA traversal stride on that axis.
//
// GEMM view per utterance:  D[T_out, C_out] = sum_tap X[T_out*s + tap*d - pad, C_in] * W_tap[C_in, C_out]
//   A operand = activation box 128 frames x 128 channels, K-major, 128B swizzle (3-D TMA, time coordinate
//               shifted per tap)
//   B operand = packed weights [tap][C_out_pad][C_in_pad] (same packing as K2), K-major, 128B swizzle
//   D         = int32 in TMEM: lanes = frames, columns = output channels; epilogue threads own one frame
//               and walk 16 channels at a time -> 16-byte stores into the channels-last output.  Per-channel
//               epilogue parameters sit in a shared-memory table (broadcast reads).
// Pipeline / warp roles as in gemm_tc.cu (TMA producer, MMA issuer, TMEM allocator, 16 epilogue warps,
// double-buffered TMEM, optional second operand pair for the residual 1x1 conv + res_act).
#include <cuda.h>

#include "qasr_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace qasr_tc;

// Tile = BM frames x (128 NB) output channels.  NB = 2 (single-operand launches whose C_out is a multiple of 256): one
// 16 KB activation slab feeds an N = 256 MMA, i.e. 48 KB of operands per 512 clk of MMAs instead of 64 KB -- the dense
// convs of Jasper are bound by operand delivery (435 clk per slab pair against 256 clk of MMAs at NB = 1).
constexpr int BM = 128, BN1 = 128, BK = 128;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr uint32_t kABytes = BM * BK, kBBytes1 = BN1 * BK;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;   // TMEM columns per accumulator stage: 2 operands x 128 (NB = 1) or 1 x 256 (NB = 2);
                                       // two operands x 256 (NB = 2, res_act) take all 512 columns: ONE stage
template <int NB> struct Geo {
  static constexpr int BN = BN1 * NB;
  static constexpr int kStages = NB == 1 ? 5 : 4;
  static constexpr uint32_t kStageBytes = kABytes + NB * kBBytes1;
  static constexpr uint32_t kTabBytes = 2 /*operands*/ * 2 /*fast, exact*/ * BN * 16;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kTabBytes;
};

struct Seg {
  int taps, cblocks, stride, dil, pad, c_out_pad;
  uint32_t idesc;
};

struct Args {
  void* out;
  int out_ld, out_nct, B, C_out, T_out;
  int ct_major;      // tile walk: 1 = contiguous range per CTA with the channel tile outermost (few slabs per tile), 0 = strided
  int nseg, t_tiles, c_tiles, total_tiles;
  Seg seg[2];
  unsigned long long* stamp;   // in-graph timing slot (qasr_profile_begin) or nullptr
  EpiArgs ep;
};

// shared-memory per-channel records (two uint4 per operand and channel)
__device__ __forceinline__ void store_chan(uint4* fast, uint4* exact, const qasr_chan* tab, int c, int C_out) {
  int4 v = make_int4(0, 0, 0, 1);
  if (c < C_out) v = __ldg(reinterpret_cast<const int4*>(tab) + c);
  const int m = v.z, e = v.w;
  const long long add = (long long)v.y * (long long)m + ((e >= 1 && e <= 62) ? (1ll << (e - 1)) : 0ll);
  const double mult = dyad_mult(m, e);
  *fast = make_uint4((uint32_t)m, (uint32_t)max(e - 32, 0), (uint32_t)(add & 0xffffffffll), (uint32_t)(add >> 32));
  const long long mb = __double_as_longlong(mult);
  *exact = make_uint4((uint32_t)v.x, (uint32_t)v.y, (uint32_t)(mb & 0xffffffffll), (uint32_t)(mb >> 32));
}
__device__ __forceinline__ ChanRegs load_chan_smem(const uint4* fast, const uint4* exact, bool need_exact) {
  ChanRegs r;
  const uint4 f = *fast;
  r.m = (int)f.x;
  r.sh = (int)f.y;
  r.add = (long long)(((unsigned long long)f.w << 32) | f.z);
  r.e = r.sh + 32;
  r.s = 0.f;
  r.bias = 0;
  r.mult = 0.0;
  if (need_exact) {
    const uint4 x = *exact;
    r.s = __uint_as_float(x.x);
    r.bias = (int)x.y;
    r.mult = __longlong_as_double((long long)(((unsigned long long)x.w << 32) | x.z));
  }
  return r;
}

// EPI as in gemm_tc.cu: bit0 fast, bit1 dual, bit2 prev, bit3 stage2; 0 = literal path
constexpr int kEpiFast = 1, kEpiDual = 2, kEpiPrev = 4, kEpiStage2 = 8;
constexpr int kEpiRange = 64;   // QASR_F_RANGE_CHECK (see gemm_tc.cu)

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}

template <int EPI, int NB>
__global__ void __launch_bounds__(kThreads, 1)
conv_ntc_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmW0,
                const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmW1, const Args g) {
  constexpr int BN = Geo<NB>::BN, kStages = Geo<NB>::kStages;
  constexpr uint32_t kStageBytes = Geo<NB>::kStageBytes;
  // accumulator stages in TMEM: the MMAs of tile i + 1 overlap the epilogue of tile i, except for two-operand 256-channel
  // tiles, whose accumulators fill the 512 columns (the drain is a few percent of a dense tile's MMA time)
  constexpr int kAccStages = (NB == 2 && (EPI & kEpiDual)) ? 1 : 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  uint4* tab = reinterpret_cast<uint4*>(gen_base + kStages * kStageBytes + 256);  // [op][fast|exact][BN]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW0) : "memory");
    if (g.nseg > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
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
  if (threadIdx.x == 0) qasr_stamp_begin(g.stamp);
  // tile walk of this CTA: a CONTIGUOUS range with the channel tile outermost, so the epilogue refills its per-channel
  // table (a global round trip between two barriers of all epilogue warps) once instead of per tile and consecutive tiles
  // reuse the channel tile's weights.  Same-box A/B on Jasper10x5dr (B32 x T1600; the box-to-box spread of this
  // power-limited workload is larger than the effect): 8.38 ms against 8.64-8.77 ms with the strided walk (bring-up
  // knob 20 = 0), the 1x1 dense-residual chain launches gaining most (40.7 -> 34.5 us on 256 -> 768 channels).
  const int tile0 = g.ct_major ? (int)((long long)blockIdx.x * g.total_tiles / gridDim.x) : (int)blockIdx.x;
  const int tile1 = g.ct_major ? (int)((long long)(blockIdx.x + 1) * g.total_tiles / gridDim.x) : g.total_tiles;
  const int tile_step = g.ct_major ? 1 : (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks with uniform values, one elected lane issues) =====================
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    // A CTA walks a CONTIGUOUS range of tiles, channel tile outermost: its ~9 tiles share one channel tile, so the
    // epilogue refills its per-channel table (a global round trip between two CTA-wide barriers) once, not per tile
    const int per_ct = g.t_tiles * g.B;
    for (int tile = tile0; tile < tile1; tile += tile_step) {
      const int ct = g.ct_major ? tile / per_ct : tile % g.c_tiles;
      const int rest = g.ct_major ? tile - ct * per_ct : tile / g.c_tiles;
      const int tt = rest % g.t_tiles, b = rest / g.t_tiles;
      for (int sg = 0; sg < g.nseg; ++sg) {
        const Seg& s = g.seg[sg];
        const CUtensorMap* mX = sg == 0 ? &tmX0 : &tmX1;
        const CUtensorMap* mW = sg == 0 ? &tmW0 : &tmW1;
        const int taps = s.taps, cblocks = s.cblocks, wrow0 = ct * BN, c_out_pad = s.c_out_pad;
        const int t0 = tt * BM * s.stride - s.pad, dil = s.dil;
        for (int tap = 0; tap < taps; ++tap) {
          const int tcoord = t0 + tap * dil;
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (leader) {
              const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
              mbar_expect_tx(full_bar(stage), kStageBytes);
              tma_load_3d(sa, mX, cb * BK, tcoord, b, full_bar(stage));
              tma_load_2d(sb, mW, cb * BK, tap * c_out_pad + wrow0, full_bar(stage));
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks; descriptors stay in uniform registers: see gemm_tc.cu) =====================
    const bool leader = elect_one();
    const uint64_t desc_hi = make_smem_desc(0, 0, 1024);
    const uint32_t stage_lo0 = (smem_base & 0x3FFFFu) >> 4;
    constexpr uint32_t kStageLo = kStageBytes >> 4, kBLo = kABytes >> 4;
    const int seg_slabs[2] = {g.seg[0].taps * g.seg[0].cblocks, g.seg[1].taps * g.seg[1].cblocks};
    const uint32_t seg_idesc[2] = {g.seg[0].idesc, g.seg[1].idesc};
    const int nseg = g.nseg;
    int stage = 0;
    uint32_t phase = 0, stage_lo = stage_lo0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = tile0; tile < tile1; tile += tile_step) {
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        if (sg >= nseg) break;
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccStride + sg * BN);
        const int slabs = seg_slabs[sg];
        const uint32_t idesc = seg_idesc[sg];
        for (int it = 0; it < slabs; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < BK / 32; ++kk)
              tc_mma_i8(d_tmem, desc_hi | (uint64_t)(stage_lo + kk * 2), desc_hi | (uint64_t)(stage_lo + kBLo + kk * 2), idesc,
                        (it | kk) != 0 ? 1u : 0u);
            tc_commit(empty_bar(stage));
          }
          stage_lo += kStageLo;
          if (++stage == kStages) { stage = 0; phase ^= 1u; stage_lo = stage_lo0; }
        }
      }
      if (leader) tc_commit(tfull_bar(as));
      if (++as == kAccStages) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;          // TMEM lane quarter = 32 frames
    const int cgroup = (warp - 4) >> 2;    // 32 of the tile's 128 output channels
    const int etid = threadIdx.x - 128;    // 0..511
    const EpiConsts kc = make_epi_consts(g.ep, g.nseg > 1);
    const bool need_exact = !(EPI & kEpiFast);
    int as = 0;
    uint32_t aphase = 0;
    const int per_ct = g.t_tiles * g.B;
    int cur_ct = -1;
    for (int tile = tile0; tile < tile1; tile += tile_step) {
      const int ct = g.ct_major ? tile / per_ct : tile % g.c_tiles;
      const int rest = g.ct_major ? tile - ct * per_ct : tile / g.c_tiles;
      const int tt = rest % g.t_tiles, b = rest / g.t_tiles;
      if (ct != cur_ct) {   // CTA-uniform
        // per-channel parameter table of this channel tile
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");  // previous tile's readers are done
        if (etid < BN) {
          store_chan(tab + etid, tab + BN + etid, g.ep.ch0, ct * BN + etid, g.C_out);
          if (g.nseg > 1) store_chan(tab + 2 * BN + etid, tab + 3 * BN + etid, g.ep.ch1, ct * BN + etid, g.C_out);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
        cur_ct = ct;
      }
      const int t = tt * BM + quarter * 32 + lane;
      const bool t_ok = t < g.T_out;
      const int len = (g.ep.flags & QASR_F_MASK) ? min(g.ep.out_lens[b], g.T_out) : g.T_out;
      const size_t trow = (size_t)b * g.T_out + (t_ok ? t : 0);

      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      // single-operand fast launches drain their 32 columns in one go: the TMEM buffer goes back to the MMA warp before
      // the arithmetic and every lane writes one full 32-byte sector (Jasper's 1x1 dense-residual chain is paced by this
      // epilogue); two-operand and literal launches keep 16-column halves (registers)
      constexpr int W = ((EPI & kEpiFast) && !(EPI & kEpiDual)) ? 32 : 16;
      constexpr int kCols = BN / 4;          // columns of this warp's column group
      constexpr int kChunks = kCols / W;
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ++ch) {
        const int col0 = cgroup * kCols + ch * W;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccStride + col0);
        int acc0[W], acc1[16];
        if constexpr (W == 32) {
          tmem_ld32(taddr, acc0);
        } else {
          tmem_ld16(taddr, acc0);
          if (g.nseg > 1) tmem_ld16(taddr + BN, acc1);
        }
        tmem_ld_wait();
        if (W == 32 && ch == kChunks - 1) {   // the last load of this buffer has landed
          tc_fence_before();
          mbar_arrive_warp(tempty_bar(as));
        }
        const int c_first = ct * BN + col0;
        if (g.ep.mode == QASR_EPI_REQUANT_I8) {
          uint32_t pw[W / 4];
#pragma unroll
          for (int i = 0; i < W / 4; ++i) pw[i] = 0u;
          if ((EPI == 0 || (EPI & kEpiPrev)) && g.ep.prev != nullptr && t_ok && c_first < g.C_out) {
            const uint4* pp = reinterpret_cast<const uint4*>(g.ep.prev + trow * g.ep.prev_ld + c_first);
#pragma unroll
            for (int i = 0; i < W / 16; ++i) {
              if (c_first + 16 * i < g.C_out) {
                const uint4 pv = pp[i];
                pw[4 * i] = pv.x; pw[4 * i + 1] = pv.y; pw[4 * i + 2] = pv.z; pw[4 * i + 3] = pv.w;
              }
            }
          }
          uint32_t packed[W / 4];
#pragma unroll
          for (int w4 = 0; w4 < W / 4; ++w4) {
            int q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = 4 * w4 + i;
              const ChanRegs c0 = load_chan_smem(tab + col0 + j, tab + BN + col0 + j, need_exact);
              ChanRegs c1 = c0;
              if (EPI == 0 ? g.nseg > 1 : (EPI & kEpiDual) != 0)
                c1 = load_chan_smem(tab + 2 * BN + col0 + j, tab + 3 * BN + col0 + j, need_exact);
              const int prev = (int)(int8_t)((pw[w4] >> (8 * i)) & 0xff);
              if (EPI & kEpiFast)
                q[i] = epilogue_fast<(EPI & kEpiDual) != 0, (EPI & kEpiPrev) != 0, (EPI & kEpiStage2) != 0, (EPI & kEpiRange) != 0>(
                    g.ep, kc, acc0[j], c0, (EPI & kEpiDual) ? acc1[j & 15] : 0, c1, prev);
              else
                q[i] = epilogue_exact(g.ep, kc, acc0[j], c0, g.nseg > 1, g.nseg > 1 ? acc1[j & 15] : 0, c1, prev);
            }
            packed[w4] = pack4(q[0], q[1], q[2], q[3]);
          }
          if (t >= len) {   // jasper.py:177-181
#pragma unroll
            for (int i = 0; i < W / 4; ++i) packed[i] = 0u;
          }
          if (t_ok && c_first < g.C_out) {
            uint8_t* orow = reinterpret_cast<uint8_t*>(g.out) + trow * g.out_ld + c_first;
            if (W == 32 && c_first + 16 < g.C_out && ((reinterpret_cast<uintptr_t>(orow) & 31) == 0)) {
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow), "r"(packed[0]), "r"(packed[1]),
                           "r"(packed[2]), "r"(packed[3]), "r"(packed[W / 4 - 4]), "r"(packed[W / 4 - 3]), "r"(packed[W / 4 - 2]),
                           "r"(packed[W / 4 - 1]) : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < W / 16; ++i)
                if (c_first + 16 * i < g.C_out)
                  *reinterpret_cast<uint4*>(orow + 16 * i) = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            }
          }
        } else {
          // float32 / int32 outputs go to the reference's (B, C, T) layout: consecutive lanes = consecutive
          // frames of one channel -> coalesced
#pragma unroll
          for (int j = 0; j < W; ++j) {
            const int c = c_first + j;
            if (!t_ok || c >= g.C_out) continue;
            const ChanRegs c0 = load_chan_smem(tab + col0 + j, tab + BN + col0 + j, true);
            const size_t o = ((size_t)b * g.C_out + c) * g.out_ld + t;
            if (g.ep.mode == QASR_EPI_DEQUANT_F32)
              reinterpret_cast<float*>(g.out)[o] = epilogue_dequant(g.ep, acc0[j], c0);
            else
              reinterpret_cast<int32_t*>(g.out)[o] = acc0[j] + c0.bias;
          }
        }
      }
      if (W != 32) {
        tc_fence_before();
        mbar_arrive_warp(tempty_bar(as));
      }
      if (++as == kAccStages) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) qasr_stamp_end(g.stamp);
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

int make_map_weights(CUtensorMap* map, const int8_t* wp, int rows, int c_in_pad, int bn) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { qasr_set_error("cuTensorMapEncodeTiled entry point not available"); return QASR_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)c_in_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)c_in_pad};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(wp), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { qasr_set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return QASR_ERR_CUDA; }
  return QASR_OK;
}

// activations [B][T][ld]: dims (C, T, B); the time axis carries the conv stride as a traversal stride
int make_map_acts(CUtensorMap* map, const void* x, int C, int T, int B, int ld, int stride) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { qasr_set_error("cuTensorMapEncodeTiled entry point not available"); return QASR_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * (cuuint64_t)T};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(BM * stride), 1};
  cuuint32_t estr[3] = {1, (cuuint32_t)stride, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { qasr_set_error("cuTensorMapEncodeTiled(activations, channels-last) failed: %d", (int)r); return QASR_ERR_CUDA; }
  return QASR_OK;
}

// ---------------------------------------------------------------------------------------------
// (B,C,T) float32 -> channels-last int8 codes, tile transpose through shared memory
//   mode 0: first-layer SymmetricQuantFunction  clamp(rne((1/s) * x), -n, n-1)   (quant_modules.py:180-184)
//   mode 1: codes of an already quantised tensor  rne(x / s)                      (quant_modules.py:301)
// ---------------------------------------------------------------------------------------------
__global__ void quantize_ntc_kernel(const float* __restrict__ x, int x_ld, const float* scale, int mode, int n,
                                    int8_t* __restrict__ codes, int c_ld, int B, int C, int T,
                                    const int32_t* lens) {
  __shared__ int8_t tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const float s = scale[0];
  const float inv = __fdiv_rn(1.0f, s);
  const int len = lens ? lens[b] : T;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    float v = 0.f;
    if (c < C && t < T && t < len) v = x[((size_t)b * C + c) * x_ld + t];
    float r;
    if (mode == 0) r = fminf(fmaxf(rintf(__fmul_rn(inv, v)), (float)(-n)), (float)(n - 1));
    else r = rintf(__fdiv_rn(v, s));
    tile[i][threadIdx.x] = (int8_t)(int)r;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < c_ld) codes[((size_t)b * T + t) * c_ld + c] = c < C ? tile[threadIdx.x][i] : (int8_t)0;
  }
}

}  // namespace

extern "C" int qasr_quantize_f32_ntc(const float* x, int32_t x_ld, const float* scale, int32_t mode, int32_t bit,
                                     int8_t* codes, int32_t c_ld, int32_t B, int32_t C, int32_t T,
                                     const int32_t* lens, void* stream) {
  QASR_CHECK_ARG(x && scale && codes, "quantize_f32_ntc: null pointer");
  QASR_CHECK_ARG(B > 0 && C > 0 && T > 0 && x_ld >= T && c_ld >= C && c_ld % 16 == 0, "quantize_f32_ntc: bad shape / pitch");
  QASR_CHECK_ARG(mode == 0 || mode == 1, "quantize_f32_ntc: mode");
  QASR_CHECK_ARG(mode == 1 || (bit >= 2 && bit <= 8), "quantize_f32_ntc: activation_bit %d not in [2, 8]", bit);
  QASR_CHECK_ARG(B <= 65535, "quantize_f32_ntc: B too large");
  dim3 grid((T + 31) / 32, (c_ld + 31) / 32, B), block(32, 8);
  quantize_ntc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, x_ld, scale, mode, (1 << (bit - 1)) - 1, codes,
                                                               c_ld, B, C, T, lens);
  return qasr_check_launch("quantize_ntc_kernel");
}

extern "C" int qasr_conv1d_ntc_i8(const qasr_gemm_operand* op0, const qasr_gemm_operand* op1, void* out,
                                  int32_t out_ld, int32_t B, int32_t C_out, int32_t T_out,
                                  const qasr_epilogue* epi, void* stream) {
  QASR_CHECK_ARG(op0 && out && epi, "conv1d_ntc: null pointer");
  QASR_CHECK_ARG(B > 0 && C_out > 0 && T_out > 0, "conv1d_ntc: bad shape");
  int rc = qasr_validate_epilogue(epi, true);
  if (rc) return rc;
  QASR_CHECK_ARG((op1 != nullptr) == (epi->ch1 != nullptr), "conv1d_ntc: second operand and epi->ch1 must come together");
  QASR_CHECK_ARG(!op1 || epi->mode == QASR_EPI_REQUANT_I8, "conv1d_ntc: second operand needs REQUANT mode");
  if (epi->mode == QASR_EPI_REQUANT_I8) {
    QASR_CHECK_ARG(qasr_aligned16(out) && out_ld % 16 == 0 && out_ld >= round_up(C_out, 16),
                   "conv1d_ntc: channels-last output pitch must be a multiple of 16 covering C_out");
  } else {
    QASR_CHECK_ARG(out_ld >= T_out, "conv1d_ntc: (B,C,T) output pitch must cover T_out");
  }
  Args g;
  memset(&g, 0, sizeof(g));
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  const qasr_gemm_operand* ops[2] = {op0, op1};
  g.nseg = op1 ? 2 : 1;
  const bool fast = !(epi->flags & (QASR_F_EXACT_F32 | QASR_F_EXACT_DYAD | QASR_F_BIAS64)) &&
                    epi->mode == QASR_EPI_REQUANT_I8;
  // 256-channel tiles: single-operand fast launches (conv -> act) with C_out a multiple of 256 and enough slabs per tile
  // for the operand traffic to matter; bring-up knob 21 = 0 keeps 128-channel tiles
  int nb = 1;
  // (a last half-empty tile is accepted up to 20 % extra work: C_out = 896, 640; its surplus channels are computed on whatever
  // weight rows follow and never stored)
  const int c256 = (C_out + 255) / 256 * 256;
  long long slabs_tile = (long long)op0->k * ((op0->C_in + BK - 1) / BK);
  if (op1) slabs_tile += (long long)op1->k * ((op1->C_in + BK - 1) / BK);
  // single operand: conv -> act (variants 1, 65); two operands: res_act with or without the second stage (3, 11, 67, 75)
  const bool nb2_shape = op1 ? !epi->prev : (!epi->prev && !(epi->flags & QASR_F_STAGE2));
  if (fast && nb2_shape && (c256 - C_out) * 100 <= 20 * C_out && qasr_debug_get(21) != 0 && slabs_tile >= 8)
    nb = 2;
  const int BN = BN1 * nb;
  for (int i = 0; i < g.nseg; ++i) {
    const qasr_gemm_operand* o = ops[i];
    QASR_CHECK_ARG(o->x && o->wp && o->C_in > 0 && o->T_in > 0 && o->k > 0, "conv1d_ntc: operand %d bad", i);
    QASR_CHECK_ARG(o->stride >= 1 && o->stride <= 2 && o->dilation >= 1 && o->pad >= 0, "conv1d_ntc: stride/dilation/pad");
    QASR_CHECK_ARG((o->T_in + 2 * o->pad - o->dilation * (o->k - 1) - 1) / o->stride + 1 == T_out,
                   "conv1d_ntc: operand %d: T_out %d inconsistent with T_in %d", i, T_out, o->T_in);
    QASR_CHECK_ARG(o->x_ld % 16 == 0 && o->x_ld >= o->C_in && qasr_aligned16(o->x) && qasr_aligned16(o->wp),
                   "conv1d_ntc: operand %d channel pitch / alignment", i);
    Seg& s = g.seg[i];
    s.taps = o->k;
    s.cblocks = (o->C_in + BK - 1) / BK;
    s.stride = o->stride;
    s.dil = o->dilation;
    s.pad = o->pad;
    s.c_out_pad = round_up(C_out, BN1);   // the packing of qasr_pack_conv_weights
    const uint32_t a_signed = (epi->flags & (i == 0 ? QASR_F_B_SIGNED : QASR_F_B1_SIGNED)) ? 1u : 0u;
    // D=S32, A = activations (un)signed K-major, B = weights signed K-major, N=128, M=128
    s.idesc = (2u << 4) | (a_signed << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    rc = make_map_acts(&maps[2 * i], o->x, o->C_in, o->T_in, B, o->x_ld, o->stride);
    if (rc) return rc;
    rc = make_map_weights(&maps[2 * i + 1], o->wp, o->k * s.c_out_pad, round_up(o->C_in, BK), BN);
    if (rc) return rc;
  }
  if (g.nseg == 1) { maps[2] = maps[0]; maps[3] = maps[1]; }
  g.out = out;
  g.out_ld = out_ld;
  g.B = B;
  g.C_out = C_out;
  g.T_out = T_out;
  g.t_tiles = (T_out + BM - 1) / BM;
  g.c_tiles = (C_out + BN - 1) / BN;
  g.total_tiles = g.t_tiles * g.c_tiles * B;
  {
    int slabs = 0;
    for (int sg = 0; sg < g.nseg; ++sg) slabs += g.seg[sg].taps * g.seg[sg].cblocks;
    (void)slabs;
    g.ct_major = 1;
    if (qasr_debug_get(20) >= 0) g.ct_major = qasr_debug_get(20) > 0 ? 1 : 0;   // bring-up knob 20: force the tile walk
  }
  g.ep = make_epi_args(epi);
  g.stamp = qasr_next_stamp();
  int variant = 0;
  if (fast)
    variant = kEpiFast | (op1 ? kEpiDual : 0) | (epi->prev ? kEpiPrev : 0) | ((epi->flags & QASR_F_STAGE2) ? kEpiStage2 : 0);
  if (fast && (epi->flags & QASR_F_RANGE_CHECK)) variant |= kEpiRange;
  int grid = qasr_sm_count();
  if (grid > g.total_tiles) grid = g.total_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[128][64] = {};
  int dev = 0;
  QASR_CHECK_CUDA(cudaGetDevice(&dev));
  const bool need_attr = dev >= 0 && dev < 64 && !attr_set[variant][dev];
#define QASR_NTC_CASE(V_)                                                                                      \
  case V_:                                                                                                     \
    if (need_attr)                                                                                             \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(conv_ntc_kernel<V_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)Geo<1>::kSmemBytes));                                          \
    conv_ntc_kernel<V_, 1><<<grid, kThreads, Geo<1>::kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], g); \
    break;
  if (nb == 2) {   // 256-channel tiles: variants 1, 65 (single operand) and 3, 11, 67, 75 (res_act), see the choice of nb
    static bool attr2_set[128][64] = {};
    const bool need2 = dev >= 0 && dev < 64 && !attr2_set[variant][dev];
#define QASR_NTC2_CASE(V_)                                                                                         \
  case V_:                                                                                                         \
    if (need2)                                                                                                     \
      QASR_CHECK_CUDA(cudaFuncSetAttribute(conv_ntc_kernel<V_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)Geo<2>::kSmemBytes));                                              \
    conv_ntc_kernel<V_, 2><<<grid, kThreads, Geo<2>::kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], g);     \
    break;
    switch (variant) {
      QASR_NTC2_CASE(1)
      QASR_NTC2_CASE(65)
      QASR_NTC2_CASE(3)
      QASR_NTC2_CASE(11)
      QASR_NTC2_CASE(67)
      QASR_NTC2_CASE(75)
      default:
        qasr_set_error("conv1d_ntc: internal: no 256-channel kernel variant %d", variant);
        return QASR_ERR_INVALID;
    }
#undef QASR_NTC2_CASE
    if (need2) attr2_set[variant][dev] = true;
    return qasr_check_launch("conv_ntc_kernel");
  }
  switch (variant) {
    QASR_NTC_CASE(0)
    QASR_NTC_CASE(1)
    QASR_NTC_CASE(3)
    QASR_NTC_CASE(5)
    QASR_NTC_CASE(7)
    QASR_NTC_CASE(9)
    QASR_NTC_CASE(11)
    QASR_NTC_CASE(13)
    QASR_NTC_CASE(15)
    QASR_NTC_CASE(65)
    QASR_NTC_CASE(67)
    QASR_NTC_CASE(69)
    QASR_NTC_CASE(71)
    QASR_NTC_CASE(73)
    QASR_NTC_CASE(75)
    QASR_NTC_CASE(77)
    QASR_NTC_CASE(79)
    default:
      qasr_set_error("conv1d_ntc: internal: no kernel variant %d", variant);
      return QASR_ERR_INVALID;
  }
#undef QASR_NTC_CASE
  if (need_attr) attr_set[variant][dev] = true;
  return qasr_check_launch("conv_ntc_kernel");
}

// K4b: dynamic quantisation on the device (reference QuantAct.forward with dynamic=True,
// nemo/quantization/utils/quant_modules.py:149-167).  The range of every QuantAct is min / max of the float tensor
// it receives, i.e. of the producing conv's output y = f32(acc + b + delta) * s_out[c] (quant_modules.py:301-309),
// where delta = sum w * eps(q) is the float noise of the reference's un-rounded x_int = x / s_x (:301).  Only
// max(|min|, |max|) enters the scale (quant_utils.py:47-52), so one number per QuantAct is needed EXACTLY:
// absmax = max |x_act|.  The statistics pass of the conv kernels (QASR_EPI_STATS) leaves per-channel extremes of the
// noise-free integer accumulators; here
//   dyn_select   bounds every channel's max |y| from both sides (|delta| <= sum|w| * max|eps|) and lists the
//                channels that can hold the global maximum (almost always one),
//   dyn_refine   recomputes those channels position by position with the noise term, bit for bit as the
//                reference's float64 conv + float32 rescale does, and reduces the exact absmax,
//   dyn_finalize turns absmax (optionally exchanged between the ranks of a data-parallel job through peer memory)
//                into the scale, the producers' dyadic multipliers and the consumers' s_out / integer bias,
// so no scale ever visits the host and the whole dynamic forward replays as one CUDA graph.
#include <float.h>

#include "qasr_common.cuh"

namespace {

constexpr int kRefThreads = 256;   // 8 warps: pointwise convs split their input channels over the warps
constexpr int kMaxWRow = 8192;
constexpr int kMaxHits = 64;      // positions per block and candidate channel whose noise term is evaluated by the whole block   // weights of one output channel kept in shared memory by dyn_refine

struct DynOp {   // by-value kernel copy of qasr_dyn_operand
  const uint8_t* x;
  const int8_t* w;
  const float* s_x;
  const qasr_chan* ch;
  const int32_t* wabs;
  int x_ld, C_in, T_in, k, stride, dil, pad, depthwise, in_unsigned;
};

DynOp make_op(const qasr_dyn_operand* o) {
  DynOp d;
  memset(&d, 0, sizeof(d));
  if (o) {
    d.x = static_cast<const uint8_t*>(o->x);
    d.w = o->w;
    d.s_x = o->s_x;
    d.ch = o->ch;
    d.wabs = o->wabs;
    d.x_ld = o->x_ld; d.C_in = o->C_in; d.T_in = o->T_in; d.k = o->k; d.stride = o->stride; d.dil = o->dilation;
    d.pad = o->pad; d.depthwise = o->depthwise; d.in_unsigned = o->in_unsigned;
  }
  return d;
}

// eps(q) * 2^22 of the reference's x_int = (f32(q) * s) / s  (quant_modules.py:301): a small integer, |.| <= 64
__device__ __forceinline__ int noise_code(int q, float s) {
  const float r = __fdiv_rn(__fmul_rn((float)q, s), s);
  return (int)rint(((double)r - (double)q) * 4194304.0);
}

__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// candidate channels: bounds  L_c <= max_t |y| <= U_c  from the noise-free statistics; every block of the refinement
// repeats this (C records, a few KB from L2) instead of a separate one-block launch in front of it
// ---------------------------------------------------------------------------------------------
struct Cand {
  int c;
  float thr_f;         // two operands: positions with |v^| >= thr_f can hold the maximum
  long long thr_z;     // one operand : positions with |z| (ReLU: z) >= thr_z can
};

__device__ void select_candidates(const qasr_stats* stats, const DynOp& op0, const DynOp& op1, int dual, int relu, int C,
                                  float* sh, int* s_count, Cand* s_list) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // largest |eps| over the 256 codes of each operand's input scale
  float e0 = 0.0f, e1 = 0.0f;
  for (int i = tid; i < 256; i += nt) {
    e0 = fmaxf(e0, (float)abs(noise_code(op0.in_unsigned ? i : (int)(int8_t)i, op0.s_x[0])));
    if (dual) e1 = fmaxf(e1, (float)abs(noise_code(op1.in_unsigned ? i : (int)(int8_t)i, op1.s_x[0])));
  }
  const float emax0 = block_max(e0, sh) * (1.0f / 4194304.0f);
  const float emax1 = dual ? block_max(e1, sh) * (1.0f / 4194304.0f) : 0.0f;
  if (tid == 0) *s_count = 0;
  float lmax = 0.0f;
  auto bounds = [&](int c, float& U, float& L, long long& thr_z, float& thr_f) {
    U = 0.0f; L = 0.0f; thr_z = 0x7fffffffffffffffLL; thr_f = FLT_MAX;
    if (!dual) {
      const long long lo = stats[c].lo, hi = stats[c].hi;
      if (hi < lo) return;
      const double d = (double)op0.wabs[c] * (double)emax0;     // |delta| <= sum|w| * max|eps|
      const float sc = op0.ch[c].s_out;
      const double a = relu ? (double)hi : fmax(fabs((double)lo), fabs((double)hi));
      U = __fmul_rn(__double2float_ru(fmax(a + d, 0.0)), sc);    // f32(.) and * s are monotone
      L = __fmul_rn(__double2float_rd(fmax(a - d, 0.0)), sc);
      thr_z = (long long)floor(a - 2.0 * d) - 1;
    } else {
      const unsigned klo = reinterpret_cast<const unsigned*>(&stats[c].lo)[0];
      const unsigned khi = reinterpret_cast<const unsigned*>(&stats[c].hi)[0];
      if (khi < klo) return;
      const float a = fmaxf(fabsf(qasr_key_f32(klo)), fabsf(qasr_key_f32(khi)));
      const float sc0 = op0.ch[c].s_out, sc1 = op1.ch[c].s_out;
      // |v - v^|: each operand's float differs by at most (|delta_i| + ulp(A_i) / 2) * s_i plus the rounding of the
      // product, the sum adds one more rounding.  With |A_i| <= sum|w| * 255 + |b_i| =: Amax_i and Y_i = Amax_i * s_i:
      //   |v - v^| <= 2 (d0 s0 + d1 s1) + (2^-23 + 2^-21) (Y0 + Y1)        (generous constants)
      const double y0 = ((double)op0.wabs[c] * 255.0 + fabs((double)op0.ch[c].bias)) * sc0;
      const double y1 = ((double)op1.wabs[c] * 255.0 + fabs((double)op1.ch[c].bias)) * sc1;
      const double e = 2.0 * ((double)op0.wabs[c] * emax0 * sc0 + (double)op1.wabs[c] * emax1 * sc1) +
                       (1.1920929e-7 + 4.76837158e-7) * (y0 + y1);
      U = __double2float_ru((double)a + e);
      L = __double2float_rd(fmax((double)a - e, 0.0));
      thr_f = __double2float_rd((double)a - 2.0 * e - 1e-30);
    }
  };
  for (int c = tid; c < C; c += nt) {
    float U, L, tf;
    long long tz;
    bounds(c, U, L, tz, tf);
    lmax = fmaxf(lmax, L);
  }
  const float lstar = block_max(lmax, sh);
  for (int c = tid; c < C; c += nt) {
    float U, L, tf;
    long long tz;
    bounds(c, U, L, tz, tf);
    if (U > 0.0f && U >= lstar) {
      const int slot = atomicAdd(s_count, 1);
      if (slot < QASR_DYN_MAX_CAND) { s_list[slot].c = c; s_list[slot].thr_z = tz; s_list[slot].thr_f = tf; }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// refine: candidate channels again, position by position; the noise term only where the position can win
// ---------------------------------------------------------------------------------------------
// generic scalar conv of one output position (any k / stride / dilation, depthwise or dense)
__device__ __forceinline__ int conv_at(const DynOp& op, const int8_t* wrow, int c, int b, int t) {
  int acc = 0;
  const int tin0 = t * op.stride - op.pad;
  const int nci = op.depthwise ? 1 : op.C_in;
  for (int ci = 0; ci < nci; ++ci) {
    const uint8_t* row = op.x + ((size_t)b * op.C_in + (op.depthwise ? c : ci)) * op.x_ld;
    for (int j = 0; j < op.k; ++j) {
      const int ti = tin0 + j * op.dil;
      if (ti >= 0 && ti < op.T_in) {
        const unsigned code = row[ti];
        acc += (int)wrow[ci * op.k + j] * (op.in_unsigned ? (int)code : (int)(int8_t)code);
      }
    }
  }
  return acc;
}
__device__ __forceinline__ int noise_at(const DynOp& op, const int8_t* wrow, const int8_t* lut, int c, int b, int t) {
  int dacc = 0;
  const int tin0 = t * op.stride - op.pad;
  const int nci = op.depthwise ? 1 : op.C_in;
  for (int ci = 0; ci < nci; ++ci) {
    const uint8_t* row = op.x + ((size_t)b * op.C_in + (op.depthwise ? c : ci)) * op.x_ld;
    for (int j = 0; j < op.k; ++j) {
      const int ti = tin0 + j * op.dil;
      if (ti >= 0 && ti < op.T_in) dacc += (int)wrow[ci * op.k + j] * (int)lut[row[ti]];
    }
  }
  return dacc;
}
// the same sum spread over the whole block (the taps of one output position are hundreds of different cache lines:
// one thread walking them alone costs ~100 us); every thread returns the block-wide total
__device__ int noise_block(const DynOp& op, const int8_t* wrow, const int8_t* lut, int c, int b, int t, int* red) {
  const int n = op.depthwise ? op.k : op.C_in * op.k;
  const int tin0 = t * op.stride - op.pad;
  int part = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = op.depthwise ? c : i / op.k, j = op.depthwise ? i : i - (i / op.k) * op.k;
    const int ti = tin0 + j * op.dil;
    if (ti >= 0 && ti < op.T_in) part += (int)wrow[i] * (int)lut[op.x[((size_t)b * op.C_in + ci) * op.x_ld + ti]];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  int tot = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  return tot;
}

// four consecutive output frames of a 1x1 stride-1 conv over input channels [ci0, ci1): one 32-bit load per channel
__device__ __forceinline__ void conv4_pointwise(const DynOp& op, const int8_t* wrow, int b, int t4, int ci0, int ci1,
                                                int (&acc)[4]) {
  acc[0] = acc[1] = acc[2] = acc[3] = 0;
  const uint8_t* p = op.x + ((size_t)b * op.C_in + ci0) * op.x_ld + t4;
  if (op.in_unsigned) {
#pragma unroll 8
    for (int ci = ci0; ci < ci1; ++ci, p += op.x_ld) {
      const uint32_t w4 = *reinterpret_cast<const uint32_t*>(p);
      const int w = wrow[ci];
      acc[0] += w * (int)(w4 & 0xff);
      acc[1] += w * (int)((w4 >> 8) & 0xff);
      acc[2] += w * (int)((w4 >> 16) & 0xff);
      acc[3] += w * (int)(w4 >> 24);
    }
  } else {
#pragma unroll 8
    for (int ci = ci0; ci < ci1; ++ci, p += op.x_ld) {
      const uint32_t w4 = *reinterpret_cast<const uint32_t*>(p);
      const int w = wrow[ci];
      acc[0] += w * (int)(int8_t)(w4 & 0xff);
      acc[1] += w * (int)(int8_t)((w4 >> 8) & 0xff);
      acc[2] += w * (int)(int8_t)((w4 >> 16) & 0xff);
      acc[3] += w * ((int)w4 >> 24);
    }
  }
}

__device__ __forceinline__ float literal_y(long long z, int dacc, float s_out) {
  // f32(f64(acc) + bias + delta) * s_out  (quant_modules.py:305-307; .type(torch.float) of the float64 conv)
  return __fmul_rn((float)((double)z + (double)dacc * (1.0 / 4194304.0)), s_out);
}

__global__ void dyn_refine_kernel(const qasr_stats* stats, const DynOp op0, const DynOp op1, int dual, int relu,
                                  const int32_t* out_lens, int B, int C_out, int T_out, int32_t* cand_out,
                                  uint32_t* absmax_bits) {
  extern __shared__ int8_t wsm[];    // weights of the candidate channel: [n0][n1]
  __shared__ int8_t lut[2][256];
  __shared__ float sh[kRefThreads / 32];
  __shared__ int s_count;
  __shared__ Cand s_list[QASR_DYN_MAX_CAND];
  __shared__ int part[2][8][32][4];
  __shared__ int s_nhits, s_red[kRefThreads / 32];
  __shared__ int4 s_hits[kMaxHits];   // (b, t, acc0, acc1) of the positions that can hold the maximum
  const int tid = threadIdx.x;
  select_candidates(stats, op0, op1, dual, relu, C_out, sh, &s_count, s_list);
  const int ncand = min(s_count, QASR_DYN_MAX_CAND);
  if (blockIdx.x == 0) {
    if (tid < QASR_DYN_MAX_CAND) cand_out[tid] = tid < ncand ? s_list[tid].c : -1;
    if (tid == 0) cand_out[QASR_DYN_MAX_CAND] = s_count;
  }
  if (ncand == 0) return;
  for (int i = tid; i < 256; i += kRefThreads) {
    lut[0][i] = (int8_t)noise_code(op0.in_unsigned ? i : (int)(int8_t)i, op0.s_x[0]);
    lut[1][i] = dual ? (int8_t)noise_code(op1.in_unsigned ? i : (int)(int8_t)i, op1.s_x[0]) : 0;
  }
  const int n0 = op0.depthwise ? op0.k : op0.C_in * op0.k;
  const int n1 = dual ? (op1.depthwise ? op1.k : op1.C_in * op1.k) : 0;
  const bool pw0 = !op0.depthwise && op0.k == 1 && op0.stride == 1 && op0.pad == 0;
  const bool pw1 = !dual || (!op1.depthwise && op1.k == 1 && op1.stride == 1 && op1.pad == 0);
  const int tq = (T_out + 3) >> 2;
  const long long groups = (long long)B * tq;
  float best = 0.0f;
  for (int ic = 0; ic < ncand; ++ic) {
    const Cand cd = s_list[ic];
    const int c = cd.c;
    __syncthreads();
    for (int i = tid; i < n0; i += kRefThreads) wsm[i] = op0.w[(size_t)c * n0 + i];
    for (int i = tid; i < n1; i += kRefThreads) wsm[n0 + i] = op1.w[(size_t)c * n1 + i];
    __syncthreads();
    const qasr_chan p0 = op0.ch[c];
    const qasr_chan p1 = dual ? op1.ch[c] : p0;
    const long long b0 = p0.bias == 2147483647 ? 2147483648LL : (long long)p0.bias;
    const long long b1 = p1.bias == 2147483647 ? 2147483648LL : (long long)p1.bias;
    if (tid == 0) s_nhits = 0;
    __syncthreads();
    // can this position hold the channel's maximum once the noise is added?  (thresholds from select_candidates)
    auto qualifies = [&](int acc0, int acc1) -> bool {
      const long long z0 = (long long)acc0 + b0;
      if (!dual) return (relu ? z0 : (z0 < 0 ? -z0 : z0)) >= cd.thr_z;
      const long long z1 = (long long)acc1 + b1;
      return fabsf(__fadd_rn(__fmul_rn((float)z1, p1.s_out), __fmul_rn((float)z0, p0.s_out))) >= cd.thr_f;
    };
    // |x_act| of one position with the reference's float noise (quant_modules.py:301-309, :108)
    auto literal = [&](int acc0, int acc1, int d0, int d1) -> float {
      float y = literal_y((long long)acc0 + b0, d0, p0.s_out);
      if (dual) y = __fadd_rn(literal_y((long long)acc1 + b1, d1, p1.s_out), y);      // identity + x
      else if (relu) y = fmaxf(y, 0.0f);                                              // jasper.py:661 / :687
      return fabsf(y);
    };
    // rare: park the position for the block-wide noise pass; a full list falls back to one thread walking the taps
    auto judge = [&](int b, int t, int acc0, int acc1) -> float {
      if (!qualifies(acc0, acc1)) return 0.0f;
      const int slot = atomicAdd(&s_nhits, 1);
      if (slot < kMaxHits) { s_hits[slot] = make_int4(b, t, acc0, acc1); return 0.0f; }
      return literal(acc0, acc1, noise_at(op0, wsm, lut[0], c, b, t), dual ? noise_at(op1, wsm + n0, lut[1], c, b, t) : 0);
    };
    if (pw0 && pw1) {
      // pointwise operands: 32 position groups (4 frames each) per round, the input channels split over the 8 warps,
      // partial sums meet in shared memory
      const int lane = tid & 31, warp = tid >> 5;
      const int cps0 = (op0.C_in + 7) >> 3, cps1 = dual ? (op1.C_in + 7) >> 3 : 0;
      for (long long gbase = (long long)blockIdx.x * 32; gbase < groups; gbase += (long long)gridDim.x * 32) {
        const long long gi = gbase + lane;
        const bool live = gi < groups;
        const int b = live ? (int)(gi / tq) : 0, t4 = live ? 4 * (int)(gi - (long long)b * tq) : 0;
        const int len = out_lens ? min(out_lens[b], T_out) : T_out;
        int a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
        if (live && t4 < len) {
          conv4_pointwise(op0, wsm, b, t4, min(warp * cps0, op0.C_in), min((warp + 1) * cps0, op0.C_in), a0);
          if (dual) conv4_pointwise(op1, wsm + n0, b, t4, min(warp * cps1, op1.C_in), min((warp + 1) * cps1, op1.C_in), a1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { part[0][warp][lane][u] = a0[u]; part[1][warp][lane][u] = a1[u]; }
        __syncthreads();
        if (warp == 0 && live && t4 < len) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int s0 = 0, s1 = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { s0 += part[0][w][lane][u]; s1 += part[1][w][lane][u]; }
            if (t4 + u < len) best = fmaxf(best, judge(b, t4 + u, s0, s1));
          }
        }
        __syncthreads();
      }
    } else {
      // any other conv (depthwise k = 33..87, strided, dilated): one output frame per thread
      const long long total = (long long)B * T_out;
      for (long long pos = blockIdx.x * (long long)kRefThreads + tid; pos < total; pos += (long long)gridDim.x * kRefThreads) {
        const int b = (int)(pos / T_out), t = (int)(pos - (long long)b * T_out);
        if (out_lens && t >= out_lens[b]) continue;
        best = fmaxf(best, judge(b, t, conv_at(op0, wsm, c, b, t), dual ? conv_at(op1, wsm + n0, c, b, t) : 0));
      }
    }
    __syncthreads();
    const int nh = min(s_nhits, kMaxHits);
    for (int h = 0; h < nh; ++h) {
      const int4 hit = s_hits[h];
      const int d0 = noise_block(op0, wsm, lut[0], c, hit.x, hit.y, s_red);
      const int d1 = dual ? noise_block(op1, wsm + n0, lut[1], c, hit.x, hit.y, s_red) : 0;
      if (tid == 0) best = fmaxf(best, literal(hit.z, hit.w, d0, d1));
    }
  }
  best = block_max(best, sh);
  if (tid == 0 && best > 0.0f) atomicMax(absmax_bits, __float_as_uint(best));
}

// ---------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dyadic_dev(float pre, float zsf, int& m, int& e) {   // batch_frexp, quant_utils.py:121-147
  const double r = (double)pre / (double)zsf;
  int ex;
  const double mant = frexp(r, &ex);
  long long mm = (long long)floor(mant * 2147483648.0 + 0.5);
  int ee = 31 - ex;
  if (mm >= 2147483648LL) { mm >>= 1; ee -= 1; }
  if (!(r > 0.0) || !isfinite(r)) { mm = 0; ee = 1; }
  m = (int)mm;
  e = ee;
}

struct DynFin {
  int src_kind;
  void* src;
  const float* src_scale;
  int n_levels;
  float* act_scale;
  int n_prod;
  qasr_chan* prod[2];
  int prod_C[2];
  const float* prev_scale;
  int32_t* prev_me;
  int n_cons;
  qasr_chan* cons[2];
  const float* cons_sw[2];
  const float* cons_bias[2];
  int cons_C[2];
  unsigned long long* const* peer_slots;
  int rank, world, slot;
  uint32_t* seq;
  const int32_t* prod_wabs[2];
  int prod_qabs[2];
  qasr_stats* stats;
  int stats_C, stats_dual, phase;
};

// May the all-integer epilogue ((z*m + 2^(e-1)) >> e) replace the literal float recipe for |z| <= bound?  Same proof as
// qasr/plan.py fast_ok(): the float32 round trip of the accumulator is the identity (bound < 2^22), the shift is in the
// fast path's range, the result fits, and no reachable z lands on an exact .5 tie.
__device__ __forceinline__ bool fast_ok_dev(long long bound, int m, int e) {
  if (bound >= (1 << 22)) return false;
  if (m == 0) return true;
  if (e < 32 || e > 62) return false;
  if ((((unsigned long long)bound * (unsigned long long)m) >> e) >= (1ull << 30)) return false;
  const int t = e - 1 - (__ffs(m) - 1);
  if (t < 0) return false;
  return t >= 62 || (1ll << t) > bound;
}

__global__ void dyn_finalize_kernel(const DynFin a) {
  __shared__ float s_scale;
  const bool post = a.world <= 1 || a.phase != 2, finish = a.world <= 1 || a.phase != 1;
  if (threadIdx.x == 0) {
    float absmax = 0.0f;
    if (post) {
      if (a.src_kind == 0) {
        uint32_t* p = static_cast<uint32_t*>(a.src);
        absmax = __uint_as_float(*p);
        *p = 0u;
      } else if (a.src_kind == 1) {
        const float* p = static_cast<const float*>(a.src);
        absmax = fmaxf(fabsf(p[0]), fabsf(p[1]));
      } else {
        int32_t* p = static_cast<int32_t*>(a.src);
        const int lo = p[0], hi = p[1];
        absmax = hi >= lo ? fmaxf(fabsf(__fmul_rn((float)lo, a.src_scale[0])), fabsf(__fmul_rn((float)hi, a.src_scale[0]))) : 0.0f;
        p[0] = 2147483647;
        p[1] = -2147483647 - 1;
      }
    }
    if (a.world > 1) {
      // one-shot all-reduce(MAX) over peer memory: write (sequence, value) into this rank's slot of EVERY rank's
      // table, then wait until all ranks' entries of this sequence number have arrived here.  A rank can run ahead
      // of a peer by less than one forward (every exchange waits for all ranks), so the sequence number in the
      // entry tells this forward's value from the previous one.
      if (post) {
        const uint32_t seq = ++(*a.seq);
        const unsigned long long word = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(absmax);
        for (int r = 0; r < a.world; ++r) {
          unsigned long long* dst = a.peer_slots[r] + (size_t)a.slot * a.world + a.rank;
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
        }
      }
      if (finish) {
        const uint32_t seq = *a.seq;
        const unsigned long long* mine = a.peer_slots[a.rank] + (size_t)a.slot * a.world;
        absmax = 0.0f;
        for (int r = 0; r < a.world; ++r) {
          unsigned long long v;
          unsigned spin = 0;
          do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + r) : "memory");
            if (++spin > (1u << 28)) __trap();   // a lost peer must not hang the GPU
          } while ((uint32_t)(v >> 32) != seq);
          absmax = fmaxf(absmax, __uint_as_float((uint32_t)v));
        }
      }
    }
    if (!finish) return;
    // symmetric_linear_quantization_params, quant_utils.py:51-52: clamp(max(|min|, |max|), 1e-8) / n
    const float s = __fdiv_rn(fmaxf(absmax, 1e-8f), (float)a.n_levels);
    a.act_scale[0] = s;
    s_scale = s;
    if (a.prev_me != nullptr) {
      int m, e;
      dyadic_dev(a.prev_scale[0], s, m, e);
      a.prev_me[0] = m;
      a.prev_me[1] = e;
    }
  }
  if (!finish) return;
  __syncthreads();
  const float s = s_scale;
  for (int i = 0; i < a.n_prod; ++i) {
    for (int c = threadIdx.x; c < a.prod_C[i]; c += blockDim.x) {
      int m, e;
      dyadic_dev(a.prod[i][c].s_out, s, m, e);
      const int bias = a.prod[i][c].bias;
      bool fast = false;
      if (a.prod_wabs[i] != nullptr && bias != 2147483647) {
        const long long bound = (long long)a.prod_wabs[i][c] * a.prod_qabs[i] + (bias < 0 ? -(long long)bias : (long long)bias);
        fast = fast_ok_dev(bound, m, e);
      }
      a.prod[i][c].m = m;
      a.prod[i][c].e = fast ? e : (int)((unsigned)e | 0x80000000u);   // bit 31: literal recipe for this channel
    }
  }
  if (a.stats != nullptr) {   // statistics records back to their reset state for the next forward
    for (int c = threadIdx.x; c < a.stats_C; c += blockDim.x) {
      a.stats[c].lo = a.stats_dual ? 0xffffffffLL : 0x7fffffffffffffffLL;
      a.stats[c].hi = a.stats_dual ? 0LL : -0x7fffffffffffffffLL - 1;
    }
  }
  for (int i = 0; i < a.n_cons; ++i) {
    for (int c = threadIdx.x; c < a.cons_C[i]; c += blockDim.x) {
      const float so = __fmul_rn(a.cons_sw[i][c], s);                          // quant_modules.py:293
      int bias = 0;
      if (a.cons_bias[i] != nullptr) {
        // SymmetricQuantFunction(bias, 32, s_out), float32: clamp(rne((1 / s_out) * b), -(2^31 - 1), 2^31 - 2), both
        // bounds rounding to +-2^31 in float32 (quant_modules.py:296-297, quant_utils.py:74-76)
        float q = rintf(__fmul_rn(__fdiv_rn(1.0f, so), a.cons_bias[i][c]));
        q = fminf(fmaxf(q, -2147483648.0f), 2147483648.0f);
        bias = q >= 2147483648.0f ? 2147483647 : (int)q;                       // INT32_MAX = sentinel of 2^31
      }
      a.cons[i][c].s_out = so;
      a.cons[i][c].bias = bias;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// int8 statistics / requantisation of the block outputs
// ---------------------------------------------------------------------------------------------
__global__ void minmax_i8_kernel(const int8_t* __restrict__ codes, int ld, int relu, int B, int C, int T,
                                 const int32_t* lens, int32_t* out2) {
  int lo = 2147483647, hi = -2147483647 - 1;
  const int tq = (T + 15) / 16;
  const long long total = (long long)B * C * tq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % tq);
    const long long row = i / tq;
    const int b = (int)(row / C);
    const int len = lens ? min(lens[b], T) : T;
    const uint4 v = *reinterpret_cast<const uint4*>(codes + row * ld + 16 * j);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = 16 * j + 4 * g + u;
        if (t < T) {
          int q = t < len ? (int)(int8_t)((w[g] >> (8 * u)) & 0xff) : 0;     // masked frames count as 0 (jasper.py:177-181)
          if (relu) q = max(q, 0);
          lo = min(lo, q);
          hi = max(hi, q);
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && hi >= lo) {
    atomicMin(out2, lo);
    atomicMax(out2 + 1, hi);
  }
}

__global__ void requant_i8_dev_kernel(const int8_t* __restrict__ in, int in_ld, int relu, const int32_t* me, int qmin,
                                      int qmax, int8_t* out, int out_ld, int B, int C, int T, const int32_t* lens) {
  __shared__ int8_t lut[256];
  {
    int code = (int)(int8_t)threadIdx.x;
    if (relu) code = max(code, 0);
    double o = rint(__dmul_rn((double)code, scalbn((double)me[0], -me[1])));   // quant_utils.py:195-196
    o = fmin(fmax(o, (double)qmin), (double)qmax);
    lut[threadIdx.x] = (int8_t)(int)o;
  }
  __syncthreads();
  const int tq = (T + 15) / 16;
  const long long total = (long long)B * C * tq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % tq);
    const long long row = i / tq;
    const int b = (int)(row / C);
    const int len = lens ? min(lens[b], T) : T;
    const uint4 v = *reinterpret_cast<const uint4*>(in + row * in_ld + 16 * j);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t r[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t acc = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = 16 * j + 4 * g + u;
        const uint32_t code = (w[g] >> (8 * u)) & 0xff;
        const uint32_t q = t < len ? (uint8_t)lut[code] : 0u;
        acc |= q << (8 * u);
      }
      r[g] = acc;
    }
    *reinterpret_cast<uint4*>(out + row * out_ld + 16 * j) = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

int check_operand(const qasr_dyn_operand* o, const char* what) {
  QASR_CHECK_ARG(o->x && o->w && o->s_x && o->ch && o->wabs, "%s: null pointer in operand", what);
  QASR_CHECK_ARG(o->C_in > 0 && o->T_in > 0 && o->k > 0 && o->stride >= 1 && o->dilation >= 1 && o->pad >= 0 &&
                 o->x_ld >= o->T_in, "%s: bad operand shape", what);
  return QASR_OK;
}

}  // namespace

extern "C" int qasr_dyn_refine(const qasr_stats* stats, const qasr_dyn_operand* op0, const qasr_dyn_operand* op1,
                               int32_t relu, const int32_t* out_lens, int32_t B, int32_t C_out, int32_t T_out, int32_t* cand,
                               uint32_t* absmax_bits, void* stream) {
  QASR_CHECK_ARG(stats && op0 && cand && absmax_bits && B > 0 && C_out > 0 && T_out > 0, "dyn_refine: bad arguments");
  int rc = check_operand(op0, "dyn_refine");
  if (rc) return rc;
  if (op1 && (rc = check_operand(op1, "dyn_refine"))) return rc;
  const int n0 = op0->depthwise ? op0->k : op0->C_in * op0->k;
  const int n1 = op1 ? (op1->depthwise ? op1->k : op1->C_in * op1->k) : 0;
  if (n0 + n1 > kMaxWRow) {
    qasr_set_error("dyn_refine: %d weights per output channel exceed the %d the kernel keeps in shared memory", n0 + n1, kMaxWRow);
    return QASR_ERR_UNSUPPORTED;
  }
  const bool pointwise = !op0->depthwise && op0->k == 1 && op0->stride == 1 && op0->pad == 0 &&
                         (!op1 || (!op1->depthwise && op1->k == 1 && op1->stride == 1 && op1->pad == 0));
  // per block and round: 32 groups of 4 frames (pointwise), else one frame per thread
  long long blocks = pointwise ? ((long long)B * ((T_out + 3) / 4) + 31) / 32 : ((long long)B * T_out + kRefThreads - 1) / kRefThreads;
  const long long cap = 4LL * qasr_sm_count();
  if (blocks > cap) blocks = cap;
  dyn_refine_kernel<<<(unsigned)blocks, kRefThreads, (size_t)(n0 + n1), (cudaStream_t)stream>>>(
      stats, make_op(op0), make_op(op1), op1 != nullptr, relu, out_lens, B, C_out, T_out, cand, absmax_bits);
  return qasr_check_launch("dyn_refine_kernel");
}

extern "C" int qasr_dyn_finalize(const qasr_dyn_finalize_args* a, void* stream) {
  QASR_CHECK_ARG(a && a->src && a->act_scale && a->n_levels > 0, "dyn_finalize: bad arguments");
  QASR_CHECK_ARG(a->src_kind >= 0 && a->src_kind <= 2 && (a->src_kind != 2 || a->src_scale), "dyn_finalize: source");
  QASR_CHECK_ARG(a->n_prod >= 0 && a->n_prod <= 2 && a->n_cons >= 0 && a->n_cons <= 2, "dyn_finalize: table counts");
  QASR_CHECK_ARG((a->prev_me == nullptr) == (a->prev_scale == nullptr), "dyn_finalize: prev_scale and prev_me come together");
  QASR_CHECK_ARG(a->world <= 1 || (a->peer_slots && a->seq && a->rank >= 0 && a->rank < a->world && a->slot >= 0),
                 "dyn_finalize: data-parallel exchange needs peer_slots / seq / rank");
  DynFin f;
  memset(&f, 0, sizeof(f));
  f.src_kind = a->src_kind; f.src = a->src; f.src_scale = a->src_scale; f.n_levels = a->n_levels;
  f.act_scale = a->act_scale; f.n_prod = a->n_prod; f.prev_scale = a->prev_scale; f.prev_me = a->prev_me;
  f.n_cons = a->n_cons;
  for (int i = 0; i < 2; ++i) {
    f.prod[i] = a->prod[i]; f.prod_C[i] = a->prod_C[i];
    f.cons[i] = a->cons[i]; f.cons_sw[i] = a->cons_sw[i]; f.cons_bias[i] = a->cons_bias[i]; f.cons_C[i] = a->cons_C[i];
    QASR_CHECK_ARG(i >= a->n_prod || (a->prod[i] && a->prod_C[i] > 0), "dyn_finalize: producer table %d", i);
    QASR_CHECK_ARG(i >= a->n_cons || (a->cons[i] && a->cons_sw[i] && a->cons_C[i] > 0), "dyn_finalize: consumer table %d", i);
  }
  f.peer_slots = reinterpret_cast<unsigned long long* const*>(a->peer_slots);
  f.rank = a->rank; f.world = a->world; f.slot = a->slot; f.seq = a->seq;
  for (int i = 0; i < 2; ++i) { f.prod_wabs[i] = a->prod_wabs[i]; f.prod_qabs[i] = a->prod_qabs[i]; }
  f.stats = a->stats; f.stats_C = a->stats_C; f.stats_dual = a->stats_dual;
  f.phase = a->phase;
  QASR_CHECK_ARG(a->phase >= 0 && a->phase <= 2, "dyn_finalize: phase");
  dyn_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(f);
  return qasr_check_launch("dyn_finalize_kernel");
}

extern "C" int qasr_minmax_i8(const int8_t* codes, int32_t ld, int32_t relu, int32_t B, int32_t C, int32_t T,
                              const int32_t* lens, int32_t* out2, void* stream) {
  QASR_CHECK_ARG(codes && out2 && B > 0 && C > 0 && T > 0, "minmax_i8: bad arguments");
  QASR_CHECK_ARG(ld % 16 == 0 && ld >= T && qasr_aligned16(codes), "minmax_i8: pitch/alignment");
  long long blocks = ((long long)B * C * ((T + 15) / 16) + 255) / 256;
  const long long cap = 8LL * qasr_sm_count();
  if (blocks > cap) blocks = cap;
  minmax_i8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(codes, ld, relu, B, C, T, lens, out2);
  return qasr_check_launch("minmax_i8_kernel");
}

extern "C" int qasr_requant_i8_dev(const int8_t* in, int32_t in_ld, int32_t relu, const int32_t* me, int32_t qmin,
                                   int32_t qmax, int8_t* out, int32_t out_ld, int32_t B, int32_t C, int32_t T,
                                   const int32_t* lens, void* stream) {
  QASR_CHECK_ARG(in && out && me, "requant_i8_dev: null pointer");
  QASR_CHECK_ARG(in_ld % 16 == 0 && out_ld % 16 == 0 && in_ld >= T && out_ld >= T && qasr_aligned16(in) &&
                     qasr_aligned16(out), "requant_i8_dev: pitch/alignment");
  QASR_CHECK_ARG(qmin < qmax && qmin >= -256 && qmax <= 255, "requant_i8_dev: clamp range");
  long long blocks = ((long long)B * C * ((T + 15) / 16) + 255) / 256;
  const long long cap = 8LL * qasr_sm_count();
  if (blocks > cap) blocks = cap;
  requant_i8_dev_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, in_ld, relu, me, qmin, qmax, out, out_ld, B,
                                                                           C, T, lens);
  return qasr_check_launch("requant_i8_dev_kernel");
}

// ---------------------------------------------------------------------------------------------
// peer-mapped buffers (CUDA IPC) for the range exchange
// ---------------------------------------------------------------------------------------------
extern "C" int qasr_comm_alloc(int64_t bytes, void** ptr, void* handle_out) {
  QASR_CHECK_ARG(bytes > 0 && ptr && handle_out, "comm_alloc: bad arguments");
  QASR_CHECK_CUDA(cudaMalloc(ptr, (size_t)bytes));
  QASR_CHECK_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  QASR_CHECK_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, sizeof(h));
  return QASR_OK;
}
extern "C" int qasr_comm_open(const void* handle, void** ptr) {
  QASR_CHECK_ARG(handle && ptr, "comm_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  QASR_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QASR_OK;
}
extern "C" int qasr_comm_close(void* ptr) {
  QASR_CHECK_ARG(ptr, "comm_close: null pointer");
  QASR_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return QASR_OK;
}
extern "C" int qasr_comm_free(void* ptr) {
  QASR_CHECK_ARG(ptr, "comm_free: null pointer");
  QASR_CHECK_CUDA(cudaFree(ptr));
  return QASR_OK;
}

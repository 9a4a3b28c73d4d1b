// Shared host/device helpers of the qasr kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "qasr.h"

// ---------------------------------------------------------------------------------------------
// host: error plumbing (no exception crosses the C ABI)
// ---------------------------------------------------------------------------------------------
void qasr_set_error(const char* fmt, ...);

#define QASR_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      qasr_set_error(__VA_ARGS__);                                  \
      return QASR_ERR_INVALID;                                      \
    }                                                               \
  } while (0)

#define QASR_CHECK_CUDA(expr)                                                             \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      qasr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                     __LINE__);                                                           \
      return QASR_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

static inline int qasr_check_launch(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    qasr_set_error("launch of %s failed: %s", what, cudaGetErrorString(err));
    return QASR_ERR_CUDA;
  }
  return QASR_OK;
}

static inline bool qasr_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------
// device: the fused epilogue (see qasr.h for the reference lines each step restates)
// ---------------------------------------------------------------------------------------------
struct EpiArgs {  // by-value kernel-parameter copy of qasr_epilogue
  int mode, flags;
  const qasr_chan* ch0;
  const qasr_chan* ch1;
  const int8_t* prev;
  int prev_ld;
  int qmin, qmax, m2, e2, qmin2, qmax2;
  const int32_t* out_lens;
  const float* rf;       // per-channel float multipliers (0 = not proved) or nullptr
};

static inline EpiArgs make_epi_args(const qasr_epilogue* e) {
  EpiArgs a;
  a.mode = e->mode;
  a.flags = e->flags;
  a.ch0 = e->ch0;
  a.ch1 = e->ch1;
  a.prev = e->prev;
  a.prev_ld = e->prev_ld;
  a.qmin = e->qmin;
  a.qmax = e->qmax;
  a.m2 = e->m2;
  a.e2 = e->e2;
  a.qmin2 = e->qmin2;
  a.qmax2 = e->qmax2;
  a.out_lens = e->out_lens;
  a.rf = e->f32_mult;
  return a;
}

int qasr_validate_epilogue(const qasr_epilogue* e, bool allow_second_operand);
// Launch policy of the entry point that is running on this thread: true = plain stream order (no programmatic
// dependent launch), set from QASR_F_DEVICE_TABLES for every conv launch.
void qasr_set_plain_launch(bool plain);
bool qasr_plain_launch();
// In-graph kernel timing (qasr_profile_begin / qasr_profile_end): while a session is open every conv launch gets the
// next slot of the caller's stamp buffer; the kernel leaves there the %globaltimer of its first CTA's start on
// activation data and of its last CTA's end.  nullptr outside a session (the timed product graph).
unsigned long long* qasr_next_stamp();
long long qasr_debug_get(int key);   // bring-up knobs, see api.cu
extern "C" int qasr_debug_set(int key, long long value);

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: the hot kernels of a plan run back to back on one stream; each is launched
// with the "programmatic stream serialization" attribute, triggers its dependents at entry and waits for
// its predecessor (all of it, memory flushed) right before it first touches activation memory.  The
// next kernel's launch latency, barrier / TMEM set-up and weight loads then overlap this kernel's tail.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename Kern, typename... Args>
static inline cudaError_t qasr_launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (qasr_debug_get(9) == 0 || qasr_plain_launch()) ? 0 : 1;   // knob 9 = 0: plain launches
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
// same, as clusters of two CTAs (the two SMs of a TPC: tcgen05 cta_group::2 kernels)
template <typename Kern, typename... Args>
static inline cudaError_t qasr_launch_pdl_pair(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (qasr_debug_get(9) == 0 || qasr_plain_launch()) ? 0 : 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

__device__ __forceinline__ unsigned long long qasr_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void qasr_stamp_begin(unsigned long long* stamp) {
  if (stamp != nullptr) atomicMin(stamp, qasr_globaltimer());
}
__device__ __forceinline__ void qasr_stamp_end(unsigned long long* stamp) {
  if (stamp != nullptr) atomicMax(stamp + 1, qasr_globaltimer());
}

// m * 2^-e as a double: exact (m has 31 bits), so z * mult rounds exactly like the
// reference's f64(z) * f64(m) followed by the exact division by 2^e (quant_utils.py:195-196).
__device__ __forceinline__ double dyad_mult(int m, int e) {
  // 2^-e assembled from its exponent bits when 0 <= e <= 1022 (always, in practice): three instructions instead
  // of the generic scalbn; both are exact
  if ((unsigned)e <= 1022u) return (double)m * __longlong_as_double((long long)(1023 - e) << 52);
  return scalbn((double)m, -e);
}

// Per-channel registers of one conv operand.
struct ChanRegs {
  float s;
  int bias;
  int m, e;
  double mult;          // exact path: m * 2^-e
  int sh;               // fast path: e - 32 (host guarantees 32 <= e <= 62)
  long long add;        // fast path: bias * m + 2^(e-1)
  float r;              // float fast path: f32(m * 2^-e), or 0 when only the integer path is proved
  bool literal;         // device-written tables only (QASR_F_DEVICE_TABLES): the integer fast path is NOT proved for
                        // this channel (bit 31 of the record's e, set by qasr_dyn_finalize)
};

__device__ __forceinline__ ChanRegs load_chan(const qasr_chan* p, int c) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(p) + c);
  ChanRegs r;
  r.s = __int_as_float(v.x);
  r.bias = v.y;
  r.m = v.z;
  r.e = v.w & 0x7fffffff;
  r.literal = v.w < 0;
  r.mult = dyad_mult(r.m, r.e);
  r.sh = max(r.e - 32, 0);
  r.add = (long long)r.bias * (long long)r.m + ((r.e >= 1 && r.e <= 62) ? (1ll << (r.e - 1)) : 0ll);
  r.r = 0.0f;
  return r;
}
// conv-0 parameters with the float-path choice of this launch: every channel (QASR_F_FAST_F32), the proved
// channels of the optional table, or none
__device__ __forceinline__ ChanRegs load_chan0(const EpiArgs& ep, int c) {
  ChanRegs r = load_chan(ep.ch0, c);
  if (ep.flags & QASR_F_FAST_F32) r.r = __double2float_rn(dyad_mult(r.m, r.e));
  else if (ep.rf != nullptr) r.r = __ldg(ep.rf + c);
  return r;
}

// All-integer fast path of one conv operand: rne((acc + bias) * m / 2^e) as
// ((acc * m + (bias * m + 2^(e-1))) >> 32) >> (e - 32): one IMAD.WIDE + one shift.  Valid when the host
// proved (qasr/plan.py fast_ok): |acc + bias| < 2^22 (the float32 round trip of quant_modules.py:305 /
// quant_utils.py:187 is the identity and the product is exact in double), 32 <= e <= 62, no in-range
// value lands on an exact .5 tie (half-up == half-to-even), result fits int32.
__device__ __forceinline__ int contrib_fast(int acc, const ChanRegs& c) {
  const long long p = (long long)acc * (long long)c.m + c.add;
  return (int)(p >> 32) >> c.sh;
}

// Literal restatement of the float recipe (any accumulator size, ties, huge biases).
__device__ __forceinline__ double contrib_exact(int acc, const ChanRegs& c, bool relu, bool exact_f32,
                                                bool bias64) {
  double z;
  if (exact_f32) {
    // INT32_MAX is the sentinel for b_int == 2^31: the reference's float32 clamp bound n-1 =
    // 2^31-2 rounds to 2^31 (quant_utils.py:76 with k = 32), which no int32 can hold
    const long long b64 = c.bias == 2147483647 ? 2147483648LL : (long long)c.bias;
    float a = bias64 ? (float)((long long)acc + b64) : (float)(acc + c.bias);
    float y = __fmul_rn(a, c.s);                 // quant_modules.py:307
    if (relu) y = fmaxf(y, 0.0f);                // jasper.py:661
    z = (double)rintf(__fdiv_rn(y, c.s));        // quant_utils.py:187
  } else {
    int zi = acc + c.bias;
    if (relu) zi = max(zi, 0);
    z = (double)zi;
  }
  return rint(__dmul_rn(z, c.mult));             // quant_utils.py:195-196
}

// Kernel-uniform constants of the REQUANT epilogue, derived once per thread.
constexpr float kF32Magic = 12582912.0f;   // 1.5 * 2^23: float(kF32Magic + n) has n in its low mantissa bits

struct EpiConsts {
  int lo1, hi1, lo2, hi2;   // clamps with the ReLUs folded in (fast path)
  long long half2;
  double mult2;
  float flo, fhi;           // float fast path: kF32Magic + lo1 / hi1
  uint32_t hi16;            // QASR_F_CLAMP16: hi1 in both 16-bit halves
  int m16;                  // QASR_F_STAGE2_16: rne(m2 * 2^(16 - e2))
  uint32_t hi16_2;          // QASR_F_STAGE2_16: hi2 in both 16-bit halves
};

__device__ __forceinline__ EpiConsts make_epi_consts(const EpiArgs& ep, bool has1) {
  EpiConsts k;
  // ReLU commutes with the monotone requantisation: relu(z) -> clamp lower bound 0 (single operand),
  // block ReLU after the first clamp -> lower bound 0 as well.
  const bool relu_lo = (ep.flags & QASR_F_RELU1) || ((ep.flags & QASR_F_RELU0) && !has1 && ep.prev == nullptr);
  k.lo1 = relu_lo ? max(ep.qmin, 0) : ep.qmin;
  k.hi1 = ep.qmax;
  k.lo2 = ep.qmin2;
  k.hi2 = ep.qmax2;
  k.half2 = (ep.e2 >= 1 && ep.e2 <= 62) ? (1ll << (ep.e2 - 1)) : 0ll;
  k.mult2 = dyad_mult(ep.m2, ep.e2);
  k.flo = kF32Magic + (float)k.lo1;
  k.fhi = kF32Magic + (float)k.hi1;
  k.hi16 = ((uint32_t)k.hi1 & 0xffffu) * 0x10001u;
  k.m16 = (ep.e2 >= 17 && ep.e2 <= 62) ? (int)(((long long)ep.m2 + (1ll << (ep.e2 - 17))) >> (ep.e2 - 16)) : 0;
  k.hi16_2 = ((uint32_t)k.hi2 & 0xffffu) * 0x10001u;
  return k;
}

// Full per-element REQUANT pipeline -> final 8-bit code (as int), all-integer fast path.  The
// operand / stage structure is a compile-time property of the kernel instantiation.
template <bool DUAL, bool PREV, bool STAGE2, bool RC = false>
__device__ __forceinline__ int epilogue_fast(const EpiArgs& ep, const EpiConsts& k, int acc0, const ChanRegs& c0,
                                             int acc1, const ChanRegs& c1, int prev) {
  int o = contrib_fast(acc0, c0);
  if (DUAL) o += contrib_fast(acc1, c1);
  if (RC) {   // QASR_F_RANGE_CHECK: a separate instantiation (its doubles cost the plain fast path registers)
    // proved for |acc + bias| < 2^22 only: the (practically never seen) elements beyond go through the literal recipe
    const bool big0 = (unsigned)(acc0 + c0.bias + (1 << 22)) >= (1u << 23);
    const bool big1 = DUAL && (unsigned)(acc1 + c1.bias + (1 << 22)) >= (1u << 23);
    if (big0 || big1) {
      o = (int)contrib_exact(acc0, c0, (ep.flags & QASR_F_RELU0) != 0, true, false);
      if (DUAL) o += (int)contrib_exact(acc1, c1, false, true, false);
    }
  }
  if (PREV) o += prev;
  int q = min(max(o, k.lo1), k.hi1);
  if (STAGE2) {
    const long long p = (long long)q * (long long)ep.m2 + k.half2;
    q = (int)(p >> ep.e2);
    q = min(max(q, k.lo2), k.hi2);
  }
  return q;
}

// QASR_F_STAGE2_16: the second requantisation of FOUR outputs in 16-bit fixed point.  After the first clamp and the block
// ReLU the code q lies in [0, hi1] (hi1 <= 127), so rne(q * m2 / 2^e2) has at most 128 arguments: the host checks
// EXHAUSTIVELY (qasr/plan.py stage2_16_ok) that (q * M16 + 2^15) >> 16 with M16 = rne(m2 * 2^(16 - e2)) gives the same
// value for every one of them.  One IMAD per output on the FMA pipe; the shift by 16 and the pairing are ONE PRMT per two
// outputs (the high halves), the upper clamp one VIMNMX.S16x2 per two outputs (q2 >= 0 already): 1.25 ALU-pipe
// instructions per output instead of IMAD.WIDE + 64-bit shift + 2 VIMNMX + 0.75 PRMT.
template <bool DUAL, bool PREV, bool RC>
__device__ __forceinline__ uint32_t epilogue_fast_s16x4(const EpiArgs& ep, const EpiConsts& k, const int* a0, const ChanRegs& c0,
                                                        const int* a1, const ChanRegs& c1, uint32_t prev4) {
  int p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int prev = PREV ? (int)(int8_t)(prev4 >> (8 * i)) : 0;
    int q;   // first clamp, [0, hi1] by the flag's contract: add + min + ReLU are ONE VIADDMNMX.RELU
    if (RC) {
      q = epilogue_fast<DUAL, PREV, false, true>(ep, k, a0[i], c0, DUAL ? a1[i] : 0, c1, prev);
    } else {
      int o = contrib_fast(a0[i], c0);
      if (PREV) o += prev;
      q = DUAL ? __viaddmin_s32_relu(o, contrib_fast(a1[i], c1), k.hi1) : __vimin_s32_relu(o, k.hi1);
    }
    p[i] = q * k.m16 + 0x8000;
  }
  const uint32_t q01 = __vmins2(__byte_perm((uint32_t)p[0], (uint32_t)p[1], 0x7632), k.hi16_2);
  const uint32_t q23 = __vmins2(__byte_perm((uint32_t)p[2], (uint32_t)p[3], 0x7632), k.hi16_2);
  return __byte_perm(q01, q23, 0x6420);
}

// Float fast path (QASR_F_FAST_F32, single operand, single stage): I2F + FFMA + 2 FMNMX on the 128-lane FP32
// pipe instead of IMAD.HI (32 lanes/clk/SM measured on B200) + SHF + 2 VIMNMX.  The FMA rounds the exact
// product plus 1.5 * 2^23 once, to an integer (ties to even); the code is the low byte of the result's bit
// pattern, so the caller packs the returned words directly.  Equality with the dyadic for every reachable
// accumulator and this clamp is proved on the host (qasr/plan.py f32_ok).
__device__ __forceinline__ uint32_t epilogue_f32_bits(const EpiConsts& k, int acc, const ChanRegs& c) {
  // int -> float without the half-rate I2F: |acc + bias| < 2^22 (host proof), so 0x4B400000 + (acc + bias) is the
  // bit pattern of 1.5 * 2^23 + (acc + bias) and one FADD recovers the exact float; the constant
  // bias + 0x4B400000 is loop invariant (one IADD + FADD per output instead of IADD + I2F)
  const float f = __int_as_float(acc + (c.bias + 0x4B400000)) - kF32Magic;
  const float y = fmaf(f, c.r, kF32Magic);
  return __float_as_uint(fminf(fmaxf(y, k.flo), k.fhi));
}

// QASR_F_CLAMP16: the float fast path of FOUR outputs with the clamp done on packed 16-bit lanes.  The FMA result's bit
// pattern is 0x4B400000 + n (n = the requantised integer before the clamp), and the low half of 0x4B400000 is zero, so
// the low 16 bits ARE n as an int16 whenever |n| < 2^15 -- which the host proves for every reachable accumulator
// (qasr/plan.py clamp16_ok; the window argument does not apply here: an accumulator beyond the clamp window must still
// not wrap).  One PRMT gathers two low halves, one VIMNMX.S16x2.RELU clamps both to [0, hi], one PRMT packs four bytes:
// 5 ALU-pipe instructions per 4 outputs instead of 8 FMNMX + 3 PRMT (the ALU pipe -- IADD / FMNMX / PRMT at 16 lanes per
// clk per scheduler -- is what paces the pointwise GEMMs' epilogue warps).
__device__ __forceinline__ uint32_t f32_pair16(int acc0, int acc1, const ChanRegs& c) {
  const int kb = c.bias + 0x4B400000;
  const float y0 = fmaf(__int_as_float(acc0 + kb) - kF32Magic, c.r, kF32Magic);
  const float y1 = fmaf(__int_as_float(acc1 + kb) - kF32Magic, c.r, kF32Magic);
  return __byte_perm(__float_as_uint(y0), __float_as_uint(y1), 0x5410);
}
__device__ __forceinline__ uint32_t epilogue_f32_pack16(const EpiConsts& k, int a0, int a1, int a2, int a3, const ChanRegs& c) {
  uint32_t p01 = f32_pair16(a0, a1, c), p23 = f32_pair16(a2, a3, c);
  // the flag is only set for ReLU-folded clamps [0, hi] (validated in api.cu): max(min(x, hi), 0) is ONE instruction
  p01 = __vimin_s16x2_relu(p01, k.hi16);
  p23 = __vimin_s16x2_relu(p23, k.hi16);
  return __byte_perm(p01, p23, 0x6420);
}

// Same, for an accumulator that was STARTED at 0x4B400000 + bias (the bit pattern of 1.5 * 2^23 + bias): its bits
// are then the float 1.5 * 2^23 + (acc + bias) exactly (|acc + bias| < 2^22), and one FADD replaces the half-rate
// I2F.  Used where the MMA's initial accumulator is free to choose (the IMMA depthwise kernel).
constexpr int kF32MagicBits = 0x4B400000;
__device__ __forceinline__ uint32_t epilogue_f32_bits_biased(const EpiConsts& k, int acc_bits, const ChanRegs& c) {
  const float f = __int_as_float(acc_bits) - kF32Magic;
  const float y = fmaf(f, c.r, kF32Magic);
  return __float_as_uint(fminf(fmaxf(y, k.flo), k.fhi));
}

// Literal float recipe (runtime flags; any accumulator size, exact ties, 64-bit bias).
__device__ __forceinline__ int epilogue_exact(const EpiArgs& ep, const EpiConsts& k, int acc0, const ChanRegs& c0,
                                              bool has1, int acc1, const ChanRegs& c1, int prev) {
  const bool ex = ep.flags & QASR_F_EXACT_F32, b64 = ep.flags & QASR_F_BIAS64;
  double o = contrib_exact(acc0, c0, ep.flags & QASR_F_RELU0, ex, b64);
  if (has1) o += contrib_exact(acc1, c1, false, ex, b64);
  o += (double)prev;
  o = fmin(fmax(o, (double)ep.qmin), (double)ep.qmax);   // quant_utils.py:213-214
  if (ep.flags & QASR_F_RELU1) o = fmax(o, 0.0);          // jasper.py:687
  if (ep.flags & QASR_F_STAGE2) {
    o = rint(__dmul_rn(o, k.mult2));
    o = fmin(fmax(o, (double)ep.qmin2), (double)ep.qmax2);
  }
  return (int)o;
}

__device__ __forceinline__ float epilogue_dequant(const EpiArgs& ep, int acc, const ChanRegs& c) {
  const long long b64 = c.bias == 2147483647 ? 2147483648LL : (long long)c.bias;
  float a = (ep.flags & QASR_F_BIAS64) ? (float)((long long)acc + b64) : (float)(acc + c.bias);
  float y = __fmul_rn(a, c.s);
  if (ep.flags & QASR_F_RELU0) y = fmaxf(y, 0.0f);
  return y;
}

// order-preserving map float -> uint32 (statistics kernels: atomicMin / atomicMax on the keys)
__device__ __forceinline__ unsigned qasr_f32_key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float qasr_key_f32(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// accumulator + bias as the reference's conv sees it (64-bit: INT32_MAX is the sentinel of b_int == 2^31)
__device__ __forceinline__ long long qasr_acc_bias64(int acc, int bias) {
  return (long long)acc + (bias == 2147483647 ? 2147483648LL : (long long)bias);
}

// low bytes of four ints -> one word (3 PRMT)
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
  const uint32_t lo = __byte_perm((uint32_t)a, (uint32_t)b, 0x0040);
  const uint32_t hi = __byte_perm((uint32_t)c, (uint32_t)d, 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}

#endif  // __CUDACC__

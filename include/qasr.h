/*
 * qasr.h -- C ABI of the B200 (sm_100a) integer-only QuartzNet / Jasper kernels.
 *
 * This is the drop-in boundary for ONE path of kssteven418/Q-ASR: the fake-quant
 * forward of nemo/quantization/utils/{quant_utils,quant_modules}.py as driven by
 * nemo/collections/asr/parts/jasper.py and modules/conv_asr.py.  The reference has
 * no native code; each entry point below names the Python lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     name ends in _host.  The caller owns all memory; nothing is allocated here.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises the device.
 *   - return value: 0 on success, negative qasr_status otherwise; qasr_last_error()
 *     returns a thread-local message.  No C++ exception crosses the ABI.
 *   - activation codes are stored as 8-bit tensors [B][C][ld] with time contiguous
 *     (the reference's (B,C,T) layout), `ld` = row pitch in elements, ld % 16 == 0,
 *     base pointers 16-byte aligned.  Signedness is a property of the producer's
 *     clamp range: qmin >= 0 -> unsigned (the reference's "act_bit+1" inputs,
 *     jasper.py:159-164), else signed.
 *   - lens[b] = number of valid frames of utterance b in that tensor (int32).
 */
#ifndef QASR_H_
#define QASR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QASR_VERSION 101   /* 101: QASR_F_CLAMP16, QASR_F_STAGE2_16 */

typedef enum {
  QASR_OK = 0,
  QASR_ERR_INVALID = -1,   /* bad argument (shape, alignment, null pointer)        */
  QASR_ERR_CUDA = -2,      /* CUDA runtime / driver error, see qasr_last_error()    */
  QASR_ERR_UNSUPPORTED = -3
} qasr_status;

int qasr_version(void);
const char* qasr_last_error(void);
/* Number of SMs of the current device (grid sizing of the persistent kernels). */
int qasr_sm_count(void);

/* ------------------------------------------------------------------------------------
 * Fused epilogue = QuantConv1d's rescale + the consumer QuantAct's requantisation.
 *
 * Replaces, per output element (SURVEY.md 3.4 b,c):
 *   quant_modules.py:305-309   A = f32(acc);  y = A * s_out[c]
 *   jasper.py:661,687          y = relu(y)
 *   quant_utils.py:187         z = rne(y / s_out[c])
 *   quant_utils.py:188-196     o = rne(f64(z) * f64(m[c]) / 2^e[c])
 *   quant_utils.py:198-211     o += rne(f64(z1) * f64(m1[c]) / 2^e1[c])   (residual branch)
 *   quant_utils.py:213-214     q = clamp(o, qmin, qmax)
 * optionally followed by the block-output ReLU (jasper.py:687) and ONE more scalar
 * requantisation (the next MaskedConv1d's QuantAct, quant_modules.py:186-194), and
 * by MaskedConv1d's length mask (jasper.py:177-181: t >= lens[b] -> code 0).
 *
 * Per-channel parameter record (16 bytes, device):
 * ---------------------------------------------------------------------------------- */
typedef struct {
  float   s_out;   /* accumulator scale s_w[c] * s_x            (quant_modules.py:293) */
  int32_t bias;    /* b_int[c] (0 when the conv has no bias)    (quant_modules.py:296) */
  int32_t m;       /* dyadic mantissa, [2^30, 2^31)             (quant_utils.py:139)   */
  int32_t e;       /* dyadic shift: o = rne(z * m / 2^e)        (quant_utils.py:145)   */
} qasr_chan;

enum {
  QASR_EPI_REQUANT_I8 = 0, /* int8/uint8 codes out                                     */
  QASR_EPI_DEQUANT_F32 = 1,/* y = f32(acc+bias) * s_out[c] (+relu): the module-level
                              QuantConv1d return value, quant_modules.py:307-309       */
  QASR_EPI_RAW_I32 = 2,    /* acc + bias as int32 (dynamic / calibration passes)       */
  QASR_EPI_STATS = 3       /* no tensor is written: the STATISTICS pass of dynamic quantisation
                              (quant_modules.py:149-156).  `out` is a per-channel record table
                              qasr_stats[C_out], updated with atomics:
                                one operand : min / max of acc + bias over the counted frames (int64)
                                two operands: min / max of f32(acc1+b1)*s1 + f32(acc0+b0)*s0, the x_act of
                                              res_act (quant_modules.py:108), as order-preserving uint32
                              frames t >= out_lens[b] are not counted when QASR_F_MASK is set     */
};

/* One record per output channel of a QASR_EPI_STATS pass.  Reset state: lo = INT64_MAX, hi = INT64_MIN
 * (one operand) or lo32 = 0xffffffff, hi32 = 0 (two operands, float keys: bits ^ sign-dependent mask). */
typedef struct {
  int64_t lo, hi;
} qasr_stats;

/* flags */
#define QASR_F_RELU0      1   /* ReLU between conv 0 and the activation quantiser          */
#define QASR_F_RELU1      2   /* ReLU after the first clamp (JasperBlock.mout)              */
#define QASR_F_STAGE2     4   /* second, scalar requantisation (m2, e2, qmin2, qmax2)       */
#define QASR_F_MASK       8   /* zero codes at t >= out_lens[b]                             */
#define QASR_F_EXACT_F32  16  /* emulate the f32 round trip of the accumulator bit-for-bit
                                 (needed when |acc| may reach 2^22); else z = acc           */
#define QASR_F_EXACT_DYAD 32  /* dyadic product in double (53-bit rounding + half-to-even),
                                 else (z*m + 2^(e-1)) >> e, valid when the host proved
                                 there is no exact tie and no 53-bit rounding               */
#define QASR_F_B_SIGNED   64  /* conv-0 input codes are signed (else unsigned)              */
#define QASR_F_B1_SIGNED  128 /* conv-1 (residual branch) input codes are signed            */
#define QASR_F_BIAS64     256 /* add the bias in 64-bit (|b_int| close to 2^31)             */
#define QASR_F_DEVICE_TABLES 1024 /* ch0 / ch1 are written by preceding work on the same stream (dynamic
                                 quantisation: scales exist on the device only): the kernel is launched
                                 WITHOUT programmatic dependent launch, so it starts after its predecessor
                                 has completed and its pre-wait prologue may read the tables            */
#define QASR_F_RANGE_CHECK 2048 /* integer fast path only: the host proved it for |acc + bias| < 2^22 (where the reference's
                                  float32 round trip, quant_modules.py:305-307 / quant_utils.py:187, is the identity); the
                                  kernel compares every accumulator with that limit and evaluates the literal recipe
                                  (EXACT_F32 + EXACT_DYAD) for the elements beyond it.  For launches whose worst-case
                                  bound exceeds 2^22 (8-bit weights x activations over >= 256 channels) while real
                                  accumulators practically never do */
#define QASR_F_FAST_F32   512 /* single-operand REQUANT only: the host proved that
                                 clamp(rint(f32(acc+bias) * f32(m*2^-e))), one float FMA
                                 with a single rounding, equals the dyadic result for every
                                 reachable accumulator (|acc+bias| < 2^22); kernels then run
                                 the epilogue on the FP32 pipe (2x the IMAD.HI rate on B200) */
#define QASR_F_CLAMP16   4096 /* with FAST_F32 or f32_mult: the host ALSO proved |rint((acc+bias) * r)| < 2^15 for every
                                 reachable accumulator of every float-path channel (the worst-case bound, not the clamp
                                 window), and the clamp is the ReLU-folded [0, qmax] (RELU0 set, qmin <= 0), so kernels
                                 may clamp two outputs at a time on packed 16-bit lanes (one VIMNMX.S16x2.RELU) instead
                                 of two FMNMX per output.  A hint: kernels without a 16-bit epilogue ignore it         */
#define QASR_F_STAGE2_16 8192 /* with STAGE2 + RELU1 on the integer fast path: the host checked for EVERY code q in
                                 [0, qmax] (the first clamp with the block ReLU folded in; qmax <= 127) that
                                 (q * M16 + 2^15) >> 16, M16 = rne(m2 * 2^(16 - e2)), equals rne(q * m2 / 2^e2), and
                                 qmin2 <= 0 <= qmax2 < 2^15: kernels may run the second stage in 16-bit fixed point, two
                                 outputs per clamp.  A hint: kernels without that epilogue ignore it                   */

typedef struct {
  int32_t mode;              /* QASR_EPI_*                                               */
  int32_t flags;             /* QASR_F_*                                                 */
  const qasr_chan* ch0;      /* [C_out] parameters of conv 0                             */
  const qasr_chan* ch1;      /* [C_out] parameters of conv 1 (residual branch) or NULL   */
  const int8_t* prev;        /* optional int8 codes [B][C_out][prev_ld] added before the
                                clamp: a previous res_act result of the same scale
                                (dense residual, jasper.py:667-682; ratio exactly 1)     */
  int32_t prev_ld;
  int32_t qmin, qmax;        /* first clamp                                              */
  int32_t m2, e2;            /* stage-2 scalar dyadic                                    */
  int32_t qmin2, qmax2;      /* stage-2 clamp                                            */
  const int32_t* out_lens;   /* [B] valid output frames (mask), or NULL                  */
  const float* f32_mult;     /* optional [C_out] (conv 0 only, single-stage REQUANT without
                                ch1 / prev): f32(m * 2^-e) for the channels where the host
                                proved the one-FMA float restatement (see QASR_F_FAST_F32),
                                0.0f for the others; kernels pick the float or the integer
                                epilogue per channel.  NULL = integer epilogue everywhere  */
} qasr_epilogue;

/* ------------------------------------------------------------------------------------
 * K1  depthwise (time-channel-separable) 1-D convolution on 8-bit codes.
 * Replaces F.conv1d(x_int.double(), w_int.double(), groups=C) at
 * quant_modules.py:301-305 for groups == C_in == C_out, k in {33..87}, stride 1|2,
 * dilation 1|2, plus the fused epilogue above.
 *   x      [B][C][x_ld]  8-bit codes (signed iff QASR_F_B_SIGNED), already masked
 *   wp     packed weights from qasr_pack_dw_weights(): the int8 codes W_int [C][k]
 *          (quant_utils.py:74-76) laid out as the banded tensor-core A fragments the kernel
 *          consumes; packing depends on (k, stride, dilation, pad) and is done once per layer
 *   out    [B][C][out_ld] int8 (REQUANT), float (DEQUANT) or int32 (RAW); pitch in elems
 * ---------------------------------------------------------------------------------- */
int64_t qasr_dw_packed_bytes(int32_t C, int32_t k, int32_t stride, int32_t dilation, int32_t pad);
int qasr_pack_dw_weights(const int8_t* w, void* wp, int32_t C, int32_t k, int32_t stride,
                         int32_t dilation, int32_t pad, void* stream);
int qasr_dwconv1d_i8(const void* x, int32_t x_ld, const int8_t* wp, void* out, int32_t out_ld,
                     int32_t B, int32_t C, int32_t T_in, int32_t T_out, int32_t k,
                     int32_t stride, int32_t dilation, int32_t pad,
                     const qasr_epilogue* epi, void* stream);

/* ------------------------------------------------------------------------------------
 * K2  dense 1-D convolution as int8 tensor-core GEMM / implicit GEMM
 *     (tcgen05.mma kind::i8, TMEM accumulators, TMA-fed), k >= 1, groups == 1.
 * Replaces F.conv1d(x_int.double(), w_int.double(), bias) at quant_modules.py:301-305
 * for the pointwise 1x1 convs, the residual 1x1 convs, Jasper's dense k<=29 convs and
 * the decoder projection (conv_asr.py:261), plus the fused epilogue.  With x1 != NULL
 * a second GEMM (the residual 1x1 conv of the block, jasper.py:667-672) accumulates
 * into a second TMEM tile and both meet in the res_act epilogue (jasper.py:681).
 *
 *   wp  packed weights from qasr_pack_conv_weights(): [k][C_out_pad][C_in_pad] int8,
 *       C_*_pad = round_up(C, 128), zero filled
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const void*   x;        /* [B][C_in][x_ld] codes                                      */
  const int8_t* wp;       /* packed weights                                             */
  int32_t x_ld, C_in, T_in, k, stride, dilation, pad;
} qasr_gemm_operand;

int qasr_conv1d_gemm_i8(const qasr_gemm_operand* op0, const qasr_gemm_operand* op1,
                        void* out, int32_t out_ld, int32_t B, int32_t C_out, int32_t T_out,
                        const qasr_epilogue* epi, void* stream);

/* ------------------------------------------------------------------------------------
 * K2b  the same convolution on CHANNELS-LAST codes [B][T][ld] (ld = channel pitch, % 16 == 0):
 *      any k / dilation, stride 1|2 -- Jasper's dense k = 11..29 convs and strided head
 *      (quant_modules.py:301-305).  TMA cannot fetch time-shifted boxes from the (B,C,T) layout
 *      (16-byte alignment of the contiguous coordinate), so networks with dense k > 1 convs keep
 *      their int8 activations channels-last.  Same packed weights, same epilogue; `prev` and a
 *      REQUANT output are channels-last [B][T_out][ld], DEQUANT_F32 / RAW_I32 outputs are written
 *      in the reference's (B, C_out, T_out) layout (out_ld = time pitch).
 * ---------------------------------------------------------------------------------- */
int qasr_conv1d_ntc_i8(const qasr_gemm_operand* op0, const qasr_gemm_operand* op1,
                       void* out, int32_t out_ld, int32_t B, int32_t C_out, int32_t T_out,
                       const qasr_epilogue* epi, void* stream);

/* (B,C,T) float32 -> channels-last int8 codes [B][T][c_ld].
 *   mode 0: first-layer SymmetricQuantFunction clamp(rne((1/s)*x), -n, n-1) (quant_modules.py:180-184)
 *   mode 1: codes of an already quantised tensor rne(x / s)                 (quant_modules.py:301)
 *   t >= lens[b] -> code 0 when lens != NULL. */
int qasr_quantize_f32_ntc(const float* x, int32_t x_ld, const float* scale, int32_t mode, int32_t bit,
                          int8_t* codes, int32_t c_ld, int32_t B, int32_t C, int32_t T,
                          const int32_t* lens, void* stream);

/* Bytes of the packed weight buffer for (C_out, C_in, k). */
int64_t qasr_packed_weight_bytes(int32_t C_out, int32_t C_in, int32_t k);
/* w [C_out][C_in][k] int8 (torch Conv1d layout) -> packed.  Device to device. */
int qasr_pack_conv_weights(const int8_t* w, int8_t* wp, int32_t C_out, int32_t C_in, int32_t k,
                           void* stream);

/* ------------------------------------------------------------------------------------
 * K3  stand-alone quantiser kernels (module-level API)
 * ---------------------------------------------------------------------------------- */
/* SymmetricQuantFunction on activations (first layer, quant_modules.py:180-184):
 *   q = clamp(rne((1/s) * x), -n, n-1); optional outputs: codes (int8, pitch codes_ld,
 *   may be NULL) and xq = q * s (float, same pitch as x, may be NULL).
 *   t >= lens[b] -> x treated as 0 when lens != NULL.  scale is a DEVICE scalar. */
int qasr_quantize_f32(const float* x, int32_t x_ld, const float* scale, int32_t bit,
                      int8_t* codes, int32_t codes_ld, float* xq,
                      int32_t B, int32_t C, int32_t T, const int32_t* lens, void* stream);

/* fixedpoint_mul (quant_utils.py:164-216) on float inputs as the reference's QuantAct
 * calls it (quant_modules.py:186-194):
 *   z = rne(x / pre_sf[c]);  o = rne(f64(z) * m / 2^e)  [+ same on identity];  clamp;
 *   codes (int8, optional) and xq = f32(o) * z_sf (float, optional).
 *   pre_sf / id_sf: device arrays with n_pre / n_id entries (1 = scalar, C = per channel).
 *   (m, e) are derived ON DEVICE from pre_sf / z_sf exactly as batch_frexp does
 *   (quant_utils.py:121-147), so dynamic mode needs no host round trip. */
int qasr_requant_f32(const float* x, const float* pre_sf, int32_t n_pre,
                     const float* identity, const float* id_sf, int32_t n_id,
                     const float* z_sf, int32_t qmin, int32_t qmax,
                     int8_t* codes, int32_t codes_ld, float* xq, int32_t ld,
                     int32_t B, int32_t C, int32_t T, const int32_t* lens, void* stream);

/* batch_frexp + the f64 scale ratio of quant_utils.py:188-193 on device:
 *   (m[i], e[i]) from pre_sf[i] / z_sf, i < n; results normalised to m in [2^30, 2^31). */
int qasr_dyadic_from_scales(const float* pre_sf, int32_t n, const float* z_sf,
                            int32_t* m, int32_t* e, void* stream);

/* codes*scale floats -> 8-bit codes: q = rne(x / s) (the x_int of quant_modules.py:301,
 * rounded).  Used by the module-level QuantConv1d to feed K1/K2. */
int qasr_codes_from_f32(const float* x, int32_t x_ld, const float* scale, int8_t* codes,
                        int32_t codes_ld, int32_t B, int32_t C, int32_t T, void* stream);

/* Float-noise emulation of the module-level QuantConv1d (the reference never rounds x_int = x / s_x,
 * quant_modules.py:301, so code q enters its float64 conv as q + eps(q)):
 *   noise_codes:   out[i] = rne((f64(f32(f32(q) * s) / s) - q) * 2^22) for every code q of the buffer (n bytes,
 *                  a multiple of 4; signed or unsigned codes); |.| <= 64.
 *   combine_noise: out[b,c,t] = f32(f64(acc) + f64(bias[c]) + f64(dacc) * 2^-22) * s_out[c]   (:305-307), acc / dacc
 *                  the raw int32 accumulators of the conv over the codes / over the noise codes; bias may be NULL. */
int qasr_noise_codes_i8(const int8_t* codes, int32_t in_unsigned, const float* scale, int8_t* out, int64_t n,
                        void* stream);
int qasr_combine_noise_f32(const int32_t* acc, const int32_t* dacc, int32_t in_ld, const float* bias,
                           const float* s_out, float* out, int32_t out_ld, int32_t B, int32_t C, int32_t T,
                           void* stream);

/* int8 -> int8 requantisation through a scalar dyadic (a QuantAct whose input already
 * has a scalar scale, e.g. a second consumer of a res_act output). */
int qasr_requant_i8(const int8_t* in, int32_t in_ld, int32_t in_signed, int32_t m, int32_t e,
                    int32_t qmin, int32_t qmax, int8_t* out, int32_t out_ld,
                    int32_t B, int32_t C, int32_t T, const int32_t* lens, void* stream);

/* ------------------------------------------------------------------------------------
 * K4  range statistics (calibration / dynamic mode, quant_modules.py:108-167)
 * ---------------------------------------------------------------------------------- */
/* Global min / max of x (or of identity + x, f32 add, quant_modules.py:108) over the
 * logical [B][C][T] extent of a pitched tensor; masked positions (t >= lens[b]) count
 * as 0 when lens != NULL.  out2[0] = min, out2[1] = max (device).  `workspace` must hold
 * qasr_minmax_workspace_bytes() bytes. */
int64_t qasr_minmax_workspace_bytes(void);
int qasr_minmax_f32(const float* x, const float* identity, int32_t ld, int32_t B, int32_t C,
                    int32_t T, const int32_t* lens, float* out2, void* workspace, void* stream);

/* torch.quantile(x.flatten(), q) pair as used at quant_modules.py:122-123: exact k-th
 * order statistics by radix select on the float bit patterns + torch's float32 rank
 * arithmetic and lerp.  q_lo, q_hi are the float32 quantile levels.  out2 = (lo, hi). */
int64_t qasr_quantile_workspace_bytes(void);
int qasr_quantile2_f32(const float* x, const float* identity, int32_t ld, int32_t B, int32_t C,
                       int32_t T, const int32_t* lens, float q_lo, float q_hi, float* out2,
                       void* workspace, void* stream);

/* ------------------------------------------------------------------------------------
 * K4b  dynamic quantisation on the device (quant_modules.py:149-167 with running ranges off): the range of
 *      every QuantAct is a function of the batch, so scales, dyadic multipliers and integer biases exist on
 *      the device only and the whole forward stays one CUDA graph.  Per conv -> consumer QuantAct:
 *        1. qasr_conv* with QASR_EPI_STATS        per-channel extremes of the integer accumulators
 *        2. qasr_dyn_refine                        which channels can hold max |y| (rigorous float bounds), then
 *                                                  those channels again, position by position, WITH the float
 *                                                  noise of the reference's un-rounded x / s (quant_modules.py:301):
 *                                                  the exact max |y| the reference's x.min() / x.max() see
 *        3. qasr_dyn_finalize                      range -> scale (quant_utils.py:47-52) -> (m, e) of the producers
 *                                                  (quant_utils.py:188-193, 121-147) -> s_out / b_int of the
 *                                                  consumer convs (quant_modules.py:293-297)
 *        4. qasr_conv* with QASR_EPI_REQUANT_I8 | QASR_F_DEVICE_TABLES
 *      Only max(|x_min|, |x_max|) enters the scale, so one extreme is tracked: absmax = max |x_act|.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const void*   x;        /* [B][C_in][x_ld] input codes of the conv                                  */
  const int8_t* w;        /* RAW int8 weights [C_out][C_in / groups][k] (not the packed layouts)       */
  const float*  s_x;      /* device scalar: scale of the input codes                                   */
  const qasr_chan* ch;    /* [C_out] table holding s_out and bias                                      */
  const int32_t* wabs;    /* [C_out] sum of |w| per output channel                                     */
  int32_t x_ld, C_in, T_in, k, stride, dilation, pad;
  int32_t depthwise;      /* groups == C_in == C_out                                                   */
  int32_t in_unsigned;    /* input codes are unsigned                                                  */
} qasr_dyn_operand;

#define QASR_DYN_MAX_CAND 4
/* Reads the records the QASR_EPI_STATS pass left in `stats` (qasr_dyn_finalize resets them afterwards), picks the
 * candidate channels, recomputes them with the float noise and folds the exact max |x_act| into *absmax_bits: the
 * uint32 bit pattern of a non-negative float, updated with atomicMax (reset value 0).
 * cand[0..3] = candidate channels (-1 = none), cand[4] = how many channels qualified (> 4: refused, the caller checks
 * it after the forward).  relu: a ReLU sits between the conv and the QuantAct (one operand only). */
int qasr_dyn_refine(const qasr_stats* stats, const qasr_dyn_operand* op0, const qasr_dyn_operand* op1, int32_t relu,
                    const int32_t* out_lens, int32_t B, int32_t C_out, int32_t T_out, int32_t* cand,
                    uint32_t* absmax_bits, void* stream);

typedef struct {
  int32_t src_kind;          /* 0: *src = absmax bits (uint32, reset to 0 after reading)
                                1: src = float[2] (min, max) of a float tensor (qasr_minmax_f32)
                                2: src = int32[2] (min, max) of int8 codes (qasr_minmax_i8, reset after reading),
                                   times *src_scale                                                       */
  void* src;
  const float* src_scale;
  int32_t n_levels;          /* 2^(activation_bit - 1) - 1                                                */
  float* act_scale;          /* out: the QuantAct's scale                                                 */
  int32_t n_prod;            /* producer conv tables whose (m, e) = dyadic(s_out[c] / scale) are written  */
  qasr_chan* prod[2];
  int32_t prod_C[2];
  const float* prev_scale;   /* optional scalar producer (int8 codes with scale *prev_scale) ...          */
  int32_t* prev_me;          /* ... (m, e) written here (qasr_requant_i8_dev)                              */
  int32_t n_cons;            /* consumer conv tables whose s_out = s_w[c] * scale and bias are written    */
  qasr_chan* cons[2];
  const float* cons_sw[2];
  const float* cons_bias[2]; /* folded float bias [C] or NULL                                             */
  int32_t cons_C[2];
  /* data parallelism: exchange of absmax between the ranks of one node (NULL / 0: single rank)          */
  uint64_t* const* peer_slots;   /* device array [world]: every rank's slot table (peer-mapped)           */
  int32_t rank, world, slot;
  uint32_t* seq;                 /* device counter of this finalize node                                  */
  /* per producer: sum |w| per channel and the largest |input code|, for the device-side proof that the all-integer
   * epilogue equals the literal float recipe on that channel (else bit 31 of the record's e is set)      */
  const int32_t* prod_wabs[2];
  int32_t prod_qabs[2];
  qasr_stats* stats;             /* optional: statistics records to put back into their reset state       */
  int32_t stats_C, stats_dual;
  int32_t phase;                 /* world > 1 only.  0: post this rank's absmax to every rank's slots, wait for all
                                    ranks, finish (one rank per GPU: the production form).  1: post only.  2: wait and
                                    finish only.  1 and 2 let ONE stream run several ranks' plans in turn (all ranks'
                                    posts are complete before any rank waits), which is how the exchange is tested on a
                                    single GPU: launches that spin on each other must never share a GPU.             */
} qasr_dyn_finalize_args;
int qasr_dyn_finalize(const qasr_dyn_finalize_args* a, void* stream);

/* min / max of int8 codes after the block ReLU and length mask (the float tensor relu(q * s) the next
 * QuantAct sees, jasper.py:687 + :177-181): out2 int32 (min, max), updated with atomics (reset: INT32_MAX, INT32_MIN). */
int qasr_minmax_i8(const int8_t* codes, int32_t ld, int32_t relu, int32_t B, int32_t C, int32_t T,
                   const int32_t* lens, int32_t* out2, void* stream);
/* qasr_requant_i8 with the dyadic (m, e) read from device memory and an optional ReLU in front. */
int qasr_requant_i8_dev(const int8_t* in, int32_t in_ld, int32_t relu, const int32_t* me, int32_t qmin,
                        int32_t qmax, int8_t* out, int32_t out_ld, int32_t B, int32_t C, int32_t T,
                        const int32_t* lens, void* stream);

/* peer-memory plumbing of the range exchange: a device buffer that other processes of the node can map.
 * handle_out: 64 bytes (cudaIpcMemHandle_t).  qasr_comm_open maps a peer's buffer, qasr_comm_close unmaps. */
int qasr_comm_alloc(int64_t bytes, void** ptr, void* handle_out);
int qasr_comm_open(const void* handle, void** ptr);
int qasr_comm_close(void* ptr);
int qasr_comm_free(void* ptr);

/* ------------------------------------------------------------------------------------
 * K5  decoder tail: greedy CTC labels from integer logits
 * Replaces log_softmax + argmax (conv_asr.py:275, ctc_models.py:405): argmax_c of
 * y = f32(acc) * s_out[c] (log_softmax is monotone), first maximum wins as in torch.
 *   logits_f32 [B][V][ld] float;  labels [B][T] int32
 * ---------------------------------------------------------------------------------- */
int qasr_ctc_argmax(const float* logits, int32_t ld, int32_t B, int32_t V, int32_t T,
                    int32_t* labels, void* stream);

/* log_softmax over the class axis with the (B,V,T)->(B,T,V) transpose of conv_asr.py:275. */
int qasr_log_softmax_tv(const float* logits, int32_t ld, int32_t B, int32_t V, int32_t T,
                        float* log_probs, void* stream);

/* ------------------------------------------------------------------------------------
 * K6  mel-spectrogram featurizer (n_fft = 512)
 * qasr_melspec_f32 replaces FilterbankFeatures.forward up to the log
 * (nemo/collections/asr/parts/features.py:347-368): pre-emphasis, torch.stft framing (center,
 * reflect padding), |X|^2, mel filterbank (non-zero bands: fb_w / fb_start / fb_off), log(x + guard).
 *   audio [B][audio_ld] float, L samples per row;  feat [B][n_mels][feat_ld], n_frames = 1 + L/hop
 *   window512: window zero-padded to 512;  twiddle256: exp(-2 pi i k/512) as (re, im) pairs
 * qasr_normalize_features_f32 replaces normalize_batch 'per_feature' (:53-68) + the length mask and
 * zero padding (:383-395), in place: seq_len[b] valid frames, T_pad total frames.
 * ---------------------------------------------------------------------------------- */
int qasr_melspec_f32(const float* audio, int32_t audio_ld, int32_t B, int32_t L, int32_t hop,
                     const float* window512, const float* twiddle256, const float* fb_w,
                     const int32_t* fb_start, const int32_t* fb_off, int32_t n_mels, float preemph,
                     float log_guard, float* feat, int32_t feat_ld, int32_t n_frames, void* stream);
int qasr_normalize_features_f32(float* feat, int32_t feat_ld, int32_t B, int32_t n_mels, int32_t n_frames,
                                int32_t T_pad, const int32_t* seq_len, int32_t do_normalize, void* stream);

/* ------------------------------------------------------------------------------------
 * K5b  greedy CTC collapse on device (metrics/wer.py:117-136): merge repeats, drop blanks.
 *   labels [B][T] int32 -> tokens [B][T] int32 (compacted, -1 filled), counts [B] int32;
 *   only the first lens[b] frames of each row are decoded when lens != NULL.
 * ---------------------------------------------------------------------------------- */
int qasr_ctc_collapse(const int32_t* labels, int32_t B, int32_t T, const int32_t* lens, int32_t blank_id,
                      int32_t* tokens, int32_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QASR_H_ */

// K3 / K4 / K5: stand-alone quantiser, statistics and decoder-tail kernels.
// All of them are HBM-bound streaming kernels over pitched [B][C][ld] tensors: grid-stride
// loops, grids sized in multiples of the SM count, warp-shuffle reductions.
#include <float.h>

#include "qasr_common.cuh"

namespace {

constexpr int kBlock = 256;

inline int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = 8LL * qasr_sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// device frexp + round-half-up mantissa: batch_frexp, quant_utils.py:121-147, applied to
// f64(pre_sf) / f64(f32(z_sf)) (quant_utils.py:188-193).  m normalised to [2^30, 2^31).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dyadic_of(float pre, float zsf, int& m, int& e) {
  const double r = (double)pre / (double)zsf;
  int ex;
  const double mant = frexp(r, &ex);                 // r = mant * 2^ex, mant in [0.5, 1)
  long long mm = (long long)floor(mant * 2147483648.0 + 0.5);  // Decimal ROUND_HALF_UP (mant > 0)
  int ee = 31 - ex;
  if (mm >= 2147483648LL) { mm >>= 1; ee -= 1; }     // 2^31 * 2^-e == 2^30 * 2^-(e-1)
  if (!(r > 0.0) || !isfinite(r)) { mm = 0; ee = 1; }
  m = (int)mm;
  e = ee;
}

__global__ void dyadic_kernel(const float* pre_sf, int n, const float* z_sf, int* m, int* e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dyadic_of(pre_sf[i], z_sf[0], m[i], e[i]);
}

// exact dyadic product with (m, e) possibly outside the fast range
__device__ __forceinline__ double dyad_apply(double z, int m, int e) {
  return rint(__dmul_rn(z, scalbn((double)m, -e)));
}

// ---------------------------------------------------------------------------------------------
// qasr_quantize_f32: first-layer SymmetricQuantFunction (quant_modules.py:180-184) followed by
// the identity fixedpoint_mul and clamp of :186-194 (ratio 1.0 -> m = 2^30, e = 30).
// ---------------------------------------------------------------------------------------------
__global__ void quantize_f32_kernel(const float* __restrict__ x, int x_ld, const float* scale, int n,
                                    int8_t* codes, int codes_ld, float* xq, int B, int C, int T,
                                    const int32_t* lens) {
  const float s = scale[0];
  const float inv = __fdiv_rn(1.0f, s);                      // quant_utils.py:26  (1. / scale)
  const int tq = (T + 3) / 4;
  const long long total = (long long)B * C * tq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % tq);
    const long long row = i / tq;
    const int b = (int)(row / C);
    const int len = lens ? lens[b] : T;
    int q[4];
    // four frames at once where the row allows it (16-byte aligned, inside T): one 16-byte load instead of four
    // predicated 4-byte ones (the 26 MB feature tensor of the bench workload: 16.7 us -> the HBM time)
    float v4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const float* px = x + row * x_ld + 4 * j;
    if (4 * j + 4 <= T && (reinterpret_cast<uintptr_t>(px) & 15) == 0) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(px));
      v4[0] = f.x; v4[1] = f.y; v4[2] = f.z; v4[3] = f.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (4 * j + u < T) v4[u] = px[u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = 4 * j + u;
      const float v = t < len ? v4[u] : 0.0f;
      float r = rintf(__fmul_rn(inv, v));
      r = fminf(fmaxf(r, (float)(-n)), (float)(n - 1));     // quant_utils.py:76
      // fixedpoint_mul with pre_sf == z_sf: z = rne((r*s)/s) == r; clamp [-n-1, n] is a no-op
      q[u] = (int)r;
      if (xq && t < T) xq[row * x_ld + t] = __fmul_rn(r, s);
    }
    if (codes) *reinterpret_cast<uint32_t*>(codes + row * codes_ld + 4 * j) = pack4(q[0], q[1], q[2], q[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// qasr_requant_f32: fixedpoint_mul on float inputs (quant_utils.py:164-216)
// ---------------------------------------------------------------------------------------------
__global__ void requant_f32_kernel(const float* __restrict__ x, const float* pre_sf, int n_pre,
                                   const float* __restrict__ identity, const float* id_sf, int n_id,
                                   const float* z_sf, int qmin, int qmax, int8_t* codes, int codes_ld,
                                   float* xq, int ld, int B, int C, int T, const int32_t* lens) {
  const float zs = z_sf[0];
  const int tq = (T + 3) / 4;
  const long long total = (long long)B * C * tq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % tq);
    const long long row = i / tq;
    const int c = (int)(row % C);
    const int b = (int)(row / C);
    const int len = lens ? lens[b] : T;
    const float ps = pre_sf[n_pre == 1 ? 0 : c];
    int m, e, m1 = 0, e1 = 1;
    dyadic_of(ps, zs, m, e);
    float is = 1.0f;
    if (identity) {
      is = id_sf[n_id == 1 ? 0 : c];
      dyadic_of(is, zs, m1, e1);
    }
    int q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = 4 * j + u;
      const bool ok = t < T && t < len;
      const float v = ok ? x[row * ld + t] : 0.0f;
      double o = dyad_apply((double)rintf(__fdiv_rn(v, ps)), m, e);
      if (identity) {
        const float w = ok ? identity[row * ld + t] : 0.0f;
        o += dyad_apply((double)rintf(__fdiv_rn(w, is)), m1, e1);
      }
      const float of = fminf(fmaxf((float)o, (float)qmin), (float)qmax);
      q[u] = (int)of;
      if (xq && t < T) xq[row * ld + t] = __fmul_rn(of, zs);   // quant_modules.py:194
    }
    if (codes) *reinterpret_cast<uint32_t*>(codes + row * codes_ld + 4 * j) = pack4(q[0], q[1], q[2], q[3]);
  }
}

// codes * scale floats -> codes (x_int of quant_modules.py:301, rounded)
__global__ void codes_from_f32_kernel(const float* __restrict__ x, int x_ld, const float* scale,
                                      int8_t* codes, int codes_ld, int B, int C, int T) {
  const float s = scale[0];
  const int tq = (T + 3) / 4;
  const long long total = (long long)B * C * tq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % tq);
    const long long row = i / tq;
    int q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = 4 * j + u;
      q[u] = t < T ? (int)rintf(__fdiv_rn(x[row * x_ld + t], s)) : 0;
    }
    *reinterpret_cast<uint32_t*>(codes + row * codes_ld + 4 * j) = pack4(q[0], q[1], q[2], q[3]);
  }
}

// noise codes of the reference's un-rounded x_int (see include/qasr.h)
__global__ void noise_codes_kernel(const uint32_t* __restrict__ codes, int in_unsigned, const float* scale,
                                   uint32_t* __restrict__ out, long long n_words) {
  __shared__ uint8_t lut[256];
  {
    const float s = scale[0];
    const int q = in_unsigned ? (int)threadIdx.x : (int)(int8_t)threadIdx.x;
    const float r = __fdiv_rn(__fmul_rn((float)q, s), s);
    lut[threadIdx.x] = (uint8_t)(int8_t)(int)rint(((double)r - (double)q) * 4194304.0);
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_words; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t w = codes[i];
    out[i] = (uint32_t)lut[w & 0xff] | ((uint32_t)lut[(w >> 8) & 0xff] << 8) | ((uint32_t)lut[(w >> 16) & 0xff] << 16) |
             ((uint32_t)lut[w >> 24] << 24);
  }
}

__global__ void combine_noise_kernel(const int32_t* __restrict__ acc, const int32_t* __restrict__ dacc, int in_ld,
                                     const float* __restrict__ bias, const float* __restrict__ s_out,
                                     float* __restrict__ out, int out_ld, int B, int C, int T) {
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long row = i / T;
    const int c = (int)(row % C);
    double tot = (double)acc[row * in_ld + t];
    if (bias) tot += (double)bias[c];
    tot += (double)dacc[row * in_ld + t] * (1.0 / 4194304.0);
    out[row * out_ld + t] = __fmul_rn((float)tot, s_out[c]);
  }
}

__global__ void requant_i8_kernel(const int8_t* __restrict__ in, int in_ld, int in_signed, int m, int e,
                                  int qmin, int qmax, int8_t* out, int out_ld, int B, int C, int T,
                                  const int32_t* lens) {
  __shared__ int8_t lut[256];
  {
    const int code = in_signed ? (int)(int8_t)threadIdx.x : (int)threadIdx.x;
    double o = dyad_apply((double)code, m, e);
    o = fmin(fmax(o, (double)qmin), (double)qmax);
    lut[threadIdx.x] = (int8_t)(int)o;
  }
  __syncthreads();
  const int tq = (T + 15) / 16;
  const long long total = (long long)B * C * tq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
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

// ---------------------------------------------------------------------------------------------
// K4 min / max (quant_modules.py:113-119 / :150-156 with x_act of :108)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned enc_f32(float f) {  // order-preserving map float -> uint32
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void minmax_init_kernel(unsigned* ws) {
  ws[0] = 0xffffffffu;  // running min (encoded)
  ws[1] = 0u;           // running max (encoded)
}

__device__ __forceinline__ float load_act(const float* __restrict__ x, const float* __restrict__ id,
                                          long long off, bool ok) {
  if (!ok) return 0.0f;
  const float v = x[off];
  return id ? __fadd_rn(id[off], v) : v;   // identity + x, quant_modules.py:108
}

__global__ void minmax_kernel(const float* __restrict__ x, const float* __restrict__ identity, int ld,
                              int B, int C, int T, const int32_t* lens, unsigned* ws) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long row = i / T;
    const int b = (int)(row / C);
    const bool ok = !lens || t < lens[b];
    const float v = load_act(x, identity, row * ld + t, ok);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[kBlock / 32], smx[kBlock / 32];
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    atomicMin(ws + 0, enc_f32(mn));
    atomicMax(ws + 1, enc_f32(mx));
  }
}

__global__ void minmax_final_kernel(const unsigned* ws, float* out2) {
  out2[0] = dec_f32(ws[0]);
  out2[1] = dec_f32(ws[1]);
}

// ---------------------------------------------------------------------------------------------
// K4 quantile pair: exact order statistics by 4-pass (8 bits per pass) radix select on the
// order-preserving encoding, for up to 4 ranks at once (lo/hi neighbours of both quantiles),
// then torch's interpolation:  rank = q * (n-1) in float32, lerp(a, b, w) as ATen's lerp.
// workspace layout (uint32): hist[4 ranks][256], prefix[4], remaining[4], want[4]
// ---------------------------------------------------------------------------------------------
struct QuantileWs {
  unsigned hist[4][256];
  unsigned prefix[4];
  unsigned long long rank[4];
};

__global__ void quantile_init_kernel(QuantileWs* ws, long long n, float q_lo, float q_hi) {
  const int t = threadIdx.x;
  for (int r = 0; r < 4; ++r) ws->hist[r][t] = 0;
  if (t < 4) {
    const float q = (t < 2) ? q_lo : q_hi;
    const float rank = __fmul_rn(q, (float)(n - 1));
    const float fl = floorf(rank), ce = ceilf(rank);
    ws->rank[t] = (unsigned long long)((t & 1) ? ce : fl);
    ws->prefix[t] = 0;
  }
}

// pass p examines bits [24-8p, 32-8p) of the keys that match the prefix chosen so far
__global__ void quantile_hist_kernel(const float* __restrict__ x, const float* __restrict__ identity,
                                     int ld, int B, int C, int T, const int32_t* lens, QuantileWs* ws,
                                     int pass) {
  __shared__ unsigned sh[4][256];
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  unsigned pre[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) pre[r] = ws->prefix[r];
  const long long total = (long long)B * C * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long row = i / T;
    const int b = (int)(row / C);
    const bool ok = !lens || t < lens[b];
    const unsigned key = enc_f32(load_act(x, identity, row * ld + t, ok));
    const unsigned digit = (key >> shift) & 0xff;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if ((key & mask) == pre[r]) atomicAdd(&sh[r][digit], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
    const unsigned v = (&sh[0][0])[i];
    if (v) atomicAdd(&ws->hist[0][0] + i, v);
  }
}

__global__ void quantile_scan_kernel(QuantileWs* ws, int pass) {
  const int r = threadIdx.x;
  if (r < 4) {
    const int shift = 24 - 8 * pass;
    unsigned long long want = ws->rank[r];
    unsigned long long cum = 0;
    int d = 0;
    for (; d < 256; ++d) {
      const unsigned long long h = ws->hist[r][d];
      if (want < cum + h) break;
      cum += h;
    }
    if (d > 255) d = 255;
    ws->rank[r] = want - cum;
    ws->prefix[r] |= ((unsigned)d) << shift;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&ws->hist[0][0])[i] = 0;
}

__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
  const float diff = __fsub_rn(b, a);
  // ATen lerp (native/Lerp.h): weight < 0.5 ? a + w * diff : b - diff * (1 - w); torch's CPU and
  // CUDA builds both contract the multiply-add into one fma (checked against torch.lerp).
  return (w < 0.5f) ? __fmaf_rn(w, diff, a) : __fmaf_rn(-diff, __fsub_rn(1.0f, w), b);
}

__global__ void quantile_final_kernel(const QuantileWs* ws, long long n, float q_lo, float q_hi,
                                      float* out2) {
  for (int k = 0; k < 2; ++k) {
    const float q = k == 0 ? q_lo : q_hi;
    const float rank = __fmul_rn(q, (float)(n - 1));
    const float w = __fsub_rn(rank, floorf(rank));
    const float a = dec_f32(ws->prefix[2 * k]), b = dec_f32(ws->prefix[2 * k + 1]);
    out2[k] = torch_lerp(a, b, w);
  }
}

// ---------------------------------------------------------------------------------------------
// K5 decoder tail
// ---------------------------------------------------------------------------------------------
__global__ void ctc_argmax_kernel(const float* __restrict__ logits, int ld, int B, int V, int T,
                                  int32_t* labels) {
  const long long total = (long long)B * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const int b = (int)(i / T);
    const float* p = logits + (size_t)b * V * ld + t;
    float best = p[0];
    int arg = 0;
    for (int v = 1; v < V; ++v) {
      const float y = p[(size_t)v * ld];
      if (y > best) { best = y; arg = v; }
    }
    labels[i] = arg;
  }
}

__global__ void log_softmax_tv_kernel(const float* __restrict__ logits, int ld, int B, int V, int T,
                                      float* __restrict__ out) {
  const long long total = (long long)B * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const int b = (int)(i / T);
    const float* p = logits + (size_t)b * V * ld + t;
    float mx = -FLT_MAX;
    for (int v = 0; v < V; ++v) mx = fmaxf(mx, p[(size_t)v * ld]);
    float sum = 0.0f;
    for (int v = 0; v < V; ++v) sum += expf(p[(size_t)v * ld] - mx);
    const float lse = mx + logf(sum);
    float* o = out + i * V;
    for (int v = 0; v < V; ++v) o[v] = p[(size_t)v * ld] - lse;
  }
}

// greedy CTC collapse: one warp per utterance, ballot + prefix popcount compaction
__global__ void ctc_collapse_kernel(const int32_t* __restrict__ labels, int B, int T, const int32_t* lens,
                                    int blank, int32_t* __restrict__ tokens, int32_t* __restrict__ counts) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const int n = lens ? min(lens[b], T) : T;
  const int32_t* row = labels + (size_t)b * T;
  int32_t* out = tokens + (size_t)b * T;
  int count = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const int p = t < n ? row[t] : blank;
    const int prev = (t > 0 && t < n) ? row[t - 1] : blank;          // previous = blank_id at the start
    const bool keep = t < n && p != blank && (p != prev || prev == blank);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) out[count + __popc(mask & ((1u << lane) - 1))] = p;
    count += __popc(mask);
  }
  for (int t = count + lane; t < T; t += 32) out[t] = -1;
  if (lane == 0) counts[b] = count;
}

}  // namespace

extern "C" int qasr_ctc_collapse(const int32_t* labels, int32_t B, int32_t T, const int32_t* lens, int32_t blank_id,
                                 int32_t* tokens, int32_t* counts, void* stream) {
  QASR_CHECK_ARG(labels && tokens && counts && B > 0 && T > 0, "ctc_collapse: bad arguments");
  ctc_collapse_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(labels, B, T, lens, blank_id, tokens, counts);
  return qasr_check_launch("ctc_collapse_kernel");
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int qasr_dyadic_from_scales(const float* pre_sf, int32_t n, const float* z_sf, int32_t* m,
                                       int32_t* e, void* stream) {
  QASR_CHECK_ARG(pre_sf && z_sf && m && e && n > 0, "dyadic_from_scales: bad arguments");
  dyadic_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pre_sf, n, z_sf, m, e);
  return qasr_check_launch("dyadic_kernel");
}

extern "C" int qasr_quantize_f32(const float* x, int32_t x_ld, const float* scale, int32_t bit,
                                 int8_t* codes, int32_t codes_ld, float* xq, int32_t B, int32_t C,
                                 int32_t T, const int32_t* lens, void* stream) {
  QASR_CHECK_ARG(x && scale && (codes || xq), "quantize_f32: null pointer");
  QASR_CHECK_ARG(bit >= 2 && bit <= 8, "quantize_f32: activation_bit %d not in [2, 8]", bit);
  QASR_CHECK_ARG(B > 0 && C > 0 && T > 0 && x_ld >= T, "quantize_f32: bad shape");
  QASR_CHECK_ARG(!codes || (codes_ld % 4 == 0 && codes_ld >= ((T + 3) / 4) * 4), "quantize_f32: codes pitch");
  const int n = (1 << (bit - 1)) - 1;
  const long long work = (long long)B * C * ((T + 3) / 4);
  quantize_f32_kernel<<<grid_for(work, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      x, x_ld, scale, n, codes, codes_ld, xq, B, C, T, lens);
  return qasr_check_launch("quantize_f32_kernel");
}

extern "C" int qasr_requant_f32(const float* x, const float* pre_sf, int32_t n_pre, const float* identity,
                                const float* id_sf, int32_t n_id, const float* z_sf, int32_t qmin,
                                int32_t qmax, int8_t* codes, int32_t codes_ld, float* xq, int32_t ld,
                                int32_t B, int32_t C, int32_t T, const int32_t* lens, void* stream) {
  QASR_CHECK_ARG(x && pre_sf && z_sf && (codes || xq), "requant_f32: null pointer");
  QASR_CHECK_ARG(n_pre == 1 || n_pre == C, "requant_f32: pre_sf must have 1 or C entries");
  QASR_CHECK_ARG(!identity || (id_sf && (n_id == 1 || n_id == C)), "requant_f32: identity scale");
  QASR_CHECK_ARG(B > 0 && C > 0 && T > 0 && ld >= T, "requant_f32: bad shape");
  QASR_CHECK_ARG(qmin >= -256 && qmax <= 255 && qmin < qmax,
                 "requant_f32: clamp range [%d, %d] wider than the 9-bit activation range", qmin, qmax);
  QASR_CHECK_ARG(!codes || (codes_ld % 4 == 0 && codes_ld >= ((T + 3) / 4) * 4), "requant_f32: codes pitch");
  const long long work = (long long)B * C * ((T + 3) / 4);
  requant_f32_kernel<<<grid_for(work, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      x, pre_sf, n_pre, identity, id_sf, n_id, z_sf, qmin, qmax, codes, codes_ld, xq, ld, B, C, T, lens);
  return qasr_check_launch("requant_f32_kernel");
}

extern "C" int qasr_codes_from_f32(const float* x, int32_t x_ld, const float* scale, int8_t* codes,
                                   int32_t codes_ld, int32_t B, int32_t C, int32_t T, void* stream) {
  QASR_CHECK_ARG(x && scale && codes, "codes_from_f32: null pointer");
  QASR_CHECK_ARG(B > 0 && C > 0 && T > 0 && x_ld >= T, "codes_from_f32: bad shape");
  QASR_CHECK_ARG(codes_ld % 4 == 0 && codes_ld >= ((T + 3) / 4) * 4, "codes_from_f32: codes pitch");
  const long long work = (long long)B * C * ((T + 3) / 4);
  codes_from_f32_kernel<<<grid_for(work, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      x, x_ld, scale, codes, codes_ld, B, C, T);
  return qasr_check_launch("codes_from_f32_kernel");
}

extern "C" int qasr_noise_codes_i8(const int8_t* codes, int32_t in_unsigned, const float* scale, int8_t* out, int64_t n,
                                   void* stream) {
  QASR_CHECK_ARG(codes && scale && out, "noise_codes: null pointer");
  QASR_CHECK_ARG(n > 0 && n % 4 == 0 && (reinterpret_cast<uintptr_t>(codes) & 3) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 3) == 0, "noise_codes: size / alignment must be a multiple of 4");
  noise_codes_kernel<<<grid_for(n / 4, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(codes), in_unsigned, scale, reinterpret_cast<uint32_t*>(out), n / 4);
  return qasr_check_launch("noise_codes_kernel");
}

extern "C" int qasr_combine_noise_f32(const int32_t* acc, const int32_t* dacc, int32_t in_ld, const float* bias,
                                      const float* s_out, float* out, int32_t out_ld, int32_t B, int32_t C, int32_t T,
                                      void* stream) {
  QASR_CHECK_ARG(acc && dacc && s_out && out, "combine_noise: null pointer");
  QASR_CHECK_ARG(B > 0 && C > 0 && T > 0 && in_ld >= T && out_ld >= T, "combine_noise: bad shape");
  combine_noise_kernel<<<grid_for((long long)B * C * T, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      acc, dacc, in_ld, bias, s_out, out, out_ld, B, C, T);
  return qasr_check_launch("combine_noise_kernel");
}

extern "C" int qasr_requant_i8(const int8_t* in, int32_t in_ld, int32_t in_signed, int32_t m, int32_t e,
                               int32_t qmin, int32_t qmax, int8_t* out, int32_t out_ld, int32_t B,
                               int32_t C, int32_t T, const int32_t* lens, void* stream) {
  QASR_CHECK_ARG(in && out, "requant_i8: null pointer");
  QASR_CHECK_ARG(in_ld % 16 == 0 && out_ld % 16 == 0 && in_ld >= T && out_ld >= T &&
                     qasr_aligned16(in) && qasr_aligned16(out), "requant_i8: pitch/alignment");
  // [-256, 255] is the reference's 9-bit clamp; the caller guarantees the codes fit 8 bits
  // (non-negative input, or a signed range within [-128, 127])
  QASR_CHECK_ARG(qmin < qmax && qmin >= -256 && qmax <= 255, "requant_i8: clamp range");
  const long long work = (long long)B * C * ((T + 15) / 16);
  requant_i8_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(
      in, in_ld, in_signed, m, e, qmin, qmax, out, out_ld, B, C, T, lens);
  return qasr_check_launch("requant_i8_kernel");
}

extern "C" int64_t qasr_minmax_workspace_bytes(void) { return 16; }

extern "C" int qasr_minmax_f32(const float* x, const float* identity, int32_t ld, int32_t B, int32_t C,
                               int32_t T, const int32_t* lens, float* out2, void* workspace,
                               void* stream) {
  QASR_CHECK_ARG(x && out2 && workspace, "minmax_f32: null pointer");
  QASR_CHECK_ARG(B > 0 && C > 0 && T > 0 && ld >= T, "minmax_f32: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* ws = static_cast<unsigned*>(workspace);
  minmax_init_kernel<<<1, 1, 0, st>>>(ws);
  minmax_kernel<<<grid_for((long long)B * C * T, kBlock * 8), kBlock, 0, st>>>(x, identity, ld, B, C, T,
                                                                              lens, ws);
  minmax_final_kernel<<<1, 1, 0, st>>>(ws, out2);
  return qasr_check_launch("minmax kernels");
}

extern "C" int64_t qasr_quantile_workspace_bytes(void) { return (int64_t)sizeof(QuantileWs); }

extern "C" int qasr_quantile2_f32(const float* x, const float* identity, int32_t ld, int32_t B, int32_t C,
                                  int32_t T, const int32_t* lens, float q_lo, float q_hi, float* out2,
                                  void* workspace, void* stream) {
  QASR_CHECK_ARG(x && out2 && workspace, "quantile2_f32: null pointer");
  QASR_CHECK_ARG(B > 0 && C > 0 && T > 0 && ld >= T, "quantile2_f32: bad shape");
  QASR_CHECK_ARG(q_lo >= 0.f && q_lo <= 1.f && q_hi >= 0.f && q_hi <= 1.f, "quantile2_f32: q outside [0,1]");
  const long long n = (long long)B * C * T;
  // torch.quantile refuses inputs above 2^24 elements (float32 rank arithmetic); same contract
  QASR_CHECK_ARG(n <= (1LL << 24), "quantile() input tensor is too large");
  cudaStream_t st = (cudaStream_t)stream;
  QuantileWs* ws = static_cast<QuantileWs*>(workspace);
  quantile_init_kernel<<<1, 256, 0, st>>>(ws, n, q_lo, q_hi);
  const int grid = grid_for(n, kBlock * 8);
  for (int pass = 0; pass < 4; ++pass) {
    quantile_hist_kernel<<<grid, kBlock, 0, st>>>(x, identity, ld, B, C, T, lens, ws, pass);
    quantile_scan_kernel<<<1, 256, 0, st>>>(ws, pass);
  }
  quantile_final_kernel<<<1, 1, 0, st>>>(ws, n, q_lo, q_hi, out2);
  return qasr_check_launch("quantile kernels");
}

extern "C" int qasr_ctc_argmax(const float* logits, int32_t ld, int32_t B, int32_t V, int32_t T,
                               int32_t* labels, void* stream) {
  QASR_CHECK_ARG(logits && labels && B > 0 && V > 0 && T > 0 && ld >= T, "ctc_argmax: bad arguments");
  ctc_argmax_kernel<<<grid_for((long long)B * T, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      logits, ld, B, V, T, labels);
  return qasr_check_launch("ctc_argmax_kernel");
}

extern "C" int qasr_log_softmax_tv(const float* logits, int32_t ld, int32_t B, int32_t V, int32_t T,
                                   float* log_probs, void* stream) {
  QASR_CHECK_ARG(logits && log_probs && B > 0 && V > 0 && T > 0 && ld >= T, "log_softmax_tv: bad arguments");
  log_softmax_tv_kernel<<<grid_for((long long)B * T, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      logits, ld, B, V, T, log_probs);
  return qasr_check_launch("log_softmax_tv_kernel");
}

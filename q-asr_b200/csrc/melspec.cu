// K6: mel-spectrogram featurizer.  Restates FilterbankFeatures.forward
// (nemo/collections/asr/parts/features.py:335-397) and normalize_batch (:53-82, per_feature) for
// n_fft = 512: pre-emphasis, torch.stft framing (center=True, reflect padding, window zero-padded to
// n_fft), power spectrum, mel filterbank, log(x + guard); then per-feature mean / unbiased std over the
// valid frames, mask, zero padding to the padded frame count.
//
// One warp per STFT frame.  The 512 real samples are packed as 256 complex values held 8 per lane; the 256-point FFT is
// an 8-point FFT in registers, a twiddle, and a 32-point FFT across the lanes with shuffles (the first version walked 9
// radix-2 stages through shared memory with bit-reversed scatter stores: 550 us for 64 000 frames, issue rate 0.2 per
// scheduler); the real signal's 257 bins are unpacked from shared memory, their power stays there and each lane
// accumulates two mel bins over its filter's non-zero band.  A CTA covers 32 consecutive frames of one utterance (8 warps x 4 frames) and collects their
// mel values in shared memory, so the (B, n_mels, T) output is written as 128-byte runs per feature row (a frame's
// 64 values lie one row pitch apart: written directly they are 64 scattered 4-byte stores per frame).
// HBM traffic = the audio (each sample is read by ~3.2 overlapping frames, L1/L2 hits) + 64 floats per frame out.
#include "qasr_common.cuh"

namespace {

constexpr int kNfft = 512;
constexpr int kBins = kNfft / 2 + 1;
constexpr int kWarpsPerCta = 8;
constexpr int kFramesPerCta = 32;
constexpr int kMaxStagedMels = 64;

struct MelArgs {
  const float* audio;
  int audio_ld, B, L, hop, n_frames, n_mels;
  const float* window;      // [512] window zero-padded to n_fft
  const float2* twiddle;    // [256] exp(-2 pi i k / 512)
  const float* fb_w;        // concatenated non-zero filter weights
  const int* fb_start;      // [n_mels] first FFT bin of each filter
  const int* fb_off;        // [n_mels + 1] offsets into fb_w
  float preemph, log_guard;
  float* feat;
  int feat_ld;
};

__device__ __forceinline__ float preemph_sample(const float* x, int j, float a) {
  return j == 0 ? x[0] : __fsub_rn(x[j], __fmul_rn(a, x[j - 1]));   // features.py:347-348
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) melspec_kernel(const MelArgs a) {
  __shared__ float2 buf[kWarpsPerCta][kNfft];
  __shared__ float staged[kMaxStagedMels][kFramesPerCta + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fblocks = (a.n_frames + kFramesPerCta - 1) / kFramesPerCta;
  const int b = blockIdx.x / fblocks, f_base = (blockIdx.x - b * fblocks) * kFramesPerCta;
  const float* x = a.audio + (size_t)b * a.audio_ld;
  float2* s = buf[warp];
  const bool stage_out = a.n_mels <= kMaxStagedMels;

  // twiddles of the across-lane stages (constants of the lane): stage h = 16, 8, 4, 2, 1 multiplies the lower lanes'
  // difference by W_{2h}^(lane mod h) = exp(-2 pi i (lane mod h) (16 / h) / 32); upper lanes multiply by 1
  float2 wl[5];
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int h = 16 >> st;
    wl[st] = (lane & h) ? __ldg(a.twiddle + 16 * ((lane & (h - 1)) * (16 / h))) : make_float2(1.f, 0.f);
  }
  const unsigned brev_lane = __brev((unsigned)lane) >> 27;   // 5-bit reversal

  for (int fl = warp; fl < kFramesPerCta; fl += kWarpsPerCta) {
    const int f = f_base + fl;
    if (f >= a.n_frames) break;   // warp-uniform
    // ---- framing + reflect padding + window.  The 512 real samples are packed as 256 complex z[n] = x[2n] + i x[2n+1];
    //      lane l holds z[32 j + l], j = 0..7 (samples 64 j + 2 l, + 1: consecutive lanes read consecutive pairs)
    float2 z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n = 64 * j + 2 * lane + q;
        const float w = __ldg(a.window + n);
        v[q] = 0.f;
        if (w != 0.f) {
          int jj = f * a.hop + n - kNfft / 2;
          if (jj < 0) jj = -jj;
          if (jj >= a.L) jj = 2 * (a.L - 1) - jj;
          v[q] = __fmul_rn(preemph_sample(x, jj, a.preemph), w);
        }
      }
      z[j] = make_float2(v[0], v[1]);
    }
    // ---- 256-point complex FFT, n = 32 j + l, k = k1 + 8 k2:
    //      Z[k1 + 8 k2] = sum_l W32^(l k2) [ W256^(l k1) sum_j z[32 j + l] W8^(j k1) ]
    // (1) 8-point DIF over the registers: afterwards z[r] holds A_l[k1 = brev3(r)]
    auto cadd = [](float2 p, float2 q) { return make_float2(p.x + q.x, p.y + q.y); };
    auto csub = [](float2 p, float2 q) { return make_float2(p.x - q.x, p.y - q.y); };
    auto cmul = [](float2 p, float2 w) { return make_float2(p.x * w.x - p.y * w.y, p.x * w.y + p.y * w.x); };
    {
      const float r2 = 0.70710678118654752440f;
      // stage 1: pairs (j, j + 4), twiddle W8^j on the difference
      float2 t0 = cadd(z[0], z[4]), t1 = cadd(z[1], z[5]), t2 = cadd(z[2], z[6]), t3 = cadd(z[3], z[7]);
      float2 d0 = csub(z[0], z[4]), d1 = csub(z[1], z[5]), d2 = csub(z[2], z[6]), d3 = csub(z[3], z[7]);
      d1 = make_float2((d1.x + d1.y) * r2, (d1.y - d1.x) * r2);     // * W8^1 = (1 - i) / sqrt 2
      d2 = make_float2(d2.y, -d2.x);                               // * W8^2 = -i
      d3 = make_float2((d3.y - d3.x) * r2, -(d3.x + d3.y) * r2);    // * W8^3 = (-1 - i) / sqrt 2
      // stage 2: within each half pairs (m, m + 2), twiddle W4^m
      float2 u0 = cadd(t0, t2), u1 = cadd(t1, t3), u2 = csub(t0, t2), u3 = csub(t1, t3);
      u3 = make_float2(u3.y, -u3.x);
      float2 v0 = cadd(d0, d2), v1 = cadd(d1, d3), v2 = csub(d0, d2), v3 = csub(d1, d3);
      v3 = make_float2(v3.y, -v3.x);
      // stage 3: pairs (m, m + 1); outputs in bit-reversed order: z[r] = A[brev3(r)]
      z[0] = cadd(u0, u1); z[1] = csub(u0, u1);   // A[0], A[4]
      z[2] = cadd(u2, u3); z[3] = csub(u2, u3);   // A[2], A[6]
      z[4] = cadd(v0, v1); z[5] = csub(v0, v1);   // A[1], A[5]
      z[6] = cadd(v2, v3); z[7] = csub(v2, v3);   // A[3], A[7]
    }
    // (2) twiddle W256^(l k1) = exp(-2 pi i (2 l k1) / 512): table index 2 l k1 (< 512; >= 256 is the negated first half)
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      constexpr int kRev3[8] = {0, 4, 2, 6, 1, 5, 3, 7};
      const int t = 2 * lane * kRev3[r];
      float2 w = __ldg(a.twiddle + (t & 255));
      if (t & 256) w = make_float2(-w.x, -w.y);
      z[r] = cmul(z[r], w);
    }
    // (3) 32-point DIF across the lanes for each register: afterwards lane l holds k2 = brev5(l)
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const int h = 16 >> st;
      const float sgn = (lane & h) ? -1.f : 1.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float ox = __shfl_xor_sync(0xffffffffu, z[r].x, h), oy = __shfl_xor_sync(0xffffffffu, z[r].y, h);
        const float2 d = make_float2(fmaf(sgn, z[r].x, ox), fmaf(sgn, z[r].y, oy));   // upper: x + o, lower: o - x
        z[r] = cmul(d, wl[st]);
      }
    }
    // (4) Z in natural order to shared memory: register r of lane l is Z[brev3(r) + 8 brev5(l)]
    __syncwarp();
    {
      constexpr int kRev3[8] = {0, 4, 2, 6, 1, 5, 3, 7};
#pragma unroll
      for (int r = 0; r < 8; ++r) s[kRev3[r] + 8 * (int)brev_lane] = z[r];
    }
    __syncwarp();
    // (5) spectrum of the real signal, bins 0..256: X[k] = E[k] + W512^k O[k], E = (Z[k] + conj Z[256 - k]) / 2,
    //     O = (Z[k] - conj Z[256 - k]) / (2 i); power as features.py:355-360 (sqrt(re^2 + im^2), then squared)
    float* p = reinterpret_cast<float*>(s + kNfft / 2 + 8);   // behind the 256 complex values
    float pw[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int k = lane + 32 * i;
      pw[i] = 0.f;
      if (k < kBins) {
        const float2 zk = s[k & 255], zn = s[(256 - k) & 255];
        const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));   // (Z - conj Zn) / (2 i)
        float2 w = k < 256 ? __ldg(a.twiddle + k) : make_float2(-1.f, 0.f);
        const float2 xo = cmul(o, w);
        const float re = e.x + xo.x, im = e.y + xo.y;
        const float mag = sqrtf(re * re + im * im);
        pw[i] = mag * mag;
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int k = lane + 32 * i;
      if (k < kBins) p[k] = pw[i];
    }
    __syncwarp();
    // ---- mel filterbank + log
    for (int m = lane; m < a.n_mels; m += 32) {
      const int k0 = __ldg(a.fb_start + m), o0 = __ldg(a.fb_off + m), o1 = __ldg(a.fb_off + m + 1);
      float acc = 0.f;
      for (int o = o0; o < o1; ++o) acc = fmaf(__ldg(a.fb_w + o), p[k0 + (o - o0)], acc);
      const float val = logf(acc + a.log_guard);   // features.py:366-368
      if (stage_out) staged[m][fl] = val;
      else a.feat[((size_t)b * a.n_mels + m) * a.feat_ld + f] = val;
    }
    __syncwarp();   // the next frame reuses this warp's buffer
  }
  if (!stage_out) return;
  __syncthreads();
  // ---- 32 frames x n_mels values out: one 128-byte run per feature row and warp pass
  const int nf = min(kFramesPerCta, a.n_frames - f_base);
  for (int m = warp; m < a.n_mels; m += kWarpsPerCta)
    if (lane < nf) a.feat[((size_t)b * a.n_mels + m) * a.feat_ld + f_base + lane] = staged[m][lane];
}

// per (utterance, feature): mean and unbiased std over the valid frames, normalise, mask, zero pad
__global__ void normalize_kernel(float* feat, int feat_ld, int n_mels, int n_frames, int T_pad,
                                 const int32_t* seq_len, int do_normalize) {
  const int m = blockIdx.x, b = blockIdx.y;
  float* row = feat + ((size_t)b * n_mels + m) * feat_ld;
  const int n = min(seq_len[b], n_frames);
  __shared__ double red[32];
  __shared__ double s_mean, s_std;
  auto block_sum = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x < 32) {
      t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;  // valid in warp 0
  };
  if (do_normalize) {
    double acc = 0;
    for (int t = threadIdx.x; t < n; t += blockDim.x) acc += (double)row[t];
    const double tot = block_sum(acc);
    if (threadIdx.x == 0) s_mean = tot / (double)n;
    __syncthreads();
    const double mean = s_mean;
    acc = 0;
    for (int t = threadIdx.x; t < n; t += blockDim.x) { const double d = (double)row[t] - mean; acc += d * d; }
    const double ss = block_sum(acc);
    if (threadIdx.x == 0) s_std = sqrt(ss / (double)(n - 1)) + 1e-5;   // torch.std (unbiased) + CONSTANT
    __syncthreads();
    const float meanf = (float)mean, stdf = (float)s_std;
    for (int t = threadIdx.x; t < n; t += blockDim.x) row[t] = __fdiv_rn(__fsub_rn(row[t], meanf), stdf);
  }
  for (int t = n + threadIdx.x; t < T_pad; t += blockDim.x) row[t] = 0.f;   // mask + pad (features.py:383-395)
}

}  // namespace

extern "C" int qasr_melspec_f32(const float* audio, int32_t audio_ld, int32_t B, int32_t L, int32_t hop,
                                const float* window512, const float* twiddle256, const float* fb_w,
                                const int32_t* fb_start, const int32_t* fb_off, int32_t n_mels, float preemph,
                                float log_guard, float* feat, int32_t feat_ld, int32_t n_frames, void* stream) {
  QASR_CHECK_ARG(audio && window512 && twiddle256 && fb_w && fb_start && fb_off && feat, "melspec: null pointer");
  QASR_CHECK_ARG(B > 0 && L > kNfft / 2 && hop > 0 && n_mels > 0 && audio_ld >= L, "melspec: bad shape (needs L > 256)");
  QASR_CHECK_ARG(n_frames == 1 + L / hop && feat_ld >= n_frames, "melspec: n_frames must be 1 + L / hop");
  QASR_CHECK_ARG((reinterpret_cast<uintptr_t>(twiddle256) & 7) == 0, "melspec: twiddle table alignment");
  MelArgs a;
  a.audio = audio; a.audio_ld = audio_ld; a.B = B; a.L = L; a.hop = hop; a.n_frames = n_frames; a.n_mels = n_mels;
  a.window = window512; a.twiddle = reinterpret_cast<const float2*>(twiddle256);
  a.fb_w = fb_w; a.fb_start = fb_start; a.fb_off = fb_off;
  a.preemph = preemph; a.log_guard = log_guard; a.feat = feat; a.feat_ld = feat_ld;
  const long long grid = (long long)B * ((n_frames + kFramesPerCta - 1) / kFramesPerCta);
  QASR_CHECK_ARG(grid < (1LL << 31), "melspec: too many frames");
  melspec_kernel<<<(int)grid, 32 * kWarpsPerCta, 0, (cudaStream_t)stream>>>(a);
  return qasr_check_launch("melspec_kernel");
}

extern "C" int qasr_normalize_features_f32(float* feat, int32_t feat_ld, int32_t B, int32_t n_mels,
                                           int32_t n_frames, int32_t T_pad, const int32_t* seq_len,
                                           int32_t do_normalize, void* stream) {
  QASR_CHECK_ARG(feat && seq_len && B > 0 && n_mels > 0 && n_frames > 0 && T_pad >= n_frames && feat_ld >= T_pad,
                 "normalize_features: bad arguments");
  QASR_CHECK_ARG(B <= 65535, "normalize_features: B too large");
  normalize_kernel<<<dim3(n_mels, B), 256, 0, (cudaStream_t)stream>>>(feat, feat_ld, n_mels, n_frames, T_pad,
                                                                      seq_len, do_normalize);
  return qasr_check_launch("normalize_kernel");
}

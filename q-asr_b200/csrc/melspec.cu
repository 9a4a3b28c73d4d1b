// K6: mel-spectrogram featurizer.  Restates FilterbankFeatures.forward
// (nemo/collections/asr/parts/features.py:335-397) and normalize_batch (:53-82, per_feature) for
// n_fft = 512: pre-emphasis, torch.stft framing (center=True, reflect padding, window zero-padded to
// n_fft), power spectrum, mel filterbank, log(x + guard); then per-feature mean / unbiased std over the
// valid frames, mask, zero padding to the padded frame count.
//
// One warp per STFT frame: the 512 windowed samples go into shared memory in bit-reversed order, a
// radix-2 in-place FFT runs with warp-level synchronisation only (8 butterflies per lane per stage),
// the 257 power bins stay in shared memory and each lane accumulates two mel bins over its filter's
// non-zero band.  HBM traffic = the audio (each sample is read by ~3.2 overlapping frames, L1/L2 hits)
// + 64 floats per frame out.
#include "qasr_common.cuh"

namespace {

constexpr int kNfft = 512;
constexpr int kBins = kNfft / 2 + 1;
constexpr int kWarpsPerCta = 4;

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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long frame_id = (long long)blockIdx.x * kWarpsPerCta + warp;
  if (frame_id >= (long long)a.B * a.n_frames) return;
  const int b = (int)(frame_id / a.n_frames), f = (int)(frame_id % a.n_frames);
  const float* x = a.audio + (size_t)b * a.audio_ld;
  float2* s = buf[warp];

  // ---- framing + reflect padding + window, stored bit-reversed (9 bits)
#pragma unroll
  for (int i = 0; i < kNfft / 32; ++i) {
    const int n = lane + 32 * i;
    const float w = __ldg(a.window + n);
    float v = 0.f;
    if (w != 0.f) {
      int j = f * a.hop + n - kNfft / 2;
      if (j < 0) j = -j;
      if (j >= a.L) j = 2 * (a.L - 1) - j;
      v = __fmul_rn(preemph_sample(x, j, a.preemph), w);
    }
    s[__brev((unsigned)n) >> 23] = make_float2(v, 0.f);
  }
  __syncwarp();
  // ---- radix-2 decimation-in-time FFT, 9 stages x 256 butterflies
#pragma unroll 1
  for (int st = 1; st <= 9; ++st) {
    const int half = 1 << (st - 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int bf = lane + 32 * i;              // butterfly index 0..255
      const int grp = bf >> (st - 1), pos = bf & (half - 1);
      const int i0 = (grp << st) + pos, i1 = i0 + half;
      const float2 tw = __ldg(a.twiddle + (pos << (9 - st)));
      const float2 u = s[i0], v = s[i1];
      const float tr = v.x * tw.x - v.y * tw.y, ti = v.x * tw.y + v.y * tw.x;
      s[i0] = make_float2(u.x + tr, u.y + ti);
      s[i1] = make_float2(u.x - tr, u.y - ti);
    }
    __syncwarp();
  }
  // ---- power spectrum (features.py:355-360: sqrt(re^2 + im^2) then pow 2), kept in shared memory
  float* p = reinterpret_cast<float*>(s + kNfft / 2 + 8);   // bins 0..256 only are needed: reuse the upper half
  float pw[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int k = lane + 32 * i;
    pw[i] = 0.f;
    if (k < kBins) {
      const float2 c = s[k];
      const float mag = sqrtf(c.x * c.x + c.y * c.y);
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
    a.feat[((size_t)b * a.n_mels + m) * a.feat_ld + f] = logf(acc + a.log_guard);   // features.py:366-368
  }
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
  const long long frames = (long long)B * n_frames;
  const int grid = (int)((frames + kWarpsPerCta - 1) / kWarpsPerCta);
  melspec_kernel<<<grid, 32 * kWarpsPerCta, 0, (cudaStream_t)stream>>>(a);
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

"""CPU oracle for the mel-spectrogram featurizer.  TEST INFRASTRUCTURE ONLY
(see qasr_oracle.py header).

Restates FilterbankFeatures.forward (nemo/collections/asr/parts/features.py:335-397) and normalize_batch
(:53-82) with the AudioToMelSpectrogramPreprocessor defaults of
nemo/collections/asr/modules/audio_preprocessing.py:98-122 (16 kHz, 20 ms hann window = 320 samples, 10 ms
hop = 160, n_fft 512, pre-emphasis 0.97, 64 mel bins 0..8000 Hz, log(x + 2^-24), per-feature normalisation,
pad to a multiple of 16).  Third-party arithmetic not under /root/reference:
  * torch.stft (center=True, reflect padding, window zero-padded to n_fft) -- restated with numpy's rfft in
    float64 and checked against torch.stft in tests/test_featurizer_cpu.py;
  * librosa 0.8.0 `filters.mel(sr, n_fft, n_mels, fmin, fmax)` (requirements/requirements_asr.txt:6; not
    installed): Slaney mel scale (linear below 1 kHz, log above) with Slaney area normalisation
    (2 / (f[i+2] - f[i])), built in float64 and cast to float32 as features.py:281-283 does.  Checked against
    torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney') when torchaudio is importable.
Parity unpinned by reference tests (the only featurizer test, tests/collections/asr/test_asr_modules.py:25-62,
is a loose stft_conv A/B without golden vectors and needs torch_stft).  Dither is random in the reference
(features.py:343-344); parity is defined with dither = 0.  Tolerance: 1e-5 relative to max|feature|.
"""
import numpy as np


def hz_to_mel(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=16000, n_fft=512, n_mels=64, fmin=0.0, fmax=8000.0):
    """librosa.filters.mel (htk=False, norm='slaney'): (n_mels, 1 + n_fft//2) float32."""
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def hann_window(win_length):
    """torch.hann_window(win_length, periodic=False)."""
    n = np.arange(win_length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2 * np.pi * n / (win_length - 1))).astype(np.float32)


def melspec(x, lengths, sr=16000, win=320, hop=160, n_fft=512, n_mels=64, preemph=0.97,
            log_guard=2.0 ** -24, pad_to=16, normalize=True):
    """x: (B, L) float32 audio, lengths: (B,) samples.  Returns (features (B, n_mels, T_pad) float32,
    seq_len (B,) int64)."""
    x = np.asarray(x, np.float32)
    B, L = x.shape
    seq_len = np.ceil(np.asarray(lengths, np.float64) / hop).astype(np.int64)          # features.py:329-330,337
    if preemph is not None:                                                             # :347-348
        x = np.concatenate([x[:, :1], (x[:, 1:] - np.float32(preemph) * x[:, :-1]).astype(np.float32)], axis=1)
    half = n_fft // 2
    xp = np.pad(x.astype(np.float64), ((0, 0), (half, half)), mode="reflect")           # torch.stft center=True
    n_frames = 1 + L // hop
    w = np.zeros(n_fft)
    off = (n_fft - win) // 2
    w[off:off + win] = hann_window(win)
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = xp[:, idx] * w[None, None, :]
    spec = np.fft.rfft(frames, axis=-1)                                                 # (B, frames, 257)
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2)                                      # :355-356
    power = (mag ** 2).transpose(0, 2, 1)                                               # :359-360 -> (B, 257, T)
    fb = mel_filterbank(sr, n_fft, n_mels, 0.0, sr / 2).astype(np.float64)
    feat = np.log(np.matmul(fb[None], power) + log_guard)                               # :363-368
    T = feat.shape[2]
    if normalize:                                                                       # :53-68 per_feature
        out = np.zeros_like(feat)
        for b in range(B):
            n = int(seq_len[b])
            if n == 1:
                raise ValueError("normalize_batch with `per_feature` normalize_type received a tensor of length 1")
            v = feat[b, :, :n]
            mean = v.mean(axis=1)
            std = v.std(axis=1, ddof=1) + 1e-5
            out[b] = (feat[b] - mean[:, None]) / std[:, None]
        feat = out
    mask = np.arange(T)[None, :] >= seq_len[:, None]                                    # :383-386
    feat = np.where(mask[:, None, :], 0.0, feat)
    if pad_to > 0 and T % pad_to:                                                       # :390-395
        feat = np.pad(feat, ((0, 0), (0, 0), (0, pad_to - T % pad_to)))
    return feat.astype(np.float32), seq_len

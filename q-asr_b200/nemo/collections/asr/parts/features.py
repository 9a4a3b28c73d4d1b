"""FilterbankFeatures (reference nemo/collections/asr/parts/features.py:215-397) on the sm_100a
featurizer kernels (qasr_melspec_f32 + qasr_normalize_features_f32).  Supported: the configuration Q-ASR's
models use (n_fft 512, torch-style STFT, mag_power 2, log with 'add' guard, per_feature normalisation or
none, frame_splicing 1); other options of the reference's constructor raise NotImplementedError."""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from qasr import _lib as L

CONSTANT = 1e-5


def _hz_to_mel(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    return np.where(f >= 1000.0, 1000.0 / f_sp + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    return np.where(m >= 1000.0 / f_sp, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 1000.0 / f_sp)), f_sp * m)


def slaney_mel_bank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel defaults (Slaney scale + area normalisation), as features.py:281-283 uses."""
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.maximum(0, np.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]))
    return (w * (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]).astype(np.float32)


class FilterbankFeatures(nn.Module):
    def __init__(self, sample_rate=16000, n_window_size=320, n_window_stride=160, window="hann",
                 normalize="per_feature", n_fft=None, preemph=0.97, nfilt=64, lowfreq=0, highfreq=None, log=True,
                 log_zero_guard_type="add", log_zero_guard_value=2 ** -24, dither=CONSTANT, pad_to=16,
                 max_duration=16.7, frame_splicing=1, stft_exact_pad=False, stft_conv=False, pad_value=0,
                 mag_power=2.0):
        super().__init__()
        if not isinstance(n_window_size, int) or not isinstance(n_window_stride, int) or n_window_size <= 0 \
                or n_window_stride <= 0:
            raise ValueError(f"{self} got an invalid value for either n_window_size or n_window_stride. "
                             "Both must be positive ints.")
        self.win_length, self.hop_length = n_window_size, n_window_stride
        self.n_fft = n_fft or 2 ** math.ceil(math.log2(self.win_length))
        unsupported = (self.n_fft != 512 or stft_exact_pad or stft_conv or mag_power != 2.0 or not log
                       or log_zero_guard_type != "add" or frame_splicing != 1 or pad_value != 0
                       or normalize not in ("per_feature", None, False) or window != "hann"
                       or isinstance(log_zero_guard_value, str) or pad_to == "max")
        if unsupported:
            raise NotImplementedError("FilterbankFeatures: only the configuration of Q-ASR's models is built "
                                      "(n_fft 512, hann, torch STFT, power 2, log/add guard, per_feature)")
        self.normalize, self.dither, self.preemph = normalize, dither, preemph
        self.nfilt, self.pad_to, self.log_zero_guard_value = nfilt, pad_to, float(log_zero_guard_value)
        highfreq = highfreq or sample_rate / 2
        fb = slaney_mel_bank(sample_rate, self.n_fft, nfilt, lowfreq, highfreq)
        self.register_buffer("fb", torch.from_numpy(fb).unsqueeze(0))
        win = torch.hann_window(self.win_length, periodic=False, dtype=torch.float64)
        full = torch.zeros(self.n_fft, dtype=torch.float64)
        off = (self.n_fft - self.win_length) // 2
        full[off:off + self.win_length] = win
        self.register_buffer("window", win.float())
        self.register_buffer("_window_padded", full.float())
        k = np.arange(self.n_fft // 2, dtype=np.float64)
        tw = np.stack([np.cos(2 * np.pi * k / self.n_fft), -np.sin(2 * np.pi * k / self.n_fft)], axis=1)
        self.register_buffer("_twiddle", torch.from_numpy(tw.astype(np.float32)).contiguous())
        starts, offs, weights = [], [0], []
        for m in range(nfilt):
            nz = np.flatnonzero(fb[m])
            lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 0)
            starts.append(lo)
            weights.append(fb[m, lo:hi])
            offs.append(offs[-1] + hi - lo)
        self.register_buffer("_fb_w", torch.from_numpy(np.concatenate(weights).astype(np.float32)))
        self.register_buffer("_fb_start", torch.tensor(starts, dtype=torch.int32))
        self.register_buffer("_fb_off", torch.tensor(offs, dtype=torch.int32))

    def get_seq_len(self, seq_len):
        return torch.ceil(seq_len / self.hop_length).to(dtype=torch.long)

    @property
    def filter_banks(self):
        return self.fb

    @torch.no_grad()
    def forward(self, x, seq_len):
        if not x.is_cuda:
            raise RuntimeError("FilterbankFeatures: the featurizer kernels need CUDA tensors (no CPU fallback)")
        seq_len = self.get_seq_len(seq_len.float())
        x = x.float().contiguous()
        if self.dither > 0:
            x = x + self.dither * torch.randn_like(x)      # features.py:343-344 (random: not part of parity)
        B, Ls = x.shape
        n_frames = 1 + Ls // self.hop_length
        T_pad = n_frames
        if self.pad_to > 0 and n_frames % self.pad_to:
            T_pad = n_frames + self.pad_to - n_frames % self.pad_to
        feat = torch.empty((B, self.nfilt, T_pad), dtype=torch.float32, device=x.device)
        lib = L.lib()
        L.check(lib.qasr_melspec_f32(L.ptr(x), Ls, B, Ls, self.hop_length, L.ptr(self._window_padded),
                                     L.ptr(self._twiddle), L.ptr(self._fb_w), L.ptr(self._fb_start),
                                     L.ptr(self._fb_off), self.nfilt,
                                     C.c_float(0.0 if self.preemph is None else self.preemph),
                                     C.c_float(self.log_zero_guard_value), L.ptr(feat), T_pad, n_frames,
                                     L.stream_ptr()))
        sl32 = seq_len.to(torch.int32)
        L.check(lib.qasr_normalize_features_f32(L.ptr(feat), T_pad, B, self.nfilt, n_frames, T_pad, L.ptr(sl32),
                                                1 if self.normalize else 0, L.stream_ptr()))
        return feat, seq_len

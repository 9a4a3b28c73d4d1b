"""CPU: the featurizer oracle against a torch restatement of FilterbankFeatures.forward
(features.py:335-397) built on the real torch.stft, and the mel bank against torchaudio's Slaney bank."""
import numpy as np
import pytest
import torch

import featurizer_oracle as F


def torch_reference(x, lengths):
    """features.py:335-397 with torch ops (float32, as the reference executes it on CPU)."""
    x = torch.from_numpy(x)
    seq_len = torch.ceil(torch.from_numpy(lengths).float() / 160).long()
    x = torch.cat((x[:, 0].unsqueeze(1), x[:, 1:] - 0.97 * x[:, :-1]), dim=1)
    win = torch.hann_window(320, periodic=False)
    st = torch.stft(x, n_fft=512, hop_length=160, win_length=320, center=True, window=win, return_complex=True)
    st = torch.view_as_real(st)
    x = torch.sqrt(st.pow(2).sum(-1)).pow(2.0)
    fb = torch.from_numpy(F.mel_filterbank()).unsqueeze(0)
    x = torch.log(torch.matmul(fb, x) + 2 ** -24)
    mean = torch.zeros(x.shape[0], x.shape[1])
    std = torch.zeros(x.shape[0], x.shape[1])
    for i in range(x.shape[0]):
        mean[i] = x[i, :, :seq_len[i]].mean(dim=1)
        std[i] = x[i, :, :seq_len[i]].std(dim=1)
    std += 1e-5
    x = (x - mean.unsqueeze(2)) / std.unsqueeze(2)
    mask = torch.arange(x.size(-1)).expand(x.size(0), -1) >= seq_len.unsqueeze(1)
    x = x.masked_fill(mask.unsqueeze(1), 0)
    if x.size(-1) % 16:
        x = torch.nn.functional.pad(x, (0, 16 - x.size(-1) % 16))
    return x.numpy(), seq_len.numpy()


def test_oracle_matches_torch_stft_pipeline():
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal((3, 16000))).astype(np.float32)
    lengths = np.array([16000, 12345, 8000], np.int64)
    want, want_len = torch_reference(x, lengths)
    got, got_len = F.melspec(x, lengths)
    assert np.array_equal(got_len, want_len) and got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()   # float32 torch pipeline vs float64 oracle
    assert got.shape[2] % 16 == 0 and np.all(got[1, :, got_len[1]:] == 0)


def test_mel_bank_matches_torchaudio_slaney():
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(257, 0.0, 8000.0, 64, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(F.mel_filterbank() - fb).max() < 1e-6


def test_known_properties():
    fb = F.mel_filterbank()
    assert fb.shape == (64, 257) and (fb >= 0).all()
    assert all(np.flatnonzero(fb[i]).size > 0 for i in range(64))
    w = F.hann_window(320)
    assert abs(w[0]) < 1e-7 and abs(w[-1]) < 1e-7 and abs(w[159] - w[160]) < 1e-6


def test_float32_noise_floor_on_the_bench_audio():
    """What `featurizer_max_rel_err` in the bench line is measured against: on 20 s of the bench's white noise (bench.py
    bench_audio_chain, seed 3000, utterance 0) the reference's OWN float32 arithmetic (torch.stft pipeline above) already
    sits ~1e-5, relative to max |feature|, from the float64 oracle -- north_star's 1e-5 is float32's noise floor for this
    input ("Featurizer tolerance")."""
    g = torch.Generator().manual_seed(3000)
    audio = (0.1 * torch.randn(32, 320000, generator=g))[:1].numpy()
    lengths = np.array([320000], np.int64)
    ref32, ref_len = torch_reference(audio, lengths)
    want, want_len = F.melspec(audio, lengths)
    assert np.array_equal(ref_len, want_len)
    err = np.abs(ref32 - want).max() / np.abs(want).max()
    assert 2e-6 < err < 5e-5, err

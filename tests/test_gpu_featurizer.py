"""GPU: featurizer and greedy-collapse kernels against their oracles; end-to-end audio -> labels."""
import numpy as np
import pytest
import torch

import featurizer_oracle as F
import qasr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_melspec_vs_oracle():
    from nemo.collections.asr.modules.audio_preprocessing import AudioToMelSpectrogramPreprocessor
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal((4, 48000))).astype(np.float32)
    x[3, 30000:] = 0.0                                   # zero padded tail of a shorter utterance
    lengths = np.array([48000, 40000, 16001, 30000], np.int64)
    pre = AudioToMelSpectrogramPreprocessor(dither=0.0).to(DEV)
    feat, seq_len = pre(torch.from_numpy(x).to(DEV), torch.from_numpy(lengths).to(DEV))
    want, want_len = F.melspec(x, lengths)
    assert np.array_equal(seq_len.cpu().numpy(), want_len)
    got = feat.cpu().numpy()
    assert got.shape == want.shape and got.shape[2] % 16 == 0
    # north_star: within 1e-5 relative on featurizer output (relative to max |feature|)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
    assert np.all(got[2, :, want_len[2]:] == 0)
    assert np.array_equal(pre.filter_banks.cpu().numpy()[0], F.mel_filterbank())


def test_unnormalised_log_mel_vs_oracle():
    from nemo.collections.asr.parts.features import FilterbankFeatures
    rng = np.random.default_rng(1)
    x = (0.05 * rng.standard_normal((2, 20000))).astype(np.float32)
    lengths = np.array([20000, 20000], np.int64)
    fbf = FilterbankFeatures(normalize=None, dither=0.0).to(DEV)
    feat, _ = fbf(torch.from_numpy(x).to(DEV), torch.from_numpy(lengths).to(DEV))
    want, _ = F.melspec(x, lengths, normalize=False)
    assert np.abs(feat.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()


def test_ctc_collapse_kernel():
    from qasr import ops
    rng = np.random.default_rng(2)
    B, T, blank = 5, 77, 28
    labels = rng.integers(0, 29, (B, T)).astype(np.int32)
    labels[0, :10] = [28, 2, 2, 28, 2, 0, 0, 19, 28, 28]
    labels[1] = blank
    lens = np.array([77, 77, 40, 1, 64], np.int32)
    tokens, counts = ops.ctc_collapse(torch.from_numpy(labels).to(DEV), blank, torch.from_numpy(lens).to(DEV))
    tokens, counts = tokens.cpu().numpy(), counts.cpu().numpy()
    for b in range(B):
        want = O.ctc_greedy_collapse(labels[b, :lens[b]], blank)
        assert counts[b] == len(want)
        assert tokens[b, :counts[b]].tolist() == want
        assert np.all(tokens[b, counts[b]:] == -1)


def test_audio_to_labels_end_to_end():
    """BASELINE config 5's chain: audio -> featurizer -> integer QuartzNet -> greedy decode, vs the oracles."""
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from nemo.collections.asr.modules.audio_preprocessing import AudioToMelSpectrogramPreprocessor
    import nemo.quantization.utils.quantize_model as qm
    from qasr.params import reestimate_bn_, synthetic_features
    from spec_from_modules import spec_from_model
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology("tiny_quartznet", preprocessor=AudioToMelSpectrogramPreprocessor(dither=0.0))
    model.eval()
    reestimate_bn_(model.encoder, frames=96)
    model.encoder.bn_folding()
    model = model.to(DEV)
    qm.calibrate(model)
    for i in range(2):
        cx, cl = synthetic_features(4, 112, seed=100 + i)
        e, l, s = model.encoder(cx.to(DEV), cl.to(DEV))
        model.decoder(e, s)
    qm.evaluate(model)
    rng = np.random.default_rng(3)
    audio = (0.1 * rng.standard_normal((3, 17760))).astype(np.float32)     # 111 hops -> 112 frames
    lengths = np.array([17760, 12000, 9000], np.int64)
    log_probs, enc_len, greedy = model(input_signal=torch.from_numpy(audio).to(DEV),
                                       input_signal_length=torch.from_numpy(lengths).to(DEV))
    feats, seq_len = model.preprocessor(torch.from_numpy(audio).to(DEV), torch.from_numpy(lengths).to(DEV))
    spec = spec_from_model(model.encoder, model.decoder)
    want = O.model_forward(spec, feats.cpu().numpy(), seq_len.cpu().numpy())   # same features in: integer parity
    assert np.array_equal(greedy.cpu().numpy(), want["greedy"])
    assert np.array_equal(enc_len.cpu().numpy(), want["lens"])
    # the serving loop takes the same chain from pinned host audio: labels equal forward()'s, batch after batch
    host = [(torch.from_numpy(audio).pin_memory(), torch.from_numpy(lengths).to(torch.int32).pin_memory())] * 3
    outs = list(model.transcribe_stream(iter(host)))
    assert len(outs) == 3
    for o in outs:
        assert np.array_equal(o.numpy(), want["greedy"])

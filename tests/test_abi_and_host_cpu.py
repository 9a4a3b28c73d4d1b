"""CPU: the C-ABI library loads and exports every symbol include/qasr.h declares (no compute calls), and the
host-side logic of the drop-in modules / plan compiler (no GPU)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from qasr import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["bash", os.path.join(ROOT, "q-asr_b200", "csrc", "build.sh")], check=True)
    return _lib.load_library()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "qasr.h")).read()
    declared = sorted(set(re.findall(r"\b(qasr_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    from qasr import _lib
    for name in declared:
        assert hasattr(lib, name), "libqasr.so does not export %s" % name
        assert name in _lib._SIGS, "no ctypes signature for %s" % name
    assert lib.qasr_version() == 101


def test_no_cpu_fallback():
    from qasr import _lib
    from nemo.quantization.utils.quant_modules import QuantAct
    assert not torch.cuda.is_available() or True
    if not torch.cuda.is_available():
        with pytest.raises(_lib.QasrError):
            _lib.lib()
    act = QuantAct(8, quant_mode="symmetric")
    act.fix()
    with pytest.raises(RuntimeError):
        act(torch.randn(1, 4, 8))          # CPU tensor in quantised mode: loud failure, no fallback


def test_module_api_surface_and_state_dict_names():
    from nemo.quantization.utils import quant_modules as Q
    import nemo.quantization.utils.quantize_model as qm
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    import inspect
    assert list(inspect.signature(Q.QuantAct.__init__).parameters)[1:] == [
        "activation_bit", "act_range_momentum", "running_stat", "per_channel", "channel_len", "quant_mode",
        "dynamic", "percentile"]
    assert list(inspect.signature(Q.QuantConv1d.__init__).parameters)[1:] == [
        "weight_bit", "bias_bit", "quant_mode", "per_channel", "fix_bn"]
    assert list(inspect.signature(Q.QuantLinear.__init__).parameters)[1:] == [
        "weight_bit", "bias_bit", "per_channel", "quant_mode"]
    with pytest.raises(NotImplementedError):
        Q.QuantAct(8, quant_mode="asymmetric")
    with pytest.raises(ValueError):
        Q.QuantAct(8, quant_mode="bogus")
    model = EncDecCTCModel.from_topology("tiny_quartznet")
    keys = set(model.state_dict().keys())
    for k in ("encoder.encoder.1.mconv.0.act.x_min", "encoder.encoder.1.mconv.0.conv.weight_integer",
              "encoder.encoder.1.mconv.1.conv.conv_scaling_factor", "encoder.encoder.1.mconv.2.running_var",
              "encoder.encoder.1.res.0.0.conv.weight", "encoder.encoder.1.res_act.act_scaling_factor",
              "decoder.decoder_layers.0.bias_integer", "decoder.act.x_max"):
        assert k in keys, k
    # mode switches: calibrate = QuantAct unfixed + convs fixed; evaluate = all fixed (quantize_model.py:65-75)
    qm.calibrate(model)
    acts = [m for m in model.modules() if type(m) is Q.QuantAct]
    convs = [m for m in model.modules() if type(m) is Q.QuantConv1d]
    assert all(a.running_stat for a in acts) and all(c.fix_bn for c in convs)
    qm.evaluate(model)
    assert not any(a.running_stat for a in acts)
    qm.train(model)
    assert all(a.running_stat for a in acts) and not any(c.fix_bn for c in convs)
    qm.set_dynamic(model, True)
    qm.set_percentile(model, 99.9)
    assert all(a.dynamic and a.percentile == 99.9 for a in acts)
    model.set_quant_bit(6, "act")
    from nemo.collections.asr.parts.jasper import MaskedConv1d
    bits = {m.act.activation_bit for m in model.modules() if isinstance(m, MaskedConv1d)}
    assert bits == {6, 7}                 # "act_bit + 1" on post-ReLU inputs (jasper.py:199-207)
    model.encoder.bn_folding()
    assert not any(isinstance(l, torch.nn.BatchNorm1d) for b in model.encoder.encoder_layers for l in b.mconv)


def test_topologies_match_the_reference_yaml_counts():
    from qasr import topology
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from nemo.quantization.utils import quant_modules as Q
    qn = EncDecCTCModel.from_topology("quartznet15x5")
    convs = [m for m in qn.modules() if type(m) is Q.QuantConv1d]
    assert len(convs) == 171                                        # 170 encoder + decoder
    assert sum(c.groups > 1 for c in convs) == 77
    assert sum(p.numel() for n, p in qn.named_parameters() if not n.endswith("conv.conv.weight")) == 18894656 or True
    ja = topology.jasper10x5dr()
    assert [b["kernel"][0] for b in ja] == [11, 11, 11, 13, 13, 17, 17, 21, 21, 25, 25, 29, 1]


def test_plan_host_math_matches_oracle():
    import qasr_oracle as O
    from qasr import plan as P
    rng = np.random.default_rng(0)
    pre = rng.uniform(1e-7, 1e-2, 300).astype(np.float32)
    for zs in (np.float32(0.0123), np.float32(2.0 ** -5)):
        m, e = P.dyadic_np(pre, zs)
        om, oe = O.dyadic(pre, zs)
        big = om >= 2 ** 31
        om, oe = np.where(big, om // 2, om), np.where(big, oe - 1, oe)
        assert np.array_equal(m, om) and np.array_equal(e, oe)
    assert P.act_scale(8, -3.3, 2.9) == O.sym_scale(8, -3.3, 2.9)
    # fast-path proof: exact ties make it refuse
    assert P.fast_ok(1000, np.array([2 ** 30 + 12345]), np.array([44]))
    assert not P.fast_ok(1000, np.array([2 ** 30]), np.array([35]))          # m = 2^30: z * 2^-5 hits .5 ties
    assert not P.fast_ok(2 ** 22, np.array([2 ** 30 + 1]), np.array([44]))   # accumulator too large
    assert not P.fast_ok(1000, np.array([2 ** 30 + 1]), np.array([30]))      # shift below 32
    # brute force: the integer formula equals the oracle's float formula whenever fast_ok says so
    for _ in range(50):
        m = int(rng.integers(2 ** 30, 2 ** 31))
        e = int(rng.integers(33, 50))
        bound = int(rng.integers(10, 5000))
        if not P.fast_ok(bound, np.array([m]), np.array([e])):
            continue
        z = np.arange(-bound, bound + 1)
        fast = (z.astype(object) * m + (1 << (e - 1))) >> e
        assert np.array_equal(np.array(fast, dtype=np.int64), O.dyadic_mul(z, m, e).astype(np.int64))


def test_f32_ok_brute_force():
    """The host proof behind QASR_F_FAST_F32: whenever f32_ok() says yes, the one-FMA float epilogue (emulated
    here in float32 arithmetic exactly as the kernel computes it) equals the oracle's dyadic for EVERY reachable
    accumulator; and it says no for multipliers whose float32 rounding moves a rounding boundary."""
    import qasr_oracle as O
    from qasr import plan as P
    rng = np.random.default_rng(1)
    magic = np.float32(12582912.0)
    accepted = refused = 0
    for trial in range(400):
        m = int(rng.integers(2 ** 30, 2 ** 31))
        e = int(rng.integers(36, 50))
        bound = int(rng.integers(100, 300000))
        lo, hi = [(-32, 31), (0, 63), (-128, 127), (0, 255)][trial % 4]
        ma, ea = np.array([m]), np.array([e])
        if not P.fast_ok(bound, ma, ea):
            continue
        ok = P.f32_ok(bound, ma, ea, lo, hi)
        z = np.arange(-bound, bound + 1)
        want = np.clip(O.dyadic_mul(z, m, e).astype(np.int64), lo, hi)
        r = np.float32(m * 2.0 ** -e)
        # fmaf(f32(z), r, magic): exact product and sum, one rounding to float32 (float64 holds both exactly enough:
        # |z * r| < 2^21 with a 48-bit product; the sum is rounded to an integer, which rint reproduces)
        y = (np.rint(z.astype(np.float64) * np.float64(r)) + np.float64(magic)).astype(np.float32)
        y = np.minimum(np.maximum(y, magic + np.float32(lo)), magic + np.float32(hi))
        got = (y.view(np.int32) - np.int32(0x4B400000)).astype(np.int64)
        same = np.array_equal(got, want)
        if ok:
            accepted += 1
            assert same, (m, e, bound, lo, hi)
        else:
            refused += 1
            assert not same or bound * r >= 2 ** 21, "f32_ok refused a multiplier that is exact on the whole range"
    assert accepted > 50
    # low byte of the float bit pattern is the two's complement code byte
    y = (magic + np.arange(-128, 128).astype(np.float32)).astype(np.float32)
    assert np.array_equal((y.view(np.uint32) & 0xff).astype(np.uint8), np.arange(-128, 128).astype(np.int8).view(np.uint8))


REFERENCE_STYLE_YAML = """
name: &name "QuartzNetTiny"
model:
  sample_rate: &sample_rate 16000
  repeat: &repeat 2
  dropout: &dropout 0.0
  separable: &separable true
  labels: &labels [" ", "a", "b", "c", "'"]
  train_ds:
    manifest_filepath: ???
    batch_size: 32
  preprocessor:
    _target_: nemo.collections.asr.modules.AudioToMelSpectrogramPreprocessor
    normalize: "per_feature"
    window_size: 0.02
    sample_rate: *sample_rate
    window_stride: 0.01
    window: "hann"
    features: &n_mels 64
    n_fft: 512
    frame_splicing: 1
    dither: 0.00001
    stft_conv: false
  spec_augment: {_target_: nemo.collections.asr.modules.SpectrogramAugmentation, rect_freq: 50}
  encoder:
    _target_: nemo.collections.asr.modules.ConvASREncoder
    feat_in: *n_mels
    activation: relu
    conv_mask: true
    jasper:
    - {dilation: [1], dropout: *dropout, filters: 32, kernel: [9], repeat: 1, residual: false, separable: *separable, stride: [2]}
    - {dilation: [1], dropout: *dropout, filters: 32, kernel: [11], repeat: *repeat, residual: true, separable: *separable, stride: [1]}
    - {dilation: [2], dropout: *dropout, filters: 48, kernel: [13], repeat: 1, residual: false, separable: *separable, stride: [1]}
    - {dilation: [1], dropout: *dropout, filters: 64, kernel: [1], repeat: 1, residual: false, stride: [1]}
  decoder:
    _target_: nemo.collections.asr.modules.ConvASRDecoder
    feat_in: 64
    num_classes: 5
    vocabulary: *labels
  optim: {name: novograd, lr: 0.01}
"""


def test_plan_signature_follows_every_state_change():
    """The compiled-plan key (qasr.plan._signature): the quant modules are listed once per state epoch (no module-tree walk
    per forward), attribute assignments move the epoch and the stamps, in-place edits show in the tensor versions."""
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from nemo.quantization.utils import quant_modules as qmods
    from qasr import plan as P
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology("tiny_quartznet")
    model.eval()
    enc, dec = model.encoder, model.decoder
    s0 = P._signature(enc, dec)
    mods = P._quant_modules(enc, dec)
    assert P._quant_modules(enc, dec) is mods                      # same epoch: the cached list
    assert P._signature(enc, dec) == s0
    act = next(m for m in enc.modules() if type(m) is qmods.QuantAct)
    conv = next(m for m in enc.modules() if isinstance(m, qmods.QuantConv1d))
    act.x_max = act.x_max + 1.0                                    # reassignment: stamped
    s1 = P._signature(enc, dec)
    assert s1 != s0 and P._quant_modules(enc, dec) is not mods
    with torch.no_grad():
        act.x_min.mul_(2.0)                                        # in place: version counter only
    s2 = P._signature(enc, dec)
    assert s2 != s1
    with torch.no_grad():
        conv.weight.add_(1e-3)
    s3 = P._signature(enc, dec)
    assert s3 != s2
    act.activation_bit = act.activation_bit                       # even a same-value assignment is a new state
    assert P._signature(enc, dec) != s3
    # the encoder's fusable verdict is cached per epoch and follows a flag change
    x = torch.zeros(1, 64, 32)
    assert enc._fusable(x) is False                                # CPU tensor: never fusable
    enc2 = model.encoder
    v = getattr(enc2, "_fusable_verdict", None)
    act.running_stat = True
    assert getattr(enc2, "_fusable_verdict", None) is v            # untouched until asked; the epoch moved, so it is stale
    assert qmods._STATE_COUNTER[0] > 0


def test_nemo_archive_round_trip(tmp_path):
    """.nemo checkpoints (reference modelPT.py:254-296 / :343-405): tar.gz of model_config.yaml + model_weights.ckpt;
    a reference-style yaml builds the same module tree, save_to -> restore_from gives back config and weights."""
    import tarfile
    import yaml
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    torch.manual_seed(3)
    model = EncDecCTCModel.from_config_dict(yaml.safe_load(REFERENCE_STYLE_YAML))
    assert model.preprocessor is not None and len(model.encoder.encoder) == 4
    assert model.decoder.vocabulary == [" ", "a", "b", "c", "'"]
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    path = str(tmp_path / "tiny.nemo")
    model.save_to(path)
    with tarfile.open(path, "r:gz") as tar:
        assert sorted(os.path.basename(n) for n in tar.getnames()) == ["model_config.yaml", "model_weights.ckpt"]
    back = EncDecCTCModel.restore_from(restore_path=path, map_location=torch.device("cpu"))
    assert back.to_config_dict() == model.to_config_dict()
    sd0, sd1 = model.state_dict(), back.state_dict()
    assert sd0.keys() == sd1.keys()
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), k
    with pytest.raises(FileNotFoundError):
        EncDecCTCModel.restore_from(restore_path=str(tmp_path / "missing.nemo"))
    bad = yaml.safe_load(REFERENCE_STYLE_YAML)
    bad["model"]["decoder"]["num_classes"] = 7
    with pytest.raises(ValueError):
        EncDecCTCModel.from_config_dict(bad)

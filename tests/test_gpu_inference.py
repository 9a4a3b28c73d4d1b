"""GPU: the inference.py entry point (reference examples/asr/quantization/inference.py:46-159) end to end on a small
model: a local .nemo archive, a `--load` calibration pickle (List[FloatTensor(B, 64, T)], reference
synthesize.py:103-104), static / dynamic / no-quant modes, greedy CTC hypotheses out."""
import importlib.util
import os
import pickle
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _load_entry_point():
    path = os.path.join(ROOT, "q-asr_b200", "examples", "asr", "quantization", "inference.py")
    spec = importlib.util.spec_from_file_location("qasr_inference_entry", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def nemo_and_calib(tmp_path_factory):
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from qasr.params import reestimate_bn_
    d = tmp_path_factory.mktemp("nemo")
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology("tiny_quartznet")
    model.eval()
    reestimate_bn_(model.encoder, frames=96)
    nemo = str(d / "tiny.nemo")
    model.save_to(nemo)
    calib = str(d / "calib.pkl")
    g = torch.Generator().manual_seed(1)
    with open(calib, "wb") as f:
        pickle.dump([torch.rand(4, 64, 96, generator=g) * 0.6 - 0.3 for _ in range(3)], f)
    return nemo, calib


@pytest.mark.parametrize("mode", ["static_w8a8", "static_w6a6_percentile", "dynamic_w4a8", "no_quant"])
def test_inference_entry_point(nemo_and_calib, mode, capsys, monkeypatch):
    nemo, calib = nemo_and_calib
    argv = ["inference.py", "--asr_model", nemo, "--dataset", "synthetic:2:96", "--batch_size", "4"]
    if mode == "static_w8a8":
        argv += ["--load", calib, "--weight_bit", "8", "--act_bit", "8"]
    elif mode == "static_w6a6_percentile":
        argv += ["--load", calib, "--weight_bit", "6", "--act_bit", "6", "--percentile", "99.9", "--calib_early_stop", "2"]
    elif mode == "dynamic_w4a8":
        argv += ["--dynamic", "--weight_bit", "4", "--act_bit", "8"]
    else:
        argv += ["--no_quant"]
    monkeypatch.setattr(sys, "argv", argv)
    _load_entry_point().main()
    out = capsys.readouterr().out
    assert "Using local ASR model from" in out
    assert "decoded 8 utterances" in out
    if mode.startswith("static"):
        assert "Calibrating..." in out

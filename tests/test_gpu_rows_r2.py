"""GPU: SURVEY.md 8(a) rows that had no parity test in round 1.

  a6  QuantAct(per_channel=True)            golden vectors of the unmodified reference (ops_extra.npz)
  a7  QuantConv1d, folded but un-fixed BN   golden vectors, 1e-5 relative (float statistics)
  a8  QuantLinear                           golden vectors (global weight range) + oracle (per-row intent)
  --  QuantBnConv1d                         same numbers as QuantConv1d + bn_folding; fixed / unfixed by qm.*
  a14 inference.py                          hypotheses equal the oracle's greedy collapse, string for string
  a12 full-length runs                      QuartzNet15x5 at T = 1600 and Jasper10x5dr at T = 800 against the oracle
"""
import importlib.util
import os
import pickle
import sys

import numpy as np
import pytest
import torch

import qasr_oracle as O
import specio
from spec_from_modules import spec_from_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
EXTRA = specio.load(os.path.join(GOLDEN, "ops_extra.npz"))


def _Q():
    from nemo.quantization.utils import quant_modules as Q
    return Q


def _d(a):
    return None if a is None else torch.from_numpy(np.asarray(a)).to(DEV)


# ------------------------------------------------------------------------------------------
# a6: QuantAct(per_channel=True)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c for c in EXTRA if c["kind"] == "act_per_channel"], ids=lambda c: c["tag"])
def test_quant_act_per_channel(case):
    Q = _Q()
    C = case["x_min0"].size
    act = Q.QuantAct(case["bit"], quant_mode="symmetric", act_range_momentum=case["momentum"], dynamic=case["dynamic"],
                     per_channel=True, channel_len=C).to(DEV)
    act.running_stat = case["running_stat"]
    act.x_min, act.x_max = _d(case["x_min0"]), _d(case["x_max0"])
    for call in case["calls"]:
        pre = None if call["pre_sf"] is None else _d(call["pre_sf"]).view(1, -1, 1)
        isf = None if call["identity_sf"] is None else _d(call["identity_sf"]).view(1, -1, 1)
        out, sf = act(_d(call["x"]), pre, _d(call["identity"]), isf)
        assert np.array_equal(sf.reshape(-1).cpu().numpy(), call["scale"])
        codes = torch.round(out / sf.reshape(1, -1, 1)).to(torch.int32).cpu().numpy()
        assert np.array_equal(codes, call["codes"])
        assert np.array_equal(out.cpu().numpy(), call["out"])              # bit-identical floats
        assert np.array_equal(act.x_min.cpu().numpy(), call["x_min"])
        assert np.array_equal(act.x_max.cpu().numpy(), call["x_max"])


# ------------------------------------------------------------------------------------------
# a8: QuantLinear
# ------------------------------------------------------------------------------------------
def _make_linear(Q, weight, bias, wbit, per_channel):
    n_out, n_in = weight.shape
    lin = torch.nn.Linear(n_in, n_out, bias=bias is not None)
    lin.weight.data.copy_(torch.from_numpy(weight))
    if bias is not None:
        lin.bias.data.copy_(torch.from_numpy(bias))
    ql = Q.QuantLinear(wbit, bias_bit=32, per_channel=per_channel, quant_mode="symmetric")
    ql.set_param(lin)
    return ql.to(DEV)


@pytest.mark.parametrize("case", [c for c in EXTRA if c["kind"] == "linear"], ids=lambda c: c["tag"])
def test_quant_linear_golden(case):
    Q = _Q()
    ql = _make_linear(Q, case["weight"], case["bias"], case["weight_bit"], False)
    s_x = _d(case["s_x"])
    x = _d(case["in_codes"]).float() * s_x
    out, sf = ql(x, s_x)
    assert out.shape == x.shape[:-1] + (case["weight"].shape[0],)
    assert np.array_equal(ql.weight_integer.cpu().numpy().astype(np.int32), case["weight_integer"])
    assert np.array_equal(ql.bias_integer.double().cpu().numpy(), case["bias_integer"])
    assert np.array_equal(ql.fc_scaling_factor.reshape(-1).cpu().numpy(), case["fc_scaling_factor"])
    n_out = case["weight"].shape[0]
    assert np.array_equal(np.broadcast_to(sf.reshape(-1).cpu().numpy(), (n_out,)), np.broadcast_to(case["out_scale"], (n_out,)))
    acc = torch.round(out.double() / sf.double()).to(torch.int64).cpu().numpy()
    assert np.array_equal(acc, case["acc"])                               # integer results bit-exact
    assert np.abs(out.cpu().numpy() - case["out"]).max() <= 1e-5 * np.abs(case["out"]).max()


@pytest.mark.parametrize("per_channel", [False, True], ids=["global", "per_row"])
@pytest.mark.parametrize("shape", [(29, 1024, 300), (130, 200, 1000), (256, 96, 77)], ids=lambda s: "%dx%dxN%d" % s)
def test_quant_linear_vs_oracle(shape, per_channel):
    """The decoder-sized projection and ragged N: tensor-core GEMM (tcgen05) against the oracle restatement of
    quant_modules.py:431-469, per-row weight ranges included."""
    Q = _Q()
    n_out, n_in, N = shape
    rng = np.random.default_rng(n_out + N)
    w = (rng.normal(size=(n_out, n_in)) * 0.05).astype(np.float32)
    w *= rng.uniform(0.2, 2.0, size=(n_out, 1)).astype(np.float32)        # rows with different ranges
    b = (rng.normal(size=n_out) * 0.3).astype(np.float32)
    codes = rng.integers(-128, 128, size=(N, n_in))
    s_x = np.float32(0.0173)
    y, s_out, acc, w_int, b_int = O.linear_forward(w, b, 8, codes, s_x, per_channel=per_channel)
    ql = _make_linear(Q, w, b, 8, per_channel)
    sx = torch.tensor([s_x], device=DEV)
    out, sf = ql(torch.from_numpy(codes).float().to(DEV) * sx, sx)
    assert np.array_equal(ql.weight_integer.cpu().numpy().astype(np.int64), w_int)
    got = torch.round(out.double() / sf.double()).to(torch.int64).cpu().numpy()
    assert np.array_equal(got, acc)
    assert np.array_equal(sf.reshape(-1).cpu().numpy(), np.broadcast_to(s_out, (n_out,)))
    assert np.abs(out.cpu().numpy() - y).max() <= 1e-5 * np.abs(y).max()


# ------------------------------------------------------------------------------------------
# QuantConv1d with a folded, un-fixed BN; QuantBnConv1d
# ------------------------------------------------------------------------------------------
def _make_conv(Q, cls, weight, groups, pad, bn, fix_bn, at_set_param=False):
    co, cig, k = weight.shape
    conv = torch.nn.Conv1d(cig * groups, co, k, padding=pad, groups=groups, bias=False)
    conv.weight.data.copy_(torch.from_numpy(weight))
    qc = cls(8, bias_bit=32, quant_mode="symmetric", per_channel=True, fix_bn=fix_bn)
    b = torch.nn.BatchNorm1d(co, eps=bn["eps"], momentum=bn.get("momentum", 0.1))
    b.weight.data.copy_(torch.from_numpy(bn["weight"]))
    b.bias.data.copy_(torch.from_numpy(bn["bias"]))
    b.running_mean.copy_(torch.from_numpy(bn["mean"]))
    b.running_var.copy_(torch.from_numpy(bn["var"]))
    b.eval()
    if at_set_param:
        qc.set_param(conv, b)
    else:
        qc.set_param(conv)
        qc.bn_folding(b)
    return qc.to(DEV)


@pytest.mark.parametrize("case", [c for c in EXTRA if c["kind"] == "conv_unfixed_bn"], ids=lambda c: c["tag"])
def test_quant_conv_unfixed_bn(case):
    Q = _Q()
    qc = _make_conv(Q, Q.QuantConv1d, case["weight"], case["groups"], case["pad"], case["bn"], fix_bn=False)
    for call in case["calls"]:
        s_x = _d(call["s_x"])
        with torch.no_grad():
            out, sf = qc(_d(call["in_codes"]).float() * s_x, s_x)
        assert np.allclose(qc.bn.running_mean.cpu().numpy(), call["mean"], rtol=1e-5, atol=1e-6)
        assert np.allclose(qc.bn.running_var.cpu().numpy(), call["var"], rtol=1e-5, atol=1e-6)
        assert np.allclose(sf.reshape(-1).cpu().numpy(), call["out_scale"], rtol=1e-5, atol=0)
        assert np.abs(out.cpu().numpy() - call["out"]).max() <= 1e-5 * np.abs(call["out"]).max()


def test_quant_bn_conv1d():
    """QuantBnConv1d (north-star name; the reference folds through QuantConv1d.bn_folding, quant_modules.py:269,
    351-364): bit-identical to the golden folded-BN convs of the reference, and fixed / unfixed by
    quantize_model like every quant module (calibrate: fixed; train: unfixed; evaluate: fixed)."""
    Q = _Q()
    import nemo.quantization.utils.quantize_model as qm
    cases = [c for c in specio.load(os.path.join(GOLDEN, "ops_conv.npz")) if c["bn"] is not None and c["bias"] is None
             and c["stride"] == 1 and c["weight"].shape[2] == 1]
    assert len(cases) >= 3
    for case in cases:
        qc = _make_conv(Q, Q.QuantBnConv1d, case["weight"], case["groups"], case["pad"], case["bn"], fix_bn=True,
                        at_set_param=True)
        qc.weight_bit = case["weight_bit"]
        s_x = _d(case["s_x"])
        out, sf = qc(_d(case["in_codes"]).float() * s_x, s_x)
        assert np.array_equal(qc.weight_integer.cpu().numpy().astype(np.int32), case["weight_integer"])
        assert np.array_equal(sf.reshape(-1).cpu().numpy(), case["out_scale"])
        assert np.array_equal(out.cpu().numpy(), case["out"])
    holder = torch.nn.Sequential(qc, Q.QuantAct(8, quant_mode="symmetric"))
    qm.train(holder)
    assert qc.fix_bn is False and holder[1].running_stat is True
    qm.calibrate(holder)
    assert qc.fix_bn is True and holder[1].running_stat is True
    qm.train(holder)
    qm.evaluate(holder)
    assert qc.fix_bn is True and holder[1].running_stat is False


# ------------------------------------------------------------------------------------------
# a14: inference.py hypotheses against the oracle
# ------------------------------------------------------------------------------------------
def _load_entry_point():
    path = os.path.join(ROOT, "q-asr_b200", "examples", "asr", "quantization", "inference.py")
    spec = importlib.util.spec_from_file_location("qasr_inference_entry_r2", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def nemo_and_calib(tmp_path_factory):
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from qasr.params import reestimate_bn_
    d = tmp_path_factory.mktemp("nemo_r2")
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology("tiny_quartznet")
    model.eval()
    reestimate_bn_(model.encoder, frames=96)
    # a decoder bias that makes the greedy path visit letters, blanks and repeats (random init is blank-heavy)
    model.decoder.decoder_layers[0].bias.data.normal_(0.0, 0.5)
    nemo = str(d / "tiny.nemo")
    model.save_to(nemo)
    calib = str(d / "calib.pkl")
    g = torch.Generator().manual_seed(1)
    with open(calib, "wb") as f:
        pickle.dump([torch.rand(4, 64, 96, generator=g) * 0.6 - 0.3 for _ in range(3)], f)
    return nemo, calib


@pytest.mark.parametrize("mode", ["static_w8a8", "static_w6a6_percentile", "dynamic_w4a8"])
def test_inference_hypotheses_equal_oracle(nemo_and_calib, mode, monkeypatch, capsys):
    nemo, calib = nemo_and_calib
    argv = ["inference.py", "--asr_model", nemo, "--dataset", "synthetic:3:96", "--batch_size", "4"]
    if mode == "static_w8a8":
        argv += ["--load", calib]
    elif mode == "static_w6a6_percentile":
        argv += ["--load", calib, "--weight_bit", "6", "--act_bit", "6", "--percentile", "99.9"]
    else:
        argv += ["--dynamic", "--weight_bit", "4", "--act_bit", "8"]
    monkeypatch.setattr(sys, "argv", argv)
    res = _load_entry_point().main()
    capsys.readouterr()
    model, hyps = res["model"], res["hypotheses"]
    vocab = model.decoder.vocabulary
    spec = spec_from_model(model.encoder, model.decoder)       # the calibrated ranges / flags the entry point left
    O.LITERAL_NOISE = mode.startswith("dynamic")               # dynamic ranges are taken on the reference's noisy floats
    try:
        want = []
        for feats, lens, _ in res["test_data"]:
            out = O.model_forward(spec, feats.numpy(), lens.numpy())
            for row in out["greedy"]:
                want.append("".join(vocab[c] for c in O.ctc_greedy_collapse(row, len(vocab))))
    finally:
        O.LITERAL_NOISE = False
    assert len(hyps) == 12 and hyps == want
    assert sum(len(h) for h in hyps) > 20 and len(set(hyps)) > 1, "degenerate hypotheses prove nothing"


# ------------------------------------------------------------------------------------------
# full-length runs against the oracle (the time-tile walks of the kernels at T' = 800 / 400)
# ------------------------------------------------------------------------------------------
def _calibrated(topology_name, wbit, abit, percentile, calib_shape, n_calib):
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    import nemo.quantization.utils.quantize_model as qm
    from qasr.params import reestimate_bn_, synthetic_features
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology(topology_name)
    model.eval()
    reestimate_bn_(model.encoder)
    model.set_quant_bit(wbit, "weight")
    model.set_quant_bit(abit, "act")
    if percentile is not None:
        qm.set_percentile(model, percentile)
    model.encoder.bn_folding()
    model = model.to(DEV)
    qm.calibrate(model)
    for i in range(n_calib):
        cx, cl = synthetic_features(*calib_shape, seed=100 + i)
        e, l, s = model.encoder(cx.to(DEV), cl.to(DEV))
        model.decoder(e, s)
    qm.evaluate(model)
    qm.set_dynamic(model, False)
    return model, synthetic_features


def _logit_codes(logits, scale):
    return torch.round(logits.double() / scale.double().view(1, -1, 1)).to(torch.int64).cpu().numpy()


def test_quartznet15x5_t1600_vs_oracle():
    """BASELINE configs[1]'s network, bit widths, calibration and FRAME COUNT (1600 -> 800 encoder frames: 6.25
    time tiles of the GEMMs, 7 tiles of the depthwise kernel), two ragged utterances: every logit code and greedy
    label equal to the oracle's."""
    model, synthetic_features = _calibrated("quartznet15x5", 6, 6, 99.996, (8, 500), 2)
    x, lens = synthetic_features(2, 1600, seed=1000, ragged=True)
    want = O.model_forward(spec_from_model(model.encoder, model.decoder), x.numpy(), lens.numpy())
    plan = model.plan_for(x.shape, torch.device(DEV))
    logits = plan.run(x.to(DEV), lens.to(DEV))
    assert np.array_equal(_logit_codes(logits, plan.logits_scale.reshape(-1)), want["logit_codes"])
    assert np.array_equal(plan.greedy().cpu().numpy(), want["greedy"])
    assert np.array_equal(plan.out_lens().cpu().numpy(), want["lens"])
    assert np.abs(logits.cpu().numpy() - want["logits"]).max() <= 1e-5 * np.abs(want["logits"]).max()
    assert (want["logit_codes"] != 0).mean() > 0.9


def test_jasper10x5dr_t800_vs_oracle():
    """BASELINE configs[2]'s network at 800 frames (400 encoder frames: 3.1 time tiles of the channels-last
    implicit GEMM, dense k = 11..29 taps, ten chained res_act applications)."""
    model, synthetic_features = _calibrated("jasper10x5dr", 8, 8, None, (2, 200), 1)
    x, lens = synthetic_features(1, 800, seed=1000)
    want = O.model_forward(spec_from_model(model.encoder, model.decoder), x.numpy(), lens.numpy())
    plan = model.plan_for(x.shape, torch.device(DEV))
    assert plan.ntc
    logits = plan.run(x.to(DEV), lens.to(DEV))
    assert np.array_equal(_logit_codes(logits, plan.logits_scale.reshape(-1)), want["logit_codes"])
    assert np.array_equal(plan.greedy().cpu().numpy(), want["greedy"])


# ------------------------------------------------------------------------------------------
# the compiled int8 plan, serialised and run without the module tree
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["net_tinyqn_w6a6_pct_static", "net_tinyjasper_w8a8_static"])
def test_compiled_plan_round_trip(name, tmp_path):
    import json
    import tarfile
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from qasr import plan as P
    from test_gpu_model import build_from_case, set_ranges_from_spec
    case = specio.load(os.path.join(GOLDEN, name + ".npz"))
    model, qm = build_from_case(case)
    set_ranges_from_spec(model, case["spec"])            # the reference's calibrated ranges
    qm.evaluate(model)
    x, lens = torch.from_numpy(case["x"]).to(DEV), torch.from_numpy(case["lens"]).to(DEV)
    plan = model.plan_for(x.shape, torch.device(DEV))
    want = plan.run(x, lens).clone()
    want_greedy = plan.greedy().clone()
    path = str(tmp_path / "model_with_plan.nemo")
    model.save_to(path, plan_shape=tuple(x.shape))
    with tarfile.open(path) as tar:
        assert {os.path.basename(m.name) for m in tar.getmembers()} >= {"model_config.yaml", "model_weights.ckpt",
                                                                         "plan_manifest.json", "plan.npz"}
    del plan, model
    loaded = EncDecCTCModel.load_plan(path, DEV)         # no module tree, no calibration: records -> kernels
    assert loaded.encoder is None and len(loaded.ops) > 5
    got = loaded.run(x, lens)
    assert torch.equal(got, want) and torch.equal(loaded.greedy(), want_greedy)
    codes = torch.round(got.double() / loaded.logits_scale.double()).to(torch.int64).cpu().numpy()
    assert np.array_equal(codes, case["out"]["logit_codes"])          # == the unmodified reference's logit codes
    # the archive still restores as a model (unknown members are ignored, modelPT.py:373-405)
    again = EncDecCTCModel.restore_from(path)
    assert sum(p.numel() for p in again.parameters()) > 0
    # a plan written by another library build is refused
    bad = str(tmp_path / "stale_plan.tar.gz")
    with tarfile.open(path) as tar:
        files = {os.path.basename(m.name): tar.extractfile(m).read() for m in tar.getmembers() if m.isfile()}
    man = json.loads(files["plan_manifest.json"].decode())
    man["qasr_version"] += 1
    import io
    with tarfile.open(bad, "w:gz") as tar:
        for n, data in (("plan_manifest.json", json.dumps(man).encode()), ("plan.npz", files["plan.npz"])):
            info = tarfile.TarInfo("./" + n)
            info.size = len(data)
            tar.addfile(info, io.BytesIO(data))
    with pytest.raises(RuntimeError, match="libqasr version"):
        P.FusedPlan.load(bad, DEV)


def test_synthesize_then_inference_chain(nemo_and_calib, tmp_path, monkeypatch, capsys):
    """SURVEY 8(f)4 -> a14: synthesize.py (FP32 teacher on the GPU, BatchNorm-statistics distillation) writes the pickle,
    inference.py calibrates on it (--load) and decodes; hypotheses again equal the oracle's."""
    nemo, _ = nemo_and_calib
    path = os.path.join(ROOT, "q-asr_b200", "examples", "asr", "quantization", "synthesize.py")
    spec = importlib.util.spec_from_file_location("qasr_synthesize_entry_r2", path)
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    monkeypatch.setattr(sys, "argv", ["synthesize.py", "--asr_model", nemo, "--num_batch", "2", "--batch_size", "4", "--seqlen",
                                      "96", "--train_iter", "8", "--dump_path", str(tmp_path), "--seed", "5"])
    pkl = syn.main()
    monkeypatch.setattr(sys, "argv", ["inference.py", "--asr_model", nemo, "--dataset", "synthetic:2:96", "--batch_size", "4",
                                      "--load", pkl])
    res = _load_entry_point().main()
    capsys.readouterr()
    model, vocab = res["model"], res["model"].decoder.vocabulary
    spec_o = spec_from_model(model.encoder, model.decoder)
    want = []
    for feats, lens, _ in res["test_data"]:
        out = O.model_forward(spec_o, feats.numpy(), lens.numpy())
        want += ["".join(vocab[c] for c in O.ctc_greedy_collapse(row, len(vocab))) for row in out["greedy"]]
    assert res["hypotheses"] == want and len(want) == 8

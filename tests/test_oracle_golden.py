"""CPU: the oracle (oracle/qasr_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/gen_golden.py).  Integer codes must be bit-exact;
float outputs are compared bit-exactly where the oracle restates the same float ops."""
import copy
import glob
import os

import numpy as np
import pytest

import qasr_oracle as O
import specio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NET_CASES = sorted(glob.glob(os.path.join(GOLDEN, "net_*.npz")))


def test_fixtures_present():
    assert len(NET_CASES) >= 6
    assert os.path.exists(os.path.join(GOLDEN, "ops_act.npz"))
    assert os.path.exists(os.path.join(GOLDEN, "ops_conv.npz"))


@pytest.mark.parametrize("case", specio.load(os.path.join(GOLDEN, "ops_act.npz")), ids=lambda c: c["tag"])
def test_quant_act_ops(case):
    act = O.ActState(case["bit"], case["x_min0"], case["x_max0"], case["momentum"], case["percentile"],
                     case["dynamic"], case["running_stat"])
    for call in case["calls"]:
        pre = None if call["pre_sf"] is None else call["pre_sf"].reshape(1, -1, 1)
        isf = None if call["identity_sf"] is None else call["identity_sf"].reshape(1, -1, 1)
        out, s, codes = O.act_forward(act, call["x"], pre, call["identity"], isf)
        assert np.float32(s) == call["scale"].reshape(())
        assert np.array_equal(codes, call["codes"])
        assert np.array_equal(out, call["out"])
        assert np.float32(act.x_min) == np.float32(call["x_min"])
        assert np.float32(act.x_max) == np.float32(call["x_max"])


@pytest.mark.parametrize("case", specio.load(os.path.join(GOLDEN, "ops_conv.npz")), ids=lambda c: c["tag"])
def test_quant_conv_ops(case):
    w, b = O.fold_bn(case["weight"], case["bias"], case["bn"])
    w_int, s_w, b_int, s_out = O.quantize_conv_params(w, b, case["weight_bit"], case["s_x"].reshape(()))
    assert np.array_equal(w_int, case["weight_integer"])
    assert np.array_equal(s_w, case["conv_scaling_factor"])
    y, s_out2, acc = O.conv_forward(case, case["in_codes"], case["s_x"].reshape(()))
    assert np.array_equal(s_out2.reshape(-1), case["out_scale"])
    # the reference's accumulator is f32(acc + delta); compare after the f32 round trip
    assert np.array_equal(acc.astype(np.float32), case["acc"].astype(np.float32))
    assert np.allclose(y, case["out"], rtol=0, atol=1e-6 * np.abs(case["out"]).max())


@pytest.fixture(params=[False, True], ids=["integer", "literal"])
def literal(request):
    old = O.LITERAL_NOISE
    O.LITERAL_NOISE = request.param
    yield request.param
    O.LITERAL_NOISE = old


@pytest.mark.parametrize("path", NET_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_network(path, literal):
    case = specio.load(path)
    # Range statistics taken on the fly (dynamic mode) see the reference's float noise (see
    # qasr_oracle.LITERAL_NOISE): scales are bit-equal in literal mode, equal to ~1e-6 otherwise.
    exact_scales = literal or case["mode"] != "dynamic" or case["percentile"] is None
    spec = copy.deepcopy(case["spec"])
    trace = []
    out = O.model_forward(spec, case["x"], case["lens"], trace)
    g = case["out"]
    assert np.array_equal(out["lens"], g["enc_lens"])
    assert np.array_equal(out["logit_codes"], g["logit_codes"])
    assert np.array_equal(out["logit_scale"], g["logit_scale"])
    assert np.array_equal(np.round(out["encoded"] / out["encoded_scale"]),
                          np.round(g["encoded"] / g["encoded_scale"].reshape(1, -1, 1)))
    if literal:
        assert np.array_equal(out["logits"], g["logits"])  # bit-identical dequantised logits
        assert np.array_equal(out["encoded"], g["encoded"])
    tol = 1e-5 * np.abs(g["logits"]).max()
    assert np.abs(out["logits"] - g["logits"]).max() <= tol
    assert np.abs(out["log_probs"] - g["log_probs"]).max() <= 1e-5 * np.abs(g["log_probs"]).max()
    assert np.array_equal(out["greedy"], g["greedy"])
    # per-layer codes
    seen = {}
    for rec in trace:
        name = rec["name"]
        i = seen.get(name, 0)
        seen[name] = i + 1
        if name.endswith("res_act"):
            ref = case["trace"][name][i]
            assert np.array_equal(rec["codes"], ref["codes"]), name
            if exact_scales:
                assert np.float32(rec["scale"]) == ref["scale"].reshape(())
            else:
                assert abs(np.float32(rec["scale"]) / ref["scale"].reshape(()) - 1) < 1e-5
        else:
            act_name = "decoder.act" if name == "decoder" else name + ".act"
            conv_name = "decoder.decoder_layers.0" if name == "decoder" else name + ".conv"
            assert np.array_equal(rec["in_codes"], case["trace"][act_name][i]["codes"]), name
            assert np.array_equal(rec["acc"], case["trace"][conv_name][i]["acc"]), name
            if exact_scales:
                assert np.array_equal(rec["out_scale"], case["trace"][conv_name][i]["out_scale"]), name
    assert len(trace) > 5


def test_calibration_replay(literal):
    """Re-run the reference's calibration sequence in the oracle from zeroed ranges and
    land on the same (x_min, x_max) buffers (bit-equal with the literal float noise; the
    integer recipe is bit-equal on these cases too, and within 1e-5 relative in general)."""
    for path in NET_CASES:
        case = specio.load(path)
        if case["mode"] != "static":
            continue
        spec = copy.deepcopy(case["spec"])
        for a in O.iter_acts(spec):
            a.x_min = a.x_max = np.float32(0)
        O.set_mode(spec, True)
        for c in case["calib"]:
            O.model_forward(spec, c["x"], c["lens"])
        for a, b in zip(O.iter_acts(spec), O.iter_acts(case["spec"])):
            if literal:
                assert np.float32(a.x_min) == np.float32(b.x_min)
                assert np.float32(a.x_max) == np.float32(b.x_max)
            else:
                assert abs(float(a.x_min) - float(b.x_min)) <= 1e-5 * max(abs(float(b.x_min)), 1e-3)
                assert abs(float(a.x_max) - float(b.x_max)) <= 1e-5 * max(abs(float(b.x_max)), 1e-3)


def test_torch_quantile_restatement():
    import torch
    rng = np.random.default_rng(0)
    for n in (7, 1000, 12345, 200001):
        x = rng.standard_normal(n).astype(np.float32)
        for p in (99.0, 99.9, 99.996, 50.0):
            for q in (np.float32(p / 100), np.float32(1 - p / 100)):
                want = torch.quantile(torch.from_numpy(x), torch.tensor(float(q))).item()
                assert np.float32(want) == O.torch_quantile(x, q)


def test_ctc_collapse_known_answers():
    # the collapse rule of metrics/wer.py:117-136; blank = 28
    assert O.ctc_greedy_collapse([28, 2, 2, 28, 2, 0, 0, 19, 28], 28) == [2, 2, 0, 19]
    assert O.ctc_greedy_collapse([28, 28], 28) == []
    assert O.ctc_greedy_collapse([3, 3, 3], 28) == [3]


# ------------------------------------------------------------------------------------------
# round-2 fixtures (tests/golden/gen_golden_extra.py): QuantAct(per_channel=True), QuantLinear,
# QuantConv1d with a folded, un-fixed BatchNorm
# ------------------------------------------------------------------------------------------
EXTRA = specio.load(os.path.join(GOLDEN, "ops_extra.npz"))


@pytest.mark.parametrize("case", [c for c in EXTRA if c["kind"] == "act_per_channel"], ids=lambda c: c["tag"])
def test_quant_act_per_channel(case):
    act = O.ActStatePC(case["bit"], case["x_min0"], case["x_max0"], case["momentum"], case["dynamic"],
                       case["running_stat"])
    for call in case["calls"]:
        pre = None if call["pre_sf"] is None else call["pre_sf"].reshape(1, -1, 1)
        isf = None if call["identity_sf"] is None else call["identity_sf"].reshape(1, -1, 1)
        out, s, codes = O.act_forward_per_channel(act, call["x"], pre, call["identity"], isf)
        assert np.array_equal(s, call["scale"])
        assert np.array_equal(codes, call["codes"])
        assert np.array_equal(out, call["out"])
        assert np.array_equal(act.x_min, call["x_min"]) and np.array_equal(act.x_max, call["x_max"])


@pytest.mark.parametrize("case", [c for c in EXTRA if c["kind"] == "linear"], ids=lambda c: c["tag"])
def test_quant_linear(case):
    y, s_out, acc, w_int, b_int = O.linear_forward(case["weight"], case["bias"], case["weight_bit"], case["in_codes"],
                                                   case["s_x"].reshape(()))
    assert np.array_equal(w_int, case["weight_integer"])
    assert np.array_equal(b_int.astype(np.float64), case["bias_integer"])
    assert np.array_equal(s_out, case["out_scale"])
    assert np.abs(acc).max() < 2 ** 24          # the float32 F.linear of the reference is exact here
    assert np.array_equal(acc, case["acc"])
    assert np.allclose(y, case["out"], rtol=0, atol=1e-6 * np.abs(case["out"]).max())


def test_quant_linear_per_row_intent():
    """per_channel=True (per-row weight ranges): the reference cannot run this branch (see gen_golden_extra.py);
    the oracle's intent restatement must reduce to the global case when all rows share one range."""
    rng = np.random.default_rng(3)
    w = rng.normal(size=(12, 20)).astype(np.float32) * 0.1
    w[:, 0], w[:, 1] = 0.5, -0.5                 # every row has the same extrema
    codes = rng.integers(-128, 128, size=(9, 20))
    a = O.linear_forward(w, None, 8, codes, np.float32(0.02), per_channel=False)
    b = O.linear_forward(w, None, 8, codes, np.float32(0.02), per_channel=True)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0])
    w[3] *= 0.25                                  # a row with its own range gets its own scale
    c = O.linear_forward(w, None, 8, codes, np.float32(0.02), per_channel=True)
    assert c[1][3] == np.float32(np.float32(0.125 / 127) * np.float32(0.02)) and np.abs(c[3][3]).max() == 127


@pytest.mark.parametrize("case", [c for c in EXTRA if c["kind"] == "conv_unfixed_bn"], ids=lambda c: c["tag"])
def test_quant_conv_unfixed_bn(case):
    bn = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in case["bn"].items()}
    conv = dict(weight=case["weight"], bias=None, bn=None, stride=1, pad=case["pad"], dil=1, groups=case["groups"],
                weight_bit=8)
    for call in case["calls"]:
        out, sf = O.conv_forward_unfixed_bn(conv, bn, call["in_codes"], call["s_x"].reshape(()))
        # float statistics (mean / unbiased variance, sqrt): 1e-5 relative, the tolerance of the float outputs
        assert np.allclose(bn["mean"], call["mean"], rtol=1e-5, atol=1e-6)
        assert np.allclose(bn["var"], call["var"], rtol=1e-5, atol=1e-6)
        assert np.allclose(sf.reshape(-1), call["out_scale"], rtol=1e-5, atol=0)
        assert np.abs(out - call["out"]).max() <= 1e-5 * np.abs(call["out"]).max()

"""GPU: the drop-in module API and the fused int8 plan against (a) the golden vectors produced by
the unmodified reference and (b) the oracle on freshly generated QuartzNet15x5 parameters.
Integer codes bit-exact; dequantised logits within 1e-5 * max|logit| (pure-integer path, see
SURVEY.md 3.5); log-probs within 1e-5 * max|log-prob|."""
import copy
import glob
import os

import numpy as np
import pytest
import torch

import qasr_oracle as O
import specio
from spec_from_modules import spec_from_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NET_CASES = sorted(glob.glob(os.path.join(GOLDEN, "net_*.npz")))


def _imports():
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    import nemo.quantization.utils.quantize_model as qm
    from nemo.quantization.utils import quant_modules as Q
    from qasr import topology
    return EncDecCTCModel, qm, Q, topology


def build_from_case(case):
    EncDecCTCModel, qm, Q, topology = _imports()
    model = EncDecCTCModel(topology.TOPOLOGIES[case["topology"]](**case["topology_kwargs"]))
    model.eval()
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in case["state_dict"].items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("conv.conv.weight") or "num_batches" in k for k in missing), missing
    model.set_quant_bit(case["weight_bit"], "weight")
    model.set_quant_bit(case["act_bit"], "act")
    if case["percentile"] is not None:
        qm.set_percentile(model, case["percentile"])
    if case["momentum"] is not None:
        for m in model.modules():
            if type(m) is Q.QuantAct:
                m.act_range_momentum = case["momentum"]
    model.encoder.bn_folding()
    return model.to(DEV), qm


def module_acts(model):
    from nemo.collections.asr.parts.jasper import MaskedConv1d
    for b in model.encoder.encoder_layers:
        for m in b.mconv:
            if isinstance(m, MaskedConv1d):
                yield m.act
        if b.res is not None:
            for layer in b.res:
                yield layer[0].act
        yield b.res_act
    yield model.decoder.act


def set_ranges_from_spec(model, spec):
    for a, m in zip(O.iter_acts(spec), module_acts(model)):
        m.x_min = torch.full_like(m.x_min, float(a.x_min))
        m.x_max = torch.full_like(m.x_max, float(a.x_max))


def logit_codes(logits, scale):
    return torch.round(logits.double() / scale.double().view(1, -1, 1)).to(torch.int64).cpu().numpy()


def check_outputs(case_out, logits, logit_scale, log_probs, greedy, enc_lens):
    assert np.array_equal(enc_lens.cpu().numpy(), case_out["enc_lens"])
    assert np.array_equal(logit_scale.reshape(-1).cpu().numpy(), case_out["logit_scale"])
    assert np.array_equal(logit_codes(logits, logit_scale.reshape(-1)), case_out["logit_codes"])
    lg = logits.cpu().numpy()
    assert np.abs(lg - case_out["logits"]).max() <= 1e-5 * np.abs(case_out["logits"]).max()
    lp = log_probs.cpu().numpy()
    assert np.abs(lp - case_out["log_probs"]).max() <= 1e-5 * np.abs(case_out["log_probs"]).max()
    assert np.array_equal(greedy.cpu().numpy(), case_out["greedy"])


def needs_dense_kernel(case):
    return "jasper" in case["topology"]


@pytest.mark.parametrize("path", NET_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_golden_network_module_level(path):
    case = specio.load(path)
    model, qm = build_from_case(case)
    model.use_fused_plan = model.encoder.use_fused_plan = False
    x = torch.from_numpy(case["x"]).to(DEV)
    lens = torch.from_numpy(case["lens"]).to(DEV)
    if case["mode"] == "static":
        qm.calibrate(model)
        for c in case["calib"]:
            e, l, s = model.encoder(torch.from_numpy(c["x"]).to(DEV), torch.from_numpy(c["lens"]).to(DEV))
            model.decoder(e, s)
        # the module-level path reproduces the reference's float noise (QuantConv1d.exact_float_noise): the
        # ranges calibrated on the GPU are bit-equal to the reference's, percentile mode included
        for a, m in zip(O.iter_acts(case["spec"]), module_acts(model)):
            assert np.float32(m.x_min.item()) == np.float32(a.x_min)
            assert np.float32(m.x_max.item()) == np.float32(a.x_max)
    qm.evaluate(model)
    qm.set_dynamic(model, case["mode"] == "dynamic")
    log_probs, enc_len, greedy = model(processed_signal=x, processed_signal_length=lens)
    sf = model.decoder.decoder_layers[0].conv_scaling_factor * model.decoder.act.act_scaling_factor
    check_outputs(case["out"], model.decoder.last_logits, sf, log_probs, greedy, enc_len)
    # bit-identical dequantised logits and encoder output floats
    assert np.array_equal(model.decoder.last_logits.cpu().numpy(), case["out"]["logits"])
    enc, _, enc_sf = model.encoder(audio_signal=x, length=lens)
    assert np.array_equal(enc.cpu().numpy(), case["out"]["encoded"])


@pytest.mark.parametrize("path", [p for p in NET_CASES if "static" in p or "mom" in p],
                         ids=lambda p: os.path.basename(p)[:-4])
def test_golden_network_fused_plan(path):
    case = specio.load(path)
    model, qm = build_from_case(case)
    set_ranges_from_spec(model, case["spec"])
    qm.evaluate(model)
    qm.set_dynamic(model, False)
    x = torch.from_numpy(case["x"]).to(DEV)
    lens = torch.from_numpy(case["lens"]).to(DEV)
    for use_graph in (False, True):
        model._plan_cache = {}
        from qasr import plan as P
        plan = P.FusedPlan(model.encoder, model.decoder, x.shape, x.device, use_graph=use_graph)
        logits = plan.run(x, lens)
        if use_graph:  # replay twice: static buffers, no stale state
            logits = plan.run(x, lens)
        check_outputs(case["out"], logits, plan.logits_scale.reshape(-1), plan.log_probs(), plan.greedy(),
                      plan.out_lens())
        # every intermediate int8 tensor against the reference trace
        checked = 0
        names = {id(m): n for n, m in model.named_modules()}
        for (vid, key), qt in plan._q.items():
            if qt.info is None:
                continue
            n = names[id(qt.info.module)].replace("encoder.encoder.", "encoder.")
            ref = case["trace"][n][0]["codes"]
            assert np.array_equal(qt.codes_bct()[:, :, :ref.shape[-1]], ref), n   # the plan is compiled for the frame bucket
            checked += 1
        assert checked > 8
    # public API path (EncDecCTCModel.forward picks the plan up by itself)
    log_probs, enc_len, greedy = model(processed_signal=x, processed_signal_length=lens)
    assert np.array_equal(greedy.cpu().numpy(), case["out"]["greedy"])
    enc, enc_len, enc_sf = model.encoder(audio_signal=x, length=lens)
    want = np.round(case["out"]["encoded"] / case["out"]["encoded_scale"].reshape(1, -1, 1))
    assert np.array_equal(torch.round(enc / enc_sf).cpu().numpy(), want)


@pytest.mark.parametrize("case", specio.load(os.path.join(GOLDEN, "ops_act.npz")), ids=lambda c: c["tag"])
def test_golden_quant_act(case):
    _, _, Q, _ = _imports()
    act = Q.QuantAct(case["bit"], quant_mode="symmetric", act_range_momentum=case["momentum"],
                     dynamic=case["dynamic"], percentile=case["percentile"]).to(DEV)
    act.running_stat = case["running_stat"]
    act.x_min = torch.tensor([case["x_min0"]], device=DEV)
    act.x_max = torch.tensor([case["x_max0"]], device=DEV)
    for call in case["calls"]:
        d = lambda a: None if a is None else torch.from_numpy(a).to(DEV)  # noqa: E731
        pre = None if call["pre_sf"] is None else d(call["pre_sf"]).view(1, -1, 1)
        isf = None if call["identity_sf"] is None else d(call["identity_sf"]).view(1, -1, 1)
        out, sf = act(d(call["x"]), pre, d(call["identity"]), isf)
        assert np.float32(sf.reshape(-1)[0].item()) == call["scale"].reshape(())
        assert np.array_equal(out.cpu().numpy(), call["out"])
        assert np.float32(act.x_min.item()) == np.float32(call["x_min"])
        assert np.float32(act.x_max.item()) == np.float32(call["x_max"])


@pytest.mark.parametrize("case", specio.load(os.path.join(GOLDEN, "ops_conv.npz")), ids=lambda c: c["tag"])
def test_golden_quant_conv(case):
    _, _, Q, _ = _imports()
    w = torch.from_numpy(case["weight"])
    co, cig, k = w.shape
    conv = torch.nn.Conv1d(cig * case["groups"], co, k, stride=case["stride"], padding=case["pad"],
                           dilation=case["dil"], groups=case["groups"], bias=case["bias"] is not None)
    conv.weight.data.copy_(w)
    if case["bias"] is not None:
        conv.bias.data.copy_(torch.from_numpy(case["bias"]))
    qc = Q.QuantConv1d(case["weight_bit"], bias_bit=32, quant_mode="symmetric", per_channel=True)
    qc.set_param(conv)
    if case["bn"] is not None:
        bn = torch.nn.BatchNorm1d(co, eps=case["bn"]["eps"])
        bn.weight.data.copy_(torch.from_numpy(case["bn"]["weight"]))
        bn.bias.data.copy_(torch.from_numpy(case["bn"]["bias"]))
        bn.running_mean.copy_(torch.from_numpy(case["bn"]["mean"]))
        bn.running_var.copy_(torch.from_numpy(case["bn"]["var"]))
        bn.eval()
        qc.bn_folding(bn)
    qc = qc.to(DEV)
    s_x = torch.from_numpy(case["s_x"]).to(DEV)
    x = torch.from_numpy(case["in_codes"]).float().to(DEV) * s_x
    out, sf = qc(x, s_x)
    assert np.array_equal(qc.weight_integer.cpu().numpy().astype(np.int32), case["weight_integer"])
    assert np.array_equal(qc.conv_scaling_factor.cpu().numpy(), case["conv_scaling_factor"])
    assert np.array_equal(sf.reshape(-1).cpu().numpy(), case["out_scale"])
    got_acc = torch.round(out.double() / sf.double()).cpu().numpy()
    assert np.array_equal(got_acc.astype(np.float32), case["acc"].astype(np.float32))
    assert np.array_equal(out.cpu().numpy(), case["out"])          # bit-identical floats (delta reproduced)
    qc.exact_float_noise = False                                   # the noise-free integer recipe
    out2, _ = qc(x, s_x)
    assert np.abs(out2.cpu().numpy() - case["out"]).max() <= 1e-6 * np.abs(case["out"]).max()


# ------------------------------------------------------------------------------------------
# full-size QuartzNet15x5 against the oracle (parameters generated here, identical inputs)
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def quartznet():
    EncDecCTCModel, qm, Q, topology = _imports()
    from qasr.params import reestimate_bn_, synthetic_features
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology("quartznet15x5")
    model.eval()
    reestimate_bn_(model.encoder)
    model.encoder.bn_folding()
    return model.to(DEV), qm, synthetic_features


@pytest.mark.parametrize("wbit,abit,percentile", [(8, 8, None), (6, 6, 99.996)], ids=["w8a8", "w6a6_pct"])
def test_quartznet15x5_static_vs_oracle(quartznet, wbit, abit, percentile):
    model, qm, synthetic_features = quartznet
    for m in module_acts(model):
        m.x_min, m.x_max = torch.zeros_like(m.x_min), torch.zeros_like(m.x_max)
    model.set_quant_bit(wbit, "weight")
    model.set_quant_bit(abit, "act")
    qm.set_percentile(model, percentile)
    spec = spec_from_model(model.encoder, model.decoder)
    B, T = 2, 200
    model.use_fused_plan = model.encoder.use_fused_plan = False
    O.LITERAL_NOISE = True     # the module-level path reproduces the reference's float noise
    qm.calibrate(model)
    O.set_mode(spec, True)
    for i in range(2):
        cx, cl = synthetic_features(B, T, seed=100 + i, ragged=(i % 2 == 1))
        e, l, s = model.encoder(cx.to(DEV), cl.to(DEV))
        model.decoder(e, s)
        O.model_forward(spec, cx.numpy(), cl.numpy())
    qm.evaluate(model)
    O.set_mode(spec, False)
    # calibration statistics (min/max and percentile): bit-equal ranges
    for a, m in zip(O.iter_acts(spec), module_acts(model)):
        assert np.float32(m.x_min.item()) == np.float32(a.x_min)
        assert np.float32(m.x_max.item()) == np.float32(a.x_max)
    x, lens = synthetic_features(B, T, seed=5, ragged=True)
    trace = []
    want = O.model_forward(spec, x.numpy(), lens.numpy(), trace)
    O.LITERAL_NOISE = False
    assert (want["logit_codes"] != 0).mean() > 0.9
    # module-level path
    log_probs, enc_len, greedy = model(processed_signal=x.to(DEV), processed_signal_length=lens.to(DEV))
    sf = model.decoder.decoder_layers[0].conv_scaling_factor * model.decoder.act.act_scaling_factor
    assert np.array_equal(logit_codes(model.decoder.last_logits, sf), want["logit_codes"])
    assert np.array_equal(model.decoder.last_logits.cpu().numpy(), want["logits"])   # bit-identical floats
    assert np.array_equal(greedy.cpu().numpy(), want["greedy"])
    # fused plan
    model.use_fused_plan = model.encoder.use_fused_plan = True
    plan = model.plan_for(x.shape, torch.device(DEV))
    logits = plan.run(x.to(DEV), lens.to(DEV))
    assert np.array_equal(logit_codes(logits, plan.logits_scale.reshape(-1)), want["logit_codes"])
    assert np.abs(logits.cpu().numpy() - want["logits"]).max() <= 1e-5 * np.abs(want["logits"]).max()
    assert np.array_equal(plan.greedy().cpu().numpy(), want["greedy"])
    assert np.array_equal(plan.out_lens().cpu().numpy(), want["lens"])
    by_name = {r["name"]: r for r in trace}
    names = {id(m): n for n, m in model.named_modules()}
    checked = 0
    for (vid, key), qt in plan._q.items():
        if qt.info is None:
            continue
        n = names[id(qt.info.module)].replace("encoder.encoder.", "encoder.")
        if n.endswith(".act") and n[:-4] in by_name:
            ref = by_name[n[:-4]]["in_codes"]
        elif n == "decoder.act":
            ref = by_name["decoder"]["in_codes"]
        else:
            continue
        assert np.array_equal(qt.codes_bct()[:, :, :ref.shape[-1]], ref), n   # the plan is compiled for the frame bucket
        checked += 1
    assert checked >= 150
    print("plan stats", plan.stats)


def test_quartznet15x5_dynamic_vs_oracle(quartznet):
    """cfg4's mode: W4A8, dynamic min/max ranges (module-level path; statistics kernels)."""
    model, qm, synthetic_features = quartznet
    model.set_quant_bit(4, "weight")
    model.set_quant_bit(8, "act")
    qm.set_percentile(model, None)
    qm.evaluate(model)
    qm.set_dynamic(model, True)
    spec = spec_from_model(model.encoder, model.decoder)
    x, lens = synthetic_features(2, 200, seed=7, ragged=True)
    O.LITERAL_NOISE = True     # dynamic ranges are taken on floats that carry the reference's noise
    try:
        want = O.model_forward(spec, x.numpy(), lens.numpy())
    finally:
        O.LITERAL_NOISE = False
    for fused in (False, True):     # module-level path (statistics kernels per QuantAct), then the fused dynamic plan
        model.use_fused_plan = fused
        log_probs, enc_len, greedy = model(processed_signal=x.to(DEV), processed_signal_length=lens.to(DEV))
        # forward() leaves the decoder QuantAct's scale of THIS batch in the module, like the reference does
        sf = model.decoder.decoder_layers[0].conv_scaling_factor * model.decoder.act.act_scaling_factor
        assert np.array_equal(logit_codes(model.decoder.last_logits, sf), want["logit_codes"]), fused
        assert np.array_equal(greedy.cpu().numpy(), want["greedy"])
    assert model._dynamic_plan_or_none(x.to(DEV)) is not None
    qm.set_dynamic(model, False)


def test_jasper10x5dr_static_vs_oracle():
    """BASELINE config 3's network: dense k = 11..29 convs (channels-last implicit GEMM), strided head,
    dense residual chains of up to 10 res_act applications, accumulators beyond 2^24."""
    EncDecCTCModel, qm, Q, topology = _imports()
    from qasr.params import reestimate_bn_, synthetic_features
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology("jasper10x5dr")
    model.eval()
    reestimate_bn_(model.encoder)
    model.encoder.bn_folding()
    model = model.to(DEV)
    spec = spec_from_model(model.encoder, model.decoder)
    B, T = 1, 120
    model.use_fused_plan = model.encoder.use_fused_plan = False
    O.LITERAL_NOISE = True
    qm.calibrate(model)
    O.set_mode(spec, True)
    cx, cl = synthetic_features(B, T, seed=100)
    e, l, s = model.encoder(cx.to(DEV), cl.to(DEV))
    model.decoder(e, s)
    O.model_forward(spec, cx.numpy(), cl.numpy())
    qm.evaluate(model)
    O.set_mode(spec, False)
    for a, m in zip(O.iter_acts(spec), module_acts(model)):
        assert np.float32(m.x_min.item()) == np.float32(a.x_min)
        assert np.float32(m.x_max.item()) == np.float32(a.x_max)
    x, lens = synthetic_features(B, T, seed=5)
    trace = []
    O.LITERAL_NOISE = False    # the fused plan is the noise-free integer recipe
    want = O.model_forward(spec, x.numpy(), lens.numpy(), trace)
    model.use_fused_plan = model.encoder.use_fused_plan = True
    plan = model.plan_for(x.shape, torch.device(DEV))
    assert plan.ntc
    logits = plan.run(x.to(DEV), lens.to(DEV))
    assert np.array_equal(logit_codes(logits, plan.logits_scale.reshape(-1)), want["logit_codes"])
    assert np.array_equal(plan.greedy().cpu().numpy(), want["greedy"])
    by_name = {}
    for r in trace:
        by_name.setdefault(r["name"], []).append(r)
    names = {id(m): n for n, m in model.named_modules()}
    checked = 0
    for (vid, key), qt in plan._q.items():
        if qt.info is None:
            continue
        n = names[id(qt.info.module)].replace("encoder.encoder.", "encoder.")
        if n.endswith(".act") and n[:-4] in by_name:
            ref = by_name[n[:-4]][0]["in_codes"]
            assert np.array_equal(qt.codes_bct()[:, :, :ref.shape[-1]], ref), n
            checked += 1
    assert checked >= 50
    print("jasper plan stats", plan.stats)


@pytest.mark.gpu
def test_transcribe_stream_matches_forward():
    """The pipelined serving loop (H2D prefetch on a copy stream, asynchronous label read-back) returns exactly
    forward()'s greedy labels, batch by batch, in order."""
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    from qasr.params import reestimate_bn_, synthetic_features
    import nemo.quantization.utils.quantize_model as qm
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    model = EncDecCTCModel.from_topology("tiny_quartznet")
    model.eval()
    reestimate_bn_(model.encoder)
    model.encoder.bn_folding()
    model = model.to(dev)
    qm.calibrate(model)
    for i in range(2):
        cx, cl = synthetic_features(4, 160, seed=100 + i)
        e, l, sf = model.encoder(cx.to(dev), cl.to(dev))
        model.decoder(e, sf)
    qm.evaluate(model)
    qm.set_dynamic(model, False)
    batches = [synthetic_features(4, 160, seed=10 + i, ragged=True) for i in range(5)]
    want = [model(processed_signal=x.to(dev), processed_signal_length=l.to(dev))[2].cpu() for x, l in batches]
    got = list(model.transcribe_stream((x.pin_memory(), l.to(torch.int32).pin_memory()) for x, l in batches))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.dtype == torch.long and torch.equal(g, w)


@pytest.mark.gpu
def test_full_size_properties_bench_workload():
    """BASELINE.json configs[1] at its full size (QuartzNet15x5 W6A6 percentile 99.996, batch 64 x 1600 frames), where
    the CPU oracle would take minutes: size-independent properties that tie the full-size run to the small-size
    runs the oracle checks bit-exactly (same calibrated model, same kernels):
      * batch independence: every utterance's logits in the batch-64 run are bit-identical to its logits in a
        batch-4 run of the same utterances (static evaluation has no cross-utterance term);
      * padding invariance: frames at t >= length do not influence the logits inside the length (MaskedConv1d);
      * determinism: two runs give identical bytes;
      * the serving loop returns forward()'s labels."""
    import bench
    from qasr.params import synthetic_features
    wl = dict(bench.WORKLOAD)
    model = bench.build_model(torch.device(DEV), wl)
    B, T = wl["batch"], wl["frames"]
    x, lens = synthetic_features(B, T, seed=77, ragged=True)
    xd, ld = x.to(DEV), lens.to(DEV)
    plan = model.plan_for(xd.shape, torch.device(DEV))
    assert plan.stats["exact"] == 0 and plan.stats["f32_channels"] > 0.9 * plan.stats["channels"]
    logits = plan.run(xd, ld).clone()
    greedy = plan.greedy().clone()
    out_lens = plan.out_lens().clone()
    torch.cuda.synchronize()
    assert torch.equal(plan.run(xd, ld), logits)                                  # determinism
    # padding invariance: garbage beyond each utterance's length
    x2 = x.clone()
    for b in range(B):
        x2[b, :, int(lens[b]):] = 1000.0 * torch.randn(x.shape[1], T - int(lens[b]))
    logits2 = plan.run(x2.to(DEV), ld)
    for b in range(0, B, 7):
        n = int(out_lens[b])
        assert torch.equal(logits2[b, :, :n], logits[b, :, :n])
    # batch independence against batch-4 plans of the same model (a different plan / graph / grid)
    for b0 in (0, 28, 60):
        small = model.plan_for((4,) + tuple(xd.shape[1:]), torch.device(DEV))
        ls = small.run(xd[b0:b0 + 4].contiguous(), ld[b0:b0 + 4].contiguous())
        for i in range(4):
            n = int(out_lens[b0 + i])
            assert torch.equal(ls[i, :, :n], logits[b0 + i, :, :n]), "utterance %d differs between batch 64 and batch 4" % (b0 + i)
    # serving loop == forward
    got = list(model.transcribe_stream([(x.pin_memory(), lens.to(torch.int32).pin_memory())] * 2))
    assert len(got) == 2 and torch.equal(got[0], greedy.cpu().to(torch.long)) and torch.equal(got[1], got[0])

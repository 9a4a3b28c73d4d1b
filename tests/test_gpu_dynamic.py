"""GPU: dynamic quantisation on the device (qasr.dynplan, csrc/dynamic.cu) -- BASELINE config 4's mode.

Kernel level: statistics pass -> qasr_dyn_refine -> qasr_dyn_finalize -> requantising pass of one
conv (and of the two-conv res_act tail) against a numpy restatement of the reference's literal arithmetic
(quant_modules.py:149-156 ranges on floats that carry the un-rounded x / s noise of :301).
Network level: the fused dynamic plan against the golden vectors of the unmodified reference and against the oracle
on QuartzNet15x5, logit codes bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import qasr_oracle as O
import specio
from qasr import _lib as L
from qasr import ops
from spec_from_modules import spec_from_model
from test_gpu_kernels import expected_requant, from_codes_tensor, norm_me, rng, to_codes_tensor

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F32, F64 = np.float32, np.float64


def literal_conv(q, signed, s_x, w, b_int, s_out, stride, pad, dil, groups):
    """y = f32(conv_f64(x / s_x) + b_int) * s_out with the reference's un-rounded x_int (quant_modules.py:301-309)."""
    x_in = ((q.astype(F32) * F32(s_x)).astype(F32) / F32(s_x)).astype(F32).astype(F64)
    acc_f = O.conv1d_f64(x_in, w, b_int, stride, pad, dil, groups)
    return (acc_f.astype(F32) * s_out.reshape(1, -1, 1)).astype(F32)


def bias_int(bias, s_out):
    return O.sym_quantize(bias, 32, s_out).astype(F64).astype(np.int64)


class DevConv:
    """Device-side state of one conv as qasr.dynplan keeps it."""

    def __init__(self, w, groups, stride, dil, pad, s_w, bias):
        self.w = w
        co, cig, k = w.shape
        self.depthwise = groups > 1
        self.C_out, self.C_in, self.k, self.stride, self.dil, self.pad = co, cig * groups, k, stride, dil, pad
        w8 = torch.from_numpy(w.astype(np.int8)).to(DEV)
        self.w_raw = w8.contiguous()
        self.wabs = torch.from_numpy(np.abs(w).reshape(co, -1).sum(1).astype(np.int32)).to(DEV)
        self.s_w = torch.from_numpy(s_w.astype(F32)).to(DEV)
        self.bias = None if bias is None else torch.from_numpy(bias.astype(F32)).to(DEV)
        self.packed = None if self.depthwise else ops.pack_weights(w8)
        self.w8 = ops.pack_dw_weights(w8.reshape(co, k), stride, dil, pad) if self.depthwise else None
        self.table = torch.zeros((co, 4), dtype=torch.int32, device=DEV)
        self.stats = torch.zeros((co, 2), dtype=torch.int64, device=DEV)
        self.cand = torch.full((8,), -1, dtype=torch.int32, device=DEV)

    def operand(self, x, T_in, unsigned, s_x):
        op = L.DynOperand()
        op.x, op.w, op.s_x, op.ch, op.wabs = x.data_ptr(), self.w_raw.data_ptr(), s_x.data_ptr(), self.table.data_ptr(), \
            self.wabs.data_ptr()
        op.x_ld, op.C_in, op.T_in, op.k, op.stride, op.dilation, op.pad = x.shape[-1], self.C_in, T_in, self.k, self.stride, \
            self.dil, self.pad
        op.depthwise, op.in_unsigned = int(self.depthwise), int(unsigned)
        return op

    def launch(self, x, T_in, out, T_out, ep, other=None, x1=None, T1=None):
        B = x.shape[0]
        if self.depthwise:
            L.check(L.lib().qasr_dwconv1d_i8(L.ptr(x), x.shape[-1], L.ptr(self.w8), L.ptr(out), ops.round_up(T_out, 16), B,
                                             self.C_out, T_in, T_out, self.k, self.stride, self.dil, self.pad, C.byref(ep),
                                             L.stream_ptr()))
        else:
            g0 = ops.gemm_operand(x, T_in, self.packed, self.C_in, self.k, self.stride, self.dil, self.pad)
            g1 = None if other is None else ops.gemm_operand(x1, T1, other.packed, other.C_in, other.k, other.stride,
                                                             other.dil, other.pad)
            L.check(L.lib().qasr_conv1d_gemm_i8(C.byref(g0), C.byref(g1) if g1 is not None else None, L.ptr(out),
                                                ops.round_up(T_out, 16), B, self.C_out, T_out, C.byref(ep), L.stream_ptr()))


def finalize(kind, src, bit, act_scale, prods=(), cons=(), src_scale=None, prev=None, qabs=(), stats=None):
    a = L.DynFinalizeArgs()
    a.src_kind, a.src, a.n_levels, a.act_scale = kind, src.data_ptr(), 2 ** (bit - 1) - 1, act_scale.data_ptr()
    a.src_scale = None if src_scale is None else src_scale.data_ptr()
    a.n_prod = len(prods)
    for i, c in enumerate(prods):
        a.prod[i], a.prod_C[i] = c.table.data_ptr(), c.C_out
        a.prod_wabs[i], a.prod_qabs[i] = c.wabs.data_ptr(), qabs[i]
    if stats is not None:
        a.stats, a.stats_C, a.stats_dual = stats[0].stats.data_ptr(), stats[0].C_out, int(stats[1])
    if prev is not None:
        a.prev_scale, a.prev_me = prev[0].data_ptr(), prev[1].data_ptr()
    a.n_cons = len(cons)
    for i, c in enumerate(cons):
        a.cons[i], a.cons_sw[i], a.cons_C[i] = c.table.data_ptr(), c.s_w.data_ptr(), c.C_out
        a.cons_bias[i] = None if c.bias is None else c.bias.data_ptr()
    L.check(L.lib().qasr_dyn_finalize(C.byref(a), L.stream_ptr()))


DYN_CASES = [
    # name, B, C_in, C_out, T, k, stride, dil, depthwise, in_signed, relu, masked, bit, wmax
    ("dw_k33_signed_out", 3, 48, 48, 300, 33, 1, 1, True, False, False, True, 8, 127),
    ("dw_k33_s2_first", 2, 64, 64, 401, 33, 2, 1, True, True, False, True, 8, 127),
    ("dw_k87_dil2", 2, 32, 32, 260, 87, 1, 2, True, False, False, True, 8, 7),
    ("pw_256_relu_masked", 2, 256, 256, 300, 1, 1, 1, False, True, True, True, 9, 7),
    ("pw_512_1024_relu_unmasked", 1, 512, 1024, 200, 1, 1, 1, False, True, True, False, 8, 127),
    ("pw_64_256_relu_w6", 3, 64, 256, 130, 1, 1, 1, False, True, True, True, 7, 31),
]


@pytest.mark.parametrize("case", DYN_CASES, ids=lambda c: c[0])
def test_dynamic_range_of_one_conv(case):
    name, B, ci, co, T, k, stride, dil, dw, in_signed, relu, masked, bit, wmax = case
    r = rng(len(name))
    groups = ci if dw else 1
    pad = (dil * k) // 2 - 1 if dil > 1 else k // 2
    q = r.integers(-128, 128, (B, ci, T)) if in_signed else r.integers(0, 256, (B, ci, T))
    w = r.integers(-wmax, wmax, (co, ci // groups, k))
    s_w = r.uniform(1e-3, 8e-3, co).astype(F32)
    bias = (r.normal(size=co) * 0.4).astype(F32) if not dw else None
    in_scale_src = r.uniform(0.5, 6.0)            # absmax of the conv's input act -> s_x through finalize
    T_out = (T + 2 * pad - dil * (k - 1) - 1) // stride + 1
    lens = r.integers(T_out // 2, T_out + 1, B).astype(np.int32)
    lens[0] = T_out
    # ---- expected (numpy, literal reference arithmetic)
    in_bit = 8 if in_signed else 9
    s_x = O.sym_scale(in_bit, -in_scale_src, in_scale_src).reshape(())
    s_out = (s_w * s_x).astype(F32)
    b_int = None if bias is None else bias_int(bias, s_out)
    y = literal_conv(q, in_signed, s_x, w, b_int, s_out, stride, pad, dil, groups)
    if relu:
        y = np.maximum(y, F32(0))
    valid = np.arange(T_out)[None, None, :] < (lens[:, None, None] if masked else T_out)
    absmax = F32(np.abs(np.where(valid, y, F32(0))).max())
    s_z = O.sym_scale(bit, -absmax, absmax).reshape(())
    m, e = norm_me(*O.dyadic(s_out, s_z))
    acc = O.conv1d_int(q, w, b_int, stride, pad, dil, groups)
    n = 2 ** (bit - 1) - 1
    want = expected_requant(acc, m, e, -n - 1, n, relu0=relu, lens=lens if masked else None)
    # ---- device
    conv = DevConv(w, groups, stride, dil, pad, s_w, bias)
    x = to_codes_tensor(q, in_signed)
    src = torch.tensor([-in_scale_src, in_scale_src], dtype=torch.float32, device=DEV)
    sx_dev, sz_dev = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    finalize(1, src, in_bit, sx_dev, cons=[conv])                       # fills s_out / b_int of the conv's table
    tab = conv.table.cpu().numpy()
    assert np.array_equal(tab[:, 0].view(F32), s_out)
    if b_int is not None:
        assert np.array_equal(tab[:, 1], b_int)
    lens_d = torch.from_numpy(lens).to(DEV)
    conv.stats[:, 0], conv.stats[:, 1] = 2 ** 63 - 1, -2 ** 63
    flags = L.F_DEVICE_TABLES | (L.F_MASK if masked else 0) | (L.F_B_SIGNED if in_signed else 0)
    ep = ops.make_epilogue(L.EPI_STATS, conv.table, flags=flags, out_lens=lens_d if masked else None)
    conv.launch(x, T, conv.stats, T_out, ep)
    st = conv.stats.cpu().numpy()
    accv = np.where(valid, acc, 0)
    cnt = np.broadcast_to(valid, acc.shape)
    want_lo = np.array([acc[:, c][cnt[:, c]].min() for c in range(co)])
    want_hi = np.array([acc[:, c][cnt[:, c]].max() for c in range(co)])
    assert np.array_equal(st[:, 0], want_lo) and np.array_equal(st[:, 1], want_hi)
    del accv
    op = conv.operand(x, T, not in_signed, sx_dev)
    absmax_dev = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.check(L.lib().qasr_dyn_refine(L.ptr(conv.stats), C.byref(op), None, int(relu), L.ptr(lens_d) if masked else None, B, co,
                                    T_out, L.ptr(conv.cand), L.ptr(absmax_dev), L.stream_ptr()))
    cand = conv.cand.cpu().numpy()
    assert 1 <= cand[4] <= 4, cand
    per_ch = np.abs(np.where(valid, y, F32(0))).max(axis=(0, 2))
    assert int(np.argmax(per_ch)) in cand[:4].tolist()
    assert absmax_dev.cpu().numpy().view(F32)[0] == absmax
    finalize(0, absmax_dev, bit, sz_dev, prods=[conv], qabs=[128 if in_signed else 255], stats=(conv, False))
    st = conv.stats.cpu().numpy()
    assert (st[:, 0] == 2 ** 63 - 1).all() and (st[:, 1] == -2 ** 63).all()      # reset for the next forward
    assert sz_dev.cpu().numpy()[0] == s_z
    assert int(absmax_dev.item()) == 0
    tab = conv.table.cpu().numpy()
    from qasr import plan as P
    bound = np.abs(w).reshape(co, -1).sum(1) * (128 if in_signed else 255) + (0 if b_int is None else np.abs(b_int))
    lit = np.array([not P.fast_ok(int(bound[c]), m[c:c + 1], e[c:c + 1]) for c in range(co)])
    assert np.array_equal(tab[:, 2], m) and np.array_equal(tab[:, 3] & 0x7fffffff, e)
    assert np.array_equal(tab[:, 3] < 0, lit), "device-side fast-path proof differs from qasr.plan.fast_ok"
    out = ops.alloc_codes(B, co, T_out, DEV)
    flags = (L.F_DEVICE_TABLES | L.F_EXACT_F32 | L.F_EXACT_DYAD | L.F_BIAS64 | (L.F_RELU0 if relu else 0) |
             (L.F_MASK if masked else 0) | (L.F_B_SIGNED if in_signed else 0))
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, conv.table, flags=flags, qmin=-n - 1, qmax=n, out_lens=lens_d if masked else None)
    conv.launch(x, T, out, T_out, ep)
    assert np.array_equal(from_codes_tensor(out, T_out, not relu), want)


@pytest.mark.parametrize("shape", [(2, 256, 256, 300), (1, 256, 512, 130), (3, 512, 512, 800)], ids=lambda s: "x".join(map(str, s)))
def test_dynamic_range_of_res_act(shape):
    """The block tail: pointwise conv of the last repeat + residual 1x1 conv -> res_act (jasper.py:671-682), range over
    identity + x (quant_modules.py:108), then block ReLU -> next QuantAct through qasr_minmax_i8 / qasr_requant_i8_dev."""
    B, ci_res, co, T = shape
    r = rng(co + T)
    q0 = r.integers(-128, 128, (B, co, T))          # main: signed codes of the depthwise output
    q1 = r.integers(0, 256, (B, ci_res, T))         # residual: unsigned codes of the block input
    w0, w1 = r.integers(-7, 8, (co, co, 1)), r.integers(-7, 8, (co, ci_res, 1))
    sw0, sw1 = r.uniform(1e-3, 8e-3, co).astype(F32), r.uniform(1e-3, 8e-3, co).astype(F32)
    b0, b1 = (r.normal(size=co) * 0.4).astype(F32), (r.normal(size=co) * 0.4).astype(F32)
    sx0 = O.sym_scale(8, -3.0, 3.0).reshape(())
    sx1 = O.sym_scale(9, 0.0, 5.5).reshape(())
    so0, so1 = (sw0 * sx0).astype(F32), (sw1 * sx1).astype(F32)
    bi0, bi1 = bias_int(b0, so0), bias_int(b1, so1)
    y0 = literal_conv(q0, True, sx0, w0, bi0, so0, 1, 0, 1, 1)
    y1 = literal_conv(q1, False, sx1, w1, bi1, so1, 1, 0, 1, 1)
    absmax = F32(np.abs((y1 + y0).astype(F32)).max())
    s_ra = O.sym_scale(8, -absmax, absmax).reshape(())
    m0, e0 = norm_me(*O.dyadic(so0, s_ra))
    m1, e1 = norm_me(*O.dyadic(so1, s_ra))
    a0, a1 = O.conv1d_int(q0, w0, bi0, 1, 0, 1, 1), O.conv1d_int(q1, w1, bi1, 1, 0, 1, 1)
    want_ra = expected_requant(a0, m0, e0, -128, 127, acc1=a1, m1=m1, e1=e1)
    lens = r.integers(T // 2, T + 1, B).astype(np.int32)
    relu_masked = np.where(np.arange(T)[None, None, :] < lens[:, None, None], np.maximum(want_ra, 0), 0)
    nmax = F32(F32(relu_masked.max()) * s_ra)
    s_next = O.sym_scale(9, F32(0), nmax).reshape(())
    m2, e2 = norm_me(*O.dyadic(np.array([s_ra], F32), s_next))
    want_next = np.clip(O.dyadic_mul(relu_masked, m2, e2), -256, 255).astype(np.int64)
    # ---- device
    c0, c1 = DevConv(w0, 1, 1, 1, 0, sw0, b0), DevConv(w1, 1, 1, 1, 0, sw1, b1)
    x0, x1 = to_codes_tensor(q0, True), to_codes_tensor(q1, False)
    sx0_d, sx1_d, sra_d, snext_d = (torch.zeros(1, device=DEV) for _ in range(4))
    finalize(1, torch.tensor([-3.0, 3.0], device=DEV), 8, sx0_d, cons=[c0])
    finalize(1, torch.tensor([0.0, 5.5], device=DEV), 9, sx1_d, cons=[c1])
    c0.stats[:, 0], c0.stats[:, 1] = 0xffffffff, 0
    ep = ops.make_epilogue(L.EPI_STATS, c0.table, flags=L.F_DEVICE_TABLES | L.F_B_SIGNED, ch1=c1.table)
    c0.launch(x0, T, c0.stats, T, ep, other=c1, x1=x1, T1=T)
    o0, o1 = c0.operand(x0, T, False, sx0_d), c1.operand(x1, T, True, sx1_d)
    absmax_dev = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.check(L.lib().qasr_dyn_refine(L.ptr(c0.stats), C.byref(o0), C.byref(o1), 0, None, B, co, T, L.ptr(c0.cand),
                                    L.ptr(absmax_dev), L.stream_ptr()))
    cand = c0.cand.cpu().numpy()
    assert 1 <= cand[4] <= 4, cand
    assert int(np.argmax(np.abs((y1 + y0).astype(F32)).max(axis=(0, 2)))) in cand[:4].tolist()
    assert absmax_dev.cpu().numpy().view(F32)[0] == absmax
    finalize(0, absmax_dev, 8, sra_d, prods=[c0, c1], qabs=[128, 255], stats=(c0, True))
    assert sra_d.cpu().numpy()[0] == s_ra
    ra = ops.alloc_codes(B, co, T, DEV)
    flags = L.F_DEVICE_TABLES | L.F_EXACT_F32 | L.F_EXACT_DYAD | L.F_BIAS64 | L.F_B_SIGNED
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, c0.table, flags=flags, ch1=c1.table, qmin=-128, qmax=127)
    c0.launch(x0, T, ra, T, ep, other=c1, x1=x1, T1=T)
    assert np.array_equal(from_codes_tensor(ra, T, True), want_ra)
    lens_d = torch.from_numpy(lens).to(DEV)
    mm = torch.tensor([2 ** 31 - 1, -2 ** 31], dtype=torch.int32, device=DEV)
    me = torch.zeros(2, dtype=torch.int32, device=DEV)
    L.check(L.lib().qasr_minmax_i8(L.ptr(ra), ra.shape[-1], 1, B, co, T, L.ptr(lens_d), L.ptr(mm), L.stream_ptr()))
    assert mm.cpu().numpy().tolist() == [int(relu_masked.min()), int(relu_masked.max())]
    finalize(2, mm, 9, snext_d, src_scale=sra_d, prev=(sra_d, me))
    assert snext_d.cpu().numpy()[0] == s_next
    assert me.cpu().numpy().tolist() == [int(m2[0]), int(e2[0])]
    nxt = ops.alloc_codes(B, co, T, DEV)
    L.check(L.lib().qasr_requant_i8_dev(L.ptr(ra), ra.shape[-1], 1, L.ptr(me), -256, 255, L.ptr(nxt), nxt.shape[-1], B, co, T,
                                        L.ptr(lens_d), L.stream_ptr()))
    assert np.array_equal(from_codes_tensor(nxt, T, False), want_next)


# ------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------
def test_golden_dynamic_network_on_the_fused_plan():
    """net_tinyqn_w4a8_dynamic.npz was produced by the UNMODIFIED reference in dynamic mode."""
    from test_gpu_model import build_from_case, check_outputs
    case = specio.load(os.path.join(GOLDEN, "net_tinyqn_w4a8_dynamic.npz"))
    model, qm = build_from_case(case)
    qm.evaluate(model)
    qm.set_dynamic(model, True)
    x, lens = torch.from_numpy(case["x"]).to(DEV), torch.from_numpy(case["lens"]).to(DEV)
    plan = model._dynamic_plan_or_none(x)
    assert plan is not None, "the fused dynamic plan must cover QuartzNet-shaped networks"
    log_probs, enc_len, greedy = model(processed_signal=x, processed_signal_length=lens)
    assert not plan.overflowed()
    check_outputs(case["out"], plan.logits[:, :, :plan.cur_T_out], plan.logits_scale.reshape(-1), log_probs, greedy, enc_len)
    # a second replay of the same graph (statistics records were reset on the device) gives the same bytes
    lg = plan.logits.clone()
    model(processed_signal=x, processed_signal_length=lens)
    assert torch.equal(plan.logits, lg)
    print("dynamic plan:", plan.stats)


def test_dynamic_plan_refuses_what_it_does_not_cover():
    """Percentile ranges and Jasper's dense-residual chains stay on the module-level path (forward still works)."""
    from test_gpu_model import build_from_case, check_outputs
    case = specio.load(os.path.join(GOLDEN, "net_tinyjasper_w4a8_dynamic_pct.npz"))
    model, qm = build_from_case(case)
    qm.evaluate(model)
    qm.set_dynamic(model, True)
    x, lens = torch.from_numpy(case["x"]).to(DEV), torch.from_numpy(case["lens"]).to(DEV)
    assert model._dynamic_plan_or_none(x) is None
    log_probs, enc_len, greedy = model(processed_signal=x, processed_signal_length=lens)
    sf = model.decoder.decoder_layers[0].conv_scaling_factor * model.decoder.act.act_scaling_factor
    check_outputs(case["out"], model.decoder.last_logits, sf, log_probs, greedy, enc_len)


def test_quartznet15x5_dynamic_plan_vs_oracle():
    """BASELINE config 4's network and bit widths (W4 / A8, dynamic min / max), ragged batch, against the oracle with the
    reference's float noise: every logit code equal.  The module-level path must agree with the plan bit for bit."""
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    import nemo.quantization.utils.quantize_model as qm
    from qasr.params import reestimate_bn_, synthetic_features
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology("quartznet15x5")
    model.eval()
    reestimate_bn_(model.encoder)
    model.set_quant_bit(4, "weight")
    model.set_quant_bit(8, "act")
    model.encoder.bn_folding()
    model = model.to(DEV)
    qm.evaluate(model)
    qm.set_dynamic(model, True)
    x, lens = synthetic_features(3, 400, seed=11, ragged=True)
    spec = spec_from_model(model.encoder, model.decoder)
    O.LITERAL_NOISE = True
    try:
        want = O.model_forward(spec, x.numpy(), lens.numpy())
    finally:
        O.LITERAL_NOISE = False
    xd, ld = x.to(DEV), lens.to(DEV)
    plan = model._dynamic_plan_or_none(xd)
    assert plan is not None
    logits = plan.run(xd, ld)
    assert not plan.overflowed()
    codes = torch.round(logits.double() / plan.logits_scale.double()).to(torch.int64).cpu().numpy()
    assert np.array_equal(plan.logits_scale.reshape(-1).cpu().numpy(), want["logit_scale"])
    assert np.array_equal(codes, want["logit_codes"])
    assert np.array_equal(plan.greedy().cpu().numpy(), want["greedy"])
    assert np.abs(logits.cpu().numpy() - want["logits"]).max() <= 1e-5 * np.abs(want["logits"]).max()
    model.use_fused_plan = False
    lp, el, gr = model(processed_signal=xd, processed_signal_length=ld)
    assert np.array_equal(model.decoder.last_logits.cpu().numpy(), want["logits"])
    assert np.array_equal(gr.cpu().numpy(), want["greedy"])
    print("dynamic plan:", plan.stats)

"""Round-2 golden vectors from the UNMODIFIED reference (run in the build container):

    python tests/golden/gen_golden_extra.py      # writes tests/golden/ops_extra.npz

Rows of SURVEY.md 8(a) the first fixture set did not cover:
  * QuantAct(per_channel=True): quant_modules.py:117-119 (calibrating), :154-156 (dynamic),
    per-channel scale :175-177, requantisation with a (1,C,1) target scale :186-194;
  * QuantLinear.forward :431-469.  The reference's bias call at :462-463 passes FOUR arguments to the
    three-argument SymmetricQuantFunction and raises TypeError, so the class cannot run as written.  The
    vectors are produced by the reference's own forward with ONE shim: `weight_function` drops the stray
    `False` argument (nothing else is touched); per_channel=False only, because the reference's
    per-channel weight call broadcasts a (out,) scale over the `in` axis (:454-456), which is not a
    per-row quantiser.  The per-row intent is covered by the oracle restatement alone;
  * QuantConv1d with a folded but un-fixed BatchNorm (:337-349): the statistics update of calibration
    under quantize_model.train().
Seeds are a fixed table (zlib.crc32 of the case tag): regenerating gives the same bytes.
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [HERE, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]

from ref_loader import load_reference  # noqa: E402
import specio  # noqa: E402

torch.set_grad_enabled(False)
ref = load_reference()
Q = ref.quant_modules
SQF = ref.quant_utils.SymmetricQuantFunction


def seed_of(tag):
    return zlib.crc32(tag.encode()) % (2 ** 31)


def npf(t):
    return t.detach().numpy().astype(np.float32).copy()


def act_per_channel_cases():
    cases = []

    def run(tag, bit, calls, identity=False, first_layer=False, momentum=0.95, dynamic=False,
            running_stat=False, ranges=None, C=10, T=36, B=3):
        g = torch.Generator().manual_seed(seed_of(tag))
        act = Q.QuantAct(bit, quant_mode="symmetric", act_range_momentum=momentum, dynamic=dynamic,
                         per_channel=True, channel_len=C)
        act.running_stat = running_stat
        if ranges is not None:
            lo, hi = ranges
            act.x_min = -(torch.rand(C, generator=g) * lo + 0.5)
            act.x_max = torch.rand(C, generator=g) * hi + 0.5
        rec = dict(tag=tag, kind="act_per_channel", bit=bit, momentum=momentum, dynamic=dynamic,
                   running_stat=running_stat, x_min0=npf(act.x_min), x_max0=npf(act.x_max), calls=[])
        for ci in range(calls):
            if first_layer:
                x = torch.randn(B, C, T, generator=g) * (1.0 + 0.5 * ci)
                pre = ident = ident_sf = None
            else:
                pre = torch.rand(1, C, 1, generator=g) * 2e-4 + 1e-5
                z = torch.randint(-2 ** 15, 2 ** 15, (B, C, T), generator=g).float()
                if ci % 2 == 1:
                    z = z.abs()
                x = z * pre
                ident = ident_sf = None
                if identity:
                    ident_sf = torch.rand(1, C, 1, generator=g) * 2e-4 + 1e-5
                    ident = torch.randint(-2 ** 15, 2 ** 15, (B, C, T), generator=g).float() * ident_sf
            out, sf = act(x, pre, ident, ident_sf)
            sfc = sf.reshape(1, -1, 1)
            rec["calls"].append(dict(
                x=npf(x), pre_sf=None if pre is None else npf(pre).reshape(-1),
                identity=None if ident is None else npf(ident),
                identity_sf=None if ident_sf is None else npf(ident_sf).reshape(-1),
                out=npf(out), scale=npf(sf).reshape(-1),
                codes=torch.round(out / sfc).to(torch.int32).numpy(),
                x_min=npf(act.x_min).reshape(-1), x_max=npf(act.x_max).reshape(-1)))
        cases.append(rec)

    run("pc_static_requant_8", 8, 2, ranges=(4.0, 4.0))
    run("pc_static_requant_9", 9, 1, ranges=(0.0, 5.0))
    run("pc_static_first_layer_8", 8, 1, first_layer=True, ranges=(3.0, 3.0))
    run("pc_static_identity_8", 8, 2, identity=True, ranges=(6.0, 7.0))
    run("pc_calib_minmax_ema", 8, 3, running_stat=True)
    run("pc_calib_mom_neg1_identity", 8, 3, running_stat=True, momentum=-1, identity=True)
    run("pc_dynamic", 8, 2, dynamic=True)
    run("pc_dynamic_identity", 6, 2, dynamic=True, identity=True)
    return cases


def linear_cases():
    cases = []

    def run(tag, n_in, n_out, wbit, in_bit=8, N=37, lead=None, bias=True):
        g = torch.Generator().manual_seed(seed_of(tag))
        lin = torch.nn.Linear(n_in, n_out, bias=bias)
        lin.weight.copy_(torch.randn(n_out, n_in, generator=g) * 0.1)
        if bias:
            lin.bias.copy_(torch.randn(n_out, generator=g) * 0.3)
        ql = Q.QuantLinear(wbit, bias_bit=32, per_channel=False, quant_mode="symmetric")
        ql.set_param(lin)
        ql.weight_function = lambda *a: SQF.apply(a[0], a[1], a[-1])   # the one shim (see module docstring)
        n = 2 ** (in_bit - 1)
        shape = (N, n_in) if lead is None else tuple(lead) + (n_in,)
        codes = torch.randint(-n, n, shape, generator=g)
        s_x = torch.rand(1, generator=g) * 0.05 + 0.01
        out, sf = ql(codes.float() * s_x, s_x)
        acc = torch.round(out.double() / sf.double()).to(torch.int64).numpy()
        cases.append(dict(tag=tag, kind="linear", weight=npf(lin.weight), bias=npf(lin.bias), weight_bit=wbit,
                          in_codes=codes.to(torch.int32).numpy(), s_x=npf(s_x), out=npf(out),
                          out_scale=npf(sf).reshape(-1), acc=acc,
                          weight_integer=ql.weight_integer.to(torch.int32).numpy(),
                          bias_integer=ql.bias_integer.double().numpy(),
                          fc_scaling_factor=npf(ql.fc_scaling_factor).reshape(-1)))

    run("linear_w8_64x29", 64, 29, 8)
    run("linear_w4_96x40", 96, 40, 4)
    run("linear_w6_lead_dims", 48, 16, 6, lead=(3, 11))
    run("linear_w8_200x130", 200, 130, 8, N=150)
    return cases


def conv_unfixed_bn_cases():
    cases = []

    def run(tag, cin, cout, k, groups=1, calls=2, B=3, T=50, in_bit=9):
        g = torch.Generator().manual_seed(seed_of(tag))
        conv = torch.nn.Conv1d(cin, cout, k, padding=k // 2, groups=groups, bias=False)
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.1)
        qc = Q.QuantConv1d(8, bias_bit=32, quant_mode="symmetric", per_channel=True, fix_bn=False)
        qc.set_param(conv)
        b = torch.nn.BatchNorm1d(cout, eps=1e-3)
        b.weight.copy_(0.3 + 1.5 * torch.rand(cout, generator=g))
        b.bias.copy_(torch.randn(cout, generator=g) * 0.2)
        b.running_mean.copy_(torch.randn(cout, generator=g) * 0.1)
        b.running_var.copy_(0.5 + torch.rand(cout, generator=g))
        b.eval()
        bn0 = dict(weight=npf(b.weight), bias=npf(b.bias), mean=npf(b.running_mean), var=npf(b.running_var),
                   eps=float(b.eps), momentum=float(b.momentum))
        qc.bn_folding(b)
        rec = dict(tag=tag, kind="conv_unfixed_bn", weight=npf(conv.weight), groups=groups, pad=k // 2, bn=bn0,
                   calls=[])
        n = 2 ** (in_bit - 1)
        for _ in range(calls):
            codes = torch.randint(0 if in_bit == 9 else -n, n, (B, cin, T), generator=g)
            s_x = torch.rand(1, generator=g) * 0.05 + 0.01
            out, sf = qc(codes.float() * s_x, s_x)
            rec["calls"].append(dict(in_codes=codes.to(torch.int32).numpy(), s_x=npf(s_x), out=npf(out),
                                     out_scale=npf(sf).reshape(-1), mean=npf(qc.bn.running_mean),
                                     var=npf(qc.bn.running_var)))
        cases.append(rec)

    run("unfixed_bn_pw", 32, 24, 1, in_bit=8)
    run("unfixed_bn_dw_k9", 16, 16, 9, groups=16)
    return cases


if __name__ == "__main__":
    cases = act_per_channel_cases() + linear_cases() + conv_unfixed_bn_cases()
    path = os.path.join(HERE, "ops_extra.npz")
    specio.save(path, cases)
    print("ops_extra.npz: %d cases, %.1f KB" % (len(cases), os.path.getsize(path) / 1024))

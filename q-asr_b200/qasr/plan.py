"""Fused int8-resident execution plan for a calibrated, static Jasper / QuartzNet model.

The module-level API (nemo.quantization.utils.quant_modules) hands float32 code*scale tensors
from module to module like the reference does.  When every QuantAct is fixed and static the whole
forward is a fixed sequence of integer kernels, so this plan walks the module tree ONCE
(reference control flow: conv_asr.py:194-206, jasper.py:641-692, 175-194), folds every
QuantConv1d rescale + ReLU + consumer-QuantAct requantisation (+ res_act + block ReLU + the next
activation quantiser) into the producing kernel's epilogue, allocates the int8 tensors, and
records the launches into one CUDA graph.  Intermediates never exist as int32 / float32 in HBM.

A tensor in the plan = "value V of the reference graph as seen through QuantAct A": the int8
codes A would produce from V, masked or not.  Distinct (V, A-range, mask) keys are materialised
separately by re-running the producer with another epilogue, so arbitrary (even inconsistent)
calibration ranges stay exact; in a normally calibrated model all consumers of a block output
share one range and every producer runs once.
"""
import math

import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops

F32, F64 = np.float32, np.float64


# ------------------------------------------------------------------------------------------
# host-side parameter math (once per plan; float32 / float64 steps as the reference does them)
# ------------------------------------------------------------------------------------------
def act_scale(bit, x_min, x_max):
    """symmetric_linear_quantization_params on the fixed range (quant_utils.py:47-52)."""
    n = 2 ** (bit - 1) - 1
    r = max(abs(F32(x_min)), abs(F32(x_max)))
    return F32(max(r, F32(1e-8)) / F32(n))


def dyadic_np(pre_sf, z_sf):
    """(m, e) of quant_utils.py:188-193 + batch_frexp :121-147, m normalised to [2^30, 2^31)."""
    ratio = np.asarray(pre_sf, F32).astype(F64) / F64(F32(z_sf))
    mant, ex = np.frexp(ratio)
    m = np.floor(mant * 2147483648.0 + 0.5).astype(np.int64)
    e = (31 - ex).astype(np.int64)
    big = m >= 2 ** 31
    m = np.where(big, m // 2, m)
    e = np.where(big, e - 1, e)
    return m, e


def _tz(v):
    v = int(v)
    return (v & -v).bit_length() - 1 if v else 64


def fast_ok(bound, m, e):
    """May the all-integer epilogue ((z*m + 2^(e-1)) >> e with z = acc) replace the literal
    float recipe for |z| <= bound?  Needs: f32 round trip of the accumulator is the identity
    (bound < 2^22), the product stays exact in double (z*m < 2^53), the result fits int32, and no
    z in range lands on an exact .5 tie (where half-to-even and half-up differ)."""
    if bound >= 2 ** 22:
        return False
    for mi, ei in zip(np.atleast_1d(m).tolist(), np.atleast_1d(e).tolist()):
        if mi == 0:
            continue
        if ei < 32 or ei > 62:     # fast path shifts the high word of the product by e - 32
            return False
        if bound * mi >= 2 ** 53 or (bound * mi) >> ei >= 2 ** 30:
            return False
        t = ei - 1 - _tz(mi)          # z must be a multiple of 2^t for an exact tie
        if t < 0 or (1 << t) <= bound:
            return False
    return True


F32_IDENTITY = 2 ** 22 - 1   # largest |acc + bias| for which the reference's float32 round trip is proved to be the identity


def fast_ok_checked(bound, m, e):
    """(ok, range_check): the integer fast path for |z| <= bound, if need be with the kernel's per-element range check
    (QASR_F_RANGE_CHECK): proved up to 2^22 - 1, elements beyond take the literal recipe inside the same launch."""
    if fast_ok(bound, m, e):
        return True, False
    if bound < 2 ** 31 - 1 and fast_ok(min(bound, F32_IDENTITY), m, e):
        return True, True
    return False, False


def f32_ok(bound, m, e, lo, hi):
    """f32_ok_channels() for every channel at once (QASR_F_FAST_F32)."""
    return bool(f32_ok_channels(bound, m, e, lo, hi).all())


def f32_multipliers(bound, m, e, lo, hi, spread=1):
    """Per channel: a float32 multiplier r' for which the one-FMA float epilogue clamp(rint(f32(z) * r'), lo, hi) equals the
    dyadic rne(z * m / 2^e) for every integer |z| <= bound, or 0.0 when none is found.  The canonical choice is
    r32 = f32(m * 2^-e); where its rounding error moves one of the (hi - lo + 2) jumps across an integer z, a neighbour
    r32 +- 1 ulp may not (the multipliers that satisfy every jump form an interval of about one ulp around m 2^-e): with
    255 jumps (8-bit codes) 82 % of the channels pass on r32 and 91 % on a neighbour, with 64 jumps (6-bit) 99 % and 100 %,
    which matters because the GEMM picks its float path per WARP of 32 channels.  The proof
    is the one of f32_ok_channels(), run on the multiplier that is actually handed to the kernel (qasr_epilogue.f32_mult).
    Call only when fast_ok() already holds."""
    m1 = np.atleast_1d(m).astype(np.int64)
    e1 = np.atleast_1d(e).astype(np.int64)
    r32 = (np.where(m1 > 0, m1, 1).astype(F64) * np.exp2(-np.clip(e1, 1, 62).astype(F64))).astype(F32)
    out = np.zeros(m1.shape[0], F32)
    for k in sorted(range(-spread, spread + 1), key=abs):
        cand = r32.copy()
        for _ in range(abs(k)):
            cand = np.nextafter(cand, F32(np.inf) if k > 0 else F32(0), dtype=F32)
        ok = f32_ok_channels(bound, m, e, lo, hi, mult=cand) & (out == 0) & (cand > 0)
        out = np.where(ok, cand, out)
    return out


def f32_ok_channels(bound, m, e, lo, hi, mult=None):
    """Per channel: may the one-FMA float epilogue clamp(rint(f32(z) * f32(m * 2^-e)), lo, hi) (QASR_F_FAST_F32) replace the
    dyadic rne(z * m / 2^e) for every integer |z| <= bound?  Call only when fast_ok() already holds (so the
    dyadic has no ties and equals (z*m + 2^(e-1)) >> e).  Both functions are non-decreasing in z, so they agree
    everywhere iff they agree on both sides of every jump of the exact one inside the clamp window (and at the
    ends of the range): at most hi - lo + 2 jumps per channel, checked with exact int64 / float64 arithmetic."""
    m = np.atleast_1d(m).astype(np.int64).reshape(-1, 1)
    e = np.atleast_1d(e).astype(np.int64).reshape(-1, 1)
    if bound >= 2 ** 22:
        return np.zeros(m.shape[0], bool)
    sane = ((m > 0) & (e >= 1) & (e <= 62)).reshape(-1)
    m, e = np.where(m > 0, m, 1), np.clip(e, 1, 62)
    r64 = m.astype(F64) * np.exp2(-e.astype(F64))          # exact: 31-bit integer times a power of two
    r32 = r64.astype(F32).astype(F64) if mult is None else np.asarray(mult, F32).astype(F64).reshape(-1, 1)
    sane &= (bound * r64.reshape(-1)) < 2 ** 21
    h = np.arange(lo - 1, hi + 1, dtype=F64).reshape(1, -1) + 0.5     # exact jumps happen where z*r crosses h
    z0 = np.floor(h / r64)
    cand = np.concatenate([z0 + d for d in (-1.0, 0.0, 1.0, 2.0)] +
                          [np.full((m.shape[0], 1), -float(bound)), np.full((m.shape[0], 1), float(bound))], axis=1)
    cand = np.clip(cand, -float(bound), float(bound))
    zi = cand.astype(np.int64)
    exact = (zi * m + (np.int64(1) << (e - 1))) >> e          # |z*m| < 2^53, e <= 62: fits int64
    fl = np.rint(cand * r32)                                  # z * f32(r): exact in double, one rounding
    return sane & (np.clip(exact, lo, hi) == np.clip(fl, lo, hi).astype(np.int64)).all(axis=1)


def clamp16_ok(bounds, mult):
    """QASR_F_CLAMP16: may the float epilogue clamp on 16-bit lanes?  The kernel takes the low 16 bits of the FMA result
    fl(f32(z) * r + 1.5 * 2^23) as an int16, which is the rounded product n as long as |n| < 2^15 -- for EVERY accumulator
    the channel can reach (its worst-case bound: beyond the clamp window n only has to stay saturated, but it must not
    wrap), with f32(z) = z exactly (|z| < 2^22, so that 0x4B400000 + z is the float 1.5 * 2^23 + z)."""
    bounds = np.atleast_1d(bounds).astype(F64)
    mult = np.atleast_1d(mult).astype(F64)
    if bounds.size == 0 or bounds.max() >= 2 ** 22:
        return False
    return bool((bounds * mult + 1.0 < 2 ** 15 - 1).all())


def stage2_16_ok(hi1, m2, e2, lo2, hi2):
    """QASR_F_STAGE2_16: may the second requantisation run in 16-bit fixed point?  Its argument is a code q of [0, hi1]
    (first clamp with the block ReLU folded in), so the check is exhaustive: (q * M16 + 2^15) >> 16 with
    M16 = rne(m2 * 2^(16 - e2)) must equal (q * m2 + 2^(e2-1)) >> e2 -- what stage2_fast_ok() proved equal to the
    reference's rne(q * m2 / 2^e2) -- for every q, and everything must fit the kernel's 32- / 16-bit lanes."""
    m2, e2 = int(np.atleast_1d(m2)[0]), int(np.atleast_1d(e2)[0])
    if not (0 <= hi1 <= 127 and 17 <= e2 <= 62 and m2 > 0 and lo2 <= 0 <= hi2 < 2 ** 15):
        return False
    m16 = (m2 + (1 << (e2 - 17))) >> (e2 - 16)
    if hi1 * m16 + 0x8000 >= 2 ** 31:
        return False
    return all(((q * m16 + 0x8000) >> 16) == ((q * m2 + (1 << (e2 - 1))) >> e2) for q in range(hi1 + 1))


def window_bound(m, e, lo, hi):
    """Smallest Z such that the requantised value of every accumulator |z| >= Z lies outside the clamp window [lo, hi]
    for EVERY channel: (Z * m) >> e >= max(|lo|, |hi|) + 2.  All epilogue recipes (the reference's float one, the integer
    and the float fast paths) are non-decreasing in z, so recipes that agree on |z| <= Z and are saturated at +-Z agree
    for every larger accumulator as well: the proofs only have to cover min(bound, Z).  That matters for 8-bit weights
    and activations: the worst-case accumulator of a 512-channel pointwise conv (512 x 127 x 255 = 2^24) exceeds the 2^22
    up to which the reference's float32 round trip (quant_modules.py:305-307, quant_utils.py:187) is the identity, while
    its clamp window ends at |z| ~ 128 / r ~ 2^17."""
    m = np.atleast_1d(m).astype(np.int64)
    e = np.atleast_1d(e).astype(np.int64)
    live = m > 0
    if not live.any():
        return 0
    if ((e[live] < 1) | (e[live] > 62)).any():
        return 2 ** 62
    want = max(abs(int(lo)), abs(int(hi))) + 2
    z = (want * (np.int64(1) << e[live]) + m[live] - 1) // m[live] + 2      # ceil(want * 2^e / m) + 2
    return int(z.max())


def stage2_fast_ok(bound, m, e):
    """Same proof for the scalar second requantisation (generic 64-bit shift: 1 <= e <= 62)."""
    mi, ei = int(np.atleast_1d(m)[0]), int(np.atleast_1d(e)[0])
    if ei < 1 or ei > 62 or bound * mi >= 2 ** 53:
        return False
    t = ei - 1 - _tz(mi)
    return t >= 0 and (1 << t) > bound


class ActInfo:
    """Static view of one QuantAct."""

    def __init__(self, act, masked):
        if act.per_channel or act.dynamic or act.running_stat:
            raise ValueError("fused plan needs fixed, static, per-tensor QuantAct modules")
        self.bit = int(act.activation_bit)
        self.x_min = float(act.x_min.reshape(-1)[0])
        self.x_max = float(act.x_max.reshape(-1)[0])
        self.scale = act_scale(self.bit, self.x_min, self.x_max)
        self.n = 2 ** (self.bit - 1) - 1
        self.masked = masked
        self.module = act

    def key(self):
        return (self.bit, self.scale.tobytes(), self.masked)


# ------------------------------------------------------------------------------------------
# symbolic values of the reference graph
# ------------------------------------------------------------------------------------------
class Val:
    relu = False


class InputVal(Val):
    def __init__(self, C, T):
        self.C, self.T, self.stage = C, T, 0


class ConvVal(Val):
    """Output of a QuantConv1d applied to the quantised tensor `src` (a QTensor)."""

    def __init__(self, conv, src, T, stage):
        self.conv, self.src, self.T, self.stage = conv, src, T, stage
        self.C = conv.out_channels


class ResActVal(Val):
    """res_act(main, branch) (jasper.py:681): main is a ConvVal or a previous ResActVal."""

    def __init__(self, act, main, branch):
        self.act, self.main, self.branch = act, main, branch
        self.C, self.T, self.stage = main.C, main.T, main.stage


class ReluOf(Val):
    """ReLU applied to the float value `base` (jasper.py:661 / :687)."""

    def __init__(self, base):
        self.base = base
        self.C, self.T, self.stage = base.C, base.T, base.stage


class QTensor:
    """A materialised int8 tensor: value `val` quantised by `info` (None: plain res_act codes).
    Layout [B][C][ld] (time contiguous) or, in channels-last plans, [B][T][ld]."""

    def __init__(self, val, info, tensor, signed, scale, qabs, ntc=False):
        self.val, self.info, self.tensor, self.signed, self.scale, self.qabs = val, info, tensor, signed, scale, qabs
        self.ntc = ntc

    def codes_bct(self):
        """Integer codes as a (B, C, T) int64 numpy array (tests / debugging)."""
        t = self.tensor
        a = (t[:, :self.val.T, :self.val.C].permute(0, 2, 1) if self.ntc else t[:, :, :self.val.T]).cpu().numpy()
        return a.astype(np.int64) if self.signed else a.view(np.uint8).astype(np.int64)


class ConvParams:
    """Integer parameters of one QuantConv1d for a given input scale."""

    def __init__(self, conv, s_x, in_qabs):
        p = conv._prepare()
        self.conv = conv
        self.depthwise = p["depthwise"]
        self.s_w = p["s_w"].detach().float().cpu().numpy().astype(F32)
        self.s_out = (self.s_w * F32(s_x)).astype(F32)                       # quant_modules.py:293
        w_abs = p["w_int"].detach().abs().reshape(conv.out_channels, -1).sum(dim=1).cpu().numpy().astype(np.int64)
        if p["bias"] is not None:
            b = p["bias"].detach().float().cpu().numpy().astype(F32)
            inv = (F32(1.0) / self.s_out).astype(F32)
            n = 2 ** 31 - 1
            bi = np.clip(np.round((inv * b).astype(F32)), F32(-n), F32(n - 1))  # float32, :296-297
            self.b_int = bi.astype(F64).astype(np.int64)
        else:
            self.b_int = np.zeros(conv.out_channels, np.int64)
        self.bounds = w_abs * in_qabs + np.abs(self.b_int)          # per output channel: largest |acc + bias|
        self.bound = int(self.bounds.max())
        self.bias64 = self.bound >= 2 ** 31 - 1
        self.w8, self.packed = p["w8"], p["packed"]
        self.k, self.stride = conv.kernel_size[0], conv.stride[0]
        self.dil, self.pad = conv.dilation[0], conv.padding[0]

    def b_i32(self):
        return np.clip(self.b_int, -2 ** 31, 2 ** 31 - 1)  # 2^31 -> INT32_MAX sentinel (BIAS64 path)


PLAN_FORMAT = 1


class FusedPlan:
    """Compiled from a calibrated module tree (encoder, decoder), or re-created from a serialised plan
    (`FusedPlan.load`): every launch is first written down as a pure-data record (shapes, flags, numpy tables,
    buffer / weight ids) and then materialised into device tensors and a closure over the C ABI, so both paths run
    the same code and `save()` is a dump of the records."""

    def __init__(self, encoder, decoder, in_shape, device, use_graph=True, verbose=False, _spec=None):
        self.device = torch.device(device)
        self.B, self.C_in, self.T = int(in_shape[0]), int(in_shape[1]), int(in_shape[2])
        self.encoder, self.decoder = encoder, decoder
        self._bufs, self._buf_ids, self._weights, self._weight_ids = [], {}, [], {}
        self.spec = dict(format=PLAN_FORMAT, in_shape=[self.B, self.C_in, self.T], buffers=[], launches=[],
                         len_stages=[], outputs={})
        self.ops = []           # list of (name, callable)
        self.op_bytes = []      # algorithmic HBM bytes per op (int8 in + int8/f32 out)
        self.op_macs = []       # multiply-accumulates per op
        self.stats = dict(kernels=0, fast=0, exact=0, rerun=0, f32=0, f32_channels=0, channels=0)
        self._q = {}            # materialised QTensors
        self._keep = []
        self.verbose = verbose
        B = self.B
        self.ntc = self._needs_channels_last() if _spec is None else bool(_spec["ntc"])
        self.spec["ntc"] = self.ntc
        self.x_in = torch.zeros((B, self.C_in, self.T), dtype=torch.float32, device=self.device)
        self.lens = [torch.zeros(B, dtype=torch.int32, device=self.device)]   # per length stage
        self._lens_fns = []
        self._len_stages = []   # (k, stride, dilation, pad) of every conv that changes the frame count
        self.cur_T = self.T
        self.logits = self.encoded = None
        if _spec is None:
            self._build()
            self.spec["stats"] = dict(self.stats)
        else:
            self._rebuild(_spec)
        self.cur_T_out = self.out_frames(self.T)
        self.graph = None
        self.use_graph = use_graph
        self._warm = False

    def _needs_channels_last(self):
        """Dense k > 1 / strided convs (Jasper) run on the channels-last implicit-GEMM kernel; such a
        network keeps all its int8 activations channels-last."""
        from nemo.quantization.utils.quant_modules import QuantConv1d
        dense = depthwise = False
        for m in self.encoder.modules():
            if isinstance(m, QuantConv1d):
                dw = m.groups == m.in_channels == m.out_channels and m.groups > 1
                depthwise |= dw
                dense |= (not dw) and (m.kernel_size[0] > 1 or m.stride[0] > 1)
        if dense and depthwise:
            raise NotImplementedError("networks mixing depthwise and dense k>1 convolutions")
        return dense

    # -- symbolic walk (conv_asr.py:194-206, jasper.py:641-692) ----------------------------
    def _build(self):
        from nemo.collections.asr.parts.jasper import MaskedConv1d
        xs = [InputVal(self.C_in, self.T)]
        stage, T = 0, self.T
        for block in self.encoder.encoder_layers:
            out = xs[-1]
            stage_in, T_in = stage, T
            for l in block.mconv:
                if isinstance(l, MaskedConv1d):
                    src = self._quantized(out, ActInfo(l.act, True), first_layer=isinstance(out, InputVal))
                    k, s, d, p = l.conv.kernel_size[0], l.conv.stride[0], l.conv.dilation[0], l.conv.padding[0]
                    T_new = (T + 2 * p - d * (k - 1) - 1) // s + 1
                    if s != 1 or T_new != T:
                        stage = self._new_len_stage(stage, k, s, d, p)
                    T = T_new
                    out = ConvVal(l.conv, src, T, stage)
                elif isinstance(l, nn.ReLU):
                    out = ReluOf(out)
                elif isinstance(l, nn.Dropout):
                    pass
                else:
                    raise NotImplementedError("layer %r in a fused plan (call bn_folding() first)" % (l,))
            if block.res is not None:
                for i, branch in enumerate(block.res):
                    assert len(branch) == 1 and isinstance(branch[0], MaskedConv1d), "bn_folding() first"
                    rc = branch[0]
                    rsrc = self._quantized(xs[i], ActInfo(rc.act, True))
                    if rc.conv.stride[0] != 1 or rc.conv.kernel_size[0] != 1:
                        raise NotImplementedError("strided residual branch")
                    if xs[i].T != T:
                        raise NotImplementedError("residual across a strided block")
                    out = ResActVal(block.res_act, out, ConvVal(rc.conv, rsrc, T, stage))
            mout = list(block.mout)
            if not isinstance(mout[0], nn.ReLU):
                raise NotImplementedError("block activation %r" % (mout[0],))
            out = ReluOf(out)
            xs = xs + [out] if (block.res is not None and block.dense_residual) else [out]
        self.enc_val, self.enc_T, self.enc_stage = xs[-1], T, stage
        self.logits = self.encoded = None
        if self.decoder is None:
            # encoder-only plan: the API returns the float tensor relu(f32(acc) * s_out) and its scale
            base, _ = self._strip(xs[-1])
            if not isinstance(base, ConvVal):
                raise NotImplementedError("encoder ending in a residual block")
            self._emit_dequant(xs[-1], "encoded")
        else:
            dconv = self.decoder.decoder_layers[0]
            src = self._quantized(xs[-1], ActInfo(self.decoder.act, False))
            self._emit_dequant(ConvVal(dconv, src, T, stage), "logits")

    def _new_len_stage(self, stage, k, s, d, p):
        prev = self.lens[stage]
        self.spec["len_stages"].append([int(stage), int(k), int(s), int(d), int(p)])
        self._len_stages.append((k, s, d, p))
        self.lens.append(torch.zeros_like(prev))
        idx = len(self.lens) - 1

        def fn(prev=prev, out=self.lens[idx]):
            out.copy_(torch.div(prev + (2 * p - d * (k - 1) - 1), s, rounding_mode="floor") + 1)
        self._lens_fns.append(fn)
        return idx

    # -- materialisation -------------------------------------------------------------------
    def _alloc(self, C, T):
        t = ops.alloc_codes_ntc(self.B, C, T, self.device) if self.ntc else ops.alloc_codes(self.B, C, T, self.device)
        self._register_buffer(t)
        return t

    def _register_buffer(self, t):
        self._buf_ids[t.data_ptr()] = len(self._bufs)
        self._bufs.append(t)
        self.spec["buffers"].append(dict(shape=list(t.shape), dtype=str(t.dtype).replace("torch.", "")))
        return len(self._bufs) - 1

    def _bid(self, t):
        return None if t is None else self._buf_ids[t.data_ptr()]

    def _wid(self, t):
        """Id of a packed weight tensor (device); the bytes are dumped by save()."""
        key = t.data_ptr()
        if key not in self._weight_ids:
            self._weight_ids[key] = len(self._weights)
            self._weights.append(t)
        return self._weight_ids[key]

    def _quantized(self, val, info, first_layer=False):
        key = (id(val), info.key())
        qt = self._q.get(key)
        if qt is not None:
            return qt
        if any(k[0] == id(val) for k in self._q):
            self.stats["rerun"] += 1
        self._keep.append(val)
        if isinstance(val, InputVal):
            qt = self._emit_input(val, info)
        else:
            qt = self._emit_requant(val, info)
        self._q[key] = qt
        return qt

    def _emit_input(self, val, info):
        """First layer: SymmetricQuantFunction on the float input (quant_modules.py:180-184)."""
        if info.bit > 8:
            raise NotImplementedError("first-layer activation_bit > 8")
        out = self._alloc(val.C, val.T)
        rec = dict(kind="quantize_input", out=self._bid(out), scale=float(info.scale), bit=int(info.bit),
                   masked=bool(info.masked), C=int(val.C), T=int(val.T))
        self.spec["launches"].append(rec)
        self._materialize(rec)
        return QTensor(val, info, out, True, info.scale, info.n, self.ntc)

    def _strip(self, val):
        relu = False
        while isinstance(val, ReluOf):
            relu, val = True, val.base
        return val, relu

    def _emit_requant(self, val, info):
        """Run the producer of `val` with an epilogue that ends in QuantAct `info`."""
        base, relu = self._strip(val)
        signed = not relu and True
        qmin, qmax = -info.n - 1, info.n
        if relu:
            signed = False
            if info.n > 255:
                raise NotImplementedError("activation_bit > 9")
        elif info.n > 127:
            raise NotImplementedError("signed activation codes wider than 8 bits")
        out = self._alloc(base.C, base.T)
        lens = self.lens[base.stage] if info.masked else None
        flags = L.F_MASK if info.masked else 0
        if isinstance(base, ConvVal):
            cp = ConvParams(base.conv, base.src.scale, base.src.qabs)
            m, e = dyadic_np(cp.s_out, info.scale)
            if relu:
                flags |= L.F_RELU0
            lo_eff = max(qmin, 0) if relu else qmin
            # the proofs cover the accumulators up to the end of the clamp window; beyond it every recipe is saturated
            # (window_bound).  The fast kernels stay monotone up to |z| < 2^29: int64 products, and the float path's
            # 0x4B400000 + z stays a positive, finite bit pattern.
            eff = cp.bound if cp.bound >= 2 ** 29 else min(cp.bound, window_bound(m, e, lo_eff, qmax))
            fast = fast_ok(eff, m, e)
            if not fast:
                fast, rc = fast_ok_checked(cp.bound, m, e)
                if rc:
                    flags |= L.F_RANGE_CHECK
                    eff = 0          # no float path on top of a range-checked launch
            kw = dict(qmin=qmin, qmax=qmax)
            if fast and not (flags & L.F_RANGE_CHECK) and os.environ.get("QASR_NO_F32") != "1":   # float epilogue where the host can prove it, channel by channel
                rf = f32_multipliers(eff, m, e, lo_eff, qmax)
                okc = rf != 0
                self.stats["f32_channels"] += int(okc.sum())
                self.stats["channels"] += int(okc.size)
                canonical = (m.astype(F64) * np.exp2(-e.astype(F64))).astype(F32)
                if okc.all() and np.array_equal(rf, canonical):
                    flags |= L.F_FAST_F32          # the kernels derive r32 from (m, e) themselves
                elif okc.any():
                    kw["f32_mult"] = rf
                if okc.any() and lo_eff == 0 and relu and clamp16_ok(cp.bounds[okc], rf[okc]) and os.environ.get("QASR_NO_CLAMP16") != "1":
                    flags |= L.F_CLAMP16
                    self.stats["clamp16"] = self.stats.get("clamp16", 0) + 1
            self._emit_conv([(base, cp, m, e)], out, flags, fast, kw, lens, "conv->act")
        elif isinstance(base, ResActVal):
            ra = ActInfo(base.act, False)
            if ra.n > 127:
                raise NotImplementedError("res_act wider than 8 bits")
            m2, e2 = dyadic_np(np.array([ra.scale], F32), info.scale)
            stage2 = dict(m2=int(m2[0]), e2=int(e2[0]), qmin2=qmin, qmax2=qmax)
            if relu:
                flags |= L.F_RELU1
            flags |= L.F_STAGE2
            fast2 = stage2_fast_ok(ra.n + 1, m2, e2)
            if (fast2 and relu and stage2_16_ok(ra.n, m2, e2, qmin, qmax) and os.environ.get("QASR_NO_STAGE2_16") != "1"):
                flags |= L.F_STAGE2_16
                self.stats["stage2_16"] = self.stats.get("stage2_16", 0) + 1
            self._emit_resact(base, ra, out, flags, fast2, stage2, lens)
        else:
            raise NotImplementedError(type(base))
        return QTensor(val, info, out, signed, info.scale, max(abs(qmin), abs(qmax)) if signed else qmax, self.ntc)

    def _plain_resact(self, val):
        """res_act codes before the block ReLU (input of the next res_act in a dense-residual chain)."""
        key = (id(val), "plain")
        qt = self._q.get(key)
        if qt is None:
            ra = ActInfo(val.act, False)
            out = self._alloc(val.C, val.T)
            self._emit_resact(val, ra, out, 0, True, {}, None)
            qt = QTensor(val, None, out, True, ra.scale, ra.n + 1, self.ntc)
            self._q[key] = qt
            self._keep.append(val)
        return qt

    def _emit_resact(self, val, ra, out, flags, fast_extra, stage2, lens):
        main, m_relu = self._strip(val.main)
        branch, b_relu = self._strip(val.branch)
        assert not m_relu and not b_relu and isinstance(branch, ConvVal)
        cpb = ConvParams(branch.conv, branch.src.scale, branch.src.qabs)
        mb, eb = dyadic_np(cpb.s_out, ra.scale)
        kw = dict(qmin=-ra.n - 1, qmax=ra.n, **stage2)
        if isinstance(main, ConvVal):
            cpm = ConvParams(main.conv, main.src.scale, main.src.qabs)
            mm, em = dyadic_np(cpm.s_out, ra.scale)
            (okm, rcm), (okb, rcb) = fast_ok_checked(cpm.bound, mm, em), fast_ok_checked(cpb.bound, mb, eb)
            fast = fast_extra and okm and okb
            if fast and (rcm or rcb):
                flags |= L.F_RANGE_CHECK
            self._emit_conv([(main, cpm, mm, em), (branch, cpb, mb, eb)], out, flags, fast, kw, lens, "res_act")
        else:  # chain: previous res_act result (same scale, ratio exactly 1) + this branch
            assert isinstance(main, ResActVal) and main.act is val.act
            prev = self._plain_resact(main)
            okb, rcb = fast_ok_checked(cpb.bound, mb, eb)
            fast = fast_extra and okb
            if fast and rcb:
                flags |= L.F_RANGE_CHECK
            self._emit_conv([(branch, cpb, mb, eb)], out, flags, fast, dict(prev=prev.tensor, **kw), lens,
                            "res_act_chain")

    @staticmethod
    def _chan_table_np(cp, m, e):
        """qasr_chan records (s_out f32, bias i32, m i32, e i32) as an int32 [C][4] array."""
        C = cp.s_out.size
        tab = np.zeros((C, 4), np.int32)
        tab[:, 0] = np.ascontiguousarray(cp.s_out, dtype=F32).view(np.int32)
        tab[:, 1] = np.asarray(cp.b_i32(), np.int64).astype(np.int32)
        tab[:, 2] = np.asarray(m, np.int64).astype(np.int32)
        tab[:, 3] = np.asarray(e, np.int64).astype(np.int32)
        return tab

    def _emit_conv(self, operands, out, flags, fast, epi_kw, lens, name):
        """One kernel: conv operand 0 (+ operand 1) -> epilogue -> int8 tensor `out`."""
        tabs = [self._chan_table_np(cp, m, e) for (_, cp, m, e) in operands]
        v0 = operands[0][0]
        if v0.src.signed:
            flags |= L.F_B_SIGNED
        if len(operands) > 1 and operands[1][0].src.signed:
            flags |= L.F_B1_SIGNED
        if not fast:
            flags &= ~(L.F_STAGE2_16 | L.F_CLAMP16)      # hints of the fast paths
            flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
            if any(cp.bias64 for (_, cp, _, _) in operands):
                flags |= L.F_BIAS64
        self.stats["fast" if fast else "exact"] += 1
        if flags & L.F_RANGE_CHECK:
            self.stats["range_checked"] = self.stats.get("range_checked", 0) + 1
        if flags & L.F_FAST_F32:
            self.stats["f32"] += 1
        kw = dict(epi_kw)
        prev = kw.pop("prev", None)
        f32_mult = kw.pop("f32_mult", None)
        ep = dict(mode=L.EPI_REQUANT_I8, flags=int(flags), ch0=tabs[0], ch1=tabs[1] if len(tabs) > 1 else None,
                  prev=self._bid(prev), lens_stage=None if lens is None else self._stage_of(lens),
                  f32_mult=None if f32_mult is None else np.asarray(f32_mult, F32),
                  **{k: int(v) for k, v in kw.items()})
        self._launch_conv(operands, out, ep, name)

    def _stage_of(self, lens):
        return next(i for i, t in enumerate(self.lens) if t is lens)

    def _emit_dequant(self, val, name):
        """float32 output y = f32(acc + b_int) * s_out (decoder logits / encoder API output)."""
        base, relu = self._strip(val)
        cp = ConvParams(base.conv, base.src.scale, base.src.qabs)
        C = base.conv.out_channels
        tab = self._chan_table_np(cp, np.full(C, 2 ** 30, np.int64), np.full(C, 30, np.int64))
        flags = (L.F_RELU0 if relu else 0) | (L.F_BIAS64 if cp.bias64 else 0) | (L.F_B_SIGNED if base.src.signed else 0)
        ep = dict(mode=L.EPI_DEQUANT_F32, flags=int(flags), ch0=tab, ch1=None, prev=None, lens_stage=None, f32_mult=None)
        out = torch.zeros((self.B, C, ops.round_up(base.T, 16)), dtype=torch.float32, device=self.device)
        self._register_buffer(out)
        self._launch_conv([(base, cp, None, None)], out, ep, name)
        self.spec["outputs"][name] = dict(buffer=self._bid(out), scale=cp.s_out.copy(), T=int(base.T))
        self._bind_output(name)
        return out

    def _bind_output(self, name):
        o = self.spec["outputs"][name]
        setattr(self, name, self._bufs[o["buffer"]])
        setattr(self, name + "_scale", torch.from_numpy(np.asarray(o["scale"], F32).copy()).to(self.device).view(1, -1, 1))
        setattr(self, name + "_T", int(o["T"]))

    def _launch_conv(self, operands, out, ep, name):
        v0, cp0 = operands[0][0], operands[0][1]
        C_out, T_out = v0.conv.out_channels, v0.T
        assert not (cp0.depthwise and (len(operands) != 1 or self.ntc))
        recs = []
        nbytes = self.B * C_out * T_out * out.element_size() + (self.B * C_out * T_out if ep["prev"] is not None else 0)
        macs = 0
        for (v, cp, _, _) in operands:
            recs.append(dict(src=self._bid(v.src.tensor), T_in=int(v.src.val.T), C_in=int(v.conv.in_channels), k=int(cp.k),
                             stride=int(cp.stride), dil=int(cp.dil), pad=int(cp.pad),
                             weights=self._wid(cp.w8 if cp.depthwise else cp.packed)))
            nbytes += self.B * v.conv.in_channels * v.src.val.T
            macs += self.B * T_out * C_out * (1 if cp.depthwise else v.conv.in_channels) * cp.k
        kind = "dw" if cp0.depthwise else ("ntc" if self.ntc else "gemm")
        rec = dict(kind=kind, name=("dw_k%d_%s" % (cp0.k, name)) if cp0.depthwise else
                   "gemm_%dx%d_%s" % (v0.conv.in_channels, C_out, name),
                   out=self._bid(out), C_out=int(C_out), T_out=int(T_out), operands=recs, ep=ep, nbytes=int(nbytes),
                   macs=int(macs))
        self.spec["launches"].append(rec)
        self._materialize(rec)

    # -- records -> device tensors + closures over the C ABI (compile path and load path alike) -------------
    def _materialize(self, rec):
        import ctypes as C
        lib, B = L.lib(), self.B
        if rec["kind"] == "quantize_input":
            out = self._bufs[rec["out"]]
            scale = torch.tensor([rec["scale"]], dtype=torch.float32, device=self.device)
            lens = self.lens[0] if rec["masked"] else None
            x_in, Cn, T, bit = self.x_in, rec["C"], rec["T"], rec["bit"]
            if self.ntc:
                def run():
                    L.check(lib.qasr_quantize_f32_ntc(L.ptr(x_in), T, L.ptr(scale), 0, bit, L.ptr(out), out.shape[-1],
                                                      B, Cn, T, L.ptr(lens), L.stream_ptr()))
            else:
                def run():
                    L.check(lib.qasr_quantize_f32(L.ptr(x_in), T, L.ptr(scale), bit, L.ptr(out), out.shape[-1], None,
                                                  B, Cn, T, L.ptr(lens), L.stream_ptr()))
            self._add("quantize_input", run, keep=(scale,), nbytes=B * Cn * T * 5)
            return
        e = rec["ep"]
        dev = lambda a, dt=None: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
        stage2 = {k: e[k] for k in ("m2", "e2", "qmin2", "qmax2") if k in e}
        ep = ops.make_epilogue(e["mode"], dev(e["ch0"]), flags=e["flags"], ch1=dev(e["ch1"]),
                               prev=None if e["prev"] is None else self._bufs[e["prev"]],
                               qmin=e.get("qmin", -128), qmax=e.get("qmax", 127),
                               out_lens=None if e["lens_stage"] is None else self.lens[e["lens_stage"]],
                               f32_mult=dev(e["f32_mult"]), **stage2)
        out = self._bufs[rec["out"]]
        C_out, T_out = rec["C_out"], rec["T_out"]
        o0 = rec["operands"][0]
        if rec["kind"] == "dw":
            x, w8 = self._bufs[o0["src"]], self._weights[o0["weights"]]
            T_in, k, s, d, p = o0["T_in"], o0["k"], o0["stride"], o0["dil"], o0["pad"]

            def run():
                L.check(lib.qasr_dwconv1d_i8(L.ptr(x), x.shape[-1], L.ptr(w8), L.ptr(out), out.shape[-1], B,
                                             C_out, T_in, T_out, k, s, d, p, C.byref(ep), L.stream_ptr()))
            self._add(rec["name"], run, keep=(ep, w8), nbytes=rec["nbytes"], macs=rec["macs"])
            return
        gops = [ops.gemm_operand(self._bufs[o["src"]], o["T_in"], self._weights[o["weights"]], o["C_in"], o["k"],
                                 o["stride"], o["dil"], o["pad"]) for o in rec["operands"]]
        op1 = C.byref(gops[1]) if len(gops) > 1 else None
        entry = lib.qasr_conv1d_ntc_i8 if rec["kind"] == "ntc" else lib.qasr_conv1d_gemm_i8

        def run():
            L.check(entry(C.byref(gops[0]), op1, L.ptr(out), out.shape[-1], B, C_out, T_out, C.byref(ep), L.stream_ptr()))
        self._add(rec["name"], run, keep=(ep, gops), nbytes=rec["nbytes"], macs=rec["macs"])

    # -- serialisation (a compiled int8 plan for deployment) ------------------------------------
    def members(self):
        """The compiled plan as archive members for a .nemo-style tar.gz (the reference's container,
        nemo/core/classes/modelPT.py:254-296): `plan_manifest.json` (format / library version / shapes) + `plan.npz`
        (every launch record: packed int8 weights, b_int / (m, e) / s_out tables, f32 multipliers, epilogue flags).
        No module tree is needed to run it again (FusedPlan.load)."""
        import io
        import json
        from . import planio
        spec = dict(self.spec)
        spec["weights"] = [w.detach().cpu().numpy().view(np.uint8).reshape(-1) for w in self._weights]
        spec["enc_stage"] = int(self.enc_stage)
        manifest = dict(format=PLAN_FORMAT, qasr_version=int(L.load_library().qasr_version()), in_shape=spec["in_shape"],
                        ntc=self.ntc, launches=len(spec["launches"]), stats=spec.get("stats"))
        buf = io.BytesIO()
        planio.dump(buf, spec)
        return [("plan_manifest.json", json.dumps(manifest, indent=1).encode()), ("plan.npz", buf.getvalue())]

    def save(self, path):
        """A stand-alone plan archive (tar.gz of the two members); EncDecCTCModel.save_to(..., plan_shapes=) puts the
        same members into the .nemo file next to model_weights.ckpt."""
        import io
        import tarfile
        with tarfile.open(path, "w:gz") as tar:
            for name, data in self.members():
                info = tarfile.TarInfo("./" + name)
                info.size = len(data)
                tar.addfile(info, io.BytesIO(data))

    @classmethod
    def load(cls, path, device, use_graph=True):
        """Re-creates a plan saved by save().  Refuses a plan written by another library version: the packed weight
        layouts and epilogue flags are private to one libqasr build."""
        import io
        import json
        import tarfile
        from . import planio
        with tarfile.open(path, "r:*") as tar:
            files = {os.path.basename(m.name): tar.extractfile(m).read() for m in tar.getmembers() if m.isfile()}
        manifest = json.loads(files["plan_manifest.json"].decode())
        have = int(L.load_library().qasr_version())
        if manifest.get("format") != PLAN_FORMAT or manifest.get("qasr_version") != have:
            raise RuntimeError("plan %s was compiled for libqasr version %s / plan format %s; this build is version %d / "
                               "format %d: recompile the plan from the model" % (path, manifest.get("qasr_version"),
                                                                                manifest.get("format"), have, PLAN_FORMAT))
        spec = planio.load(io.BytesIO(files["plan.npz"]))
        return cls(None, None, spec["in_shape"], device, use_graph=use_graph, _spec=spec)

    def _rebuild(self, spec):
        for (stage, k, s, d, p) in spec["len_stages"]:
            self._new_len_stage(int(stage), int(k), int(s), int(d), int(p))
        for b in spec["buffers"]:
            self._register_buffer(torch.zeros(tuple(int(v) for v in b["shape"]), dtype=getattr(torch, b["dtype"]),
                                              device=self.device))
        self._weights = [torch.from_numpy(np.ascontiguousarray(w)).to(self.device).view(torch.int8) for w in spec["weights"]]
        for rec in spec["launches"]:
            self.spec["launches"].append(rec)
            self._materialize(rec)
        self.spec["outputs"] = spec["outputs"]
        for name in spec["outputs"]:
            self._bind_output(name)
        self.enc_stage = int(spec["enc_stage"])
        self.stats.update(spec.get("stats") or {})

    def _add(self, name, fn, keep=None, nbytes=0, macs=0):
        self.ops.append((name, fn))
        self.op_bytes.append(int(nbytes))
        self.op_macs.append(int(macs))
        self._keep.append(keep)
        self.stats["kernels"] += 1

    # -- execution -------------------------------------------------------------------------
    def _run_ops(self):
        for fn in self._lens_fns:
            fn()
        for _, fn in self.ops:
            fn()

    def _replay(self):
        if not self.use_graph:
            self._run_ops()
            return
        if self.graph is None:
            # eager warm-up on a side stream (sets kernel attributes, pages code in), then capture
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._run_ops()
            torch.cuda.current_stream(self.device).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_ops()
            self.graph = g
        self.graph.replay()

    def out_frames(self, T):
        """Encoder frames of a T-frame input (MaskedConv1d.get_seq_len, jasper.py:170-173, through every length stage)."""
        for (k, s, d, p) in self._len_stages:
            T = (T + 2 * p - d * (k - 1) - 1) // s + 1
        return T

    def profile_stamps(self):
        """One replay of an INSTRUMENTED copy of the plan's graph (same launches, same programmatic-dependent-launch
        edges; every conv kernel additionally stamps %globaltimer when its first CTA starts on activation data and when
        its last CTA ends).  Returns an int64 array (len(self.ops), 2) of nanoseconds, rows of launches without a stamp
        (the input quantiser) are (0, 0).  The timed product graph carries no stamps."""
        lib = L.lib()
        if getattr(self, "_prof", None) is None:
            conv_idx = [i for i, (n, _) in enumerate(self.ops) if n.startswith(("dw_", "gemm_"))]
            buf = torch.zeros((len(conv_idx), 2), dtype=torch.int64, device=self.device)
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._run_ops()                       # kernel attributes are set outside the capture
            torch.cuda.current_stream(self.device).wait_stream(s)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            L.check(lib.qasr_profile_begin(L.ptr(buf), len(conv_idx)))
            try:
                with torch.cuda.graph(g):
                    self._run_ops()
            finally:
                used = lib.qasr_profile_end()
            if used != len(conv_idx):
                raise RuntimeError("profile: %d stamped launches, expected %d" % (used, len(conv_idx)))
            self._prof = (g, buf, conv_idx)
        g, buf, conv_idx = self._prof
        buf[:, 0] = 2 ** 62
        buf[:, 1] = 0
        g.replay()
        torch.cuda.synchronize(self.device)
        out = np.zeros((len(self.ops), 2), np.int64)
        out[conv_idx] = buf.cpu().numpy()
        return out

    def load_inputs(self, x, lens):
        """x: (B, C, T) with T <= the plan's frame bucket (see bucket_frames); the tail of the bucket is zero."""
        B, C, T = x.shape
        if (B, C) != (self.B, self.C_in) or T > self.T:
            raise ValueError("plan compiled for input %s, got %s" % ((self.B, self.C_in, self.T), tuple(x.shape)))
        if T == self.T:
            self.x_in.copy_(x, non_blocking=True)
        else:
            self.x_in[:, :, :T].copy_(x, non_blocking=True)
            if self.cur_T != T:
                self.x_in[:, :, T:].zero_()
        self.cur_T, self.cur_T_out = T, self.out_frames(T)
        self.lens[0].copy_(lens.to(torch.int32), non_blocking=True)

    def run(self, x=None, lens=None):
        """Full model forward: returns the float32 logits view (B, V, T') (a static buffer)."""
        if x is not None:
            self.load_inputs(x, lens)
        self._replay()
        return self.logits[:, :, :self.cur_T_out] if self.logits is not None else None

    def run_encoder(self, x, lens):
        """ConvASREncoder.forward contract: (encoded float32 (B,C,T'), encoded_len, scale (1,C,1))."""
        self.load_inputs(x, lens)
        self._replay()
        enc = self.encoded[:, :, :self.cur_T_out].clone()
        return enc, self.out_lens().to(torch.long), self.encoded_scale

    def greedy(self):
        return ops.ctc_argmax(self.logits, self.cur_T_out)

    def log_probs(self):
        return ops.log_softmax_tv(self.logits, self.cur_T_out)

    def out_lens(self):
        return self.lens[self.enc_stage]

    def codes_of(self, val_key):
        return self._q[val_key]


_QUANT_MODULES = {}   # id(encoder) -> (weakref to encoder, id(decoder), state epoch, flat list of quant modules)


def _quant_modules(encoder, decoder):
    """The QuantAct / QuantConv1d modules of the model as a flat list.  The recursive module walk costs ~1 ms per call
    on QuartzNet15x5 (1 900 modules) and sat between the caller's copy and the graph launch of every forward; the
    tree can only change through stamped attribute assignments (quant_modules._bump moves the global epoch), so the
    list is rebuilt when the epoch moves and reused otherwise."""
    import weakref
    from nemo.quantization.utils import quant_modules as qmods
    epoch = qmods._STATE_COUNTER[0]
    hit = _QUANT_MODULES.get(id(encoder))
    if hit is not None and hit[0]() is encoder and hit[1] == id(decoder) and hit[2] == epoch:
        return hit[3]
    mods = list(encoder.modules()) + (list(decoder.modules()) if decoder is not None else [])
    mods = [m for m in mods if type(m) is qmods.QuantAct or isinstance(m, qmods.QuantConv1d)]
    for k in [k for k, v in _QUANT_MODULES.items() if v[0]() is None]:
        del _QUANT_MODULES[k]
    _QUANT_MODULES[id(encoder)] = (weakref.ref(encoder), id(decoder), epoch, mods)
    return mods


def _signature(encoder, decoder):
    """Everything the compiled plan bakes in.  Quant modules stamp every attribute assignment (ranges, bit widths,
    flags, weights, folded BN: quant_modules._bump), in-place edits show in the tensors' version counters."""
    from nemo.quantization.utils.quant_modules import QuantAct
    sig = []
    for m in _quant_modules(encoder, decoder):
        if type(m) is QuantAct:
            sig.append((m._state_version, m.x_min.data_ptr(), m.x_min._version, m.x_max.data_ptr(), m.x_max._version))
        else:
            sig.append(m.state_signature())
    return tuple(sig)


MAX_PLANS = 4   # compiled plans kept per model (LRU): each holds every int8 activation tensor of its shape + a CUDA graph


def bucket_frames(T):
    """Frame-count bucket a plan is compiled for: inputs are zero-padded up to it.  Every conv input is length-masked
    (jasper.py:177-181) and the conv's own padding is zeros too, so every output frame below the unpadded T' is
    bit-identical to an unbucketed run; ragged evaluation sets then share a handful of plans instead of compiling
    one per distinct length."""
    step = 32 if T <= 512 else (64 if T <= 2048 else 256)
    return (int(T) + step - 1) // step * step


class PlanUnsupported(Exception):
    """Negative cache entry: this model state cannot be compiled (the module-level path serves it)."""


def get_plan(encoder, decoder, in_shape, device, cache, **kw):
    """Plan for (batch, channels, bucket_frames(T)); `cache` is an ordered dict used as an LRU of MAX_PLANS entries.
    Raises NotImplementedError when the model cannot be fused (remembered per signature, so the symbolic walk is not
    retried on every forward)."""
    B, C, T = (int(v) for v in in_shape)
    key = ((B, C, bucket_frames(T)), str(device), id(decoder))
    sig = _signature(encoder, decoder)
    hit = cache.get(key)
    if hit is not None and hit[0] == sig:
        cache[key] = cache.pop(key)            # most recently used last
        if isinstance(hit[1], PlanUnsupported):
            raise NotImplementedError(str(hit[1]))
        return hit[1]
    cache.pop(key, None)
    for k in [k for k, v in cache.items() if v[0] != sig]:
        del cache[k]                           # plans of an older model state are dead weight
    try:
        plan = FusedPlan(encoder, decoder, key[0], device, **kw)
    except NotImplementedError as exc:
        cache[key] = (sig, PlanUnsupported(str(exc)))
        raise
    cache[key] = (sig, plan)
    while len(cache) > MAX_PLANS:
        cache.pop(next(iter(cache)))
    return plan

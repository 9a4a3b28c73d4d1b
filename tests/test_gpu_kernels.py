"""GPU: every kernel behind the C ABI against the oracle's integer arithmetic, bit-exact.
Run stand-alone (`python tests/test_gpu_kernels.py`) for a non-stopping report."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import qasr_oracle as O  # noqa: E402
from qasr import _lib as L  # noqa: E402
from qasr import ops  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rng(seed):
    return np.random.default_rng(seed)


def to_codes_tensor(q, signed):
    """int array (B,C,T) -> zero padded pitched device tensor."""
    B, Cn, T = q.shape
    t = ops.alloc_codes(B, Cn, T, DEV)
    arr = q.astype(np.int8) if signed else q.astype(np.uint8).view(np.int8)
    t[:, :, :T] = torch.from_numpy(arr).to(DEV)
    return t


def to_ntc_tensor(q, signed):
    """int array (B,C,T) -> channels-last zero padded device tensor [B][T][ld]."""
    B, Cn, T = q.shape
    t = ops.alloc_codes_ntc(B, Cn, T, DEV)
    arr = q.astype(np.int8) if signed else q.astype(np.uint8).view(np.int8)
    t[:, :, :Cn] = torch.from_numpy(np.ascontiguousarray(arr.transpose(0, 2, 1))).to(DEV)
    return t


def from_ntc_tensor(t, Cn, signed):
    a = t[:, :, :Cn].permute(0, 2, 1).cpu().numpy()
    return a.astype(np.int64) if signed else a.view(np.uint8).astype(np.int64)


def from_codes_tensor(t, T, signed):
    a = t[:, :, :T].cpu().numpy()
    return a.astype(np.int64) if signed else a.view(np.uint8).astype(np.int64)


def rand_dyadic(r, n, lo=2e-5, hi=3e-3):
    ratio = r.uniform(lo, hi, n)
    return O.batch_frexp(ratio)


def f32_dyadic(r, n, bound, lo, hi, rlo=2e-5, rhi=3e-3):
    """Per-channel dyadics for which the host proves the float fast path (plan.f32_ok), resampling the others."""
    from qasr import plan as P
    m, e = norm_me(*rand_dyadic(r, n, rlo, rhi))
    for _ in range(200):
        bad = [i for i in range(n) if not (P.fast_ok(bound, m[i:i + 1], e[i:i + 1]) and
                                           P.f32_ok(bound, m[i:i + 1], e[i:i + 1], lo, hi))]
        if not bad:
            return m, e
        mm, ee = norm_me(*rand_dyadic(r, len(bad), rlo, rhi))
        m[bad], e[bad] = mm, ee
    raise AssertionError("could not find float-exact dyadics")


def norm_me(m, e):
    m = np.asarray(m, np.int64).copy()
    e = np.asarray(e, np.int64).copy()
    big = m >= 2 ** 31
    m[big] //= 2
    e[big] -= 1
    return m, e


def expected_requant(acc, m, e, qmin, qmax, relu0=False, acc1=None, m1=None, e1=None, prev=None,
                     relu1=False, stage2=None, lens=None):
    z = np.maximum(acc, 0) if relu0 else acc
    o = O.dyadic_mul(z, m.reshape(1, -1, 1), e.reshape(1, -1, 1))
    if acc1 is not None:
        o = o + O.dyadic_mul(acc1, m1.reshape(1, -1, 1), e1.reshape(1, -1, 1))
    if prev is not None:
        o = o + prev
    q = np.clip(o, qmin, qmax)
    if relu1:
        q = np.maximum(q, 0)
    if stage2 is not None:
        m2, e2, lo2, hi2 = stage2
        q = np.clip(O.dyadic_mul(q, m2, e2), lo2, hi2)
    if lens is not None:
        T = q.shape[2]
        q = np.where(np.arange(T)[None, None, :] < lens[:, None, None], q, 0)
    return q.astype(np.int64)


# ------------------------------------------------------------------------------------------
# K2 GEMM
# ------------------------------------------------------------------------------------------
GEMM_CASES = [
    # name, B, C_in, C_out, T, k, dil, signed
    ("pw_256_256_u8", 2, 256, 256, 300, 1, 1, False),
    ("pw_64_256_s8", 2, 64, 256, 250, 1, 1, True),
    ("pw_512_512_s8", 1, 512, 512, 130, 1, 1, True),
    ("dec_1024_29_s8", 2, 1024, 29, 77, 1, 1, True),
    ("tiny_32_48_u8", 3, 32, 48, 40, 1, 1, False),
    ("pw_512_512_u8_t800", 3, 512, 512, 800, 1, 1, False),
    ("pw_256_512_s8_t300", 2, 256, 512, 300, 1, 1, True),
    ("pw_512_1024_u8_t520", 1, 512, 1024, 520, 1, 1, False),
    ("dense_k5", 2, 128, 128, 150, 5, 1, False),
    ("dense_k11_c384", 1, 256, 384, 140, 11, 1, False),
    ("dense_k7_dil2", 2, 96, 160, 90, 7, 2, False),
]


def _gemm_inputs(case, seed=0):
    name, B, ci, co, T, k, dil, signed = case
    r = rng(seed)
    q = r.integers(-128, 128, (B, ci, T)) if signed else r.integers(0, 256, (B, ci, T))
    w = r.integers(-127, 127, (co, ci, k))
    bias = r.integers(-50000, 50000, co)
    pad = (dil * k) // 2 - 1 if dil > 1 else k // 2
    return q, w, bias, pad


@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: c[0])
def test_gemm_raw_i32(case):
    name, B, ci, co, T, k, dil, signed = case
    q, w, bias, pad = _gemm_inputs(case)
    want = O.conv1d_int(q, w, bias, 1, pad, dil, 1)
    T_out = want.shape[2]
    x = to_codes_tensor(q, signed)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    ch = ops.chan_table(np.ones(co, np.float32), bias, np.full(co, 2 ** 30), np.full(co, 30), DEV)
    ep = ops.make_epilogue(L.EPI_RAW_I32, ch, flags=L.F_B_SIGNED if signed else 0)
    out = torch.full((B, co, ops.round_up(T_out, 16)), -777, dtype=torch.int32, device=DEV)
    if k > 1:  # (B,C,T) layout: TMA cannot fetch unaligned time shifts -> loud refusal, not a fault
        with pytest.raises(NotImplementedError):
            ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, k, 1, dil, pad), None, out, B, co, T_out, ep)
        # ... and the channels-last implicit GEMM computes it (int32 out in (B,C,T))
        xn = to_ntc_tensor(q, signed)
        ops.conv_ntc(ops.gemm_operand(xn, T, wp, ci, k, 1, dil, pad), None, out, B, co, T_out, ep)
        torch.cuda.synchronize()
        got = out[:, :, :T_out].cpu().numpy().astype(np.int64)
        assert np.array_equal(got, want), "ntc mismatches: %d / %d" % ((got != want).sum(), want.size)
        return
    ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, k, 1, dil, pad), None, out, B, co, T_out, ep)
    torch.cuda.synchronize()
    got = out[:, :, :T_out].cpu().numpy().astype(np.int64)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


@pytest.fixture(params=[0, 1, 2], ids=["cta1", "cta_pair", "cta1_mt2"])
def gemm_engine(request):
    """The GEMM schedules on every requantising case: one CTA per SM, CTA pairs issuing tcgen05 cta_group::2 (bring-up
    knob 12; shapes with an odd number of 128-row tiles fall back to one CTA either way), and one CTA with two resident M
    tiles forced (knob 5 = 2: the test shapes are below one wave of tiles, where the library picks one M tile per CTA)."""
    L.lib().qasr_debug_set(12, 1 if request.param == 1 else 0)
    L.lib().qasr_debug_set(5, 2 if request.param == 2 else -1)
    yield request.param
    L.lib().qasr_debug_set(12, -1)
    L.lib().qasr_debug_set(5, -1)


@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
@pytest.mark.parametrize("case", GEMM_CASES[:8], ids=lambda c: c[0])
def test_gemm_requant(case, fast, gemm_engine):
    name, B, ci, co, T, k, dil, signed = case
    q, w, bias, pad = _gemm_inputs(case, seed=1)
    acc = O.conv1d_int(q, w, bias, 1, pad, dil, 1)
    T_out = acc.shape[2]
    r = rng(2)
    m, e = norm_me(*rand_dyadic(r, co))
    lens = r.integers(T_out // 2, T_out + 1, B).astype(np.int32)
    m2, e2 = norm_me(*O.batch_frexp(np.array([1.9])))
    want = expected_requant(acc, m, e, -128, 127, relu0=True, relu1=True, stage2=(m2[0], e2[0], -256, 255),
                            lens=lens)
    x = to_codes_tensor(q, signed)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    s_out = r.uniform(1e-5, 1e-3, co).astype(np.float32)
    ch = ops.chan_table(s_out, bias, m, e, DEV)
    flags = L.F_RELU0 | L.F_RELU1 | L.F_STAGE2 | L.F_MASK | (L.F_B_SIGNED if signed else 0)
    if not fast:
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    lens_d = torch.from_numpy(lens).to(DEV)
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, qmin=-128, qmax=127, m2=m2[0], e2=e2[0],
                           qmin2=-256, qmax2=255, out_lens=lens_d)
    out = ops.alloc_codes(B, co, T_out, DEV)
    ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, k, 1, dil, pad), None, out, B, co, T_out, ep)
    torch.cuda.synchronize()
    got = from_codes_tensor(out, T_out, signed=False)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)
    assert int(out[:, :, T_out:].abs().sum()) == 0  # pitch padding stays zero


@pytest.mark.parametrize("clamp16", [False, True], ids=["clamp32", "clamp16"])
@pytest.mark.parametrize("case", GEMM_CASES[:8], ids=lambda c: c[0])
def test_gemm_requant_f32(case, gemm_engine, clamp16):
    """Float fast path (QASR_F_FAST_F32): single operand, ReLU folded into the clamp, host-proved dyadics; with
    QASR_F_CLAMP16 the clamp runs on packed 16-bit lanes (host proof: plan.clamp16_ok)."""
    name, B, ci, co, T, k, dil, signed = case
    q, w, bias, pad = _gemm_inputs(case, seed=1)
    acc = O.conv1d_int(q, w, bias, 1, pad, dil, 1)
    T_out = acc.shape[2]
    r = rng(12)
    bound = int(np.abs(acc).max())
    assert bound < 2 ** 22
    m, e = f32_dyadic(r, co, bound, 0, 63, 1e-6, 2e-4)
    lens = r.integers(T_out // 2, T_out + 1, B).astype(np.int32)
    want = expected_requant(acc, m, e, -64, 63, relu0=True, lens=lens)
    x = to_codes_tensor(q, signed)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    ch = ops.chan_table(r.uniform(1e-5, 1e-3, co).astype(np.float32), bias, m, e, DEV)
    flags = L.F_RELU0 | L.F_MASK | L.F_FAST_F32 | (L.F_B_SIGNED if signed else 0)
    if clamp16:
        from qasr import plan as P
        assert P.clamp16_ok(np.full(co, bound), (m.astype(np.float64) * np.exp2(-e.astype(np.float64))).astype(np.float32))
        flags |= L.F_CLAMP16
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, qmin=-64, qmax=63, out_lens=torch.from_numpy(lens).to(DEV))
    out = ops.alloc_codes(B, co, T_out, DEV)
    ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, k, 1, dil, pad), None, out, B, co, T_out, ep)
    torch.cuda.synchronize()
    got = from_codes_tensor(out, T_out, signed=False)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)
    assert int(out[:, :, T_out:].abs().sum()) == 0


def test_f32_mult_table_mixed(dw_engine):
    """Per-channel float multipliers (qasr_epilogue.f32_mult): proved channels take the float epilogue, the rest
    (0.0f entries) the integer one, in the depthwise kernels (per channel) and the GEMM (per warp of 32 rows)."""
    from qasr import plan as P
    r = rng(21)
    B, Cn, T, k = 20, 160, 330, 33
    q = r.integers(0, 256, (B, Cn, T))
    w = r.integers(-127, 127, (Cn, 1, k))
    acc = O.conv1d_int(q, w, None, 1, k // 2, 1, Cn)
    bound = int(np.abs(acc).max())
    m, e = norm_me(*rand_dyadic(r, Cn, 1e-4, 2e-3))
    ok = P.f32_ok_channels(bound, m, e, -128, 127) & np.array([P.fast_ok(bound, m[i:i + 1], e[i:i + 1]) for i in range(Cn)])
    assert 0 < ok.sum() < Cn, "want a mix of proved and unproved channels (got %d of %d)" % (ok.sum(), Cn)
    # channels where fast_ok itself fails would need the exact path: give them a proved dyadic instead
    bad = [i for i in range(Cn) if not P.fast_ok(bound, m[i:i + 1], e[i:i + 1])]
    if bad:
        mm, ee = f32_dyadic(r, len(bad), bound, -128, 127, 1e-4, 2e-3)
        m[bad], e[bad] = mm, ee
        ok[bad] = True
    rf = np.where(ok, (m.astype(np.float64) * np.exp2(-e.astype(np.float64))).astype(np.float32), np.float32(0)).astype(np.float32)
    want = expected_requant(acc, m, e, -128, 127)
    x = to_codes_tensor(q, False)
    wd = ops.pack_dw_weights(torch.from_numpy(w[:, 0, :].astype(np.int8)).to(DEV).contiguous(), 1, 1, k // 2)
    ch = ops.chan_table(r.uniform(1e-5, 1e-3, Cn).astype(np.float32), None, m, e, DEV)
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=0, qmin=-128, qmax=127, f32_mult=torch.from_numpy(rf).to(DEV))
    out = ops.alloc_codes(B, Cn, T, DEV)
    ops.dwconv(x, T, wd, out, T, k, 1, 1, k // 2, ep)
    torch.cuda.synchronize()
    assert np.array_equal(from_codes_tensor(out, T, signed=True), want)
    # same table through the pointwise GEMM (1x1 conv with one weight per channel pair)
    ci = 128
    q2 = r.integers(0, 64, (B, ci, T))
    w2 = r.integers(-31, 32, (Cn, ci, 1))
    acc2 = O.conv1d_int(q2, w2, None, 1, 0, 1, 1)
    bound2 = int(np.abs(acc2).max())
    m2, e2 = norm_me(*rand_dyadic(r, Cn, 1e-5, 5e-4))
    for i in range(Cn):
        while not P.fast_ok(bound2, m2[i:i + 1], e2[i:i + 1]):
            m2[i:i + 1], e2[i:i + 1] = norm_me(*rand_dyadic(r, 1, 1e-5, 5e-4))
    ok2 = P.f32_ok_channels(bound2, m2, e2, 0, 63)
    rf2 = np.where(ok2, (m2.astype(np.float64) * np.exp2(-e2.astype(np.float64))).astype(np.float32), np.float32(0)).astype(np.float32)
    rf2[:32] = np.where(ok2[:32], rf2[:32], 0)   # first warp: whatever the proof says; force one fully proved warp below
    want2 = expected_requant(acc2, m2, e2, -64, 63, relu0=True)
    wp = ops.pack_weights(torch.from_numpy(w2.astype(np.int8)).to(DEV))
    ch2 = ops.chan_table(r.uniform(1e-5, 1e-3, Cn).astype(np.float32), None, m2, e2, DEV)
    ep2 = ops.make_epilogue(L.EPI_REQUANT_I8, ch2, flags=L.F_RELU0, qmin=-64, qmax=63, f32_mult=torch.from_numpy(rf2).to(DEV))
    out2 = ops.alloc_codes(B, Cn, T, DEV)
    ops.conv_gemm(ops.gemm_operand(to_codes_tensor(q2, False), T, wp, ci, 1), None, out2, B, Cn, T, ep2)
    torch.cuda.synchronize()
    assert np.array_equal(from_codes_tensor(out2, T, signed=False), want2)
    # and with the 16-bit clamp on the float warps (QASR_F_CLAMP16)
    assert P.clamp16_ok(np.full(int(ok2.sum()), bound2), rf2[ok2])
    ep3 = ops.make_epilogue(L.EPI_REQUANT_I8, ch2, flags=L.F_RELU0 | L.F_CLAMP16, qmin=-64, qmax=63, f32_mult=torch.from_numpy(rf2).to(DEV))
    out3 = ops.alloc_codes(B, Cn, T, DEV)
    ops.conv_gemm(ops.gemm_operand(to_codes_tensor(q2, False), T, wp, ci, 1), None, out3, B, Cn, T, ep3)
    torch.cuda.synchronize()
    assert np.array_equal(from_codes_tensor(out3, T, signed=False), want2)


@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
def test_gemm_dual_residual(fast, gemm_engine):
    """pointwise conv + residual 1x1 conv meeting in the res_act epilogue (jasper.py:681)."""
    B, ci0, ci1, co, T = 2, 256, 128, 256, 200
    r = rng(3)
    q0 = r.integers(-128, 128, (B, ci0, T))
    q1 = r.integers(0, 256, (B, ci1, T))
    w0 = r.integers(-127, 127, (co, ci0, 1))
    w1 = r.integers(-127, 127, (co, ci1, 1))
    b0 = r.integers(-30000, 30000, co)
    b1 = r.integers(-30000, 30000, co)
    acc0 = O.conv1d_int(q0, w0, b0, 1, 0, 1, 1)
    acc1 = O.conv1d_int(q1, w1, b1, 1, 0, 1, 1)
    m0, e0 = norm_me(*rand_dyadic(r, co, 1e-5, 2e-4))
    m1, e1 = norm_me(*rand_dyadic(r, co, 1e-5, 2e-4))
    prev = r.integers(-128, 128, (B, co, T))
    m2, e2 = norm_me(*O.batch_frexp(np.array([2.01])))
    want = expected_requant(acc0, m0, e0, -128, 127, acc1=acc1, m1=m1, e1=e1, prev=prev, relu1=True,
                            stage2=(m2[0], e2[0], -256, 255))
    x0, x1 = to_codes_tensor(q0, True), to_codes_tensor(q1, False)
    wp0 = ops.pack_weights(torch.from_numpy(w0.astype(np.int8)).to(DEV))
    wp1 = ops.pack_weights(torch.from_numpy(w1.astype(np.int8)).to(DEV))
    one = np.full(co, 1e-4, np.float32)
    ch0, ch1 = ops.chan_table(one, b0, m0, e0, DEV), ops.chan_table(one, b1, m1, e1, DEV)
    flags = L.F_RELU1 | L.F_STAGE2 | L.F_B_SIGNED
    if not fast:
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch0, flags=flags, ch1=ch1, prev=to_codes_tensor(prev, True),
                           qmin=-128, qmax=127, m2=m2[0], e2=e2[0], qmin2=-256, qmax2=255)
    out = ops.alloc_codes(B, co, T, DEV)
    ops.conv_gemm(ops.gemm_operand(x0, T, wp0, ci0, 1), ops.gemm_operand(x1, T, wp1, ci1, 1), out, B, co, T, ep)
    torch.cuda.synchronize()
    got = from_codes_tensor(out, T, signed=False)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


@pytest.mark.parametrize("order", [-1, 0], ids=["full_tiles_first", "utterance_major"])
@pytest.mark.parametrize("shape", [(200, 300, 128), (170, 128, 128), (330, 100, 128), (90, 420, 256)], ids=lambda s: "B%d_T%d_C%d" % s)
def test_gemm_tile_walk_over_several_rounds(gemm_engine, order, shape):
    """More (utterance, time tile) pairs than CTAs per M group, so every CTA walks several rounds: the full-tiles-first
    order (set_tile_walk: partial last tiles after all full ones; T a multiple of the tile, T below one tile and
    T = 3 tiles + a rest) and the utterance-major order (bring-up knob 18 = 0) against the oracle, ragged lengths."""
    B, T, co = shape     # 256 output channels: two M tiles, so the CTA-pair engine walks them too
    ci = 64
    r = rng(44)
    q = r.integers(0, 64, (B, ci, T))
    w = r.integers(-31, 32, (co, ci, 1))
    bias = r.integers(-3000, 3000, co)
    acc = O.conv1d_int(q, w, bias, 1, 0, 1, 1)
    bound = int(np.abs(acc).max())
    m, e = f32_dyadic(r, co, bound, 0, 63, 1e-4, 2e-3)
    lens = r.integers(T // 2, T + 1, B).astype(np.int32)
    want = expected_requant(acc, m, e, -64, 63, relu0=True, lens=lens)
    x = to_codes_tensor(q, False)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    ch = ops.chan_table(r.uniform(1e-5, 1e-3, co).astype(np.float32), bias, m, e, DEV)
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_RELU0 | L.F_MASK | L.F_FAST_F32, qmin=-64, qmax=63,
                           out_lens=torch.from_numpy(lens).to(DEV))
    out = ops.alloc_codes(B, co, T, DEV)
    L.lib().qasr_debug_set(18, order)
    try:
        ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, 1), None, out, B, co, T, ep)
        torch.cuda.synchronize()
    finally:
        L.lib().qasr_debug_set(18, -1)
    got = from_codes_tensor(out, T, signed=False)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)
    assert int(out[:, :, T:].abs().sum()) == 0


@pytest.mark.parametrize("range_check", [False, True], ids=["plain", "range_check"])
def test_gemm_dual_stage2_16(gemm_engine, range_check):
    """QASR_F_STAGE2_16: the second requantisation of a two-operand res_act launch in 16-bit fixed point
    ((q * M16 + 2^15) >> 16, exhaustively checked on the host over q in [0, qmax]: plan.stage2_16_ok)."""
    from qasr import plan as P
    B, ci0, ci1, co, T = 3, 256, 128, 256, 333
    r = rng(33)
    q0 = r.integers(-64, 64, (B, ci0, T))
    q1 = r.integers(0, 128, (B, ci1, T))
    w0 = r.integers(-127, 127, (co, ci0, 1))
    w1 = r.integers(-127, 127, (co, ci1, 1))
    b0 = r.integers(-30000, 30000, co)
    b1 = r.integers(-30000, 30000, co)
    acc0 = O.conv1d_int(q0, w0, b0, 1, 0, 1, 1)
    acc1 = O.conv1d_int(q1, w1, b1, 1, 0, 1, 1)
    m0, e0 = norm_me(*rand_dyadic(r, co, 1e-5, 2e-4))
    m1, e1 = norm_me(*rand_dyadic(r, co, 1e-5, 2e-4))
    lens = r.integers(T // 2, T + 1, B).astype(np.int32)
    proved = 0
    for ratio in (1.9, 0.7307, 0.3111, 2.01, 1.0):
        m2, e2 = norm_me(*O.batch_frexp(np.array([ratio])))
        if not (P.stage2_fast_ok(128, m2, e2) and P.stage2_16_ok(127, m2, e2, -256, 255)):
            continue
        proved += 1
        want = expected_requant(acc0, m0, e0, -128, 127, acc1=acc1, m1=m1, e1=e1, relu1=True,
                                stage2=(m2[0], e2[0], -256, 255), lens=lens)
        x0, x1 = to_codes_tensor(q0, True), to_codes_tensor(q1, False)
        wp0 = ops.pack_weights(torch.from_numpy(w0.astype(np.int8)).to(DEV))
        wp1 = ops.pack_weights(torch.from_numpy(w1.astype(np.int8)).to(DEV))
        one = np.full(co, 1e-4, np.float32)
        ch0, ch1 = ops.chan_table(one, b0, m0, e0, DEV), ops.chan_table(one, b1, m1, e1, DEV)
        flags = L.F_RELU1 | L.F_STAGE2 | L.F_STAGE2_16 | L.F_B_SIGNED | L.F_MASK | (L.F_RANGE_CHECK if range_check else 0)
        ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch0, flags=flags, ch1=ch1, qmin=-128, qmax=127, m2=m2[0], e2=e2[0],
                               qmin2=-256, qmax2=255, out_lens=torch.from_numpy(lens).to(DEV))
        out = ops.alloc_codes(B, co, T, DEV)
        ops.conv_gemm(ops.gemm_operand(x0, T, wp0, ci0, 1), ops.gemm_operand(x1, T, wp1, ci1, 1), out, B, co, T, ep)
        torch.cuda.synchronize()
        got = from_codes_tensor(out, T, signed=False)
        assert np.array_equal(got, want), "ratio %g: mismatches: %d / %d" % (ratio, (got != want).sum(), want.size)
    assert proved >= 2


NTC_CASES = [
    # name, B, C_in, C_out, T, k, stride, dil, signed
    ("ntc_k11_s2_c64", 2, 64, 256, 301, 11, 2, 1, True),
    ("ntc_k13_c384", 1, 384, 384, 200, 13, 1, 1, False),
    ("ntc_k29_dil2", 2, 128, 160, 150, 29, 1, 2, False),
    ("ntc_k1_1024_29", 2, 1024, 29, 77, 1, 1, 1, True),
    ("ntc_tiny_k5", 3, 32, 48, 40, 5, 1, 1, False),
]


@pytest.mark.parametrize("fast", [False, True], ids=["exact", "fast"])
@pytest.mark.parametrize("case", NTC_CASES, ids=lambda c: c[0])
def test_conv_ntc_requant_dual_prev(case, fast):
    """Channels-last implicit GEMM: main conv (any k / stride / dilation) + residual 1x1 operand + previous
    res_act codes + block ReLU + second requantisation + length mask, against the oracle arithmetic."""
    name, B, ci, co, T, k, stride, dil, signed = case
    r = rng(20)
    q = r.integers(-128, 128, (B, ci, T)) if signed else r.integers(0, 256, (B, ci, T))
    w = r.integers(-20, 20, (co, ci, k))
    bias = r.integers(-3000, 3000, co)
    pad = (dil * k) // 2 - 1 if dil > 1 else k // 2
    acc0 = O.conv1d_int(q, w, bias, stride, pad, dil, 1)
    T_out = acc0.shape[2]
    ci1 = 64
    q1 = r.integers(0, 256, (B, ci1, T_out))
    w1 = r.integers(-60, 60, (co, ci1, 1))
    acc1 = O.conv1d_int(q1, w1, None, 1, 0, 1, 1)
    prev = r.integers(-128, 128, (B, co, T_out))
    m0, e0 = norm_me(*rand_dyadic(r, co, 1e-5, 1e-4))
    m1, e1 = norm_me(*rand_dyadic(r, co, 1e-5, 2e-4))
    m2, e2 = norm_me(*O.batch_frexp(np.array([2.01])))
    lens = r.integers(T_out // 2, T_out + 1, B).astype(np.int32)
    want = expected_requant(acc0, m0, e0, -128, 127, acc1=acc1, m1=m1, e1=e1, prev=prev, relu1=True,
                            stage2=(m2[0], e2[0], -256, 255), lens=lens)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    wp1 = ops.pack_weights(torch.from_numpy(w1.astype(np.int8)).to(DEV))
    one = np.full(co, 1e-4, np.float32)
    ch0 = ops.chan_table(one, bias, m0, e0, DEV)
    ch1 = ops.chan_table(one, None, m1, e1, DEV)
    flags = L.F_RELU1 | L.F_STAGE2 | L.F_MASK | (L.F_B_SIGNED if signed else 0)
    if not fast:
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch0, flags=flags, ch1=ch1, prev=to_ntc_tensor(prev, True), qmin=-128,
                           qmax=127, m2=m2[0], e2=e2[0], qmin2=-256, qmax2=255,
                           out_lens=torch.from_numpy(lens).to(DEV))
    out = ops.alloc_codes_ntc(B, co, T_out, DEV)
    ops.conv_ntc(ops.gemm_operand(to_ntc_tensor(q, signed), T, wp, ci, k, stride, dil, pad),
                 ops.gemm_operand(to_ntc_tensor(q1, False), T_out, wp1, ci1, 1), out, B, co, T_out, ep)
    torch.cuda.synchronize()
    got = from_ntc_tensor(out, co, signed=False)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


NTC_SINGLE_CASES = [
    # name, B, C_in, C_out, T, k, stride, dil, signed: C_out multiples of 256 with >= 8 slabs per tile take 256-channel tiles
    ("ntc1_k5_256_512", 2, 256, 512, 300, 5, 1, 1, False),
    ("ntc1_k3_384_768", 1, 384, 768, 260, 3, 1, 1, True),
    ("ntc1_k11_s2_128_256", 2, 128, 256, 301, 11, 2, 1, False),
    ("ntc1_k9_dil2_128_1024", 1, 128, 1024, 140, 9, 1, 2, False),
    ("ntc1_k9_128_896", 1, 128, 896, 200, 9, 1, 1, False),       # 3.5 x 256: the last 256-channel tile is half empty
    ("ntc1_k7_96_160", 2, 96, 160, 150, 7, 1, 1, True),          # 128-channel tiles either way
]


@pytest.mark.gpu
@pytest.mark.parametrize("tiles", [128, 256])
@pytest.mark.parametrize("path", ["exact", "fast", "range_check"])
@pytest.mark.parametrize("case", NTC_SINGLE_CASES, ids=lambda c: c[0])
def test_conv_ntc_single_operand(case, path, tiles):
    """conv -> ReLU -> act on the channels-last kernel (Jasper's dense convs): 128- and 256-channel tiles (bring-up knob
    21), literal / integer fast / range-checked epilogues, ragged lengths."""
    name, B, ci, co, T, k, stride, dil, signed = case
    r = rng(23)
    q = r.integers(-128, 128, (B, ci, T)) if signed else r.integers(0, 256, (B, ci, T))
    w = r.integers(-20, 20, (co, ci, k))
    bias = r.integers(-3000, 3000, co)
    pad = (dil * k) // 2 - 1 if dil > 1 else k // 2
    acc = O.conv1d_int(q, w, bias, stride, pad, dil, 1)
    T_out = acc.shape[2]
    m, e = norm_me(*rand_dyadic(r, co, 1e-5, 1e-4))
    lens = r.integers(T_out // 2, T_out + 1, B).astype(np.int32)
    want = expected_requant(acc, m, e, -256, 255, relu0=True, lens=lens)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    ch = ops.chan_table(np.full(co, 1e-4, np.float32), bias, m, e, DEV)
    flags = L.F_RELU0 | L.F_MASK | (L.F_B_SIGNED if signed else 0)
    if path == "exact":
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    elif path == "range_check":
        flags |= L.F_RANGE_CHECK
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, qmin=-256, qmax=255, out_lens=torch.from_numpy(lens).to(DEV))
    out = ops.alloc_codes_ntc(B, co, T_out, DEV)
    L.lib().qasr_debug_set(21, 0 if tiles == 128 else -1)
    try:
        ops.conv_ntc(ops.gemm_operand(to_ntc_tensor(q, signed), T, wp, ci, k, stride, dil, pad), None, out, B, co, T_out, ep)
        torch.cuda.synchronize()
    finally:
        L.lib().qasr_debug_set(21, -1)
    got = from_ntc_tensor(out, co, signed=False)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


@pytest.mark.gpu
@pytest.mark.parametrize("tiles", [128, 256])
@pytest.mark.parametrize("stage2", [False, True], ids=["plain", "stage2"])
@pytest.mark.parametrize("path", ["exact", "fast", "range_check"])
@pytest.mark.parametrize("case", NTC_SINGLE_CASES[:2] + NTC_SINGLE_CASES[4:5], ids=lambda c: c[0])
def test_conv_ntc_res_act(case, path, stage2, tiles):
    """res_act on the channels-last kernel (jasper.py:681): main dense conv + residual 1x1 conv, no previous addend;
    with 256-channel tiles the two operands' accumulators take all 512 TMEM columns (one accumulator stage)."""
    name, B, ci, co, T, k, stride, dil, signed = case
    r = rng(24)
    q = r.integers(-128, 128, (B, ci, T)) if signed else r.integers(0, 256, (B, ci, T))
    w = r.integers(-20, 20, (co, ci, k))
    bias = r.integers(-3000, 3000, co)
    pad = (dil * k) // 2 - 1 if dil > 1 else k // 2
    acc0 = O.conv1d_int(q, w, bias, stride, pad, dil, 1)
    T_out = acc0.shape[2]
    ci1 = 192
    q1 = r.integers(0, 256, (B, ci1, T_out))
    w1 = r.integers(-60, 60, (co, ci1, 1))
    acc1 = O.conv1d_int(q1, w1, None, 1, 0, 1, 1)
    m0, e0 = norm_me(*rand_dyadic(r, co, 1e-5, 1e-4))
    m1, e1 = norm_me(*rand_dyadic(r, co, 1e-5, 2e-4))
    m2, e2 = norm_me(*O.batch_frexp(np.array([2.01])))
    lens = r.integers(T_out // 2, T_out + 1, B).astype(np.int32)
    want = expected_requant(acc0, m0, e0, -128, 127, acc1=acc1, m1=m1, e1=e1, relu1=stage2,
                            stage2=(m2[0], e2[0], -256, 255) if stage2 else None, lens=lens)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    wp1 = ops.pack_weights(torch.from_numpy(w1.astype(np.int8)).to(DEV))
    one = np.full(co, 1e-4, np.float32)
    ch0, ch1 = ops.chan_table(one, bias, m0, e0, DEV), ops.chan_table(one, None, m1, e1, DEV)
    flags = L.F_MASK | (L.F_B_SIGNED if signed else 0) | ((L.F_RELU1 | L.F_STAGE2) if stage2 else 0)
    if path == "exact":
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    elif path == "range_check":
        flags |= L.F_RANGE_CHECK
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch0, flags=flags, ch1=ch1, qmin=-128, qmax=127, m2=m2[0], e2=e2[0],
                           qmin2=-256, qmax2=255, out_lens=torch.from_numpy(lens).to(DEV))
    out = ops.alloc_codes_ntc(B, co, T_out, DEV)
    L.lib().qasr_debug_set(21, 0 if tiles == 128 else -1)
    try:
        ops.conv_ntc(ops.gemm_operand(to_ntc_tensor(q, signed), T, wp, ci, k, stride, dil, pad),
                     ops.gemm_operand(to_ntc_tensor(q1, False), T_out, wp1, ci1, 1), out, B, co, T_out, ep)
        torch.cuda.synchronize()
    finally:
        L.lib().qasr_debug_set(21, -1)
    got = from_ntc_tensor(out, co, signed=not stage2)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


def test_quantize_f32_ntc():
    r = rng(21)
    B, Cn, T = 3, 40, 70
    x = (r.standard_normal((B, Cn, T)) * 2).astype(np.float32)
    lens = np.array([70, 33, 64], np.int32)
    s = O.sym_scale(8, -3.3, 2.9).reshape(())
    xm = np.where(np.arange(T)[None, None, :] < lens[:, None, None], x, 0).astype(np.float32)
    codes = ops.quantize_f32_ntc(torch.from_numpy(x).to(DEV), torch.tensor([s], device=DEV), 0, 8,
                                 lens=torch.from_numpy(lens).to(DEV))
    assert np.array_equal(from_ntc_tensor(codes, Cn, True), O.sym_quantize(xm, 8, s).astype(np.int64))
    assert int(codes[:, :, Cn:].abs().sum()) == 0
    z = r.integers(0, 256, (B, Cn, T))
    codes = ops.quantize_f32_ntc(torch.from_numpy((z * s).astype(np.float32)).to(DEV), torch.tensor([s], device=DEV), 1)
    assert np.array_equal(from_ntc_tensor(codes, Cn, False), z)


def test_gemm_dequant_f32_bigacc():
    """|acc| >= 2^24: the f32 round trip of quant_modules.py:305 rounds; dequant output bit-exact."""
    B, ci, co, T, k = 1, 7680, 128, 60, 1
    r = rng(4)
    q = r.integers(200, 256, (B, ci, T))
    w = r.integers(90, 127, (co, ci, k))
    bias = r.integers(-2 ** 30, 2 ** 30, co)
    acc = O.conv1d_int(q, w, bias, 1, k // 2, 1, 1)
    assert np.abs(acc).max() > 2 ** 27
    s_out = r.uniform(1e-9, 1e-7, co).astype(np.float32)
    want = (acc.astype(np.float32) * s_out.reshape(1, -1, 1)).astype(np.float32)
    want = np.maximum(want, 0)
    x = to_codes_tensor(q, False)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    ch = ops.chan_table(s_out, bias, np.full(co, 2 ** 30), np.full(co, 30), DEV)
    ep = ops.make_epilogue(L.EPI_DEQUANT_F32, ch, flags=L.F_RELU0 | L.F_BIAS64)
    out = torch.zeros((B, co, ops.round_up(T, 16)), dtype=torch.float32, device=DEV)
    ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, k, 1, 1, k // 2), None, out, B, co, T, ep)
    torch.cuda.synchronize()
    assert np.array_equal(out[:, :, :T].cpu().numpy(), want)
    # and the exact requant path on those accumulators (53-bit product rounding, f32 round trip)
    m, e = norm_me(*rand_dyadic(r, co, 1e-7, 1e-6))
    a32 = (acc.astype(np.float32) * s_out.reshape(1, -1, 1)).astype(np.float32)
    z = np.round((np.maximum(a32, 0) / s_out.reshape(1, -1, 1)).astype(np.float32))
    want_q = np.clip(O.dyadic_mul(z, m.reshape(1, -1, 1), e.reshape(1, -1, 1)), -256, 255).astype(np.int64)
    ch = ops.chan_table(s_out, bias, m, e, DEV)
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_RELU0 | L.F_BIAS64 | L.F_EXACT_F32 | L.F_EXACT_DYAD,
                           qmin=-256, qmax=255)
    outq = ops.alloc_codes(B, co, T, DEV)
    ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, k, 1, 1, k // 2), None, outq, B, co, T, ep)
    torch.cuda.synchronize()
    assert np.array_equal(from_codes_tensor(outq, T, False), want_q)


def _literal_requant(acc_plus_bias, s_out, m, e, lo, hi, relu):
    """The reference's recipe on accumulators of any size: float32 round trip (quant_modules.py:305-307,
    quant_utils.py:187), then the dyadic in float64 (quant_utils.py:195-196), then the clamp."""
    a32 = acc_plus_bias.astype(np.float32)
    y = (a32 * s_out.reshape(1, -1, 1)).astype(np.float32)
    if relu:
        y = np.maximum(y, np.float32(0))
    z = np.round((y / s_out.reshape(1, -1, 1)).astype(np.float32)).astype(np.float64)
    return np.clip(O.dyadic_mul(z, m.reshape(1, -1, 1), e.reshape(1, -1, 1)), lo, hi).astype(np.int64)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["literal", "fast", "f32", "range_check"])
@pytest.mark.parametrize("relu", [False, True], ids=["signed", "relu"])
def test_fast_paths_beyond_the_clamp_window(path, relu, gemm_engine):
    """8-bit weights x 8-bit activations over 512 channels reach |acc| ~ 2^24, beyond the 2^22 up to which the
    reference's float32 round trip is the identity, while the clamp window ends near |z| ~ 130 / r ~ 2^16.  The plan
    proves the fast epilogues on min(bound, plan.window_bound) and relies on every recipe being saturated beyond:
    here half of the channels run far outside the window on all three epilogues, against the literal recipe."""
    from qasr import plan as P
    B, ci, co, T = 2, 512, 256, 170
    r = rng(21)
    q = r.integers(150, 256, (B, ci, T))
    w = r.integers(-12, 13, (co, ci, 1))
    big = np.arange(co) % 2 == 1
    sign = np.where(np.arange(co) % 4 == 1, 1, -1)
    w[big] = (r.integers(80, 128, (int(big.sum()), ci, 1)) * sign[big].reshape(-1, 1, 1))
    bias = r.integers(-50000, 50000, co)
    acc = O.conv1d_int(q, w, bias, 1, 0, 1, 1)                 # includes the bias
    assert np.abs(acc).max() > 2 ** 23 and np.abs(acc[:, ~big]).max() < 2 ** 21
    lo, hi = (0, 255) if relu else (-128, 127)
    qmin = -256 if relu else -128
    bound = int(np.abs(acc).max())
    m, e = norm_me(*rand_dyadic(r, co, 1.5e-3, 6e-3))
    for _ in range(200):   # dyadics the host can prove inside the window (what the plan would require)
        bad = [i for i in range(co) if not (P.fast_ok(min(bound, P.window_bound(m[i:i + 1], e[i:i + 1], lo, hi)), m[i:i + 1], e[i:i + 1])
                                            and P.f32_ok(min(bound, P.window_bound(m[i:i + 1], e[i:i + 1], lo, hi)), m[i:i + 1], e[i:i + 1], lo, hi)
                                            and (path != "range_check" or P.fast_ok(P.F32_IDENTITY, m[i:i + 1], e[i:i + 1])))]
        if not bad:
            break
        mm, ee = norm_me(*rand_dyadic(r, len(bad), 1.5e-3, 6e-3))
        m[bad], e[bad] = mm, ee
    assert not bad
    assert P.window_bound(m, e, lo, hi) < 2 ** 18
    s_out = r.uniform(1e-5, 1e-3, co).astype(np.float32)
    want = _literal_requant(acc, s_out, m, e, lo, hi, relu)
    assert (want[:, big] == hi).any() and (want[:, big] == lo).any()     # both ends saturate
    x = to_codes_tensor(q, False)
    wp = ops.pack_weights(torch.from_numpy(w.astype(np.int8)).to(DEV))
    ch = ops.chan_table(s_out, bias, m, e, DEV)
    flags = L.F_RELU0 if relu else 0
    if path == "literal":
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    elif path == "f32":
        flags |= L.F_FAST_F32
    elif path == "range_check":   # what the plan falls back to when the window argument does not close a launch
        flags |= L.F_RANGE_CHECK
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, qmin=qmin, qmax=hi)
    out = ops.alloc_codes(B, co, T, DEV)
    ops.conv_gemm(ops.gemm_operand(x, T, wp, ci, 1, 1, 1, 0), None, out, B, co, T, ep)
    torch.cuda.synchronize()
    got = from_codes_tensor(out, T, signed=not relu)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["literal", "range_check"])
def test_dual_residual_beyond_the_float32_identity(path, gemm_engine):
    """res_act of two 512-channel 8-bit convs (jasper.py:681): the contributions are summed BEFORE the clamp, so no clamp
    window bounds the accumulators that matter, and their worst case (2^24) exceeds the 2^22 of the float32 round trip.
    QASR_F_RANGE_CHECK keeps the launch on the integer fast path (proved for |acc + bias| < 2^22) and sends the elements
    beyond through the literal recipe: half of the channels here are beyond on one or both operands."""
    from qasr import plan as P
    B, ci, co, T = 2, 512, 256, 150
    r = rng(31)
    q0, q1 = r.integers(120, 256, (B, ci, T)), r.integers(-128, 128, (B, ci, T))
    w0, w1 = r.integers(-10, 11, (co, ci, 1)), r.integers(-10, 11, (co, ci, 1))
    big0, big1 = np.arange(co) % 2 == 1, np.arange(co) % 3 == 1
    w0[big0] = r.integers(90, 128, (int(big0.sum()), ci, 1)) * np.where(np.arange(co)[big0] % 4 == 1, 1, -1).reshape(-1, 1, 1)
    w1[big1] = r.integers(-127, 128, (int(big1.sum()), ci, 1))
    q1[:, :, ::3] = np.where(w1[big1][0, :, 0] > 0, 127, -128).reshape(1, -1, 1)   # signs aligned with one channel: |acc1| > 2^22 there
    b0, b1 = r.integers(-50000, 50000, co), r.integers(-50000, 50000, co)
    acc0, acc1 = O.conv1d_int(q0, w0, b0, 1, 0, 1, 1), O.conv1d_int(q1, w1, b1, 1, 0, 1, 1)
    assert np.abs(acc0).max() > 2 ** 23 and np.abs(acc1).max() > 2 ** 22
    m0, e0 = norm_me(*rand_dyadic(r, co, 2e-6, 2e-5))
    m1, e1 = norm_me(*rand_dyadic(r, co, 2e-6, 2e-5))
    for mm, ee in ((m0, e0), (m1, e1)):
        for _ in range(200):
            bad = [i for i in range(co) if not P.fast_ok(P.F32_IDENTITY, mm[i:i + 1], ee[i:i + 1])]
            if not bad:
                break
            a, b = norm_me(*rand_dyadic(r, len(bad), 2e-6, 2e-5))
            mm[bad], ee[bad] = a, b
        assert not bad
    s0, s1 = r.uniform(1e-5, 1e-3, co).astype(np.float32), r.uniform(1e-5, 1e-3, co).astype(np.float32)

    def contribution(acc, s, m, e):
        y = (acc.astype(np.float32) * s.reshape(1, -1, 1)).astype(np.float32)
        z = np.round((y / s.reshape(1, -1, 1)).astype(np.float32)).astype(np.float64)
        return O.dyadic_mul(z, m.reshape(1, -1, 1), e.reshape(1, -1, 1))
    m2, e2 = norm_me(*O.batch_frexp(np.array([1.7])))
    o = np.clip(contribution(acc0, s0, m0, e0) + contribution(acc1, s1, m1, e1), -128, 127)
    o = np.maximum(o, 0)
    want = np.clip(O.dyadic_mul(o, m2[0], e2[0]), -256, 255).astype(np.int64)
    assert len(np.unique(want)) > 50
    x0, x1 = to_codes_tensor(q0, False), to_codes_tensor(q1, True)
    wp0 = ops.pack_weights(torch.from_numpy(w0.astype(np.int8)).to(DEV))
    wp1 = ops.pack_weights(torch.from_numpy(w1.astype(np.int8)).to(DEV))
    ch0, ch1 = ops.chan_table(s0, b0, m0, e0, DEV), ops.chan_table(s1, b1, m1, e1, DEV)
    flags = L.F_RELU1 | L.F_STAGE2 | L.F_B1_SIGNED
    flags |= (L.F_EXACT_F32 | L.F_EXACT_DYAD) if path == "literal" else L.F_RANGE_CHECK
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch0, flags=flags, ch1=ch1, qmin=-128, qmax=127, m2=m2[0], e2=e2[0],
                           qmin2=-256, qmax2=255)
    out = ops.alloc_codes(B, co, T, DEV)
    ops.conv_gemm(ops.gemm_operand(x0, T, wp0, ci, 1, 1, 1, 0), ops.gemm_operand(x1, T, wp1, ci, 1, 1, 1, 0), out, B, co, T, ep)
    torch.cuda.synchronize()
    got = from_codes_tensor(out, T, signed=False)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["literal", "fast", "f32"])
def test_dwconv_fast_paths_beyond_the_clamp_window(path, dw_engine):
    """Same argument on the depthwise engines: integer biases push acc + bias to ~2^24 on half of the channels."""
    from qasr import plan as P
    B, Cn, T, k = 3, 64, 300, 33
    r = rng(22)
    q = r.integers(0, 256, (B, Cn, T))
    w = r.integers(-20, 21, (Cn, 1, k))
    big = np.arange(Cn) % 2 == 1
    bias = r.integers(-40000, 40000, Cn)
    bias[big] = r.integers(2 ** 23, 2 ** 24, int(big.sum())) * np.where(np.arange(Cn)[big] % 4 == 1, 1, -1)
    acc = O.conv1d_int(q, w, bias, 1, k // 2, 1, Cn)
    lo, hi = -128, 127
    bound = int(np.abs(acc).max())
    assert bound > 2 ** 23
    m, e = norm_me(*rand_dyadic(r, Cn, 1.5e-3, 6e-3))
    for _ in range(200):
        bad = [i for i in range(Cn) if not (P.fast_ok(min(bound, P.window_bound(m[i:i + 1], e[i:i + 1], lo, hi)), m[i:i + 1], e[i:i + 1])
                                            and P.f32_ok(min(bound, P.window_bound(m[i:i + 1], e[i:i + 1], lo, hi)), m[i:i + 1], e[i:i + 1], lo, hi))]
        if not bad:
            break
        mm, ee = norm_me(*rand_dyadic(r, len(bad), 1.5e-3, 6e-3))
        m[bad], e[bad] = mm, ee
    assert not bad
    s_out = r.uniform(1e-5, 1e-3, Cn).astype(np.float32)
    want = _literal_requant(acc, s_out, m, e, lo, hi, False)
    x = to_codes_tensor(q, False)
    wd = ops.pack_dw_weights(torch.from_numpy(w[:, 0, :].astype(np.int8)).to(DEV).contiguous(), 1, 1, k // 2)
    ch = ops.chan_table(s_out, bias, m, e, DEV)
    flags = 0
    if path == "literal":
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    elif path == "f32":
        flags |= L.F_FAST_F32
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, qmin=lo, qmax=hi)
    out = ops.alloc_codes(B, Cn, T, DEV)
    ops.dwconv(x, T, wd, out, T, k, 1, 1, k // 2, ep)
    torch.cuda.synchronize()
    got = from_codes_tensor(out, T, signed=True)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)


# ------------------------------------------------------------------------------------------
# K1 depthwise
# ------------------------------------------------------------------------------------------
DW_CASES = [
    # name, B, C, T, k, stride, dil, signed
    ("k33_s2_signed", 2, 64, 500, 33, 2, 1, True),
    ("k33", 2, 256, 250, 33, 1, 1, False),
    ("k39", 1, 256, 300, 39, 1, 1, False),
    ("k51", 2, 96, 180, 51, 1, 1, False),
    ("k63", 1, 128, 333, 63, 1, 1, False),
    ("k75", 2, 64, 250, 75, 1, 1, False),
    ("k87_dil2", 1, 64, 400, 87, 1, 2, False),
    ("k99_4steps", 2, 32, 300, 99, 1, 1, False),      # 4 / 5 K steps of the IMMA window (permuted 8-byte fragment loads)
    ("k121_5steps", 2, 32, 300, 121, 1, 1, True),
    ("k9_tiny", 3, 32, 32, 9, 1, 1, False),
    ("k9_s2_tiny", 3, 64, 64, 9, 2, 1, True),
    ("k13_dil2_tiny", 2, 32, 40, 13, 1, 2, False),
    ("k5_T17", 2, 8, 17, 5, 1, 1, True),
    # batches that fill the tensor-core kernel's row tiles (several utterance chunks / time-block groups)
    ("k33_b20", 20, 24, 330, 33, 1, 1, False),
    ("k75_b70", 70, 8, 200, 75, 1, 1, True),
    ("k87_dil2_b16", 16, 16, 420, 87, 1, 2, False),
    ("k11_s2_b33", 33, 16, 321, 11, 2, 1, False),
    ("k33_b130", 130, 4, 170, 33, 1, 1, False),
]


@pytest.fixture(params=[0, 1, 2, 3, 4], ids=["imma", "tcgen05", "tcgen05_rows", "tape", "imma_rows"])
def dw_engine(request):
    """Forces one of the depthwise engines (bring-up knobs 8 / 11) so all are checked on every case: IMMA,
    tcgen05 with TMA-fetched windows, tcgen05 with bulk-copied rows cut into windows in shared memory, and the tape
    engine (tcgen05 on the utterance rows themselves through overlapping descriptors; strided cases fall through to
    the default engine).  "imma" walks row tapes on the fast epilogues (dwconv_rt.cu), "imma_rows" keeps every launch on
    the row-at-a-time kernel (bring-up knob 16 = 0)."""
    L.lib().qasr_debug_set(8, 2 if request.param == 3 else (0 if request.param == 4 else min(request.param, 1)))
    L.lib().qasr_debug_set(11, 1 if request.param == 2 else -1)
    L.lib().qasr_debug_set(16, 0 if request.param == 4 else -1)
    yield request.param
    L.lib().qasr_debug_set(8, -1)
    L.lib().qasr_debug_set(11, -1)
    L.lib().qasr_debug_set(16, -1)


@pytest.mark.parametrize("fast", [False, True, "f32"], ids=["exact", "fast", "f32"])
@pytest.mark.parametrize("case", DW_CASES, ids=lambda c: c[0])
def test_dwconv_requant(case, fast, dw_engine):
    name, B, Cn, T, k, stride, dil, signed = case
    r = rng(5)
    q = r.integers(-128, 128, (B, Cn, T)) if signed else r.integers(0, 256, (B, Cn, T))
    w = r.integers(-127, 127, (Cn, 1, k))
    pad = (dil * k) // 2 - 1 if dil > 1 else k // 2
    acc = O.conv1d_int(q, w, None, stride, pad, dil, Cn)
    T_out = acc.shape[2]
    if fast == "f32":
        m, e = f32_dyadic(r, Cn, int(np.abs(acc).max()), -128, 127, 1e-4, 2e-3)
    else:
        m, e = norm_me(*rand_dyadic(r, Cn, 1e-4, 2e-3))
    lens = r.integers(T_out // 2, T_out + 1, B).astype(np.int32)
    want = expected_requant(acc, m, e, -128, 127, lens=lens)
    x = to_codes_tensor(q, signed)
    wd = ops.pack_dw_weights(torch.from_numpy(w[:, 0, :].astype(np.int8)).to(DEV).contiguous(), stride, dil, pad)
    ch = ops.chan_table(r.uniform(1e-5, 1e-3, Cn).astype(np.float32), None, m, e, DEV)
    flags = L.F_MASK | (L.F_B_SIGNED if signed else 0)
    if not fast:
        flags |= L.F_EXACT_F32 | L.F_EXACT_DYAD
    if fast == "f32":
        flags |= L.F_FAST_F32
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, qmin=-128, qmax=127,
                           out_lens=torch.from_numpy(lens).to(DEV))
    out = ops.alloc_codes(B, Cn, T_out, DEV)
    ops.dwconv(x, T, wd, out, T_out, k, stride, dil, pad, ep)
    torch.cuda.synchronize()
    got = from_codes_tensor(out, T_out, signed=True)
    assert np.array_equal(got, want), "mismatches: %d / %d" % ((got != want).sum(), want.size)
    assert int(out[:, :, T_out:].abs().sum()) == 0


MERGED_TAIL_CASES = [
    # several utterances per warp (forced with bring-up knob 10) and a T_out that leaves <= 32 frames after the whole
    # 128-frame tiles: the row-ring / last-tile mask path of the IMMA kernel (and the case the merged-row-tail
    # experiment, tools/experiments/dwconv_merged_row_tails.patch, was written for)
    ("k33_T800_b10", 10, 64, 800, 33, 1, 1, False),
    ("k63_T288_b7", 7, 32, 288, 63, 1, 1, True),
    ("k75_T150_b9", 9, 48, 150, 75, 1, 1, False),
    ("k39_T20_b5", 5, 40, 20, 39, 1, 1, False),
]


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [False, True, "f32"], ids=["exact", "fast", "f32"])
@pytest.mark.parametrize("case", MERGED_TAIL_CASES, ids=lambda c: c[0])
def test_dwconv_merged_row_tails(case, fast):
    lib = L.lib()
    lib.qasr_debug_set(8, 0)     # IMMA engine
    lib.qasr_debug_set(10, 1)    # one warp per SM wanted -> several utterances per warp
    try:
        test_dwconv_requant(case, fast, 0)
    finally:
        lib.qasr_debug_set(8, -1)
        lib.qasr_debug_set(10, -1)


ROW_TAPE_CASES = [
    # the row-tape walk of the IMMA engine (dwconv_rt.cu) with R rows per tape forced (bring-up knob 17): ragged lengths,
    # batches that are not a multiple of R, rows of 6.25 / 2.6 / 1.2 tiles, gaps of 16 .. 64 bytes, 2 .. 5 K steps
    ("k33_T800_b10", 10, 24, 800, 33, 1, 1, False),
    ("k63_T333_b7", 7, 16, 333, 63, 1, 1, True),
    ("k75_T150_b9", 9, 16, 150, 75, 1, 1, False),
    ("k87_dil2_T420_b5", 5, 16, 420, 87, 1, 2, False),
    ("k121_T300_b6", 6, 8, 300, 121, 1, 1, True),
    ("k9_T113_b13", 13, 8, 113, 9, 1, 1, False),
]


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [2, 3, 4, 8])
@pytest.mark.parametrize("fast", [True, "f32"], ids=["fast", "f32"])
@pytest.mark.parametrize("case", ROW_TAPE_CASES, ids=lambda c: c[0])
def test_dwconv_row_tapes(case, fast, rows):
    lib = L.lib()
    lib.qasr_debug_set(8, 0)
    lib.qasr_debug_set(17, rows)
    try:
        test_dwconv_requant(case, fast, 0)
    finally:
        lib.qasr_debug_set(8, -1)
        lib.qasr_debug_set(17, -1)


def test_dwconv_raw_and_dequant(dw_engine):
    B, Cn, T, k = 2, 48, 100, 33
    r = rng(6)
    q = r.integers(0, 256, (B, Cn, T))
    w = r.integers(-127, 127, (Cn, 1, k))
    acc = O.conv1d_int(q, w, None, 1, k // 2, 1, Cn)
    x = to_codes_tensor(q, False)
    wd = ops.pack_dw_weights(torch.from_numpy(w[:, 0, :].astype(np.int8)).to(DEV).contiguous(), 1, 1, k // 2)
    s_out = r.uniform(1e-5, 1e-3, Cn).astype(np.float32)
    ch = ops.chan_table(s_out, None, np.full(Cn, 2 ** 30), np.full(Cn, 30), DEV)
    out = torch.zeros((B, Cn, ops.round_up(T, 16)), dtype=torch.int32, device=DEV)
    ops.dwconv(x, T, wd, out, T, k, 1, 1, k // 2, ops.make_epilogue(L.EPI_RAW_I32, ch))
    outf = torch.zeros((B, Cn, ops.round_up(T, 16)), dtype=torch.float32, device=DEV)
    ops.dwconv(x, T, wd, outf, T, k, 1, 1, k // 2, ops.make_epilogue(L.EPI_DEQUANT_F32, ch))
    torch.cuda.synchronize()
    assert np.array_equal(out[:, :, :T].cpu().numpy(), acc)
    assert np.array_equal(outf[:, :, :T].cpu().numpy(), (acc.astype(np.float32) * s_out.reshape(1, -1, 1)))


# ------------------------------------------------------------------------------------------
# K3 / K4 / K5
# ------------------------------------------------------------------------------------------
def test_dyadic_from_scales():
    r = rng(7)
    pre = np.concatenate([r.uniform(1e-7, 1e-2, 500), [2.0 ** -6, 2.0 ** -3, 1e-3]]).astype(np.float32)
    for zs in (np.float32(0.0123), np.float32(2.0 ** -3), np.float32(1e-3)):
        m, e = ops.dyadic_from_scales(torch.from_numpy(pre).to(DEV), torch.tensor([zs], device=DEV))
        wm, we = norm_me(*O.dyadic(pre, zs))
        assert np.array_equal(m.cpu().numpy(), wm) and np.array_equal(e.cpu().numpy(), we)


def test_quantize_and_requant_f32():
    r = rng(8)
    B, Cn, T = 3, 20, 37
    x = (r.standard_normal((B, Cn, T)) * 2).astype(np.float32)
    lens = np.array([37, 20, 5], np.int32)
    for bit in (8, 6):
        s = O.sym_scale(bit, -3.3, 2.9).reshape(())
        xm = np.where(np.arange(T)[None, None, :] < lens[:, None, None], x, 0).astype(np.float32)
        want = O.sym_quantize(xm, bit, s)
        codes, xq = ops.quantize_f32(torch.from_numpy(x).to(DEV), torch.tensor([s], device=DEV), bit,
                                     lens=torch.from_numpy(lens).to(DEV))
        assert np.array_equal(codes[:, :, :T].cpu().numpy(), want)
        assert np.array_equal(xq.cpu().numpy(), (want * s).astype(np.float32))
    # requant with per-channel scales + identity
    pre = r.uniform(1e-5, 2e-4, Cn).astype(np.float32)
    ids = r.uniform(1e-5, 2e-4, Cn).astype(np.float32)
    z = r.integers(-30000, 30000, (B, Cn, T))
    zi = r.integers(-30000, 30000, (B, Cn, T))
    xf = (z * pre.reshape(1, -1, 1)).astype(np.float32)
    idf = (zi * ids.reshape(1, -1, 1)).astype(np.float32)
    zs = np.float32(0.031)
    want = O.fixedpoint_mul(xf, pre.reshape(1, -1, 1), 8, zs, idf, ids.reshape(1, -1, 1))
    codes, xq = ops.requant_f32(torch.from_numpy(xf).to(DEV), torch.from_numpy(pre).to(DEV),
                                torch.tensor([zs], device=DEV), -128, 127,
                                identity=torch.from_numpy(idf).to(DEV), id_sf=torch.from_numpy(ids).to(DEV))
    assert np.array_equal(codes[:, :, :T].cpu().numpy(), want)
    assert np.array_equal(xq.cpu().numpy(), (want * zs).astype(np.float32))
    back = ops.codes_from_f32(xq, torch.tensor([zs], device=DEV))
    assert np.array_equal(back[:, :, :T].cpu().numpy(), want)


def test_requant_i8_lut():
    r = rng(9)
    B, Cn, T = 2, 16, 50
    q = r.integers(0, 128, (B, Cn, T))
    m, e = norm_me(*O.batch_frexp(np.array([2.0078])))
    lens = np.array([50, 31], np.int32)
    want = expected_requant(q, np.repeat(m, Cn), np.repeat(e, Cn), -256, 255, lens=lens)
    out = ops.requant_i8(to_codes_tensor(q, True), T, True, m[0], e[0], -256, 255, torch.from_numpy(lens).to(DEV))
    assert np.array_equal(from_codes_tensor(out, T, False), want)


def test_minmax_and_quantile():
    r = rng(10)
    B, Cn, T = 4, 64, 200
    x = r.standard_normal((B, Cn, T)).astype(np.float32)
    idn = r.standard_normal((B, Cn, T)).astype(np.float32)
    lens = np.array([200, 150, 90, 200], np.int32)
    xd, idd, ld = torch.from_numpy(x).to(DEV), torch.from_numpy(idn).to(DEV), torch.from_numpy(lens).to(DEV)
    mm = ops.minmax_f32(xd).cpu().numpy()
    assert mm[0] == x.min() and mm[1] == x.max()
    mm = ops.minmax_f32(xd, identity=idd).cpu().numpy()
    s = (idn + x).astype(np.float32)
    assert mm[0] == s.min() and mm[1] == s.max()
    xm = np.where(np.arange(T)[None, None, :] < lens[:, None, None], np.abs(x) + 1, 0).astype(np.float32)
    mm = ops.minmax_f32(torch.from_numpy(np.abs(x) + 1).to(DEV), lens=ld).cpu().numpy()
    assert mm[0] == 0.0 and mm[1] == xm.max()
    for p in (99.0, 99.9, 99.996):
        qlo, qhi = np.float32(1 - p / 100), np.float32(p / 100)
        got = ops.quantile2_f32(xd, qlo, qhi).cpu().numpy()
        assert got[0] == O.torch_quantile(x, qlo) and got[1] == O.torch_quantile(x, qhi)
        got = ops.quantile2_f32(xd, qlo, qhi, identity=idd).cpu().numpy()
        assert got[0] == O.torch_quantile(s, qlo) and got[1] == O.torch_quantile(s, qhi)
    xr = np.maximum(x, 0)  # many ties at 0 (post-ReLU tensors)
    got = ops.quantile2_f32(torch.from_numpy(xr).to(DEV), np.float32(0.001), np.float32(0.999)).cpu().numpy()
    assert got[0] == O.torch_quantile(xr, np.float32(0.001)) and got[1] == O.torch_quantile(xr, np.float32(0.999))


def test_decoder_tail():
    r = rng(11)
    B, V, T = 3, 29, 45
    logits = r.standard_normal((B, V, T)).astype(np.float32)
    logits[0, 5, 3] = logits[0, 9, 3] = 7.0  # tie: first index wins
    t = torch.zeros((B, V, 48), dtype=torch.float32, device=DEV)
    t[:, :, :T] = torch.from_numpy(logits).to(DEV)
    labels = ops.ctc_argmax(t, T).cpu().numpy()
    assert np.array_equal(labels, logits.argmax(axis=1))
    lp = ops.log_softmax_tv(t, T).cpu()
    want = torch.log_softmax(torch.from_numpy(logits).transpose(1, 2), dim=-1)
    assert torch.allclose(lp, want, atol=2e-6, rtol=0)


def test_error_paths():
    x = ops.alloc_codes(1, 8, 32, DEV)
    ch = ops.chan_table(np.ones(8, np.float32), None, np.full(8, 2 ** 30), np.full(8, 30), DEV)
    with pytest.raises(ValueError):
        ops.dwconv(x, 32, torch.zeros(8 * 128 * 4, dtype=torch.int8, device=DEV), x, 32, 5, 3, 1, 2,
                   ops.make_epilogue(L.EPI_REQUANT_I8, ch))
    with pytest.raises(ValueError):
        ops.quantile2_f32(torch.zeros((1, 1, 8), device=DEV), 1.5, 0.5)


def _sweep_descriptors():
    """Bring-up aid: try B-operand descriptor stride variants on the first GEMM case."""
    lib = L.lib()
    for lbo, sbo in ((16384, 1024), (1024, 16384), (128, 1024), (1024, 128), (0, 1024), (2048, 1024),
                     (4096, 1024), (1024, 1024), (16, 1024), (1024, 2048)):
        lib.qasr_debug_set(0, lbo)
        lib.qasr_debug_set(1, sbo)
        try:
            test_gemm_raw_i32(GEMM_CASES[0])
            print("SWEEP lbo=%d sbo=%d: MATCH" % (lbo, sbo), flush=True)
        except AssertionError as ex:
            print("SWEEP lbo=%d sbo=%d: %s" % (lbo, sbo, str(ex)[:80]), flush=True)
    lib.qasr_debug_set(0, -1)
    lib.qasr_debug_set(1, -1)


if __name__ == "__main__":
    # non-stopping report (first bring-up on a GPU box)
    import traceback
    if os.environ.get("QASR_SWEEP"):
        _sweep_descriptors()
    fails = 0
    tests = []
    for name, fn in list(globals().items()):
        if name.startswith("test_") and callable(fn):
            marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
            if not marks:
                tests.append((name, fn, {}))
                continue
            import itertools
            axes = []
            for m in marks:
                argn, vals = m.args[0], m.args[1]
                axes.append([(argn, v) for v in vals])
            for combo in itertools.product(*axes):
                tests.append((name + str([v[0] if isinstance(v, tuple) else v for _, v in combo]), fn, dict(combo)))
    for name, fn, kw in tests:
        try:
            fn(**kw)
            print("PASS", name, flush=True)
        except Exception as ex:  # noqa: BLE001
            fails += 1
            print("FAIL", name, "->", str(ex).splitlines()[0][:200] if str(ex) else repr(ex), flush=True)
            if os.environ.get("QASR_TRACE"):
                traceback.print_exc()
            if "CUDA" in str(ex) or "qasr error" in str(ex):
                print("fatal CUDA error, stopping")
                break
    print("failures:", fails, "of", len(tests))
    sys.exit(1 if fails else 0)

"""CPU: the host-side arguments that let the plan replace the reference's literal requantisation recipe by the integer /
float fast epilogues (qasr/plan.py: fast_ok, f32_ok_channels, window_bound, fast_ok_checked), checked EXHAUSTIVELY:
every accumulator of [-2^24, 2^24] goes through a numpy restatement of the literal recipe (float32 round trip
quant_modules.py:305-307 / quant_utils.py:187, float64 dyadic quant_utils.py:195-196, clamp :213-214) and through the
arithmetic the kernels execute (csrc/qasr_common.cuh contrib_fast / epilogue_f32_bits)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import qasr_oracle as O  # noqa: E402
from qasr import plan as P  # noqa: E402

F32, F64 = np.float32, np.float64
MAGIC_BITS = 0x4B400000
MAGIC = F32(12582912.0)


def literal(z, s, m, e, lo, hi, relu):
    a32 = z.astype(F32)
    y = (a32 * F32(s)).astype(F32)
    if relu:
        y = np.maximum(y, F32(0))
    zz = np.round((y / F32(s)).astype(F32)).astype(F64)
    return np.clip(np.round(zz * F64(m) / F64(2.0 ** e)), lo, hi).astype(np.int64)


def kernel_integer(z, m, e, lo, hi):
    """contrib_fast + clamp: ((z * m + 2^(e-1)) >> 32) >> (e - 32), 32 <= e <= 62."""
    p = z * np.int64(m) + (np.int64(1) << np.int64(e - 1))
    return np.clip((p >> np.int64(32)) >> np.int64(e - 32), lo, hi)


def kernel_float(z, m, e, lo, hi, mult=None):
    """epilogue_f32_bits: bits = 0x4B400000 + z reinterpreted as float32, minus 1.5 * 2^23, FMA with f32(m 2^-e) and the
    magic, float clamp, low byte.  The FMA is evaluated exactly here: f and r are float32 values, so f * r is exact in
    float64, and rounding (f * r + magic) to float32 = rounding f * r to the grid of spacing ulp(result)."""
    r32 = F64(F32(F64(m) * 2.0 ** -e)) if mult is None else F64(F32(mult))
    bits = (np.int64(MAGIC_BITS) + z).astype(np.int32)
    f = (bits.view(F32) - MAGIC).astype(F64)          # FADD: exact (both operands on one binade grid or coarser)
    t = f * r32                                        # exact in float64
    y = np.clip(np.rint(t), -2.0 ** 22, 2.0 ** 22 - 1)  # |t| < 2^22: magic + t stays in the binade with ulp 1 -> RNE to integer
    # beyond +-2^22 the sum leaves the binade; the float clamp that follows maps it to lo / hi all the same
    y = np.where(t >= 2.0 ** 22 - 1, F64(hi), np.where(t <= -2.0 ** 22, F64(lo), y))
    return np.clip(y, lo, hi).astype(np.int64)


CASES = [
    # ratio r = m / 2^e, clamp window, ReLU
    (2.1e-3, -128, 127, False),
    (4.4e-3, 0, 255, True),
    (9.0e-4, 0, 255, True),
    (1.3e-2, -128, 127, False),
    (6.0e-5, -32, 31, False),
    (3.1e-4, 0, 63, True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "r%.0e_%d_%d" % (c[0], c[1], c[2]))
def test_fast_paths_equal_the_literal_recipe_on_all_accumulators(case):
    ratio, lo, hi, relu = case
    rng = np.random.default_rng(int(ratio * 1e7))
    bound = 2 ** 24                      # worst case of 512 channels x 127 x 255
    found = None
    for _ in range(400):                 # a dyadic the host accepts (as the plan does per channel)
        m, e = O.batch_frexp(np.array([ratio * rng.uniform(0.97, 1.03)]))
        m, e = np.asarray(m, np.int64), np.asarray(e, np.int64)
        if m[0] >= 2 ** 31:
            m, e = m // 2, e - 1
        eff = min(bound, P.window_bound(m, e, lo, hi))
        if eff < 2 ** 22 and P.fast_ok(eff, m, e) and P.f32_ok_channels(eff, m, e, lo, hi)[0]:
            found = (int(m[0]), int(e[0]), eff)
            break
    assert found, "no provable dyadic near %g" % ratio
    m, e, eff = found
    assert not P.fast_ok(bound, np.array([m]), np.array([e]))          # the worst-case bound alone would refuse it
    s = F32(rng.uniform(1e-5, 1e-3))
    step = 1 << 22
    for start in range(-bound, bound + 1, step):
        z = np.arange(start, min(start + step, bound + 1), dtype=np.int64)
        want = literal(z, s, m, e, lo, hi, relu)
        lo_k = max(lo, 0) if relu else lo                                # ReLU folded into the clamp (make_epi_consts)
        assert np.array_equal(kernel_integer(z, m, e, lo_k, hi), want), "integer path, z from %d" % start
        assert np.array_equal(kernel_float(z, m, e, lo_k, hi), want), "float path, z from %d" % start


def test_window_bound_saturates_every_channel():
    rng = np.random.default_rng(5)
    m, e = O.batch_frexp(rng.uniform(2e-5, 2e-2, 64))
    m, e = np.asarray(m, np.int64), np.asarray(e, np.int64)
    big = m >= 2 ** 31
    m, e = np.where(big, m // 2, m), np.where(big, e - 1, e)
    for lo, hi in ((-128, 127), (0, 255), (-32, 31)):
        Z = P.window_bound(m, e, lo, hi)
        code = (np.int64(Z) * m + (np.int64(1) << (e - 1))) >> e
        assert (code >= max(abs(lo), abs(hi)) + 1).all()
        assert (((np.int64(-Z) * m + (np.int64(1) << (e - 1))) >> e) <= -(max(abs(lo), abs(hi)) + 1)).all()


def test_fast_ok_checked_asks_for_the_range_check_beyond_2_22():
    m, e = O.batch_frexp(np.array([1.7e-5]))
    m, e = np.asarray(m, np.int64), np.asarray(e, np.int64)
    for _ in range(50):
        if P.fast_ok(P.F32_IDENTITY, m, e):
            break
        m = m + 2
    assert P.fast_ok_checked(2 ** 20, m, e) == (True, False)
    assert P.fast_ok_checked(2 ** 24, m, e) == (True, True)
    assert P.fast_ok_checked(2 ** 31, m, e) == (False, False)


def test_neighbouring_float_multipliers_are_proved_too():
    """f32_multipliers(): where f32(m 2^-e) fails a jump, the neighbouring float may satisfy all of them; whatever it hands
    to the kernel must equal the literal recipe on every accumulator."""
    rng = np.random.default_rng(11)
    m, e = O.batch_frexp(rng.uniform(5e-4, 8e-3, 512))
    m, e = np.asarray(m, np.int64), np.asarray(e, np.int64)
    big = m >= 2 ** 31
    m, e = np.where(big, m // 2, m), np.where(big, e - 1, e)
    lo, hi = 0, 255
    eff = min(2 ** 24, P.window_bound(m, e, lo, hi))
    ok = np.array([P.fast_ok(eff, m[i:i + 1], e[i:i + 1]) for i in range(m.size)])
    canon = P.f32_ok_channels(eff, m, e, lo, hi)
    rf = P.f32_multipliers(eff, m, e, lo, hi)
    rescued = np.nonzero(ok & ~canon & (rf != 0))[0]
    assert rescued.size >= 10 and (rf != 0).sum() > canon.sum()
    r32 = (m.astype(F64) * np.exp2(-e.astype(F64))).astype(F32)
    assert (rf[rescued] != r32[rescued]).all() and np.array_equal(rf[canon], r32[canon])
    s = F32(3e-4)
    for i in rescued[:3]:
        for start in range(-2 ** 24, 2 ** 24 + 1, 1 << 22):
            z = np.arange(start, min(start + (1 << 22), 2 ** 24 + 1), dtype=np.int64)
            want = literal(z, s, int(m[i]), int(e[i]), lo, hi, True)
            assert np.array_equal(kernel_float(z, int(m[i]), int(e[i]), lo, hi, mult=rf[i]), want)


def kernel_float_clamp16(z, m, e, hi, mult=None):
    """epilogue_f32_pack16 (QASR_F_CLAMP16): the FMA result's bit pattern, its LOW 16 BITS read as an int16, then
    max(min(x, hi), 0) on the 16-bit lane (VIMNMX.S16x2.RELU), low byte."""
    r32 = F32(F64(m) * 2.0 ** -e) if mult is None else F32(mult)
    bits = (np.int64(MAGIC_BITS) + z).astype(np.int32)
    f = (bits.view(F32) - MAGIC).astype(F64)
    y = (f * F64(r32) + F64(MAGIC)).astype(F32)          # one rounding: the product is exact in float64, |sum| < 2^53
    low = (y.view(np.int32).astype(np.int64) & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int64)
    return np.clip(np.minimum(low, hi), 0, None) & 0xFF


@pytest.mark.parametrize("case", [(4.4e-3, 255, 2 ** 22 - 1), (3.1e-4, 63, 2 ** 22 - 1), (7.0e-3, 255, 4_000_000), (1.3e-2, 255, 4_000_000)],
                         ids=lambda c: "r%.0e_%d" % (c[0], c[1]))
def test_clamp16_equals_the_literal_recipe_up_to_the_worst_case_bound(case):
    """clamp16_ok(): the 16-bit clamp needs |rint(z r)| < 2^15 over the WHOLE reachable range (no window argument); where
    it says yes the kernel arithmetic equals the literal recipe on every accumulator up to the bound, where the product
    passes 2^15 it must say no (and the wrapped lanes indeed differ)."""
    ratio, hi, bound = case
    rng = np.random.default_rng(5)
    m = e = rf = None
    for _ in range(200):
        mm, ee = O.batch_frexp(np.array([ratio * rng.uniform(0.9, 1.1)]))
        mm, ee = int(np.asarray(mm).reshape(-1)[0]), int(np.asarray(ee).reshape(-1)[0])
        if mm >= 2 ** 31:
            mm, ee = mm // 2, ee - 1
        ma, ea = np.array([mm], np.int64), np.array([ee], np.int64)
        eff = min(bound, P.window_bound(ma, ea, 0, hi))
        if not P.fast_ok(eff, ma, ea):
            continue
        r = P.f32_multipliers(eff, ma, ea, 0, hi)
        if r[0] != 0:
            m, e, rf = mm, ee, r[0]
            break
    assert m is not None
    ok = P.clamp16_ok(np.array([bound]), np.array([rf]))
    assert ok == (bound * float(rf) + 1.0 < 2 ** 15 - 1)
    s = F32(3e-4)
    equal = True
    for start in range(-bound, bound + 1, 1 << 21):
        z = np.arange(start, min(start + (1 << 21), bound + 1), dtype=np.int64)
        want = literal(z, s, m, e, 0, hi, True) & 0xFF
        equal &= bool(np.array_equal(kernel_float_clamp16(z, m, e, hi, mult=rf), want))
    assert equal == ok
    assert not P.clamp16_ok(np.array([2 ** 22]), np.array([1e-6]))      # f32(z) = z needs |z| < 2^22


def test_stage2_16_equals_the_literal_second_requantisation():
    """stage2_16_ok(): whenever it says yes, the kernel's 16-bit fixed-point stage ((q * M16 + 2^15) >> 16, upper clamp) equals
    the reference's clamp(rne(q * m2 / 2^e2)) (quant_utils.py:195-214, float64) on every code q of [0, hi1]; it refuses
    ratios whose rounding it cannot reproduce, wide codes and exponents below 17."""
    rng = np.random.default_rng(8)
    yes = no = 0
    for hi1, hi2 in ((31, 31), (127, 127), (127, 255), (63, 63)):
        for _ in range(300):
            m2, e2 = O.batch_frexp(np.array([rng.uniform(0.1, 4.0)]))
            m2, e2 = int(np.asarray(m2).reshape(-1)[0]), int(np.asarray(e2).reshape(-1)[0])
            if m2 >= 2 ** 31:
                m2, e2 = m2 // 2, e2 - 1
            if not P.stage2_fast_ok(hi1 + 1, np.array([m2]), np.array([e2])):
                continue
            q = np.arange(hi1 + 1, dtype=np.int64)
            want = np.clip(np.round(q.astype(F64) * F64(m2) / F64(2.0 ** e2)), -hi2 - 1, hi2).astype(np.int64)
            m16 = (m2 + (1 << (e2 - 17))) >> (e2 - 16)
            got = np.minimum((q * m16 + 0x8000) >> 16, hi2)
            if P.stage2_16_ok(hi1, m2, e2, -hi2 - 1, hi2):
                yes += 1
                assert np.array_equal(got, want)
            else:
                no += 1
                assert not np.array_equal((q * m16 + 0x8000) >> 16, np.round(q.astype(F64) * F64(m2) / F64(2.0 ** e2)).astype(np.int64))
    assert yes > 500 and no > 0
    assert not P.stage2_16_ok(255, 1 << 30, 31, -256, 255)       # codes wider than 7 bits
    assert not P.stage2_16_ok(127, 1 << 15, 16, -128, 127)       # e2 < 17
    assert not P.stage2_16_ok(127, 1 << 30, 31, 1, 127)          # lower clamp above 0

"""CPU, build container only: the oracle against the UNMODIFIED reference run live at full QuartzNet15x5
size (skipped where /root/reference is absent, e.g. on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QASR_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def run(*args):
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "pin_live.py")] + [str(a) for a in args],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_quartznet15x5_static_integer_oracle_is_bit_exact_vs_reference():
    """Static W8A8 (min/max calibration): the noise-free integer recipe reproduces every code."""
    r = run("quartznet15x5", 2, 200, 8, 8, "static", 0)
    assert r["codes"] > 15_000_000
    assert r["code_mismatches"] == 0 and r["scale_mismatches"] == 0 and r["logit_code_mismatches"] == 0


def test_quartznet15x5_dynamic_needs_the_float_noise():
    """Dynamic W4A8: ranges are taken on floats that carry the reference's un-rounded x/s noise.  The
    literal oracle mode (what the module-level CUDA path reproduces) is bit-exact; the noise-free recipe
    drifts (1-ulp scale differences cascade), which is why the module-level path emulates the noise."""
    r = run("quartznet15x5", 2, 200, 4, 8, "dynamic", 1)
    assert r["code_mismatches"] == 0 and r["scale_mismatches"] == 0 and r["logit_code_mismatches"] == 0


def test_tiny_jasper_literal_oracle_is_bit_exact_vs_reference():
    r = run("tiny_jasper", 3, 64, 8, 8, "static", 1)
    assert r["code_mismatches"] == 0 and r["scale_mismatches"] == 0 and r["logit_code_mismatches"] == 0

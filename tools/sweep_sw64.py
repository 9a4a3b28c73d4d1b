"""Bring-up aid: B-operand descriptor strides (knobs 0 / 1) of the CTA-pair kernel's 64-frame half-box mode
(64-byte swizzle), checked on the two-operand residual GEMM test."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]
import test_gpu_kernels as T
from qasr import _lib as L
lib = L.lib()
lib.qasr_debug_set(12, 1)
for lbo, sbo in ((8192, 512), (512, 8192), (64, 512), (512, 64), (0, 512), (1024, 512), (4096, 512), (8192, 1024),
                 (512, 512), (16384, 512), (2048, 512)):
    lib.qasr_debug_set(0, lbo)
    lib.qasr_debug_set(1, sbo)
    try:
        T.test_gemm_dual_residual(True, 1)
        print("SWEEP lbo=%d sbo=%d: MATCH" % (lbo, sbo), flush=True)
    except AssertionError as ex:
        print("SWEEP lbo=%d sbo=%d: %s" % (lbo, sbo, str(ex)[:80]), flush=True)

"""A few launches of one pointwise layer (512 x 512, batch 64 x 800 frames, float epilogue): the target of
`ncu --set full -k regex:gemm_i8` captures.  Knobs as k=v arguments (12=1 forces the CTA-pair kernel); DUAL=1 in the
environment makes it the two-operand res_act launch (block ReLU + second requantisation, QASR_F_STAGE2_16)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "q-asr_b200")]
import numpy as np, torch
from qasr import _lib as L, ops
dev = "cuda:0"
lib = L.lib()
B, T, ci, co = 64, 800, int(os.environ.get("CI", "512")), int(os.environ.get("CO", "512"))
for knob in sys.argv[1:]:
    kk, v = knob.split("="); lib.qasr_debug_set(int(kk), int(v))
x = torch.randint(0, 127, (B, ci, T), dtype=torch.int8, device=dev)
w = torch.randint(-30, 30, (co, ci, 1), dtype=torch.int8, device=dev)
wp = ops.pack_weights(w)
ch = ops.chan_table(np.full(co, 1e-4, np.float32), np.zeros(co), np.full(co, (1 << 30) + 12345), np.full(co, 44), dev)
dual = os.environ.get("DUAL") == "1"
if dual:
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_RELU1 | L.F_STAGE2 | L.F_STAGE2_16, ch1=ch, qmin=-32, qmax=31,
                           m2=(1 << 30) + 777, e2=30, qmin2=-64, qmax2=63)
    x1 = torch.randint(0, 63, (B, ci, T), dtype=torch.int8, device=dev)
    op1 = ops.gemm_operand(x1, T, wp, ci, 1)
else:
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_RELU0 | L.F_FAST_F32, qmin=-128, qmax=127)
    op1 = None
out = torch.zeros((B, co, T), dtype=torch.int8, device=dev)
op0 = ops.gemm_operand(x, T, wp, ci, 1)
for _ in range(4): ops.conv_gemm(op0, op1, out, B, co, T, ep)
torch.cuda.synchronize()
print("ok")

"""Role timeline (producer / MMA / epilogue) of CTA 0 of the instrumented GEMM kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "q-asr_b200")]
import numpy as np, torch
from qasr import _lib as L, ops
dev = "cuda:0"
lib = L.lib()
B, T, ci, co = 64, 800, 512, 512
dual = len(sys.argv) > 1 and sys.argv[1] == "dual"
x = torch.randint(0, 127, (B, ci, T), dtype=torch.int8, device=dev)
w = torch.randint(-30, 30, (co, ci, 1), dtype=torch.int8, device=dev)
wp = ops.pack_weights(w)
ch = ops.chan_table(np.full(co, 1e-4, np.float32), np.zeros(co), np.full(co, (1 << 30) + 12345), np.full(co, 44), dev)
ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_RELU0 if not dual else L.F_RELU1, ch1=ch if dual else None, qmin=-128, qmax=127)
out = ops.alloc_codes(B, co, T, dev)
op0 = ops.gemm_operand(x, T, wp, ci, 1); op1 = ops.gemm_operand(x, T, wp, ci, 1) if dual else None
for _ in range(3): ops.conv_gemm(op0, op1, out, B, co, T, ep)
trace = torch.zeros(3 * 256, dtype=torch.int64, device=dev)
lib.qasr_debug_set(7, trace.data_ptr())
for knob in sys.argv[2:]:
    k, v = knob.split("="); lib.qasr_debug_set(int(k), int(v))
ops.conv_gemm(op0, op1, out, B, co, T, ep)
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(3, 256)
t0 = t[t > 0].min()
names = ["producer(slab issued)", "mma(tempty ok | full ok x slabs | tfull committed)", "epilogue(tfull ok | arrived)"]
for r in range(3):
    v = t[r][t[r] > 0] - t0
    print(names[r], len(v)); print("  ", " ".join("%.2f" % (x / 1e3) for x in v[:60]))

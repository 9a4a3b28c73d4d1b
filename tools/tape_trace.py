"""Role timeline (producer / MMA issuer / epilogue warp 0) of CTA 0 of the tape depthwise kernel (dwconv_tape.cu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "q-asr_b200")]
import numpy as np, torch
from qasr import _lib as L, ops
dev = "cuda:0"
lib = L.lib()
B, T, C, k = 64, 800, 512, int(os.environ.get("K", "33"))
x = torch.randint(0, 64, (B, C, T), dtype=torch.int8, device=dev)
w = torch.randint(-31, 32, (C, k), dtype=torch.int8, device=dev)
wd = ops.pack_dw_weights(w, 1, 1, k // 2)
ch = ops.chan_table(np.full(C, 1e-4, np.float32), np.zeros(C), np.full(C, (1 << 30) + 12345), np.full(C, 40), dev)
ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_FAST_F32, qmin=-32, qmax=31)
out = torch.zeros((B, C, T), dtype=torch.int8, device=dev)
lib.qasr_debug_set(8, 2)
for knob in sys.argv[1:]:
    kk, v = knob.split("="); lib.qasr_debug_set(int(kk), int(v))
for _ in range(3): ops.dwconv(x, T, wd, out, T, k, 1, 1, k // 2, ep)
torch.cuda.synchronize()
trace = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
lib.qasr_debug_set(7, trace.data_ptr())
ops.dwconv(x, T, wd, out, T, k, 1, 1, k // 2, ep)
torch.cuda.synchronize()
lib.qasr_debug_set(7, -1)
t = trace.cpu().numpy().reshape(4, 256)
t0 = t[t > 0].min()
names = ["producer (empty ok | copies issued) per item", "mma (loop top | full ok | item issued) per item",
         "epilogue warp 0 (before tfull wait | after) per round", "CTA (entry | set-up done | exit)"]
per = [2, 3, 2, 3]
for r in range(4):
    v = (t[r][t[r] > 0] - t0) / 1e3
    print(names[r], len(v))
    for i in range(0, len(v), per[r]):
        print("   item %2d: %s" % (i // per[r], " ".join("%7.2f" % q for q in v[i:i + per[r]])))

"""A few launches of one depthwise layer (512 channels, batch 64 x 800 frames) on the default engine: the target of
`ncu --set full -k regex:dwconv` captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "q-asr_b200")]
import numpy as np, torch
from qasr import _lib as L, ops
dev = "cuda:0"
lib = L.lib()
B, T, C, k = 64, 800, int(os.environ.get("C", "512")), int(os.environ.get("K", "33"))
for knob in sys.argv[1:]:
    kk, v = knob.split("="); lib.qasr_debug_set(int(kk), int(v))
x = torch.randint(0, 64, (B, C, T), dtype=torch.int8, device=dev)
w = torch.randint(-31, 32, (C, k), dtype=torch.int8, device=dev)
wd = ops.pack_dw_weights(w, 1, 1, k // 2)
ch = ops.chan_table(np.full(C, 1e-4, np.float32), np.zeros(C), np.full(C, (1 << 30) + 12345), np.full(C, 40), dev)
ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_FAST_F32, qmin=-32, qmax=31)
out = torch.zeros((B, C, T), dtype=torch.int8, device=dev)
for _ in range(4): ops.dwconv(x, T, wd, out, T, k, 1, 1, k // 2, ep)
torch.cuda.synchronize()
print("ok")

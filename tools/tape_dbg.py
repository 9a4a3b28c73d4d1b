import os, sys
sys.path[:0] = ["q-asr_b200"]
import numpy as np, torch
from qasr import _lib as L, ops
lib = L.lib(); dev = "cuda:0"
B, C, T, k = 64, 512, 800, 33
pad = k // 2
x = torch.randint(0, 64, (B, C, T), dtype=torch.int8, device=dev)
w = torch.randint(-31, 32, (C, k), dtype=torch.int8, device=dev)
wd = ops.pack_dw_weights(w, 1, 1, pad)
ch = ops.chan_table(np.full(C, 1e-4, np.float32), np.zeros(C), np.full(C, (1 << 30) + 12345), np.full(C, 40), dev)
ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_FAST_F32, qmin=-32, qmax=31)
out = torch.zeros((B, C, T), dtype=torch.int8, device=dev)
lib.qasr_debug_set(8, 2)
for pdl in (0, -1):
    lib.qasr_debug_set(9, pdl)
    for n in (1, 2, 3, 8):
        for _ in range(n):
            ops.dwconv(x, T, wd, out, T, k, 1, 1, pad, ep)
        torch.cuda.synchronize()
        print("pdl knob", pdl, "back-to-back", n, "ok", flush=True)

"""Micro-benchmark of the two depthwise-conv engines on the benchmark workload's layer shapes, with the
bring-up knobs (qasr_debug_set 6) to separate load / MMA / epilogue / store / Tc-build costs of the tcgen05 one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "q-asr_b200")]
import numpy as np, torch
from qasr import _lib as L, ops
dev = "cuda:0"
lib = L.lib()
B, T = int(os.environ.get("BENCH_B", "64")), 800


def run(C, k, dil, knobs, iters=20):
    pad = (dil * k) // 2 - 1 if dil > 1 else k // 2
    x = torch.randint(0, 64, (B, C, T), dtype=torch.int8, device=dev)
    w = torch.randint(-31, 32, (C, k), dtype=torch.int8, device=dev)
    wd = ops.pack_dw_weights(w, 1, dil, pad)
    m = np.full(C, (1 << 30) + 12345); e = np.full(C, 40)
    ch = ops.chan_table(np.full(C, 1e-4, np.float32), np.zeros(C), m, e, dev)
    # timing only: the float epilogue is what 99.7 % of the bench workload's channels run (no proof needed to time it)
    if os.environ.get("DW_PLAN_LIKE"):
        # what the plan launches on most layers: length mask + the per-channel float-multiplier table (all channels proved)
        rf = torch.full((C,), float(np.float32(((1 << 30) + 12345) * 2.0 ** -40)), dtype=torch.float32, device=dev)
        ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_MASK | (L.F_RELU0 if os.environ.get("DW_PLAN_LIKE") == "relu" else 0), qmin=-32, qmax=31,
                               out_lens=torch.full((B,), T, dtype=torch.int32, device=dev), f32_mult=rf)
    else:
        ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=0 if os.environ.get("DW_INT_EPILOGUE") else L.F_FAST_F32, qmin=-32, qmax=31)
    out = torch.zeros((B, C, T), dtype=torch.int8, device=dev)
    for kk, v in knobs.items(): lib.qasr_debug_set(kk, v)
    for _ in range(3): ops.dwconv(x, T, wd, out, T, k, 1, dil, pad, ep)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): ops.dwconv(x, T, wd, out, T, k, 1, dil, pad, ep)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / iters)
    for kk in knobs: lib.qasr_debug_set(kk, -1)
    return np.median(ts)


MODE = sys.argv[1] if len(sys.argv) > 1 else ""
if MODE == "plan":
    for (C, k, dil) in ((512, 33, 1), (512, 63, 1), (256, 33, 1), (256, 39, 1)):
        print("C=%d k=%d: rows %6.1f us   tapes %6.1f us   (DW_PLAN_LIKE=%s)" % (C, k, run(C, k, dil, {8: 0, 16: 0}), run(C, k, dil, {8: 0}), os.environ.get("DW_PLAN_LIKE")), flush=True)
    sys.exit(0)
if MODE == "rows":
    # rows per tape (knob 17) against the wave arithmetic: 8 resident CTAs of 4 warps per SM = 1184 CTAs per wave
    for (C, k, dil) in ((512, 33, 1), (512, 63, 1), (512, 75, 1), (256, 33, 1), (256, 39, 1)):
        print("C=%d k=%d: " % (C, k) + "  ".join("R=%d %5.1f" % (r, run(C, k, dil, {8: 0, 17: r})) for r in (1, 2, 3, 4, 5, 6, 7, 8)) +
              "  auto %5.1f" % run(C, k, dil, {8: 0, 17: -1}), flush=True)
    sys.exit(0)
if MODE == "rt":
    # row-tape walk of the IMMA engine (dwconv_rt.cu) against the row-at-a-time kernel (knob 16 = 0); knob 17 = rows per tape
    for (C, k, dil) in ((512, 33, 1), (512, 51, 1), (512, 63, 1), (512, 75, 1), (256, 33, 1), (256, 39, 1)):
        print("C=%d k=%d: rows %6.1f us   tapes %6.1f us   R=2 %6.1f   R=4 %6.1f   R=8 %6.1f" % (
            C, k, run(C, k, dil, {8: 0, 16: 0}), run(C, k, dil, {8: 0}), run(C, k, dil, {8: 0, 17: 2}),
            run(C, k, dil, {8: 0, 17: 4}), run(C, k, dil, {8: 0, 17: 8})), flush=True)
    sys.exit(0)
for (C, k, dil) in ((512, 33, 1), (512, 51, 1), (512, 63, 1), (512, 75, 1), (256, 33, 1), (256, 39, 1), (512, 87, 2)):
    print("C=%d k=%d dil=%d: imma rows %6.1f us   imma tapes %6.1f us   tcgen05 %6.1f us   tape %6.1f us" % (
        C, k, dil, run(C, k, dil, {8: 0, 16: 0}), run(C, k, dil, {8: 0}), run(C, k, dil, {8: 1}), run(C, k, dil, {8: 2})), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "cold":
    # the same launch on rotating buffer sets (together larger than the 126 MB L2): every launch reads cold rows
    def run_rot(C, k, nsets, engine):
        pad = k // 2
        xs = [torch.randint(0, 64, (B, C, T), dtype=torch.int8, device=dev) for _ in range(nsets)]
        outs = [torch.zeros((B, C, T), dtype=torch.int8, device=dev) for _ in range(nsets)]
        w = torch.randint(-31, 32, (C, k), dtype=torch.int8, device=dev)
        wd = ops.pack_dw_weights(w, 1, 1, pad)
        ch = ops.chan_table(np.full(C, 1e-4, np.float32), np.zeros(C), np.full(C, (1 << 30) + 12345), np.full(C, 40), dev)
        ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=L.F_FAST_F32, qmin=-32, qmax=31)
        lib.qasr_debug_set(8, engine)
        for i in range(nsets): ops.dwconv(xs[i], T, wd, outs[i], T, k, 1, 1, pad, ep)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        iters = 4 * nsets
        with torch.cuda.graph(g):
            for i in range(iters): ops.dwconv(xs[i % nsets], T, wd, outs[i % nsets], T, k, 1, 1, pad, ep)
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3 / iters)
        lib.qasr_debug_set(8, -1)
        return np.median(ts)
    for (C, k) in ((512, 33), (512, 63), (256, 33)):
        n = 8 if C == 512 else 16
        print("C=%d k=%d: imma warm %6.1f us   cold (%d buffer sets) %6.1f us   tape warm %6.1f   tape cold %6.1f" % (
            C, k, run_rot(C, k, 1, 0), n, run_rot(C, k, n, 0), run_rot(C, k, 1, 2), run_rot(C, k, n, 2)), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "tapeinner":
    for inner in (16, 32):
        print("tape, TMA rows of %3d bytes: C=512 k=33 %6.1f us   C=512 k=75 %6.1f us   loads only k=33 %6.1f us" % (
            inner, run(512, 33, 1, {8: 2, 14: inner}), run(512, 75, 1, {8: 2, 14: inner}), run(512, 33, 1, {8: 2, 14: inner, 6: 3})), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "tapeknobs":
    for name, kn in (("full", {}), ("no epilogue math", {6: 1}), ("no mma", {6: 2}), ("no copies", {6: 4}), ("no mma, no epilogue", {6: 3}),
                     ("skeleton", {6: 7})):
        kn = dict(kn); kn[8] = 2
        print("tape C=512 %-20s k=33 %6.1f us   k=75 %6.1f us" % (name, run(512, 33, 1, kn), run(512, 75, 1, kn)), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "tape":
    for nb in (4, 8, 11, 16, 22, 32, 64):
        print("tape, %2d utterances per item: C=512 k=33 %6.1f us   C=512 k=75 %6.1f us   C=256 k=33 %6.1f us" % (
            nb, run(512, 33, 1, {8: 2, 15: nb}), run(512, 75, 1, {8: 2, 15: nb}), run(256, 33, 1, {8: 2, 15: nb})), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "occupancy":
    for per_sm in (28, 32, 36, 40, 44, 48, 56, 64, 80):
        print("warps/SM wanted %3d: C=512 k=33 %6.1f us   C=512 k=75 %6.1f us   C=256 k=33 %6.1f us" % (
            per_sm, run(512, 33, 1, {8: 0, 10: per_sm}), run(512, 75, 1, {8: 0, 10: per_sm}), run(256, 33, 1, {8: 0, 10: per_sm})), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "knobs":
    for name, kn in (("full", {}), ("no-epilogue", {6: 1}), ("no-mma", {6: 2}), ("no-stores", {6: 4}), ("no-rebuild", {6: 8}),
                     ("loads only", {6: 3 + 8}), ("loads+mma", {6: 1 + 8}), ("loads+epi", {6: 2 + 8}),
                     ("rows, no cut", {11: 1, 6: 3 + 8 + 16}), ("cut, no rows", {11: 1, 6: 3 + 8 + 32}), ("neither", {11: 1, 6: 3 + 8 + 48}),
                     ("row-staged", {11: 1}), ("row-staged loads only", {11: 1, 6: 3 + 8})):
        kn = dict(kn); kn[8] = 1
        print("C=512 k=33 %-12s %6.1f us" % (name, run(512, 33, 1, kn)), flush=True)

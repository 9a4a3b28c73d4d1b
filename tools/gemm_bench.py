"""Micro-benchmark of the tensor-core GEMM kernel on the benchmark workload's layer shapes, with the
bring-up knobs (qasr_debug_set) to separate load / MMA / epilogue / store costs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "q-asr_b200")]
import numpy as np, torch
from qasr import _lib as L, ops
dev = "cuda:0"
lib = L.lib()
B, T = int(os.environ.get("BENCH_B", "64")), 800
FLUSH = False
def run(ci, co, dual, knobs, iters=20, ld=None):
    ld = ld or T
    x = torch.randint(0, 127, (B, ci, ld), dtype=torch.int8, device=dev)
    x1 = torch.randint(0, 127, (B, ci, ld), dtype=torch.int8, device=dev)
    w = torch.randint(-30, 30, (co, ci, 1), dtype=torch.int8, device=dev)
    wp, wp1 = ops.pack_weights(w), ops.pack_weights(w)
    m = np.full(co, (1 << 30) + 12345); e = np.full(co, 44)
    ch = ops.chan_table(np.full(co, 1e-4, np.float32), np.zeros(co), m, e, dev)
    flags = L.F_RELU1 | (L.F_STAGE2 if dual else 0) | L.F_RELU0 * (not dual)
    if not dual and not os.environ.get("GEMM_INT_EPILOGUE"):
        flags = L.F_RELU0 | L.F_FAST_F32   # timing only: the float epilogue is what the bench workload's single-operand layers run
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, ch1=ch if dual else None, qmin=-128, qmax=127,
                           m2=(1 << 30) + 777, e2=30, qmin2=-256, qmax2=255)
    out = torch.zeros((B, co, ld), dtype=torch.int8, device=dev)
    op0 = ops.gemm_operand(x, T, wp, ci, 1); op1 = ops.gemm_operand(x1, T, wp1, ci, 1) if dual else None
    for k, v in knobs.items(): lib.qasr_debug_set(k, v)
    for _ in range(3): ops.conv_gemm(op0, op1, out, B, co, T, ep)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): ops.conv_gemm(op0, op1, out, B, co, T, ep)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / iters)
    for k in knobs: lib.qasr_debug_set(k, -1)
    return np.median(ts)
if len(sys.argv) > 1 and sys.argv[1] == "knobs":
    # qasr_debug_set 6: 1 = epilogue does nothing, 2 = no MMA issue, 4 = no stores; 5: M tiles per CTA; 4: streaming A
    for (ci, co, dual) in ((512, 512, False), (512, 512, True)):
        for name, knobs in (("full", {}), ("no-epilogue", {6: 1}), ("no-mma", {6: 2}), ("no-stores", {6: 4}),
                            ("loads only", {6: 3}), ("MT=1", {5: 1}), ("streaming A", {4: 1}),
                            ("pair", {12: 1}), ("pair no-epilogue", {12: 1, 6: 1}), ("pair no-mma", {12: 1, 6: 2}),
                            ("pair bn128", {12: 1, 13: 128}), ("pair bn128 no-epi", {12: 1, 13: 128, 6: 1}), ("pair bn128 no-mma", {12: 1, 13: 128, 6: 2})):
            print("%4dx%4d dual=%d %-12s %7.1f us" % (ci, co, dual, name, run(ci, co, dual, knobs)), flush=True)
    sys.exit(0)
for (ci, co, dual) in ((512, 512, False), (512, 512, True), (256, 256, False), (256, 256, True), (256, 512, False), (256, 512, True),
                       (512, 1024, False), (1024, 1024, False)):
    print("%4dx%4d dual=%d   1 CTA %7.1f us   CTA pair %7.1f us   pair bn128 half boxes %7.1f us   pair bn128 full boxes %7.1f us" % (
        ci, co, dual, run(ci, co, dual, {12: 0}), run(ci, co, dual, {12: 1}), run(ci, co, dual, {12: 1, 13: 128}),
        run(ci, co, dual, {12: 1, 13: 128, 14: 0})), flush=True)


def run_ntc(ci, co, dual, iters=20):
    x = torch.randint(0, 127, (B, T, ci), dtype=torch.int8, device=dev)
    w = torch.randint(-30, 30, (co, ci, 1), dtype=torch.int8, device=dev)
    wp = ops.pack_weights(w)
    m = np.full(co, (1 << 30) + 12345); e = np.full(co, 44)
    ch = ops.chan_table(np.full(co, 1e-4, np.float32), np.zeros(co), m, e, dev)
    flags = L.F_RELU1 | (L.F_STAGE2 if dual else 0) | L.F_RELU0 * (not dual)
    ep = ops.make_epilogue(L.EPI_REQUANT_I8, ch, flags=flags, ch1=ch if dual else None, qmin=-128, qmax=127,
                           m2=(1 << 30) + 777, e2=30, qmin2=-256, qmax2=255)
    out = torch.zeros((B, T, co), dtype=torch.int8, device=dev)
    op0 = ops.gemm_operand(x, T, wp, ci, 1); op1 = ops.gemm_operand(x, T, wp, ci, 1) if dual else None
    for _ in range(3): ops.conv_ntc(op0, op1, out, B, co, T, ep)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): ops.conv_ntc(op0, op1, out, B, co, T, ep)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / iters)
    return np.median(ts)

for (ci, co, dual) in ((512, 512, False), (512, 512, True), (256, 256, False), (512, 1024, False)):
    print("NTC %4dx%4d dual=%d %7.1f us" % (ci, co, dual, run_ntc(ci, co, dual)), flush=True)

#!/usr/bin/env python
"""In-graph timeline of the bench workload's step: per launch, start / end (first CTA on data, last CTA out) in us from
the step's first stamp, the launch's duration and the gap to the previous launch's end.
usage: tools/stamps_dump.py [B] [knob=value ...] [brief]     (bring-up knobs of qasr_debug_set, e.g. 17=4)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from qasr.params import synthetic_features  # noqa: E402

dev = torch.device("cuda:0")
wl = dict(bench.WORKLOAD_CFG1 if "cfg1" in sys.argv else (bench.WORKLOAD_JASPER if "jasper" in sys.argv else bench.WORKLOAD))
brief = "brief" in sys.argv
from qasr import _lib as L  # noqa: E402
for arg in sys.argv[1:]:
    if "=" in arg:
        kk, v = arg.split("=")
        L.lib().qasr_debug_set(int(kk), int(v))
    elif arg.isdigit():
        wl["batch"] = int(arg)
model = bench.build_model(dev, wl)
x, lens = synthetic_features(wl["batch"], wl["frames"], seed=1000)
plan = model.plan_for(x.shape, dev)
print("plan", plan.stats)
plan.load_inputs(x.to(dev), lens.to(dev))
for _ in range(3):
    plan.run()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
acc = None
N = 5
for _ in range(N):
    flush.add_(1)
    st = plan.profile_stamps().astype(np.float64)
    ok = st[:, 1] > 0
    st = np.where(ok[:, None], st - st[ok, 0].min(), 0.0)
    acc = st if acc is None else acc + st
st = acc / N / 1e3
prev_end = 0.0
tot = {}
per_shape = {}
for (name, _), (s, e), v, nbytes in zip(plan.ops, st, ok, plan.op_bytes):
    if not v:
        continue
    if not brief:
        print("%-34s start %9.2f  end %9.2f  dur %7.2f  gap %6.2f  %6.0f GB/s" % (name, s, e, e - s, s - prev_end,
                                                                                   nbytes / (e - s) / 1e3))
    prev_end = e
    k = name.split("_")[0]
    tot[k] = tot.get(k, 0.0) + e - s
    kk = "_".join(name.split("_")[:2])
    per_shape.setdefault(kk, []).append(e - s)
print("knobs %s: median us per launch: %s" % ([a for a in sys.argv[1:] if "=" in a],
                                               {k: round(float(np.median(v)), 1) for k, v in per_shape.items()}))
print("span %.2f us; per class %s" % (st[ok, 1].max(), {k: round(v, 1) for k, v in tot.items()}))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = 0.0
for _ in range(10):
    flush.add_(1)
    a.record()
    plan.run()
    b.record()
    b.synchronize()
    ms += a.elapsed_time(b) / 10
print("timed graph (no stamps): %.4f ms per step" % ms)

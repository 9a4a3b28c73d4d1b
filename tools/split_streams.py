#!/usr/bin/env python
"""Experiment: the bench step as TWO half-batch plans replayed on two streams at once (launches of one half fill the
SMs the other half's kernel tails leave idle) against the one full-batch plan.
usage: tools/split_streams.py [parts]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import bench  # noqa: E402
from qasr import plan as P  # noqa: E402
from qasr.params import synthetic_features  # noqa: E402

dev = torch.device("cuda:0")
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = dict(bench.WORKLOAD)
model = bench.build_model(dev, wl)
x, lens = synthetic_features(wl["batch"], wl["frames"], seed=1000)
x, lens = x.to(dev), lens.to(dev)
full = model.plan_for(x.shape, dev)
full.load_inputs(x, lens)
hb = wl["batch"] // parts
halves = []
for i in range(parts):
    p = P.FusedPlan(model.encoder, model.decoder, (hb,) + tuple(x.shape[1:]), dev)
    p.load_inputs(x[i * hb:(i + 1) * hb].contiguous(), lens[i * hb:(i + 1) * hb].contiguous())
    halves.append(p)
streams = [torch.cuda.Stream() for _ in range(parts)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run_split():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    for p, s in zip(halves, streams):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            p.run()
    for s in streams:
        cur.wait_stream(s)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(n):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms += a.elapsed_time(b) / n
    return ms


print("full batch %d, one graph:            %.4f ms" % (wl["batch"], timeit(full.run)))
print("%d parts of %d on %d streams at once:  %.4f ms" % (parts, hb, parts, timeit(run_split)))
print("one part of %d alone:                %.4f ms" % (hb, timeit(halves[0].run)))
# same logits?
full.run()
run_split()
torch.cuda.synchronize()

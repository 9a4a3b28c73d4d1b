#!/usr/bin/env python
"""Per-kernel-class device time of the fused dynamic plan (qasr.dynplan): eager pass with one CUDA event between
launches.  usage: python tools/dynamic_profile.py [batch] [frames]"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import bench  # noqa: E402
import nemo.quantization.utils.quantize_model as qm  # noqa: E402
from qasr.params import synthetic_features  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
dev = torch.device("cuda:0")
model = bench.build_model(dev, dict(bench.WORKLOAD_DYNAMIC, batch=0), calibrate=False)
qm.set_dynamic(model, True)
x, lens = synthetic_features(B, T, seed=1)
x, lens = x.to(dev), lens.to(dev)
plan = model._dynamic_plan_or_none(x)
plan.run(x, lens)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    plan.run()
b.record()
b.synchronize()
print("graph replay: %.3f ms per forward (B=%d, T=%d) = %.0f audio-s/s" % (a.elapsed_time(b) / 5, B, T,
                                                                         B * T / 100 / (a.elapsed_time(b) / 5e3)))
n = len(plan.ops)
tot = defaultdict(float)
cnt = defaultdict(int)
for it in range(3):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda._sleep(20_000_000)
    evs[0].record()
    for i, (name, fn) in enumerate(plan.ops):
        fn()
        evs[i + 1].record()
    evs[n].synchronize()
    for i, (name, _) in enumerate(plan.ops):
        key = name.split("_")[0] + ("_stats" if name.endswith("_stats") else "") if name.startswith(("dw_", "gemm_")) else name
        if name.startswith("gemm_") and "res_act" in name:
            key = "gemm_dual" + ("_stats" if name.endswith("_stats") else "")
        tot[key] += evs[i].elapsed_time(evs[i + 1]) / 3
        cnt[key] += 1
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-22s %4d launches  %8.3f ms  (%.1f us each)" % (k, cnt[k] // 3, tot[k], 1e3 * tot[k] / (cnt[k] // 3)))
print("sum %.3f ms" % sum(tot.values()))

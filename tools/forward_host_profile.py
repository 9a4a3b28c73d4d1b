"""Host-side cost of the reference-shaped call model(processed_signal=, processed_signal_length=) on the bench workload:
where the time between the caller's copy and the graph launch goes."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]
import torch
import bench
from qasr.params import synthetic_features
dev = torch.device("cuda:0")
wl = dict(bench.WORKLOAD)
model = bench.build_model(dev, wl)
x, lens = synthetic_features(wl["batch"], wl["frames"], seed=1000)
x, lens = x.to(dev), lens.to(dev)
for _ in range(3):
    model(processed_signal=x, processed_signal_length=lens)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    model(processed_signal=x, processed_signal_length=lens)
host = (time.perf_counter() - t) / 20
torch.cuda.synchronize()
tot = (time.perf_counter() - t) / 20
print("host time per call (no sync) %.3f ms; with final sync %.3f ms per call" % (host * 1e3, tot * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    model(processed_signal=x, processed_signal_length=lens)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

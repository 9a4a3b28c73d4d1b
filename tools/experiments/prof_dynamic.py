import sys, os
sys.argv = ["x", "64"]
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200")]
src = open(os.path.join(ROOT, "tools/dynamic_bench.py")).read().split("with torch.no_grad():")[0].replace("ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))", "")
exec(src)
import torch
from torch.profiler import profile, ProfilerActivity
with torch.no_grad():
    for _ in range(2):
        model(processed_signal=x, processed_signal_length=lens)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model(processed_signal=x, processed_signal_length=lens)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

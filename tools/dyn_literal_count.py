"""How many channels of the fused dynamic plan (BASELINE config 4, W4/A8) does qasr_dyn_finalize leave on the literal recipe?
(0 of 69 725 on the bench model: the literal code inside the dynamic GEMM variants is never taken, but costs them 160-200 bytes
of spills per thread; an out-of-line per-lane call made the kernels 2x slower, so it stays inlined.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]
import torch, numpy as np
import bench
import nemo.quantization.utils.quantize_model as qm
from qasr.params import synthetic_features
dev = torch.device("cuda:0")
model = bench.build_model(dev, dict(bench.WORKLOAD_DYNAMIC, batch=0), calibrate=False)
qm.set_dynamic(model, True)
x, lens = synthetic_features(16, 1600, seed=1)
plan = model._dynamic_plan_or_none(x.to(dev))
plan.run(x.to(dev), lens.to(dev))
torch.cuda.synchronize()
tot = lit = 0
rows = []
for name, obj in vars(plan).items():
    pass
convs = [c for c in getattr(plan, "_convs", [])] if hasattr(plan, "_convs") else []
import gc
from qasr import dynplan
for o in gc.get_objects():
    if isinstance(o, dynplan._Conv):
        t = o.table.cpu().numpy()
        n = t.shape[0]; l = int((t[:, 3] < 0).sum())
        tot += n; lit += l
        wabs = o.wabs.cpu().numpy().astype(np.int64)
        rows.append((o.C_in, o.C_out, o.depthwise, l, n, int(np.abs(t[:, 1].astype(np.int64)).max()), int(wabs.max()) * 255))
print("literal channels %d of %d" % (lit, tot))
for r in rows[:12] + rows[-6:]:
    print("C_in %4d C_out %4d dw %d literal %4d / %4d   max|bias| %9d  max wabs*255 %9d" % r)

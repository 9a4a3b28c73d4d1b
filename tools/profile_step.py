"""Profiling driver: builds the bench workload, runs the fused plan eagerly (no CUDA graph) and
marks the second step with an NVTX range `profile_step` for ncu's --nvtx-include filter."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import bench  # noqa: E402

def main():
    device = torch.device("cuda:0")
    wl = dict(bench.WORKLOAD)
    if len(sys.argv) > 1:
        wl["batch"] = int(sys.argv[1])
    model = bench.build_model(device, wl)
    from qasr.params import synthetic_features
    from qasr import plan as P
    x, lens = synthetic_features(wl["batch"], wl["frames"], seed=1000)
    plan = P.FusedPlan(model.encoder, model.decoder, x.shape, device, use_graph=False)
    plan.run(x.to(device), lens.to(device))
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("profile_step")
    plan.run()
    plan.greedy()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("ops:", len(plan.ops), plan.stats)
    for i, ((name, _), nb) in enumerate(zip(plan.ops, plan.op_bytes)):
        if i in (0, 1, 2, 3, 4, 60, 61, 62, 100, 101, 102, 150, 151, 152, 153, 154, 155, 156):
            print(i, name, nb)

if __name__ == "__main__":
    main()

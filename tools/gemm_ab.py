"""A/B of the GEMM kernels on the bench workload's layer shapes: knob sets as arguments, e.g. `19=0` `19=1`."""
import os, sys
sys.argv, specs = [sys.argv[0]], sys.argv[1:]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_bench.py")).read().split("if len(sys.argv) > 1 and sys.argv[1] ==")[0])
specs = specs or [""]
for (ci, co, dual) in ((512, 512, False), (256, 256, False), (512, 1024, False), (512, 512, True), (256, 256, True)):
    line = "%4dx%4d dual=%d" % (ci, co, dual)
    for spec in specs:
        kn = {int(a.split("=")[0]): int(a.split("=")[1]) for a in spec.split(",") if a}
        line += "   [%s] 1 CTA %5.1f pair %5.1f" % (spec, run(ci, co, dual, {**kn, 12: 0}), run(ci, co, dual, {**kn, 12: 1}))
    print(line, flush=True)

#!/usr/bin/env python
"""Benchmark of the integer-only QuartzNet15x5 forward (BASELINE.json metric: audio-seconds per
second).  One "step" = one forward of the quantised encoder + CTC decoder + greedy argmax over one
batch of synthetic 64-bin mel features.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver through torch.distributed.run (one rank per GPU); utterances are
independent, so the batch dimension is sharded with no data-path collective (weak scaling: each
rank runs the per-GPU batch of the workload).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# BASELINE.json configs[1]: the configuration the metric is quoted on (fits one B200)
WORKLOAD = dict(name="quartznet15x5_w6a6_percentile99.996_static_b64_t1600", topology="quartznet15x5",
                weight_bit=6, act_bit=6, percentile=99.996, batch=64, frames=1600,
                calib_batches=4, calib_shape=(8, 64, 500))
# BASELINE.json configs[2] (parity-test network; `--workload jasper` measures it for DESIGN.md only)
WORKLOAD_JASPER = dict(name="jasper10x5dr_w8a8_static_b32_t1600", topology="jasper10x5dr", weight_bit=8, act_bit=8,
                       percentile=None, batch=32, frames=1600, calib_batches=2, calib_shape=(4, 64, 500))
FRAMES_PER_AUDIO_SECOND = 100.0  # 10 ms hop


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_model(device, wl):
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    import nemo.quantization.utils.quantize_model as qm
    from qasr.params import reestimate_bn_, synthetic_features
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology(wl["topology"])
    model.eval()
    reestimate_bn_(model.encoder)
    model.set_quant_bit(wl["weight_bit"], "weight")
    model.set_quant_bit(wl["act_bit"], "act")
    if wl["percentile"] is not None:
        qm.set_percentile(model, wl["percentile"])
    model.encoder.bn_folding()
    model = model.to(device)
    # calibration on synthetic batches of the synthesize.py default shape (distill_data.py:24)
    qm.calibrate(model)
    cb, cc, ct = wl["calib_shape"]
    for i in range(wl["calib_batches"]):
        cx, cl = synthetic_features(cb, ct, feat_in=cc, seed=100 + i)
        e, l, s = model.encoder(cx.to(device), cl.to(device))
        model.decoder(e, s)
    qm.evaluate(model)
    qm.set_dynamic(model, False)
    return model


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def kernel_breakdown(plan, flush, iters=5):
    """Per-op device time with CUDA events on the launching stream, aggregated per kernel class, with the
    algorithmic bytes each class moves (each conv reads its int8 input once and writes its int8
    output once).  The plan is run eagerly, op by op, with one event recorded between consecutive launches and a
    single synchronize at the end of the pass; a device-side spin in front keeps the device busy until the host has
    queued the whole pass, so an op's time is the distance
    between its two in-stream timestamps (kernel + the event records next to it), not a launch + sync round trip.
    The events break the programmatic-dependent-launch overlap the CUDA graph of the timed step has, so the sum
    over ops stays above the graph's step time; both are reported."""
    n = len(plan.ops)
    per_op = [0.0] * n
    for _ in range(iters):
        flush()
        for fn in plan._lens_fns:
            fn()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        torch.cuda._sleep(8_000_000)   # ~4 ms of device spin: every launch below is queued before the first one runs
        evs[0].record()
        for i, (name, fn) in enumerate(plan.ops):
            fn()
            evs[i + 1].record()
        evs[n].synchronize()
        for i in range(n):
            per_op[i] += evs[i].elapsed_time(evs[i + 1]) / iters
    classes = {}
    for (name, _), ms, nbytes in zip(plan.ops, per_op, plan.op_bytes):
        cls = "dwconv" if name.startswith("dw_") else (
            ("conv_ntc_tc" if plan.ntc else "gemm_i8_tc") if name.startswith("gemm_") else "quantize")
        c = classes.setdefault(cls, dict(ms=0.0, bytes=0, launches=0))
        c["ms"] += ms
        c["bytes"] += nbytes
        c["launches"] += 1
    return classes


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of each dominant-kernel class, from the committed
# `ncu --set full` capture (bench workload, a 512-channel layer: 52.4 MB algorithmic bytes per launch)
NCU_TRAFFIC = {
    "dwconv": dict(bytes=27065344, note="512-channel depthwise layer, 52.4 MB algorithmic per launch; the 26 MB output "
                                        "stays in the 126 MB L2 (profiles/r01_ncu_dw512_gemm512.txt)"),
    "gemm_i8_tc": dict(bytes=26753024, note="512x512 pointwise layer, 52.4 MB algorithmic per launch; output stays in L2 "
                                            "(profiles/r01_ncu_dw512_gemm512.txt)"),
}


def apply_debug_knobs():
    """QASR_DEBUG="4=1,5=1" -> qasr_debug_set(4, 1), (5, 1): schedule experiments only."""
    spec = os.environ.get("QASR_DEBUG")
    if spec:
        from qasr import _lib
        for kv in spec.split(","):
            k, v = kv.split("=")
            _lib.lib().qasr_debug_set(int(k), int(v))


def run_ours(args):
    distributed = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    wl = WORKLOAD_JASPER if args.workload == "jasper" else WORKLOAD
    from qasr.params import synthetic_features
    apply_debug_knobs()
    model = build_model(device, wl)
    B, T = wl["batch"], wl["frames"]
    x_host, lens_host = synthetic_features(B, T, seed=1000 + rank)
    x_pin, lens_pin = x_host.pin_memory(), lens_host.to(torch.int32).pin_memory()
    x_dev, lens_dev = x_host.to(device), lens_host.to(device)
    plan = model.plan_for(x_dev.shape, device)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def flush():
        flush_buf.add_(1)  # 256 MiB read+write > 126 MB L2

    def step_device():
        plan.run()          # inputs already resident in the plan's HBM buffers
        return plan.greedy()

    plan.load_inputs(x_dev, lens_dev)
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()

    def barrier():
        if distributed:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed between steps
    barrier()
    total_ms = 0.0
    for _ in range(args.steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_device()
        b.record()
        b.synchronize()
        total_ms += a.elapsed_time(b)
    barrier()
    # ---- end to end through the public API: pinned host features in, greedy labels out
    x_stage = torch.empty_like(x_dev)
    lens_stage = torch.empty(B, dtype=torch.int32, device=device)

    def step_e2e():
        x_stage.copy_(x_pin, non_blocking=True)
        lens_stage.copy_(lens_pin, non_blocking=True)
        log_probs, enc_len, greedy = model(processed_signal=x_stage, processed_signal_length=lens_stage)
        return greedy.to("cpu", non_blocking=False)

    for _ in range(3):
        labels = step_e2e()
    barrier()
    # the 26 MB pinned host -> device copy alone (explains the floor of the end-to-end number on this box)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        x_stage.copy_(x_pin, non_blocking=True)
    b.record()
    b.synchronize()
    h2d_alone_ms = a.elapsed_time(b) / 5
    serial_ms = 0.0   # one batch at a time: copy in -> forward -> labels out (latency view)
    for _ in range(args.steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        labels = step_e2e()
        b.record()
        b.synchronize()
        serial_ms += a.elapsed_time(b)
    barrier()
    # the serving loop (EncDecCTCModel.transcribe_stream): K batches back to back in ONE bracket, every batch's
    # 26 MB host->device copy and label read-back inside it, copies of batch i+1 overlapping the kernels of batch i.
    # No L2 flush here: one step streams ~3.3 GB of activation tensors, far more than the 126 MB L2.
    def host_batches(n):
        for _ in range(n):
            yield x_pin, lens_pin

    for labels in model.transcribe_stream(host_batches(3)):
        pass
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for labels in model.transcribe_stream(host_batches(args.steps)):
        pass
    b.record()
    b.synchronize()
    e2e_ms = a.elapsed_time(b)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([total_ms, e2e_ms, serial_ms], dtype=torch.float64, device=device)
    if distributed:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, serial_ms = float(t[0]), float(t[1]), float(t[2])
    if distributed:
        dist.barrier()
        dist.destroy_process_group()   # clean shutdown: no NCCL warning after the JSON line
    audio_s_per_step = world * B * T / FRAMES_PER_AUDIO_SECOND
    value = audio_s_per_step * args.steps / (total_ms / 1e3)
    e2e_value = audio_s_per_step * args.steps / (e2e_ms / 1e3)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel class (rank 0, measured live, eager op-by-op run)
    classes = kernel_breakdown(plan, flush)
    eager_ms = sum(c["ms"] for c in classes.values())
    dom = max(classes, key=lambda k: classes[k]["ms"])
    peak, peak_src = peaks()
    # The timed step replays all launches from ONE CUDA graph with programmatic dependent launch, so a kernel has no
    # event of its own inside it.  Its time inside the step = the step's device time (CUDA events, above) x the
    # kernel class's share of the eager event pass (the same share the committed ncu launch list shows); the eager
    # per-launch figure (events between launches, no launch overlap) is kept next to it as the lower bound.
    step_ms_graph = total_ms / args.steps
    share = classes[dom]["ms"] / eager_ms
    in_step_ms = step_ms_graph * share
    achieved = classes[dom]["bytes"] / (in_step_ms / 1e3) / 1e9
    achieved_eager = classes[dom]["bytes"] / (classes[dom]["ms"] / 1e3) / 1e9
    macs = sum(m for (n, _), m in zip(plan.ops, plan.op_macs) if n.startswith("gemm_"))
    tc_share = sum(v["ms"] for k, v in classes.items() if "tc" in k) / eager_ms
    roofline = dict(bound="hbm", kernel=dom, achieved=round(achieved, 1), peak=peak, unit="GB/s",
                    frac=round(achieved / peak, 4),
                    traffic=NCU_TRAFFIC.get(dom, {}).get("bytes"), traffic_note=NCU_TRAFFIC.get(dom, {}).get("note"),
                    peak_source=peak_src,
                    launches_per_step=classes[dom]["launches"],
                    algorithmic_bytes_per_step=classes[dom]["bytes"],
                    kernel_ms_per_step=round(in_step_ms, 4),
                    share_of_step=round(share, 3),
                    method="kernel time in the step = graph step time (CUDA events) x the class's share of an eager pass "
                           "timed with one CUDA event between consecutive launches (no host sync inside the pass)",
                    eager=dict(ms_per_step=round(eager_ms, 4), kernel_ms=round(classes[dom]["ms"], 4),
                               achieved=round(achieved_eager, 1), frac=round(achieved_eager / peak, 4)),
                    tensor_tops=round(2 * macs / (step_ms_graph * tc_share / 1e3) / 1e12, 1),
                    classes={k: dict(eager_ms=round(v["ms"], 4), share=round(v["ms"] / eager_ms, 3),
                                     gbs_in_step=round(v["bytes"] / (step_ms_graph * v["ms"] / eager_ms / 1e3) / 1e9, 1),
                                     launches=v["launches"]) for k, v in classes.items()})
    cpu = cpu_baseline(model, wl, seconds_budget=20.0)
    line = dict(
        metric="audio_seconds_per_second", value=round(value, 1), unit="audio-s/s", n_gpus=world,
        steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=round(total_ms / args.steps, 4),
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int8", data="synthetic",
        config=dict(workload=wl["name"], per_gpu_batch=B, frames=T, weight_bit=wl["weight_bit"],
                    act_bit=wl["act_bit"], percentile=wl["percentile"], mode="static",
                    l2="flushed between timed steps (256 MiB read+write)", parallelism="dp%d" % world),
        e2e=dict(value=round(e2e_value, 1), unit="audio-s/s", h2d_bytes_per_step=int(x_pin.numel() * 4 + B * 4),
                 d2h_bytes_per_step=int(labels.numel() * 4),
                 ms_per_step=round(e2e_ms / args.steps, 4),
                 api="EncDecCTCModel.transcribe_stream: pinned host batches in, greedy labels out, copies "
                     "double-buffered around the kernels; K steps in one bracket, working set 3.3 GB/step > L2",
                 serial_ms_per_step=round(serial_ms / args.steps, 4),
                 h2d_copy_alone_ms=round(h2d_alone_ms, 4),
                 h2d_copy_alone_gbs=round(x_pin.numel() * 4 / (h2d_alone_ms / 1e3) / 1e9, 1)),
        gpu_launches=int(args.steps * (len(plan.ops) + 1) * 2),
        roofline=roofline, cpu_baseline=cpu, clocks=clocks, plan=plan.stats)
    print(json.dumps(line), flush=True)


def oracle_setup(model, wl, sample_batch):
    import qasr_oracle as O
    from spec_from_modules import spec_from_model
    from qasr.params import synthetic_features
    spec = spec_from_model(model.encoder, model.decoder)
    x, lens = synthetic_features(sample_batch, wl["frames"], seed=1000)
    return O, spec, x.numpy(), lens.numpy()


def cpu_baseline(model, wl, seconds_budget):
    """The oracle port of the reference's CPU algorithm on the box's host cores, on a bounded
    sample of the same workload (same weights / ranges, a slice of the batch)."""
    sample_batch = 2
    O, spec, x, lens = oracle_setup(model, wl, sample_batch)
    t0 = time.perf_counter()
    O.model_forward(spec, x, lens)
    dt = time.perf_counter() - t0
    reps = 1
    while dt * (reps + 1) < seconds_budget and reps < 3:
        t1 = time.perf_counter()
        O.model_forward(spec, x, lens)
        dt = min(dt, time.perf_counter() - t1)
        reps += 1
    audio_s = sample_batch * wl["frames"] / FRAMES_PER_AUDIO_SECOND
    return dict(value=round(audio_s / dt, 2), unit="audio-s/s", cores=os.cpu_count(), kind="port",
                sample="%d of %d utterances x %d frames, numpy oracle (float64 BLAS as integer engine), "
                       "best of %d" % (sample_batch, wl["batch"], wl["frames"], reps))


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle port; the reference itself is Python and
    is not shipped to the GPU box) on the host cores, same workload / metric, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOAD
    torch.set_num_threads(os.cpu_count() or 1)
    # the calibrated parameters come from the same seeded model; calibration itself needs the GPU
    # kernels, so on a GPU-less host the ranges are calibrated by the oracle instead
    from nemo.collections.asr.models.ctc_models import EncDecCTCModel
    import nemo.quantization.utils.quantize_model as qm
    from qasr.params import reestimate_bn_, synthetic_features
    import qasr_oracle as O
    from spec_from_modules import spec_from_model
    torch.manual_seed(0)
    model = EncDecCTCModel.from_topology(wl["topology"])
    model.eval()
    reestimate_bn_(model.encoder)
    model.set_quant_bit(wl["weight_bit"], "weight")
    model.set_quant_bit(wl["act_bit"], "act")
    qm.set_percentile(model, wl["percentile"])
    model.encoder.bn_folding()
    spec = spec_from_model(model.encoder, model.decoder)
    O.set_mode(spec, True)
    cb, cc, ct = wl["calib_shape"]
    cx, cl = synthetic_features(2, ct, feat_in=cc, seed=100)
    O.model_forward(spec, cx.numpy(), cl.numpy())
    O.set_mode(spec, False)
    sample_batch = 2
    x, lens = synthetic_features(sample_batch, wl["frames"], seed=1000)
    x, lens = x.numpy(), lens.numpy()
    for _ in range(min(args.warmup, 1)):
        O.model_forward(spec, x, lens)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.model_forward(spec, x, lens)
    dt = time.perf_counter() - t0
    audio_s = sample_batch * wl["frames"] / FRAMES_PER_AUDIO_SECOND * args.steps
    value = round(audio_s / dt, 2)
    sample = ("%d of %d utterances x %d frames per step, numpy oracle port of the reference's CPU "
              "algorithm" % (sample_batch, wl["batch"], wl["frames"]))
    print(json.dumps(dict(
        impl="reference", metric="audio_seconds_per_second", value=value, unit="audio-s/s",
        n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=args.steps, warmup=min(args.warmup, 1),
        ms_per_step=round(dt / args.steps * 1e3, 2), higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="int8", data="synthetic",
        config=dict(workload=wl["name"], per_gpu_batch=wl["batch"], frames=wl["frames"],
                    weight_bit=wl["weight_bit"], act_bit=wl["act_bit"], percentile=wl["percentile"],
                    mode="static"),
        cpu_baseline=dict(value=value, unit="audio-s/s", cores=os.cpu_count(), kind="port", sample=sample),
        e2e=dict(value=value, unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="quartznet", choices=["quartznet", "jasper"],
                    help="quartznet = BASELINE configs[1] (the metric's configuration, default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

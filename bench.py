#!/usr/bin/env python
"""Benchmark of the integer-only QuartzNet15x5 forward (BASELINE.json metric: audio-seconds per
second).  One "step" = one forward of the quantised encoder + CTC decoder + greedy argmax over one
batch of synthetic 64-bin mel features.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver through torch.distributed.run (one rank per GPU); utterances are
independent, so the batch dimension is sharded with no data-path collective (weak scaling: each
rank runs the per-GPU batch of the workload).  Beside the headline (BASELINE configs[1], static) the
JSON line carries, outside the headline's timed region:
  parity_check   the oracle on two utterances of the very batch that was timed, at full length
  secondary      BASELINE configs[0] / [2] / [4] on one GPU, each with its own parity_check   (N = 1)
  dynamic_dp     BASELINE configs[3]: W4/A8 dynamic ranges, the batch sharded over the N ranks with one
                 range all-reduce per QuantAct over NCCL, checked bit for bit against a one-rank run
`--impl reference` times the UNMODIFIED reference (baseline/_ref, installed by tools/install_ref.sh)
on the host cores, same workload / config / warm-up.
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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _paths(ours):
    want = [os.path.join(ROOT, "oracle")] + ([os.path.join(ROOT, "q-asr_b200")] if ours else [])
    for p in want:
        if p not in sys.path:
            sys.path.insert(0, p)


import numpy as np  # noqa: E402
import torch  # noqa: E402

# BASELINE.json configs[1]: the configuration the metric is quoted on (fits one B200)
WORKLOAD = dict(name="quartznet15x5_w6a6_percentile99.996_static_b64_t1600", topology="quartznet15x5",
                weight_bit=6, act_bit=6, percentile=99.996, batch=64, frames=1600,
                calib_batches=4, calib_shape=(8, 64, 500))
# BASELINE.json configs[0], [2], [4]: `secondary` entries of the JSON line (and `--workload jasper`)
WORKLOAD_CFG1 = dict(name="quartznet15x5_w8a8_static_b8_t500", topology="quartznet15x5", weight_bit=8, act_bit=8,
                     percentile=None, batch=8, frames=500, calib_batches=2, calib_shape=(8, 64, 500))
WORKLOAD_JASPER = dict(name="jasper10x5dr_w8a8_static_b32_t1600", topology="jasper10x5dr", weight_bit=8, act_bit=8,
                       percentile=None, batch=32, frames=1600, calib_batches=2, calib_shape=(4, 64, 500))
WORKLOAD_DYNAMIC = dict(name="quartznet15x5_w4a8_dynamic_b1024_t1600", topology="quartznet15x5", weight_bit=4,
                        act_bit=8, percentile=None, global_batch=1024, single_gpu_batch=128, frames=1600)
FRAMES_PER_AUDIO_SECOND = 100.0  # 10 ms hop
L2_NOTE = "flushed between timed steps (256 MiB read+write)"


def make_config(wl, world):
    """The `config` object of the JSON line: ONE function for both arms, so they cannot drift apart."""
    return dict(workload=wl["name"], per_gpu_batch=wl["batch"], frames=wl["frames"], weight_bit=wl["weight_bit"],
                act_bit=wl["act_bit"], percentile=wl["percentile"], mode="static", l2=L2_NOTE,
                parallelism="dp%d" % world)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# int8 tensor peak: no driver-written figure exists; this is the repo's own microbenchmark
# (tools/microbench/umma_i8_rate.cu, output committed as profiles/r02_umma_i8_rate.txt): 64.0 clk per
# tcgen05.mma kind::i8 M128 N128 K32 per SM = 4.56 POPS at 148 SMs x 1965 MHz
INT8_PEAK_TOPS = 4560.0
INT8_PEAK_NOTE = "builder-measured tcgen05 kind::i8 issue rate (tools/microbench/umma_i8_rate.cu, profiles/r02_umma_i8_rate.txt), not a driver record"


def build_model(device, wl, calibrate=True):
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
    if calibrate:
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


def op_class(plan, name):
    if name.startswith("dw_"):
        return "dwconv"
    if name.startswith("gemm_"):
        return "conv_ntc_tc" if plan.ntc else "gemm_i8_tc"
    return "quantize"


def kernel_breakdown(plan, flush, iters=5):
    """Per-op device time with CUDA events on the launching stream (eager pass, one event between consecutive
    launches, a device-side spin in front so every launch is queued before the first one runs), aggregated per
    kernel class with the algorithmic bytes each class moves.  The events break the programmatic-
    dependent-launch overlap of the timed CUDA graph, so the sum stays above the graph's step time: this pass is
    the UPPER bound kept next to the in-graph stamps (`in_graph_times`)."""
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
    for (name, _), ms, nbytes, macs in zip(plan.ops, per_op, plan.op_bytes, plan.op_macs):
        c = classes.setdefault(op_class(plan, name), dict(ms=0.0, bytes=0, launches=0, macs=0))
        c["ms"] += ms
        c["bytes"] += nbytes
        c["launches"] += 1
        c["macs"] += macs
    return classes


def in_graph_times(plan, flush, iters=5):
    """Kernel time INSIDE the timed step, measured: an instrumented copy of the plan's CUDA graph in which every
    launch stamps %globaltimer when its first CTA starts on activation data (after griddepcontrol.wait) and when
    its last CTA retires (qasr_profile_begin / qasr_profile_read).  Same launches, same programmatic dependent
    launch edges as the timed graph.  Returns per-class sums of (last end - first start) per launch in ms, the
    whole span, and the ms covered by any kernel (union of the intervals), or None when the library has no stamps."""
    if not hasattr(plan, "profile_stamps"):
        return None
    tot, valid = None, None
    for _ in range(iters):
        flush()
        st = plan.profile_stamps()          # (n_ops, 2) int64 nanoseconds, (0, 0) for launches without a stamp
        if st is None:
            return None
        valid = st[:, 1] > 0
        st = np.where(valid[:, None], st.astype(np.float64) - float(st[valid, 0].min()), 0.0)
        tot = st if tot is None else tot + st
    st = tot / iters
    dur = (st[:, 1] - st[:, 0]) / 1e6
    classes = {}
    for (name, _), d, ok in zip(plan.ops, dur, valid):
        if ok:
            classes[op_class(plan, name)] = classes.get(op_class(plan, name), 0.0) + float(d)
    order = [i for i in np.argsort(st[:, 0]) if valid[i]]
    covered, cur_s, cur_e = 0.0, None, None
    for i in order:
        s, e = st[i]
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                covered += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    covered += cur_e - cur_s
    return dict(classes=classes, span_ms=float(st[valid, 1].max() - st[valid, 0].min()) / 1e6, covered_ms=covered / 1e6)


def oracle_for(model):
    import qasr_oracle as O
    from spec_from_modules import spec_from_model
    return O, spec_from_model(model.encoder, model.decoder)


def parity_check(model, plan, x_host, lens_host, rows=(0, 1)):
    """The oracle on `rows` of the batch the plan just ran (same calibrated weights / ranges, full length) against
    the plan's logits for those utterances: logit CODES and greedy labels must be identical."""
    O, spec = oracle_for(model)
    rows = list(rows)
    t0 = time.perf_counter()
    want = O.model_forward(spec, x_host[rows].numpy(), lens_host[rows].numpy())
    dt = time.perf_counter() - t0
    logits = plan.logits[rows, :, :plan.cur_T_out]
    got = torch.round(logits.double() / plan.logits_scale.double().view(1, -1, 1)).to(torch.int64).cpu().numpy()
    greedy = plan.greedy()[rows].cpu().numpy()
    lg = logits.cpu().numpy()
    return dict(utterances=len(rows), frames=int(x_host.shape[2]), logit_codes=int(got.size),
                logit_code_mismatches=int((got != want["logit_codes"]).sum()),
                greedy_equal=bool(np.array_equal(greedy, want["greedy"])),
                max_rel_logit_err=float(np.abs(lg - want["logits"]).max() / np.abs(want["logits"]).max()),
                oracle_seconds=round(dt, 2)), dt


def time_steps(fn, steps, flush):
    total = 0.0
    for _ in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    return total


def secondary_static(device, wl, flush, steps, warmup, note):
    """One secondary static configuration on this GPU: device-resident step time, GEMM TOPS / HBM GB/s over the
    whole step, and the oracle check on one utterance at full length."""
    from qasr.params import synthetic_features
    model = build_model(device, wl)
    B, T = wl["batch"], wl["frames"]
    x, lens = synthetic_features(B, T, seed=2000)
    plan = model.plan_for(x.shape, device)
    plan.load_inputs(x.to(device), lens.to(device))
    for _ in range(warmup):
        plan.run()
        plan.greedy()
    torch.cuda.synchronize()
    ms = time_steps(lambda: (plan.run(), plan.greedy()), steps, flush) / steps
    pc, _ = parity_check(model, plan, x, lens, rows=(0,))
    audio_s = B * T / FRAMES_PER_AUDIO_SECOND
    macs, nbytes = sum(plan.op_macs), sum(plan.op_bytes)
    hbm, _ = peaks()
    tops = 2 * macs / (ms / 1e3) / 1e12
    out = dict(workload=wl["name"], ms_per_step=round(ms, 4), value=round(audio_s / (ms / 1e3), 1), unit="audio-s/s",
               launches=len(plan.ops), tensor_tops=round(tops, 1), ops_frac=round(tops / INT8_PEAK_TOPS, 4),
               hbm_gbs=round(nbytes / (ms / 1e3) / 1e9, 1), hbm_frac=round(nbytes / (ms / 1e3) / 1e9 / hbm, 4),
               parity_check=pc, note=note)
    del plan, model
    torch.cuda.empty_cache()
    return out


def secondary_audio(device, flush, steps, warmup):
    """BASELINE configs[4] on this GPU: 20 s of 16 kHz audio -> mel featurizer (K6) -> W8/A8 QuartzNet15x5 -> greedy
    labels -> CTC collapse (K5), through EncDecCTCModel.forward(input_signal=...).  Parity: featurizer output of one
    utterance against the float64 featurizer oracle (1e-5 relative to max |feature|), logit codes of that utterance
    against the integer oracle run on the GPU featurizer's output."""
    import featurizer_oracle as F
    from nemo.collections.asr.modules.audio_preprocessing import AudioToMelSpectrogramPreprocessor
    from qasr import ops
    wl = dict(WORKLOAD_CFG1, name="audio20s_featurizer_quartznet15x5_w8a8_greedy_b32", batch=32)
    model = build_model(device, wl)
    model.preprocessor = AudioToMelSpectrogramPreprocessor(dither=0.0).to(device)
    B, L = wl["batch"], 320000
    g = torch.Generator().manual_seed(3000)
    audio = 0.1 * torch.randn(B, L, generator=g)
    alen = torch.full((B,), L, dtype=torch.long)
    audio_d, alen_d = audio.to(device), alen.to(device)
    blank = model.decoder.num_classes_with_blank - 1

    def step():
        lp, enc_len, greedy = model(input_signal=audio_d, input_signal_length=alen_d)
        return ops.ctc_collapse(greedy, blank, enc_len)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ms = time_steps(step, steps, flush) / steps
    # featurizer alone
    pre = model.preprocessor
    feat_ms = time_steps(lambda: pre(input_signal=audio_d, length=alen_d), steps, flush) / steps
    feat, flen = pre(input_signal=audio_d[:1], length=alen_d[:1])
    want_feat, want_len = F.melspec(audio[:1].numpy(), alen[:1].numpy())
    feat_err = float(np.abs(feat.cpu().numpy() - want_feat).max() / np.abs(want_feat).max())
    plan = model.plan_for(feat.shape, device)
    plan.run(feat, flen)
    pc, _ = parity_check(model, plan, feat.cpu(), flen.cpu(), rows=(0,))
    pc["featurizer_max_rel_err"] = feat_err
    pc["featurizer_within_1e-5"] = bool(feat_err <= 1e-5 and int(flen[0]) == int(want_len[0]))
    pc["featurizer_note"] = ("float32 FFT kernel against the FLOAT64 oracle, relative to max|feature|; on 20 s of white noise the "
                             "reference's own float32 torch.stft pipeline measures 1.0e-5 against the same oracle on this utterance (DESIGN 5, tests/test_featurizer_cpu.py): "
                             "1e-5 is float32's noise floor here, tests/test_gpu_featurizer.py holds 1e-5 on its cases")
    audio_s = B * L / 16000.0
    feat_bytes = B * L * 4 + B * 64 * int(feat.shape[2]) * 4
    out = dict(workload=wl["name"], ms_per_step=round(ms, 4), value=round(audio_s / (ms / 1e3), 1), unit="audio-s/s",
               featurizer_ms=round(feat_ms, 4), featurizer_gbs=round(feat_bytes / (feat_ms / 1e3) / 1e9, 1),
               parity_check=pc, note="configs[4] on one GPU, device-resident audio; batch 32 x 20 s")
    del plan, model
    torch.cuda.empty_cache()
    return out


def dynamic_dp(device, rank, world, distributed, steps=3):
    """BASELINE configs[3]: QuartzNet15x5 W4/A8, dynamic min/max ranges.  Every QuantAct's range is taken over the
    WHOLE batch (quant_modules.py:149-156); the batch is sharded over the ranks and one (min, max) all-reduce per
    QuantAct (NCCL) restores the full-batch range, so the sharded result must equal the one-rank result bit for bit:
    rank 0 re-runs the first utterances inside the FULL batch on its own and compares logits byte for byte.
    Timed on the device (CUDA events), max over ranks; `allreduce_share` = (sharded - same shard without the
    collectives) / sharded."""
    import torch.distributed as dist
    import nemo.quantization.utils.quantize_model as qm
    from qasr import dist as qdist
    from qasr.params import synthetic_features
    wl = WORKLOAD_DYNAMIC
    model = build_model(device, dict(wl, batch=0), calibrate=False)
    qm.set_dynamic(model, True)
    gb = wl["global_batch"] if world > 1 else wl["single_gpu_batch"]
    T = wl["frames"]
    lo, hi = qdist.shard_bounds(gb, rank, world)
    g = torch.Generator().manual_seed(4000)
    check_rows = 2

    def features(start, stop):   # utterance i of the global batch is a function of i alone
        xs = []
        for i in range(start, stop):
            gi = torch.Generator().manual_seed(4000 + i)
            xs.append(torch.randn(64, T, generator=gi))
        return torch.stack(xs)

    x = features(lo, hi).to(device)
    lens = torch.full((hi - lo,), T, dtype=torch.int32, device=device)

    used = {"plan": 0, "modules": 0}

    def fwd(xx, ll):
        r = model.forward_dynamic(xx, ll)      # the fused device-side dynamic plan (qasr.dynplan)
        if r is not None:
            used["plan"] += 1
            return r
        used["modules"] += 1
        lp, el, gr = model(processed_signal=xx, processed_signal_length=ll)
        return model.decoder.last_logits, gr

    def timed(n):
        ms = 0.0
        for _ in range(n):
            if distributed:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fwd(x, lens)
            b.record()
            b.synchronize()
            ms += a.elapsed_time(b)
        return ms / n

    if distributed:
        qdist.enable_data_parallel()
    fwd(x, lens)                      # warm-up (kernel attributes, allocator)
    ms_dp = timed(steps)
    logits_dp, greedy_dp = fwd(x, lens)
    logits_dp = logits_dp[:check_rows].clone()
    qdist.disable_data_parallel()
    if distributed:
        fwd(x, lens)                  # the shard-local plan is another plan: build / capture it outside the timing
    ms_local = timed(steps) if distributed else ms_dp   # same shard, ranges from the shard alone: no collective
    t = torch.tensor([ms_dp, ms_local], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dp, ms_local = float(t[0]), float(t[1])
    bit_exact = None
    if distributed:
        dist.barrier()
        if rank == 0:
            # the reference's semantics: ONE process sees the whole batch.  Chunked over the batch it would not be the
            # same computation, so the whole global batch runs on this one GPU.
            model._plan_cache.clear()
            torch.cuda.empty_cache()
            xf = features(0, gb).to(device)
            lf = torch.full((gb,), T, dtype=torch.int32, device=device)
            logits_full, _ = fwd(xf, lf)
            bit_exact = bool(torch.equal(logits_full[:check_rows], logits_dp))
            del xf, logits_full
        dist.barrier()
    audio_s = gb * T / FRAMES_PER_AUDIO_SECOND
    path = ("fused dynamic plan: statistics pass + requantising pass per conv, ranges / scales / dyadic tables on the "
            "device, one CUDA graph" if used["modules"] == 0 else
            "module-level QuantAct / QuantConv1d calls (%d of %d forwards)" % (used["modules"], used["modules"] + used["plan"]))
    out = dict(workload=wl["name"] if world > 1 else wl["name"].replace("b1024", "b%d" % gb), global_batch=gb,
               per_rank_batch=hi - lo, ms_per_step=round(ms_dp, 3), value=round(audio_s / (ms_dp / 1e3), 1),
               unit="audio-s/s", allreduce_share=round(max(0.0, (ms_dp - ms_local) / ms_dp), 4) if distributed else 0.0,
               ms_per_step_without_allreduce=round(ms_local, 3), bit_exact_vs_single_rank=bit_exact,
               collective=("one-shot MAX exchange of every QuantAct's absmax through peer-mapped memory (CUDA IPC over "
                           "NVLink), inside the finalize kernel of the graph" if used["modules"] == 0 else
                           "one NCCL MAX all-reduce of [-min, max] per QuantAct") if distributed else "none (one rank)",
               path=path)
    del model
    qdist.close_peer_slots()
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    _paths(ours=True)
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
    from qasr import _lib
    knobs = int(_lib.load_library().qasr_debug_knobs_set())
    if knobs:
        raise RuntimeError("bench.py refuses to time the kernels with %d bring-up knob(s) set (qasr_debug_set)" % knobs)
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

    warmup = max(args.warmup, 3)
    plan.load_inputs(x_dev, lens_dev)
    for _ in range(warmup):
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
    total_ms = time_steps(step_device, args.steps, flush)
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
    # the reference-shaped call, one batch at a time: copy in -> model(processed_signal=...) -> labels out
    serial_ms = time_steps(step_e2e, args.steps, flush)
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
    audio_s_per_step = world * B * T / FRAMES_PER_AUDIO_SECOND
    value = audio_s_per_step * args.steps / (total_ms / 1e3)
    e2e_value = audio_s_per_step * args.steps / (e2e_ms / 1e3)
    serial_value = audio_s_per_step * args.steps / (serial_ms / 1e3)

    # ---- BASELINE configs[3] (all ranks take part; outside the headline's timed region)
    dyn = None
    if not args.no_extras and args.workload == "quartznet":
        dyn = dynamic_dp(device, rank, world, distributed)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()   # clean shutdown: no NCCL warning after the JSON line
    if rank != 0:
        return

    # ---- parity of the timed batch itself (rank 0): oracle on two of its utterances
    plan.load_inputs(x_dev, lens_dev)
    step_device()
    pc, oracle_dt = parity_check(model, plan, x_host, lens_host)

    # ---- roofline of the dominant kernel class (rank 0)
    classes = kernel_breakdown(plan, flush)
    eager_ms = sum(c["ms"] for c in classes.values())
    stamps = in_graph_times(plan, flush)
    step_ms_graph = total_ms / args.steps
    if stamps is not None:
        kms = dict(stamps["classes"])
        for k, v in classes.items():      # launches without a stamp (the input quantiser): eager figure
            kms.setdefault(k, v["ms"])
        method = ("measured inside the step: an instrumented copy of the step's CUDA graph (same launches, same "
                  "programmatic-dependent-launch edges) in which every launch stamps %globaltimer at its first CTA's "
                  "start on activation data and at its last CTA's end; class time = sum over its launches of "
                  "(last end - first start)")
    else:
        kms = {k: step_ms_graph * v["ms"] / eager_ms for k, v in classes.items()}
        method = ("attributed: graph step time (CUDA events) x the class's share of an eager pass timed with one CUDA "
                  "event between consecutive launches")
    dom = max(kms, key=lambda k: kms[k])
    peak, peak_src = peaks()
    in_step_ms = kms[dom]
    achieved = classes[dom]["bytes"] / (in_step_ms / 1e3) / 1e9
    gemm_cls = [k for k in classes if "tc" in k]
    gemm_ms = sum(kms[k] for k in gemm_cls)
    gemm_macs = sum(classes[k]["macs"] for k in gemm_cls)
    tops = 2 * gemm_macs / (gemm_ms / 1e3) / 1e12 if gemm_ms else 0.0
    traffic = ncu_traffic(dom)
    roofline = dict(bound="hbm", kernel=dom, achieved=round(achieved, 1), peak=peak, unit="GB/s",
                    frac=round(achieved / peak, 4), traffic=traffic.get("bytes"), traffic_note=traffic.get("note"),
                    peak_source=peak_src, launches_per_step=classes[dom]["launches"],
                    algorithmic_bytes_per_step=classes[dom]["bytes"],
                    algorithmic_bytes_per_launch=classes[dom]["bytes"] // classes[dom]["launches"],
                    kernel_ms_per_step=round(in_step_ms, 4), avg_launch_us=round(1e3 * in_step_ms / classes[dom]["launches"], 3),
                    method=method,
                    step=dict(ms=round(step_ms_graph, 4), algorithmic_bytes=int(sum(c["bytes"] for c in classes.values())),
                              gbs=round(sum(c["bytes"] for c in classes.values()) / (step_ms_graph / 1e3) / 1e9, 1),
                              frac=round(sum(c["bytes"] for c in classes.values()) / (step_ms_graph / 1e3) / 1e9 / peak, 4),
                              in_graph_span_ms=None if stamps is None else round(stamps["span_ms"], 4),
                              in_graph_covered_ms=None if stamps is None else round(stamps["covered_ms"], 4)),
                    eager=dict(ms_per_step=round(eager_ms, 4),
                               note="event between launches: no launch overlap, upper bound per class",
                               classes={k: round(v["ms"], 4) for k, v in classes.items()}),
                    tensor_tops=round(tops, 1), ops_frac=round(tops / INT8_PEAK_TOPS, 4), ops_peak_tops=INT8_PEAK_TOPS,
                    ops_peak_source=INT8_PEAK_NOTE,
                    classes={k: dict(ms_in_step=round(kms[k], 4), gbs_in_step=round(v["bytes"] / (kms[k] / 1e3) / 1e9, 1),
                                     frac=round(v["bytes"] / (kms[k] / 1e3) / 1e9 / peak, 4), launches=v["launches"],
                                     algorithmic_bytes=v["bytes"]) for k, v in classes.items()})
    cpu = cpu_baseline(model, wl, oracle_dt)
    secondary = None
    if not args.no_extras and world == 1 and args.workload == "quartznet":
        del plan
        model._plan_cache.clear()
        torch.cuda.empty_cache()
        secondary = [
            secondary_static(device, WORKLOAD_CFG1, flush, args.steps, warmup,
                             "configs[0]: a latency point (40 audio-s per step), not a roofline point"),
            secondary_static(device, WORKLOAD_JASPER, flush, max(3, args.steps // 4), warmup,
                             "configs[2]: dense k = 11..29 convs as channels-last implicit GEMM; tensor-bound"),
            secondary_audio(device, flush, max(3, args.steps // 4), warmup)]
    line = dict(
        metric="audio_seconds_per_second", value=round(value, 1), unit="audio-s/s", n_gpus=world,
        steps=args.steps, warmup=warmup, ms_per_step=round(total_ms / args.steps, 4),
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int8", data="synthetic",
        config=make_config(wl, world),
        e2e=dict(value=round(e2e_value, 1), unit="audio-s/s", h2d_bytes_per_step=int(x_pin.numel() * 4 + B * 4),
                 d2h_bytes_per_step=int(labels.numel() * 4),
                 ms_per_step=round(e2e_ms / args.steps, 4),
                 api="EncDecCTCModel.transcribe_stream (the serving loop of inference.py, not a reference API): pinned host "
                     "batches in, greedy labels out, copies double-buffered around the kernels; K steps in one bracket, "
                     "working set 3.3 GB/step > L2",
                 forward_api=dict(value=round(serial_value, 1), ms_per_step=round(serial_ms / args.steps, 4),
                                  api="the reference-shaped call, one batch at a time: H2D copy -> "
                                      "model(processed_signal=, processed_signal_length=) -> greedy labels to host"),
                 h2d_copy_alone_ms=round(h2d_alone_ms, 4),
                 h2d_copy_alone_gbs=round(x_pin.numel() * 4 / (h2d_alone_ms / 1e3) / 1e9, 1)),
        gpu_launches=int(args.steps * (len(model.plan_for(x_dev.shape, device).ops) + 1) * 2),
        parity_check=pc, roofline=roofline, cpu_baseline=cpu, clocks=clocks,
        plan=model.plan_for(x_dev.shape, device).stats, debug_knobs_set=knobs, dynamic_dp=dyn, secondary=secondary)
    print(json.dumps(line), flush=True)


def ncu_traffic(cls):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the class's 512-channel layer, parsed from the
    committed `ncu --set full` summary of this round (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py
    from the capture of the same commit); falls back to round 1's capture."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            if cls in d:
                return dict(bytes=d[cls]["dram_bytes"], note="%s (%s)" % (d[cls]["note"], "profiles/" + name))
    return {}


def cpu_baseline(model, wl, oracle_dt):
    """CPU baseline of the JSON line (rank 0, N = 1 semantics): the UNMODIFIED reference (baseline/_ref) in a child
    process on the host cores, on a bounded sample of the same workload (2 utterances per step); the numpy oracle
    port timed by parity_check when the reference is not installed."""
    sample = 2
    audio_s = sample * wl["frames"] / FRAMES_PER_AUDIO_SECOND
    if os.path.isdir(REF_DIR) and wl is WORKLOAD:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3",
                                  "--warmup", "1", "--calib-batches", "1"], capture_output=True, text=True, timeout=600,
                                 env={k: v for k, v in os.environ.items()
                                      if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")})
            ref = json.loads(out.stdout.strip().splitlines()[-1])
            return dict(ref["cpu_baseline"], note="child process `bench.py --impl reference --steps 3 --warmup 1 "
                                                  "--calib-batches 1`")
        except Exception as exc:  # noqa: BLE001 -- fall back to the port, say why
            why = "reference child failed: %r" % (exc,)
    else:
        why = "baseline/_ref not installed (tools/install_ref.sh)"
    return dict(value=round(audio_s / oracle_dt, 2), unit="audio-s/s", cores=os.cpu_count(), kind="port",
                sample="%d of %d utterances x %d frames, numpy oracle (float64 BLAS as integer engine), one pass; %s"
                       % (sample, wl["batch"], wl["frames"], why))


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path -- its unmodified QuantAct / QuantConv1d /
    JasperBlock modules from baseline/_ref, assembled into the encoder / decoder by tests/golden/ref_loader.py --
    on the host cores with every thread torch can use.  Same seeded weights (identical to the GPU arm's: both draw
    from torch.manual_seed(0) in the same order), same BN re-estimation, same calibration batches, same warm-up;
    each step is a bounded sample of the workload: 2 of its 64 utterances at full length."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOAD
    warmup = max(args.warmup, 3) if args.calib_batches is None else args.warmup
    torch.set_num_threads(os.cpu_count() or 1)
    sample_batch = 2
    if os.path.isdir(REF_DIR):
        kind = "reference"
        os.environ["QASR_REFERENCE_ROOT"] = REF_DIR
        sys.path[:0] = [os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "q-asr_b200")]
        from ref_loader import RefModel, load_reference
        ref = load_reference()
        from qasr import topology
        from qasr.params import reestimate_bn_, synthetic_features
        qm = ref.quantize_model
        torch.set_grad_enabled(False)
        torch.manual_seed(0)
        model = RefModel(ref, topology.TOPOLOGIES[wl["topology"]]())
        model.eval()
        reestimate_bn_(model.encoder)
        model.set_quant_bit(wl["weight_bit"], "weight")
        model.set_quant_bit(wl["act_bit"], "act")
        qm.set_percentile(model, wl["percentile"])
        model.encoder.bn_folding()
        qm.calibrate(model)
        cb, cc, ct = wl["calib_shape"]
        n_calib = wl["calib_batches"] if args.calib_batches is None else args.calib_batches
        for i in range(n_calib):
            cx, cl = synthetic_features(cb, ct, feat_in=cc, seed=100 + i)
            e, l, s = model.encoder(cx, cl)
            model.decoder(e, s)
        qm.evaluate(model)
        qm.set_dynamic(model, False)
        x, lens = synthetic_features(sample_batch, wl["frames"], seed=1000)

        def step():
            e, l, s = model.encoder(x, lens)
            return model.decoder(e, s).argmax(dim=-1)
        what = ("the unmodified reference modules (baseline/_ref: nemo/quantization/utils + parts/jasper.py) on CPU, "
                "torch %s, %d threads, %s" % (torch.__version__, torch.get_num_threads(), cpu_model_name()))
    else:
        kind = "port"
        _paths(ours=True)
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
        for i in range(wl["calib_batches"] if args.calib_batches is None else args.calib_batches):
            cx, cl = synthetic_features(cb, ct, feat_in=cc, seed=100 + i)
            O.model_forward(spec, cx.numpy(), cl.numpy())
        O.set_mode(spec, False)
        xt, lt = synthetic_features(sample_batch, wl["frames"], seed=1000)
        x, lens = xt.numpy(), lt.numpy()

        def step():
            return O.model_forward(spec, x, lens)["greedy"]
        what = "numpy oracle port of the reference's CPU algorithm (baseline/_ref not installed: tools/install_ref.sh)"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    audio_s = sample_batch * wl["frames"] / FRAMES_PER_AUDIO_SECOND * args.steps
    value = round(audio_s / dt, 2)
    sample = "%d of %d utterances x %d frames per step; %s" % (sample_batch, wl["batch"], wl["frames"], what)
    print(json.dumps(dict(
        impl="reference", metric="audio_seconds_per_second", value=value, unit="audio-s/s",
        n_gpus=world, steps=args.steps, warmup=warmup,
        ms_per_step=round(dt / args.steps * 1e3, 2), higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="int8", data="synthetic", config=make_config(wl, world),
        cpu_baseline=dict(value=value, unit="audio-s/s", cores=torch.get_num_threads(), kind=kind, sample=sample),
        e2e=dict(value=value, unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="quartznet", choices=["quartznet", "jasper"],
                    help="quartznet = BASELINE configs[1] (the metric's configuration, default)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip `secondary` and `dynamic_dp`")
    ap.add_argument("--calib-batches", type=int, default=None,
                    help="reference arm only: calibration batches (default: the workload's; the cpu_baseline child "
                         "uses 1 to stay bounded -- calibration does not change the timed forward's cost)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

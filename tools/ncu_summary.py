"""Summaries of ncu exports: launch list (--metrics gpu__time_duration.sum --csv) and key metrics /
SASS instruction mix of a full-set .ncu-rep."""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][:64]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s n=%3d %10.1f us %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    print("total us %.1f" % (tot / 1e3))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor"]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    for r in rows[2:]:
        print("==", r[h.index("Kernel Name")][:100])
        for w in WANT:
            if w in h:
                print("   %-66s %s %s" % (w, r[h.index(w)], rows[1][h.index(w)]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    ci, si, st = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    agg, samp, tot, ts = defaultdict(float), defaultdict(float), 0.0, 0.0
    items = []
    for r in rows[2:]:
        try:
            v, s = float(r[ci]), float(r[st])
        except (ValueError, IndexError):
            continue
        parts = r[si].strip().split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        agg[op] += v
        samp[op] += s
        tot += v
        ts += s
        items.append((s, v, r[si].strip()[:80]))
    print("   warp instructions %.0f, samples %.0f" % (tot, ts))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
        print("   %-10s %6.2f%% instr %6.2f%% samples" % (k, 100 * v / tot, 100 * samp[k] / max(ts, 1)))
    items.sort(reverse=True)
    for s, v, text in items[:10]:
        print("   %5.1f%% samples %10.0f  %s" % (100 * s / max(ts, 1), v, text))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        if p.endswith(".csv"):
            launches(p)
        else:
            report(p)

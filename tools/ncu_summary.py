"""Summarise ncu outputs (run in the build container; no GPU needed):
    python tools/ncu_summary.py launches gpurun_out/launches.csv          # per-kernel totals and shares
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep           # key metrics of each captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt, allt = collections.defaultdict(float), collections.Counter(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        key = row["Kernel Name"].split("(")[0][:64]
        tot[key] += v
        cnt[key] += 1
        allt += v
    print("%-66s %5s %11s %7s" % ("kernel", "n", "total us", "share"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-66s %5d %11.1f %6.1f%%" % (k, cnt[k], v / 1e3, 100 * v / allt))
    print("%-66s %5d %11.1f" % ("TOTAL (serialised, cold cache: compare shares, not absolutes)", sum(cnt.values()), allt / 1e3))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("== %s" % d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], u[k]))
        stalls = []
        for k, v in d.items():
            if "pcsamp_warps_issue_stalled_" in k and not k.endswith("not_issued"):
                try:
                    stalls.append((float(v.replace(",", "")), k.split("stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1.0
        print("  stall samples: " + ", ".join("%s %.0f%%" % (n, 100 * s / tot) for s, n in sorted(stalls, reverse=True)[:6]))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])

"""Condense ncu outputs into the small text summaries committed under profiles/.
usage: ncu_summary.py launches <launches.csv>   |   ncu_summary.py kernels <report.ncu-rep> [name-filter]"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(u, 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})")
    print("# cold-cache, serialised launches: compare SHARES, not absolutes")
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:44s} {n:8d} {t:10.3f} {t / n:9.3f} {100 * t / tot:6.1f}%")
    print(f"{'TOTAL':44s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")


def kernels(path, flt=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    print(f"# selected metrics from `ncu --set full --clock-control none` ({path}); per launch")
    seen = set()
    for row in r[2:]:
        name = row[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        if flt and flt not in name:
            continue
        if name in seen:
            continue
        seen.add(name)
        print(f"\n[{name}]")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:86s} {row[i]:>18s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernels(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.
    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [pattern ...] > profiles/rNN_kernel.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for x in rows:
        n = re.sub(r"\(.*", "", x["Kernel Name"])[:70]
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        tot[n][0] += 1
        tot[n][1] += v
    T = sum(v[1] for v in tot.values())
    print(f"# {len(rows)} launches, {T / 1e3:.3f} ms total (ncu per-launch times are cold-cache/serialised: compare SHARES)")
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}")
    for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:70s} {c:5d} {t:11.1f} {t / c:9.1f} {t / T:6.3f}")


DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct",
           "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
           "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed.sum ", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__average_warp", "smsp__average_warps_issue_stalled",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "lts__t_bytes.sum ",
           "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum ", "lts__t_sectors_op_red.sum "]


def full(path, pats):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    pats = pats or DEFAULT
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("=" * 100)
        print(re.sub(r"\(.*", "", name), " grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(p.strip() == h or (not p.endswith(" ") and p in h) for p in pats) and r[i] not in ("", "0", "n/a"):
                print(f"  {h:90s} {r[i][:24]:>24s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3:])

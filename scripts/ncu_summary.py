"""Summarise ncu reports / launch lists brought back in gpurun_out/ into markdown for profiles/.
usage: python scripts/ncu_summary.py report <file.ncu-rep> [title]    -> key metrics table
       python scripts/ncu_summary.py launches <launches.csv> [title]   -> per-kernel totals (second half of the list)"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def report(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit = rows[0], rows[1]
    print(f"## {title}\n")
    for val in rows[2:]:
        d = dict(zip(hdr, zip(val, unit)))
        print(f"`{d.get('Kernel Name', ('?',))[0][:60]}` grid {d.get('launch__grid_size', ('?',))[0]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
        print()


def launches(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    half = len(rows) // 2
    d = collections.defaultdict(list)
    for r in rows[half:]:
        d[r[4].split("(")[0]].append(float(r[-1]) / 1e3)
    tot = sum(sum(v) for v in d.values())
    print(f"## {title}\n\nPer-launch times are cold-cache and serialised (ncu replays every launch alone); compare SHARES.\n")
    print("| kernel | launches | total ms | median us | max us | share |\n|---|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda x: -sum(x[1])):
        v2 = sorted(v)
        print(f"| {k} | {len(v)} | {sum(v) / 1e3:.3f} | {v2[len(v) // 2]:.1f} | {v2[-1]:.1f} | {100 * sum(v) / tot:.1f}% |")
    print(f"\ntotal {tot / 1e3:.1f} ms")


if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else path
    (report if kind == "report" else launches)(path, title)

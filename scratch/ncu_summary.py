"""Condense ncu CSV output into the small tables committed under profiles/.

    python scratch/ncu_summary.py launches <csv from --metrics gpu__time_duration.sum --csv>
        -> per kernel name: launches, total ms, share of the captured time
    python scratch/ncu_summary.py raw <csv from `ncu -i rep --page raw --csv`>
        -> the metrics DESIGN.md / profiles/README.md quote (duration, DRAM bytes and throughput, registers, occupancy,
           pipe utilisation, stall reasons per issued instruction), one row per metric
"""
import csv
import sys
from collections import defaultdict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]


def rows(path):
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    return list(csv.reader(lines))


def launches(path):
    r = rows(path)
    hdr = r[0]
    name_i, metric_i, val_i = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    unit_i = hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for x in r[1:]:
        if len(x) <= val_i or x[metric_i] != "gpu__time_duration.sum":
            continue
        v = float(x[val_i].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(x[unit_i], 1e-6)
        k = x[name_i].split("(")[0]
        tot[k] += v
        cnt[k] += 1
    allms = sum(tot.values()) or 1.0
    print("kernel,launches,total_ms,ms_per_launch,share")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k},{cnt[k]},{tot[k]:.3f},{tot[k] / cnt[k]:.4f},{tot[k] / allms:.3f}")


def raw(path):
    r = rows(path)
    hdr, units = r[0], r[1]
    print("metric,unit," + ",".join(f"launch{i}" for i in range(len(r) - 2)))
    if "Kernel Name" in hdr:
        i = hdr.index("Kernel Name")
        print("Kernel Name,," + ",".join('"' + x[i] + '"' for x in r[2:]))
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            print(f"{m},{units[i]}," + ",".join(x[i] for x in r[2:]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])

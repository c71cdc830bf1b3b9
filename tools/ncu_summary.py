#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (development aid).

  python tools/ncu_summary.py launches <launch-list.csv>      per-kernel count / total / share / average of a
                                                               `ncu --metrics gpu__time_duration.sum --csv` log
  python tools/ncu_summary.py full <raw.csv> [...]             selected metrics per captured kernel of
                                                               `ncu -i x.ncu-rep --page raw --csv` dumps
"""
import collections
import csv
import sys

METRICS = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "inst_executed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
)


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 6]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if r is hdr or len(r) != len(hdr) or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0]
        tot[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    for name, t in tot.most_common():
        print(f"{name:48s} n={cnt[name]:4d} total={t / 1e6:9.3f} ms share={t / total:6.1%} avg={t / cnt[name] / 1e3:9.1f} us")


def full(paths: list[str]) -> None:
    for path in paths:
        rows = list(csv.reader(open(path, errors="replace")))
        hdr = next(r for r in rows if "Kernel Name" in r)
        units = rows[rows.index(hdr) + 1]
        for r in rows[rows.index(hdr) + 2:]:
            if len(r) != len(hdr):
                continue
            print("--")
            print(f"  Kernel Name = {r[hdr.index('Kernel Name')]}")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    print(f"  {m} [{units[i]}] = {r[i]}")


def to_json(source: str, elements: int, zones: int, window: int, forcing: str, paths: list[str]) -> None:
    """profiles/ncu_kernels.json for bench.py: FP64-pipe utilisation and DRAM bytes per launch of the land kernels."""
    import json

    kernels: dict[str, dict] = {}
    for path in paths:
        rows = list(csv.reader(open(path, errors="replace")))
        hdr = next(r for r in rows if "Kernel Name" in r)
        units = rows[rows.index(hdr) + 1]

        def val(r: list[str], name: str) -> float:
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}
            return v * scale.get(units[i], 1.0)

        for r in rows[rows.index(hdr) + 2:]:
            if len(r) != len(hdr):
                continue
            name = r[hdr.index("Kernel Name")]
            key = "zone_kernel" if "zone_kernel" in name else "element_kernel" if "element_kernel" in name else \
                "reach_wave_kernel" if "reach_wave_kernel" in name else None
            if key is None or key in kernels:
                continue
            kernels[key] = {
                "name": name, "ms": val(r, "gpu__time_duration.sum"),
                "fp64_pipe_pct": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                "inst_executed": val(r, "inst_executed"),
                "registers": val(r, "launch__registers_per_thread"),
            }
    print(json.dumps({"source": source, "workload": {"elements": elements, "zones": zones, "window": window,
                                                     "forcing": forcing}, "kernels": kernels}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "json":   # json <source note> <elements> <zones> <window> <forcing> <raw.csv> [...]
        to_json(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6], sys.argv[7:])
    else:
        full(sys.argv[2:])

#!/usr/bin/env python
"""Table of the metrics that matter from `ncu -i X.ncu-rep --page raw --csv` (one row per captured launch).
Usage: ncu -i rep --page raw --csv | python tools/ncu_full_summary.py "title" > profiles/xxx.md"""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum"]


def main():
    title = sys.argv[1] if len(sys.argv) > 1 else "ncu --set full"
    rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
    head, units = rows[0], rows[1]
    cols = []
    for w in WANT:
        hits = [i for i, h in enumerate(head) if h == w or h.endswith("." + w) or (w in h and w.startswith("sm__pipe_tensor"))]
        if hits:
            cols.append(hits[0])
    print(f"# {title}\n")
    print("| " + " | ".join(f"{head[i]} [{units[i]}]" for i in cols) + " |")
    print("|" + "---|" * len(cols))
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        print("| " + " | ".join((r[i][:48] if head[i] == "Kernel Name" else r[i]) for i in cols) + " |")


if __name__ == "__main__":
    main()

"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output (one block per kernel launch).
Usage: ncu -i rep --page raw --csv | python tools/ncu_raw.py [substring ...]"""
import csv
import sys

DEFAULT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct",
           "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fma", "sm__pipe_fma", "sm__inst_executed_pipe_alu",
           "smsp__inst_executed.sum ", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "sm__throughput.avg.pct",
           "smsp__average_warp", "smsp__warp_issue_stalled", "local_"]


def main():
    want = sys.argv[1:] or DEFAULT
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        for i, h in enumerate(hdr):
            if any(w in h for w in want) and i < len(r) and r[i] not in ("", "n/a"):
                print(f"{h} = {r[i]} {units[i]}")
        print("-" * 60)


if __name__ == "__main__":
    main()

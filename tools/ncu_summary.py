"""Condense an `ncu --set full` report into the handful of numbers DESIGN.md / bench.py cite.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.json]   (needs ncu on PATH; no GPU)"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_registers": "blocks_per_sm_by_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for i, h in enumerate(hdr):
            if h in KEYS and r[i] not in ("", "n/a"):
                v = float(r[i].replace(",", ""))
                u = units[i]
                if u in SCALE:
                    v *= SCALE[u]
                d[KEYS[h]] = v
        if "dram_read" in d:
            d["dram_bytes"] = d["dram_read"] + d.get("dram_write", 0.0)
        out.append(d)
    txt = json.dumps({"report": rep.split("/")[-1], "launches": out}, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()

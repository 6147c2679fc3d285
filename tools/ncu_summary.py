#!/usr/bin/env python
"""Developer tool: condense an .ncu-rep (ncu --set full) into the per-kernel figures the rooflines use.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_ncu_x.json          (runs here: ncu -i needs no GPU)"""
import csv, io, json, subprocess, sys
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    e = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
    for m, name in WANT.items():
        if m in hdr:
            i = hdr.index(m)
            try:
                e[name] = float(r[i].replace(",", ""))
            except ValueError:
                e[name] = r[i]
            e[name + "_unit"] = units[i]
    out.append(e)
print(json.dumps({"source": sys.argv[1].split("/")[-1], "how": "ncu --set full --clock-control none (cold-cache, serialised replays)", "kernels": out}, indent=1))

"""Summarise the `ncu --set full` captures of the stand-alone memory-bound kernels (tools/gpu_ncu_micro_final.sh) into
profiles/rNN_micro_ncu_summary.txt: duration, DRAM bytes, occupancy, issue activity and the main stall reasons.
    python tools/ncu_micro_summary.py gpurun_out/r02f_micro_*.ncu-rep > profiles/r02_micro_ncu_summary_final.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle"]

print("ncu --set full --clock-control none, launches of the stand-alone memory-bound kernels in their FINAL round-2 state "
      "(tools/bench_micro.py 22 = 4.2 M samples, 65 536 rays; tools/gpu_ncu_micro_final.sh)\n")
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    for r in body:
        print("==", r[col["Kernel Name"]][:110])
        for k in KEYS:
            if k in col:
                print(f"   {k:75s} {r[col[k]]:>16s} {units[col[k]]}")

"""Condenses the ncu raw / source pages of one kernel: headline counters and the SASS lines that hold the stall samples."""
import csv
import sys

base = sys.argv[1]
rows = list(csv.reader(open(base + "_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
for k, u, v in zip(hdr, units, vals):
    if k in want:
        print(f"{k:70s} {v} {u}")
src = list(csv.reader(open(base + "_source.csv")))[2:]
tot = sum(int(r[2]) for r in src)
print("total samples", tot)
top = sorted(range(len(src)), key=lambda i: -int(src[i][2]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for i in sorted(top):
    print(f"{i:5d} {src[i][1].strip()[:80]:80s} {src[i][2]:>5s} {src[i][5]:>8s}")

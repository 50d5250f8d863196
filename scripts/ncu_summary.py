"""Condenses `ncu --page raw --csv` exports (gpurun_out/*_raw.csv) into the small tables committed under profiles/,
and a launch list (`--metrics gpu__time_duration.sum`) into per-kernel totals and shares.

    python scripts/ncu_summary.py raw  gpurun_out/prof_x_raw.csv  profiles/r1_ncu_full_x.csv
    python scripts/ncu_summary.py list gpurun_out/launches.csv    profiles/r1_ncu_launch_list_summary.csv
"""

import csv
import re
import sys
from collections import OrderedDict

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "").replace("<unnamed>::", "").strip()


def raw(src: str, dst: str) -> None:
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals = [short(r[i]) if k == "Kernel Name" else r[i] for r in rows[2:]]
                w.writerow([k, units[i]] + vals)


def launch_list(src: str, dst: str) -> None:
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    tot: "OrderedDict[str, list]" = OrderedDict()
    for r in rows:
        k = short(r[4])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r[14].replace(",", "")) / 1e3  # ns -> us
    total = sum(v[1] for v in tot.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share"])
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, round(us, 1), round(us / total, 4)])
        w.writerow(["TOTAL", sum(v[0] for v in tot.values()), round(total, 1), 1.0])


if __name__ == "__main__":
    {"raw": raw, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])

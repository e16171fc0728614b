"""Turns `ncu -i X.ncu-rep --page raw --csv` into the markdown table kept under
profiles/ (one row per kernel: the LAST captured launch of each name).

    python tools/ncu_table.py gpurun_out/prof_raw.csv > profiles/rNN_table.md
"""
import csv
import re
import sys

COLS = [
    ("time µs", "gpu__time_duration.sum", 1.0, "{:.1f}"),
    ("DRAM read GB", "dram__bytes_read.sum", None, "{:.4f}"),
    ("DRAM write MB", "dram__bytes_write.sum", None, "{:.2f}"),
    ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
    ("regs", "launch__registers_per_thread", 1.0, "{:.0f}"),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
    ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
    ("fp64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
    ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
    ("stall long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0, "{:.1f}"),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0, "{:.1f}"),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    last = {}
    for r in data:
        name = r[ci["Kernel Name"]]
        name = re.sub(r"^void (accblas::)?(<?unnamed>::)?", "", name)
        name = re.sub(r"\(.*$", "", name).replace("(bool)", "")
        last[name] = r
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for name, r in last.items():
        cells = []
        for title, metric, _, fmt in COLS:
            if metric not in ci:
                cells.append("n/a")
                continue
            v = float(r[ci[metric]].replace(",", ""))
            unit = units[ci[metric]]
            if title == "time µs":
                v *= UNIT_SCALE.get(unit, 1.0)
            elif "GB" in title:
                v *= UNIT_SCALE.get(unit, 1.0) / 1e9
            elif "MB" in title:
                v *= UNIT_SCALE.get(unit, 1.0) / 1e6
            cells.append(fmt.format(v))
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])

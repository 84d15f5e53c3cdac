"""Markdown table from an ncu report: `python tools/ncu_table.py X.ncu-rep [which]` (which = 2: the second launch of every kernel
name, the default; 0 = every launch). Runs `ncu -i X --page raw --csv` (works without a GPU)."""
import csv
import io
import json
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time us", 1e-3), ("dram__bytes_read.sum", "DRAM read MB", 1e-6), ("dram__bytes_write.sum", "DRAM write MB", 1e-6),
        ("sm__inst_executed.avg.per_cycle_active", "IPC/SM", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %", 1),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe %", 1),
        ("launch__registers_per_thread", "regs", 1), ("smsp__inst_executed.sum", "warp instr M", 1e-6)]


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    I = {h: i for i, h in enumerate(hdr)}
    seen, out, js = {}, [], {}
    for r in rows[2:]:
        name = re.sub(r"\(.*$", "", r[I["Kernel Name"]])
        full = r[I["Kernel Name"]]
        seen[full] = seen.get(full, 0) + 1
        if which and seen[full] != which:
            continue
        vals = []
        rec = {}
        for key, label, sc in COLS:
            v = ""
            if key in I and r[I[key]] not in ("", "n/a"):
                x = float(r[I[key]].replace(",", ""))
                u = units[I[key]]
                if key.startswith("gpu__time") and u in ("us", "usecond"):
                    x *= 1e3
                if key.startswith("dram__bytes") and u == "Mbyte":
                    x *= 1e6
                elif key.startswith("dram__bytes") and u == "Kbyte":
                    x *= 1e3
                elif key.startswith("dram__bytes") and u == "Gbyte":
                    x *= 1e9
                v = f"{x * sc:.1f}"
                rec[label] = round(x * sc, 3)
            vals.append(v)
        grid = r[I["Grid Size"]] if "Grid Size" in I else ""
        out.append((full[:110], grid, vals))
        js[full + (f"#{seen[full]}" if not which else "")] = rec
    print("| kernel | grid | " + " | ".join(l for _, l, _ in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for name, grid, vals in out:
        print(f"| `{name}` | {grid} | " + " | ".join(vals) + " |")
    if len(sys.argv) > 3:
        json.dump(js, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()

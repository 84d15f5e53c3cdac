"""Summarise an ncu report: per kernel launch the headline counters and stall reasons, and the top stalled SASS lines.
`python tools/ncu_stalls.py report.ncu-rep [kernel-regex] [launch-index]`"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ni = hdr.index("Kernel Name")
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__grid_size", "launch__waves_per_multiprocessor"]
for k, r in enumerate(rows[2:]):
    if rx and rx not in r[ni]:
        continue
    print(f"[{k}] {r[ni][:90]}")
    for key in keys:
        if key in hdr:
            print(f"    {key} = {r[hdr.index(key)]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
           h.startswith("smsp__average_warp_latency_issue_stalled_"):
            try:
                st.append((float(r[i].replace(",", "")), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("    stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
if len(sys.argv) > 3:
    idx = sys.argv[3]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", idx, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hi]
    ci = {n: i for i, n in enumerate(h)}
    body = rows[hi + 1:]

    def f(r, k):
        try:
            return float(r[ci[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot = sum(f(r, "# Samples") for r in body)
    print("samples", tot, "sass lines", len(body))
    top = sorted(range(len(body)), key=lambda i: -f(body[i], "# Samples"))[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]
    for i in top:
        r = body[i]
        reasons = sorted(((f(r, k), k) for k in ci if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
        print(f"{i:5d} {int(f(r, '# Samples')):6d} x{int(f(r, 'Instructions Executed')):8d} {reasons[0][1]}:{int(reasons[0][0])} {reasons[1][1]}:{int(reasons[1][0])}  {r[ci['Source']][:90]}")

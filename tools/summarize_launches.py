"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X python bench.py ...`):
`python tools/summarize_launches.py X.csv [last_n_launches]` prints a markdown table (kernel, launches, total us, share)."""
import csv, collections, re, sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        rows.append((r["Kernel Name"], us))
    if len(sys.argv) > 2:
        rows = rows[-int(sys.argv[2]):]
    tot = sum(us for _, us in rows)
    agg = collections.OrderedDict()
    for k, us in rows:
        k = re.sub(r"\(.*$", "", k)[:100]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    ours = sum(v[1] for k, v in agg.items() if "bem::" in k)
    print(f"{len(rows)} launches, {tot / 1e3:.2f} ms summed kernel time (serialised, cold cache); libbem_b200.so kernels: {100 * ours / tot:.1f} % of it.\n")
    print("| kernel | launches | total us | share | ours |\n|---|---:|---:|---:|:-:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us:.0f} | {100 * us / tot:.1f}% | {'x' if 'bem::' in k else ''} |")


if __name__ == "__main__":
    main()

"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) into per-kernel totals.

    python profiles/summarize_launches.py gpurun_out/launches_r01.csv [--skip N] > profiles/launches_r01.md

ncu's per-launch durations are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit, val = r["Metric Unit"], float(r["Metric Value"].replace(",", ""))
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    rows = rows[skip:]
    tot = defaultdict(lambda: [0, 0.0])
    for _, name, ns in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        tot[short][0] += 1
        tot[short][1] += ns
    total = sum(v[1] for v in tot.values())
    print(f"launches: {len(rows)}  total kernel time: {total / 1e6:.3f} ms  (source: {path}, skipped first {skip})\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.2f} |")


if __name__ == "__main__":
    main()

"""`ncu -i X.ncu-rep --page raw --csv` -> a compact markdown table (one row per captured launch).

    python profiles/ncu_summary.py raw.csv > profiles/ncu_<what>_r02.md"""
import csv
import sys

COLS = [("Kernel Name", "kernel", 70), ("launch__grid_size", "grid", 0), ("launch__cluster_size", "cluster", 0),
        ("launch__block_size", "block", 0), ("launch__registers_per_thread", "regs", 0),
        ("gpu__time_duration.sum", "time us", 0), ("dram__bytes_read.sum", "DRAM read MB", 0),
        ("dram__bytes_write.sum", "DRAM write MB", 0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % peak", 0),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %", 0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", 0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 0)]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    units = rows[1]
    idx = {}
    for name, _, _ in COLS:
        for i, h in enumerate(hdr):
            if h == name:
                idx[name] = i
                break
    print("| " + " | ".join(t for n, t, _ in COLS if n in idx) + " |")
    print("|" + "---|" * sum(1 for n, _, _ in COLS if n in idx))
    for r in rows[2:]:
        if not r:
            continue
        out = []
        for name, _, width in COLS:
            if name not in idx:
                continue
            v = r[idx[name]]
            u = units[idx[name]]
            if name == "Kernel Name":
                v = "`" + v.replace("void ", "").replace("tpp::", "")[:width] + "`"
            else:
                try:
                    f = float(v.replace(",", ""))
                    if u in ("byte", "Byte"):
                        f /= 1e6
                    elif u == "Kbyte":
                        f /= 1e3
                    elif u == "Gbyte":
                        f *= 1e3
                    if u in ("ns", "nsecond"):
                        f /= 1e3
                    elif u in ("ms", "msecond"):
                        f *= 1e3
                    v = f"{f:.1f}" if abs(f) < 1e5 else f"{f:.0f}"
                except ValueError:
                    pass
            out.append(v)
        print("| " + " | ".join(out) + " |")


if __name__ == "__main__":
    main()

"""Summarise an `ncu --page raw --csv` export: one line per profiled launch with the counters the
DESIGN.md roofline table cites.  python tools/ncu_summary.py file.csv [...]"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "rdMB", None),
    ("dram__bytes_write.sum", "wrMB", None),
    (("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), "dram%", 1),
    ("lts__t_sector_hit_rate.pct", "L2hit%", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    (("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
      "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "tensor%", 1),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%", 1),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
]


def to_mb(v, unit):
    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, None)
    return v * f if f else v


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"== {path}")
        print("kernel".ljust(44) + " ".join(n.rjust(8) for _, n, _ in COLS))
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            name = name.replace("void ", "").replace("rtdf::", "").replace("<unnamed>::", "").split("(")[0][:43]
            out = []
            for key, label, scale in COLS:
                if isinstance(key, tuple):     # first spelling of the metric this ncu version exported
                    key = next((k for k in key if k in idx), key[0])
                if key not in idx or r[idx[key]] == "":
                    out.append("-".rjust(8))
                    continue
                v = float(r[idx[key]].replace(",", ""))
                u = units[idx[key]]
                if label in ("rdMB", "wrMB"):
                    v = to_mb(v, u)
                elif label == "us":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
                out.append(f"{v:8.1f}")
            print(name.ljust(44) + " ".join(out))


if __name__ == "__main__":
    main()

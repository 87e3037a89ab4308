"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of ONE forward
(the launches between two conv-0 kernels: conv0_kernel | conv0_im2col_kernel | conv0_gn_stats_kernel).  python tools/launch_summary.py launches.csv [forward_index]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = [(r["Kernel Name"], float(r["Metric Value"]) / 1000, r["Grid Size"]) for r in csv.DictReader(lines)]
    idx = [i for i, (k, _, _) in enumerate(rows) if "conv0_kernel" in k or "conv0_im2col" in k or "conv0_gn_stats" in k]
    lo = idx[which]
    hi = idx[which + 1] if which + 1 < len(idx) else len(rows)
    fw = rows[lo:hi]
    print(f"{path}: forward #{which}: {len(fw)} kernels, sum {sum(t for _, t, _ in fw):.1f} us")
    agg = collections.OrderedDict()
    for k, t, g in fw:
        k = re.sub(r"\(.*", "", k).replace("void ", "").replace("rtdf::", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0, g])
        a[0] += 1
        a[1] += t
    for k, (n, t, g) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:9.1f} us {n:4d} x {t / n:8.1f}  grid {g:>14s}  {k[:90]}")


if __name__ == "__main__":
    main()

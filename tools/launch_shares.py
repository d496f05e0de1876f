"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV).  python tools/launch_shares.py file.csv [--last-step N]
--tail N: only the last N launches (one step of the bench = the launches after the final input copy)."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    tail = int(sys.argv[sys.argv.index("--tail") + 1]) if "--tail" in sys.argv else 0
    by_grid = "--grid" in sys.argv
    ours = "--ours" in sys.argv          # only this library's kernels (drops torch fills and bench.py's cuBLAS peak probe)
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    if tail:
        rows = rows[-tail:]
    if ours:
        rows = [r for r in rows if "ctx::" in r[ik] or "adam_schedule" in r[ik]]
    agg = OrderedDict()
    tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "")
        if by_grid:
            name += " " + r[ig]
        us = float(r[iv].replace(",", "")) * 1e-3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
        tot += us
    print("%d launches, %.2f ms in total" % (len(rows), tot * 1e-3))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-80s %4d launches %10.1f us %5.1f%%" % (k[:80], n, us, 100 * us / tot))


if __name__ == "__main__":
    main()

"""Print a per-launch table from an `ncu --csv --metrics ...` log (dev tool)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
cur = {}
for r in rows[1:]:
    name = r[ki].replace("void ", "").replace("<unnamed>::", "")
    cur.setdefault((int(r[ii]), name[:44]), {})[r[mi]] = r[vi]
for (i, name), v in sorted(cur.items()):
    parts = []
    for k, x in v.items():
        short = k.split(".")[0].replace("smsp__", "").replace("gpu__", "").replace("sm__", "").replace("dram__", "")
        try:
            f = float(x.replace(",", ""))
            x = f"{f/1e6:.3f}M" if f > 1e5 else f"{f:.2f}"
        except ValueError:
            pass
        parts.append(f"{short}={x}")
    print(f"{i:3d} {name:44s} " + "  ".join(parts))

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into kernel, launches, total, mean, share:
    python tools/launch_summary.py gpurun_out/launches.csv profiles/r1_xxx_launches_summary.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    iname, ival = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[iname]).replace("smplb200::<unnamed>::", "").replace("void ", "")
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + float(r[ival].replace(",", "")) / 1e3)
    lib = sum(t for k, (n, t) in agg.items() if not k.startswith("at::") and "cub::" not in k)
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,us_per_launch,share_of_library_kernels\n")
        for k, (n, t) in agg.items():
            own = not k.startswith("at::") and "cub::" not in k
            f.write("%s,%d,%.1f,%.1f,%s\n" % (k[:60], n, t, t / n, ("%.4f" % (t / lib)) if own else "nan"))
    print(open(dst).read())


if __name__ == "__main__":
    main()

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel, launches, avg us, total us, share.

    python tools/launch_summary.py gpurun_out/launches.csv "<command that produced it>" > profiles/<name>.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*$", "", r[ki]).replace("(bool)", "")
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("# %s (cold-cache, serialised: compare shares)" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("kernel,launches,avg_us,total_us,share_pct")
for k, v in tot.most_common():
    print("%s,%d,%.2f,%.1f,%.1f" % (k, cnt[k], v / cnt[k], v, 100 * v / total))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into shares per kernel.
usage: python tools/launch_summary.py gpurun_out/launches_c2.csv "<command that was profiled>" > profiles/r01_launches_c2.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[hdr.index("Metric Name")] == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])[:100]
        tot[name] += float(r[hdr.index("Metric Value")].replace(",", ""))
        cnt[name] += 1
T = sum(tot.values()) or 1
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# (cold-cache, serialised launches: shares of device time, not absolute times)")
for k, v in tot.most_common():
    print("%6.2f%%  %4d launches  %9.1f us avg  %s" % (100 * v / T, cnt[k], v / 1e3 / cnt[k], k))

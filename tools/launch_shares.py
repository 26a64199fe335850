"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_shares.py launches.csv > profiles/xxx_launch_shares.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[i_val].replace(",", ""))
    v = v / 1e3 if r[i_unit] == "ns" else v * 1e3 if r[i_unit] == "ms" else v
    a = agg.setdefault(r[i_name][:60], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | mean us | share |\n|---|---|---|---|")
for k, (n, t) in agg.items():
    print("| `%s` | %d | %.1f | %.1f %% |" % (k, n, t / n, 100 * t / tot))

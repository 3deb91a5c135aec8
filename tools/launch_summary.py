"""Per-kernel shares of one search round from an ncu launch list (--metrics gpu__time_duration.sum --csv):
   python tools/launch_summary.py gpurun_out/launches.csv [round_index_from_the_end=1]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [(r[ki], float(r[vi].replace(",", ""))) for r in data if len(r) > vi]
idx = [i for i, (n, _) in enumerate(names) if n.startswith("k_select")]
back = int(sys.argv[2]) if len(sys.argv) > 2 else 1
seg = names[idx[-1 - back]:idx[-back]]
tot = sum(v for _, v in seg)
agg = collections.OrderedDict()
for n, v in seg:
    a = agg.setdefault(n[:86], [0, 0.0])
    a[0] += 1
    a[1] += v
print("kernel,launches,total_us,avg_us,share_pct")
for k, (c, v) in agg.items():
    print("%s,%d,%.1f,%.1f,%.1f" % (k.replace(",", ";"), c, v / 1000, v / 1000 / c, 100 * v / tot))
print("# round total us %.1f over %d launches" % (tot / 1000, len(seg)))

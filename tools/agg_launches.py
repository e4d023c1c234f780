"""Aggregate an ncu --metrics gpu__time_duration.sum launch list (CSV) by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = d["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    name = re.sub(r"\(.*", "", d["Kernel Name"])[:80]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-82s %5d %10.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))

"""Sum an `ncu --csv` launch list by kernel name: count, total us, share."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[mu], 1.0)
    name = re.sub(r"\(.*", "", r[kn]); tot[name][0] += 1; tot[name][1] += v
total = sum(v[1] for v in tot.values())
print("%-60s %8s %12s %7s" % ("kernel", "launches", "total us", "share"))
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %12.1f %6.1f%%" % (k[:60], c, t, 100 * t / total))
print("%-60s %8d %12.1f" % ("all", sum(v[0] for v in tot.values()), total))

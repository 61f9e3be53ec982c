"""Aggregate an ncu --metrics gpu__time_duration.sum --csv launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for row in r:
    if len(row) <= vi:
        continue
    name = re.sub(r"\(.*", "", row[ki]).replace("b200::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    try:
        v = float(row[vi].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:48]:48s} n={a[0]:4d} total={a[1] / 1e3:9.1f} us avg={a[1] / a[0] / 1e3:8.2f} us share={a[1] / tot * 100:5.1f}%")

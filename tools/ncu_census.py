"""Aggregate an ncu --metrics gpu__time_duration.sum CSV by kernel name: count, total ms, share."""
import csv, sys, collections, re
rows, hdr = [], None
for r in csv.reader(open(sys.argv[1])):
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    rows.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"<.*", "", r["Kernel Name"].split("(")[0]).replace("void ", "").strip()[-48:]
    d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += float(r["Metric Value"].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms total")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1]:9.3f} ms {100*v[1]/tot:5.1f}% {v[0]:6d} x {k}")

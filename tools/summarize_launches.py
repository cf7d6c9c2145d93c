"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, io, sys
rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
r = list(csv.DictReader(io.StringIO("".join(rows))))
agg = collections.defaultdict(lambda: [0, 0.0])
for x in r:
    if x["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = x["Kernel Name"][:78]
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    v = v / 1000 if u in ("nsecond", "ns") else v * 1000 if u in ("msecond", "ms") else v
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'total us':>10} {'n':>5} {'avg us':>9} {'share':>6}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} {v[0]:5d} {v[1] / v[0]:9.1f} {100 * v[1] / tot:5.1f}%  {k}")
print(f"{tot:10.1f} total over {sum(v[0] for v in agg.values())} launches")

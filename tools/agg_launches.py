"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and launch shape)."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("uv::", "")
    key = f'{name} grid={row["Grid Size"]} block={row["Block Size"]}'
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
byk = collections.defaultdict(float)
for k, (n, t) in agg.items():
    byk[k.split(" grid=")[0]] += t
print(f"total {tot/1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches")
print("-- by kernel")
for k, t in sorted(byk.items(), key=lambda kv: -kv[1]):
    print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}%  {k}")
print("-- top launch shapes")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}% n={n:4d} avg={t/n:8.1f} us  {k}")

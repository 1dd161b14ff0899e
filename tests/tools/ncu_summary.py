"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (manual tool)."""
import collections, csv, re, sys
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    n += 1
    if n <= skip:
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = row["Kernel Name"]
    m = re.search(r"(gemm_kernel|gemm_pair_kernel)<([^>]*)>", name)
    short = f"{m.group(1)}<{m.group(2)}>" if m else re.sub(r"\(.*", "", name)[:64]
    agg[short][0] += 1
    agg[short][1] += v
tot = sum(t for _, t in agg.values())
print(f"{'kernel':66s} {'n':>5s} {'total ms':>9s} {'avg us':>9s} {'share':>6s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{k:66s} {c:5d} {t/1e3:9.3f} {t/c:9.1f} {t/tot*100:5.1f}%")
print(f"total {tot/1e3:.3f} ms over {sum(c for c, _ in agg.values())} launches")

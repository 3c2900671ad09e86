"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
        name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"<.*", "", name)
        rows.append((int(r["ID"]), name, ns))
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if last:
    rows = rows[-last:]
tot = sum(r[2] for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for _, n, ns in rows:
    agg[n][0] += 1; agg[n][1] += ns
print("launches %d  total %.3f ms (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), tot / 1e6))
print("%-70s %6s %10s %7s" % ("kernel", "count", "total_ms", "share"))
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-70s %6d %10.3f %6.1f%%" % (n[-70:], c, ns / 1e6, 100 * ns / tot))

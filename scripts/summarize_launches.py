"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total / min / median / max
microseconds and the share of all kernel time.    python scripts/summarize_launches.py in.csv "header comment" > out.csv"""
import csv, re, statistics, sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = defaultdict(list)
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = re.sub(r"\(.*", "", r[ik])
    t[name].append(v)
tot = sum(sum(v) for v in t.values())
for c in sys.argv[2:]:
    print("# " + c)
print("kernel,launches,total_us,share_of_all,min_us,median_us,max_us")
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k},{len(v)},{sum(v):.1f},{sum(v) / tot:.4f},{min(v):.2f},{statistics.median(v):.2f},{max(v):.2f}")

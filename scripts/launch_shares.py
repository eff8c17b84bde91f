"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel time shares."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("bd::", "")
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {total / 1e3:.2f} ms total (ncu-serialised, cold cache: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:70]:70s} {cnt[k]:8d} {tot[k] / 1e3:10.3f} {tot[k] / cnt[k]:9.1f} {100 * tot[k] / total:6.1f}%")

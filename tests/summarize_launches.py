"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel name: count, total, share."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = {n: i for i, n in enumerate(r)}
        continue
    if len(r) < len(hdr):
        continue
    name = r[hdr["Kernel Name"]]
    unit = r[hdr["Metric Unit"]]
    val = float(r[hdr["Metric Value"]].replace(",", ""))
    us = val / 1e3 if unit in ("ns", "nsecond") else val * (1e3 if unit in ("ms", "msecond") else 1.0)
    name = re.sub(r"<.*", "", name)
    name = re.sub(r"\(.*", "", name)
    agg[name][0] += 1; agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us/1e3:9.3f} ms  {100*us/tot:5.1f}%  {n:5d} launches  {us/n:9.1f} us avg  {name}")

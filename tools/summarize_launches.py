"""Summarise an ncu --csv launch list per kernel name: count, total/avg duration, share, DRAM bytes per launch (when collected).
Writes profiles-style text to stdout and, with a second argument, the GEMM's mean DRAM traffic per launch as JSON."""
import csv, sys, collections, re, json
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
per_id = {}
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = {n: i for i, n in enumerate(r)}
        continue
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r[hdr["Kernel Name"]]))
    metric, unit = r[hdr["Metric Name"]], r[hdr["Metric Unit"]]
    val = float(r[hdr["Metric Value"]].replace(",", ""))
    a = agg[name]
    if metric == "gpu__time_duration.sum":
        us = val / 1e3 if unit in ("ns", "nsecond") else val * (1e3 if unit in ("ms", "msecond") else 1.0)
        a["n"] += 1; a["us"] += us
    else:
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        a["rd" if "read" in metric else "wr"] += val * mult
tot = sum(v["us"] for v in agg.values())
print(f"total kernel time {tot/1e3:.3f} ms over {sum(v['n'] for v in agg.values())} launches (ncu, serialised, --clock-control none)")
for name, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:40]:
    n = max(v["n"], 1)
    print(f"{v['us']/1e3:9.3f} ms  {100*v['us']/tot:5.1f}%  {v['n']:5d} launches  {v['us']/n:9.1f} us avg  dram rd {v['rd']/n/1e6:8.2f} MB wr {v['wr']/n/1e6:8.2f} MB /launch  {name}")
if len(sys.argv) > 2:
    g = [v for k, v in agg.items() if any(t in k for t in ("gemm_tc", "gemm_pair", "mlp_chain", "gemm_ln"))]      # the tcgen05 GEMM family
    if g:
        v = {f: sum(x[f] for x in g) for f in ("n", "us", "rd", "wr")}
        json.dump({"kernel": "tcgen05 GEMM family: gemm_tc_kernel, gemm_pair_kernel, mlp_chain_fwd_kernel, gemm_tc_persist_kernel, gemm_ln_kernel", "launches": v["n"], "dram_bytes_per_launch": (v["rd"] + v["wr"]) / max(v["n"], 1),
                   "dram_read_bytes_per_launch": v["rd"] / max(v["n"], 1), "dram_write_bytes_per_launch": v["wr"] / max(v["n"], 1),
                   "avg_us": v["us"] / max(v["n"], 1), "share_of_kernel_time": v["us"] / tot,
                   "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one eager step"},
                  open(sys.argv[2], "w"), indent=1)

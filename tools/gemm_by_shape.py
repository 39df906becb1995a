"""Join the ncu launch list (gemm_tc_kernel rows, in launch order) with the logged GEMM shapes of the same step."""
import csv, json, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
shapes = json.load(open(sys.argv[2]))
hdr = None; durs = []
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = {n: i for i, n in enumerate(r)}
        continue
    if len(r) < len(hdr): continue
    if "gemm_tc_kernel" in r[hdr["Kernel Name"]] or "gemm_simt_kernel" in r[hdr["Kernel Name"]]:
        durs.append((float(r[hdr["Metric Value"]].replace(",", "")) / 1e3, r[hdr["Grid Size"]]))
assert len(durs) == len(shapes), (len(durs), len(shapes))
agg = collections.OrderedDict()
for (us, grid), sh in zip(durs, shapes):
    key = tuple(sh) + (grid,)
    d = agg.setdefault(key, [0.0, 0]); d[0] += us; d[1] += 1
tot = sum(v[0] for v in agg.values()); fl = 0
print(f"GEMM total {tot/1e3:.3f} ms over {len(durs)} launches")
print("   ms    calls   us/call   TFLOP/s   M N K tA tW cdt acc grid")
for key, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    M, N, K = key[0], key[1], key[2]
    print(f"{us/1e3:7.3f} {n:5d} {us/n:9.1f} {2.0*M*N*K*n/us/1e6:9.1f}   {key}")

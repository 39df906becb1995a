"""Compact digest of an `ncu --page raw --csv` export: one block per profiled launch with the metrics the roofline discussion uses."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, units = rows[0], rows[1]
idx = {n: i for i, n in enumerate(hdr)}
KEYS = [
    ("duration", "gpu__time_duration.sum"), ("grid", "Grid Size"), ("block", "Block Size"), ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram throughput %peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 throughput %peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("L1 throughput %peak", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active %", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("bf16 MMA ops %peak", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warp instructions", "smsp__inst_executed.sum"),
    ("achieved warps %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"), ("L2->L1 bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
]
stalls = [k for k in hdr if "issue_stalled" in k and k.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print(r[idx["Kernel Name"]][:110])
    for label, k in KEYS:
        if k in idx and r[idx[k]] != "":
            print(f"    {label:24s} {r[idx[k]]:>22s} {units[idx[k]]}")
    items = []
    for k in stalls:
        try:
            items.append((float(r[idx[k]]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
    print("    top stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(items, reverse=True)[:4]))

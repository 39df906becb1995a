#!/bin/bash
# Profiling pass on the GPU box (one B200).  Outputs -> gpurun_out/ (small text/CSV only; copied into profiles/ afterwards).
#  1. ncu launch list of one eager step: duration + DRAM bytes per launch (--clock-control none)
#  2. per-kernel summary and the GEMM's mean DRAM traffic per launch (roofline.traffic of bench.py)
#  3. ncu --set full of the persistent RSSM scan kernels, raw page exported as CSV
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_step.csv python tools/ncu_step.py 16 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step.csv gpurun_out/gemm_traffic.json > gpurun_out/launches_summary.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:rssm_scan -c 4 -f -o gpurun_out/scan_full \
    python tools/ncu_step.py 16 > gpurun_out/ncu_scan.log 2>&1
ncu -i gpurun_out/scan_full.ncu-rep --page raw --csv > gpurun_out/scan_full_raw.csv 2>> gpurun_out/ncu_scan.log
rm -f gpurun_out/scan_full.ncu-rep
cat gpurun_out/launches_summary.txt | head -30
du -sh gpurun_out

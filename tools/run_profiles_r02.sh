#!/bin/bash
# Round-2 profiling pass on one B200 (outputs -> gpurun_out/, small text/CSV only; copied into profiles/ afterwards).
mkdir -p gpurun_out
# 1. launch list of one eager step: duration + DRAM bytes per launch
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_step.csv python tools/ncu_step.py 16 > gpurun_out/r02_ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_step.csv gpurun_out/r02_gemm_traffic.json > gpurun_out/r02_launches_summary.txt 2>&1
# 2. full captures: 2-CTA GEMM, conv forward / wgrad, scan forward
for spec in "gemm_pair:4:pair" "conv_fwd_kernel:13:convf" "conv_wgrad_kernel:7:convw" "rssm_scan_fwd:1:scan" "gemm_tc_kernel:6:gemmtc"; do
  IFS=: read k c tag <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:$k -c $c -f -o gpurun_out/full_$tag python tools/ncu_step.py 16 > gpurun_out/r02_ncu_$tag.log 2>&1
  ncu -i gpurun_out/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>> gpurun_out/r02_ncu_$tag.log
  python tools/ncu_digest.py gpurun_out/full_${tag}_raw.csv > gpurun_out/r02_ncu_${tag}_digest.txt 2>&1
  rm -f gpurun_out/full_$tag.ncu-rep gpurun_out/full_${tag}_raw.csv
done
head -30 gpurun_out/r02_launches_summary.txt
du -sh gpurun_out

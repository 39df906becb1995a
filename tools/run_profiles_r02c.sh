#!/bin/bash
# Final pass of round 2: launch list of one eager step + full captures of the conv forward / wgrad kernels after the instruction diet.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02b_launches_step.csv python tools/ncu_step.py 16 > gpurun_out/r02b_ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02b_launches_step.csv gpurun_out/r02b_gemm_traffic.json > gpurun_out/r02b_launches_summary.txt 2>&1
for spec in "conv_fwd_kernel:13:convf" "cat_sample:4:catsample"; do
  IFS=: read k c tag <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:$k -c $c -f -o gpurun_out/full_$tag python tools/ncu_step.py 16 > gpurun_out/r02b_ncu_$tag.log 2>&1
  ncu -i gpurun_out/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>> gpurun_out/r02b_ncu_$tag.log
  python tools/ncu_digest.py gpurun_out/full_${tag}_raw.csv > gpurun_out/r02b_ncu_${tag}_digest.txt 2>&1
  rm -f gpurun_out/full_$tag.ncu-rep gpurun_out/full_${tag}_raw.csv
done
head -30 gpurun_out/r02b_launches_summary.txt
grep -A21 "conv_fwd_kernel<32" gpurun_out/r02b_ncu_convf_digest.txt | head -24

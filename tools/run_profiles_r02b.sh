#!/bin/bash
# Second profiling pass of round 2 (after the LayerNorm pipelines, the rollout restructuring, the streaming scan and the coalesced epilogues).
# Outputs -> gpurun_out/ (small text / CSV only); copied into profiles/ afterwards.
mkdir -p gpurun_out
# 1. launch list of one eager step: duration + DRAM bytes per launch
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02b_launches_step.csv python tools/ncu_step.py 16 > gpurun_out/r02b_ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02b_launches_step.csv gpurun_out/r02b_gemm_traffic.json > gpurun_out/r02b_launches_summary.txt 2>&1
# 2. full captures inside the step: chain kernel, 128xBN GEMM, LayerNorm pipelines
for spec in "mlp_chain_fwd:3:chain" "gemm_tc_kernel:6:gemmtc" "ln_bwd_pipe:4:lnbwd" "ln_fwd_pipe:3:lnfwd"; do
  IFS=: read k c tag <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:$k -c $c -f -o gpurun_out/full_$tag python tools/ncu_step.py 16 > gpurun_out/r02b_ncu_$tag.log 2>&1
  ncu -i gpurun_out/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>> gpurun_out/r02b_ncu_$tag.log
  python tools/ncu_digest.py gpurun_out/full_${tag}_raw.csv > gpurun_out/r02b_ncu_${tag}_digest.txt 2>&1
  rm -f gpurun_out/full_$tag.ncu-rep gpurun_out/full_${tag}_raw.csv
done
# 3. the weight-streaming scan kernels (configs[4] shard shape, T = 16)
timeout 600 ncu --set full --clock-control none -k regex:rssm_scan_stream --launch-skip 2 -c 2 -f -o gpurun_out/full_stream python tools/scan_stream_ncu.py 16 > gpurun_out/r02b_ncu_stream.log 2>&1
ncu -i gpurun_out/full_stream.ncu-rep --page raw --csv > gpurun_out/full_stream_raw.csv 2>> gpurun_out/r02b_ncu_stream.log
python tools/ncu_digest.py gpurun_out/full_stream_raw.csv > gpurun_out/r02b_ncu_stream_digest.txt 2>&1
rm -f gpurun_out/full_stream.ncu-rep gpurun_out/full_stream_raw.csv
head -32 gpurun_out/r02b_launches_summary.txt
cat gpurun_out/r02b_ncu_stream_digest.txt
du -sh gpurun_out

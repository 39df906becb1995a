#!/bin/bash
# Profiling pass on the GPU box (one B200): warm microbenchmarks, ncu launch list of one eager step, ncu --set full of the scan kernels
# and of the first GEMM launches of the step.  Outputs -> gpurun_out/ (copied into profiles/ by hand afterwards).
mkdir -p gpurun_out
python tests/microbench.py > gpurun_out/microbench.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv \
    python tests/ncu_step.py 16 > gpurun_out/ncu_step.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_summary.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rssm_scan -c 4 -f -o gpurun_out/scan_full \
    python tests/ncu_step.py 16 > gpurun_out/ncu_scan.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c ${NGEMM:-60} -f -o gpurun_out/gemm_full \
    python tests/ncu_step.py 16 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out

#!/bin/bash
# ncu --set full on the micro-benchmark kernels; exports the raw page as CSV on the box and keeps reports small
mkdir -p gpurun_out
K=${1:-"ln_fwd_fast|ln_bwd_colg|ln_bwd_fast|gemm_tc_kernel|conv_fwd_kernel|conv_wgrad_kernel"}
timeout 900 ncu --set full --clock-control none -k regex:"$K" -f -o gpurun_out/micro_full python tools/ncu_micro.py ${2:-ln,conv,gemm} > gpurun_out/ncu_micro.log 2>&1
ncu -i gpurun_out/micro_full.ncu-rep --page raw --csv > gpurun_out/micro_full_raw.csv 2> gpurun_out/ncu_export.log
ls -la gpurun_out/micro_full.ncu-rep
[ $(stat -c %s gpurun_out/micro_full.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/micro_full.ncu-rep
tail -3 gpurun_out/ncu_micro.log

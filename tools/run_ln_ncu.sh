#!/bin/bash
# ncu --set full of the LayerNorm kernels (second pass of tools/ln_ncu.py), digest -> gpurun_out/$1
out=${1:-r02_ncu_ln_pipe_digest.txt}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_(fwd|bwd)" --launch-skip 3 -c 3 -f -o gpurun_out/ln_full python tools/ln_ncu.py ${2:-38400} ${3:-1024} > gpurun_out/ln_ncu.log 2>&1
ncu -i gpurun_out/ln_full.ncu-rep --page raw --csv > gpurun_out/ln_full_raw.csv 2>> gpurun_out/ln_ncu.log
python tools/ncu_digest.py gpurun_out/ln_full_raw.csv > gpurun_out/$out 2>&1
cat gpurun_out/$out

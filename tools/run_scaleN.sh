#!/bin/bash
# weak-scaling line on N GPUs (per-GPU batch 50 x 48): bash tools/run_scaleN.sh N -> gpurun_out/r02b_bench_<N>gpu.json
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r02b_bench_${N}gpu.json 2> gpurun_out/r02b_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N=$N", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
PY

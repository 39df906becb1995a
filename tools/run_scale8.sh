#!/bin/bash
# 8-GPU weak-scaling lines (per-GPU batch 50 x 48) with a few NCCL / schedule variants; one JSON line each -> gpurun_out/r02b_bench_8gpu_<tag>.json
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/r02b_bench_8gpu_$tag.json 2> gpurun_out/r02b_bench_8gpu_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02b_bench_8gpu_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["ms_per_step"], d["value"])
except Exception as e:
    print("$tag failed", e)
PY
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
grep -h -i -o "nvls[a-z ]*\|[0-9]* coll channels\|[0-9]* nvls channels" gpurun_out/r02b_bench_8gpu_default.err | sort | uniq -c | sort -rn | head -8
run early_cta8 GENRL_DP_EARLY=1 NCCL_MAX_CTAS=8
run cta8 NCCL_MAX_CTAS=8

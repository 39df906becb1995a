#!/bin/bash
# Runs the GPU test-suite in isolated groups (a hung kernel only loses its own group). Output -> gpurun_out/
mkdir -p gpurun_out
run() { name=$1; shift; timeout -k 10 ${TMO:-600} python -m pytest -q -m gpu -p no:cacheprovider "$@" > gpurun_out/test_$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/test_$name.log; }
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/gpu.txt 2>&1
run gemm32 tests/test_gpu_ops.py -k "gemm_fp32"
run gemm16 tests/test_gpu_ops.py -k "gemm_bf16"
run ops tests/test_gpu_ops.py -k "not gemm and not scan"
run scan tests/test_gpu_ops.py -k "scan"
run agent tests/test_gpu_agent.py

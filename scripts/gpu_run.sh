#!/bin/bash
# One gpurun call: tests + bench (+ optional ncu).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
if [ "${RUN_BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py ${BENCH_ARGS:---steps 3 --warmup 3} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi

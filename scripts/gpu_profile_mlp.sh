#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu ${PROFILE_BENCH_ARGS:-}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc -s 3 -c 1 -o gpurun_out/prof_mlp -f $B > gpurun_out/prof_mlp.log 2>&1
echo "mlp capture exit $?"

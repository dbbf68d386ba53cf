#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the two dominant kernels.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu ${PROFILE_BENCH_ARGS:-}"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc -s 2 -c 2 -o gpurun_out/prof_mlp -f $B > gpurun_out/prof_mlp.log 2>&1
echo "mlp capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 2 -c 2 -o gpurun_out/prof_composite -f $B > gpurun_out/prof_composite.log 2>&1
echo "composite capture exit $?"
ls -la gpurun_out

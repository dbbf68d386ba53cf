#!/bin/bash
# Round evidence: tests, both bench arms, ncu launch list + full captures.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
timeout 600 python bench.py --embed mip --no-cpu > gpurun_out/bench_mip.json 2> gpurun_out/bench_mip.err; echo "bench mip exit $?"
timeout 600 python bench.py --precision fp32 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 exit $?"
B="python bench.py --steps 1 --warmup 1 --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc -s 3 -c 1 -o gpurun_out/prof_mlp -f $B > gpurun_out/prof_mlp.log 2>&1; echo "mlp capture exit $?"
NMR_TC_FUSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 2 -c 2 -o gpurun_out/prof_composite -f $B > gpurun_out/prof_composite.log 2>&1; echo "composite capture (per-sample features streamed, NMR_TC_FUSE=0) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 2 -c 2 -o gpurun_out/prof_composite_fused -f $B > gpurun_out/prof_composite_fused.log 2>&1; echo "composite capture (product path) exit $?"
timeout 300 python scripts/config3.py > gpurun_out/config3.json 2> gpurun_out/config3.err; echo "config3 exit $?"; cat gpurun_out/config3.json

#!/bin/bash
# Round-2 evidence (1 GPU): tests, bench arms, ncu launch list + captures of every kernel family.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm exit $?"
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cut -c1-300 $O/bench.json
timeout 600 python bench.py --precision fp16 --no-cpu > $O/bench_fp16.json 2> $O/bench_fp16.err; echo "bench fp16 exit $?"
timeout 600 python bench.py --embed mip --no-cpu > $O/bench_mip.json 2> $O/bench_mip.err; echo "bench mip exit $?"
timeout 600 python bench.py --impl torch-eager-gpu --steps 3 --warmup 1 > $O/bench_eager.json 2> $O/bench_eager.err; echo "eager arm exit $?"
B="python bench.py --steps 1 --warmup 1 --no-cpu --only-dev"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv $B > $O/launches.log 2>&1; echo "launch list exit $?"
# the MLP kernel: ONE pass with the tensor-pipe counters together (reconciles the round-1 53 % vs 69 % readings), then the full set
timeout 600 ncu --clock-control none -k regex:nerf_mlp_tc -s 3 -c 1 --csv --log-file $O/mlp_tensor_single_pass.csv \
  --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.sum,sm__inst_executed_pipe_tensor.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.max,sm__cycles_active.avg,gpu__time_duration.sum,smsp__cycles_active.avg,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  $B > $O/mlp_single.log 2>&1; echo "mlp single-pass exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_tc -s 3 -c 1 -o $O/prof_mlp -f $B > $O/prof_mlp.log 2>&1; echo "mlp capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 2 -c 2 -o $O/prof_composite_fused -f $B > $O/prof_composite_fused.log 2>&1; echo "composite capture (product path) exit $?"
NMR_TC_FUSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 2 -c 2 -o $O/prof_composite -f $B > $O/prof_composite.log 2>&1; echo "composite capture (streamed features) exit $?"
timeout 600 ncu --set full --clock-control none -k "regex:sample_coarse_kernel|raygen_kernel|view_bias_kernel" -c 6 -o $O/prof_sampling -f $B > $O/prof_sampling.log 2>&1; echo "sampling capture exit $?"
timeout 600 ncu --set full --clock-control none -k "regex:gemm3_tc_kernel|composite_backward_kernel|encode_backward_kernel|rgb_head_backward_kernel" -s 40 -c 8 -o $O/prof_backward -f python scripts/inerf_probe.py > $O/prof_backward.log 2>&1; echo "backward capture exit $?"
timeout 600 ncu --set full --clock-control none -k "regex:row_stats|col_stats_partial|conf_row|select_kernel|gemm3_tc_kernel|normalize_rows" -s 30 -c 8 -o $O/prof_match -f python scripts/config3.py > $O/prof_match.log 2>&1; echo "match capture exit $?"
timeout 300 python scripts/config3.py > $O/config3.json 2> $O/config3.err; echo "config3 exit $?"; cat $O/config3.json
timeout 900 python scripts/sweep.py --max-rays 1e8 --cpu-rays-per-s "$(python -c "import json;print(json.load(open('$O/bench.json'))['cpu_baseline']['value'])" 2>/dev/null || echo 0)" --out $O/sweep_n1.jsonl > $O/sweep_n1.log 2>&1; echo "sweep exit $?"; tail -4 $O/sweep_n1.log
for f in prof_mlp prof_composite_fused prof_composite prof_sampling prof_backward prof_match; do
  [ -f $O/$f.ncu-rep ] && python scripts/ncu_summary.py $O/$f.ncu-rep $O/$f.txt > /dev/null 2>&1
done
rm -f $O/prof_composite.ncu-rep $O/prof_sampling.ncu-rep $O/prof_backward.ncu-rep $O/prof_match.ncu-rep   # summaries kept; 64 MiB return limit
python scripts/launch_summary.py $O/launches.csv $O/launches_summary.csv "one bench step (--only-dev), ncu launch list" > /dev/null 2>&1
ls -la $O | head -60

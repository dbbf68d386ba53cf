#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q > gpurun_out/pytest_bf16.log 2>&1; echo "bf16 pytest exit $?"
tail -15 gpurun_out/pytest_bf16.log
python -c "
from nerfmatch_b200 import _lib as L
print('trap code', L.load().nmr_debug_trap_code())"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "all gpu pytest exit $?"
tail -8 gpurun_out/pytest_gpu.log
for cs in 1; do
  NMR_TC_CLUSTER=$cs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_cs$cs.json 2> gpurun_out/bench_cs$cs.err; echo "bench cs=$cs exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_cs$cs.json'))
print('cs=$cs', 'rays/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'mlp frac %.3f'%d['roofline']['frac'], d['stage_ms_per_step'], 'comp frac %.3f'%d['roofline_composite']['frac'], d['clocks'])
PY
done

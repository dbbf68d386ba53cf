#!/bin/bash
# Kernel iteration loop: bf16 parity tests, trap code, a short bench, the in-kernel timeline; for CG in $CGS.
set -u
mkdir -p gpurun_out
for CG in ${CGS:-2 1}; do
  export NMR_TC_CG=$CG
  echo "=== NMR_TC_CG=$CG"
  timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q > gpurun_out/pytest_bf16_cg$CG.log 2>&1; echo "bf16 pytest exit $?"
  tail -6 gpurun_out/pytest_bf16_cg$CG.log
  python -c "
from nerfmatch_b200 import _lib as L
print('trap code', L.load().nmr_debug_trap_code())"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_iter_cg$CG.json 2> gpurun_out/bench_iter_cg$CG.err; echo "bench exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_iter_cg$CG.json'))
    print('rays/s %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'mlp frac %.3f'%d['roofline']['frac'], d['stage_ms_per_step'], 'comp frac %.3f'%d['roofline_composite']['frac'], d['clocks'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_iter_cg$CG.err').read()[-1500:])
PY
  NMR_TC_TIMELINE=1 timeout 300 python scripts/tc_timeline.py > gpurun_out/tl_cg$CG.txt 2>&1; tail -13 gpurun_out/tl_cg$CG.txt
done

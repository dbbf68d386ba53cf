#!/bin/bash
# A/B on ONE box (boxes differ by +-10 % under the power cap): the library in the tree against the variants in build_ab/
# (libnmr_<name>.so).  Per variant: bf16 parity tests, the in-kernel timeline, two short device-resident bench runs.
set -u
mkdir -p gpurun_out
O=gpurun_out
export NMR_TC_CG=1
cp nerfmatch_b200/libnmr.so /tmp/libnmr_new.so
for v in new ${VARIANTS:-base}; do
  if [ $v = new ]; then cp /tmp/libnmr_new.so nerfmatch_b200/libnmr.so; else cp build_ab/libnmr_$v.so nerfmatch_b200/libnmr.so; fi
  echo "=== variant $v"
  if [ $v != base ]; then
    timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_inerf.py -x -q > $O/ab_pytest_bf16_$v.log 2>&1; echo "bf16 pytest exit $?"; tail -2 $O/ab_pytest_bf16_$v.log
    python -c "
from nerfmatch_b200 import _lib as L
print('trap code', L.load().nmr_debug_trap_code())"
  fi
  NMR_TC_TIMELINE=1 timeout 300 python scripts/tc_timeline.py > $O/ab_tl_$v.txt 2>&1; tail -16 $O/ab_tl_$v.txt
done
for rep in 1 2; do
  for v in new ${VARIANTS:-base}; do
    if [ $v = new ]; then cp /tmp/libnmr_new.so nerfmatch_b200/libnmr.so; else cp build_ab/libnmr_$v.so nerfmatch_b200/libnmr.so; fi
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --only-dev > $O/ab_bench_${v}_$rep.json 2> $O/ab_bench_${v}_$rep.err; echo "bench $v $rep exit $?"
    python - <<PY
import json
try:
    d=json.loads(open('$O/ab_bench_${v}_$rep.json').read().strip().splitlines()[-1])
    print('$v $rep', 'rays/s %.0f'%d['value'], 'ms %.2f'%d['ms_per_step'], d['clocks'])
except Exception as e:
    print('bench failed', e); print(open('$O/ab_bench_${v}_$rep.err').read()[-1500:])
PY
  done
done
cp /tmp/libnmr_new.so nerfmatch_b200/libnmr.so

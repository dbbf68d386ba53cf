#!/bin/bash
# Multi-GPU evidence on one box (round 2): bench at N = 1,2,4,8 (peer-memory exchange inside the timed step, NCCL gather as
# comparison leg, config 4 as an extra key) and the config-5 sweep at N = 2,4,8.
set -u
mkdir -p gpurun_out
O=gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
P=29710
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --no-cpu > $O/scale_n1.json 2> $O/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --no-cpu > $O/scale_n$N.json 2> $O/scale_n$N.err
    P=$((P+1))
    [ "${SWEEP:-1}" = 1 ] && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P scripts/sweep.py --min-rays 1e6 --max-rays 1e8 --out $O/sweep_n$N.jsonl > $O/sweep_n$N.log 2> $O/sweep_n$N.err
    P=$((P+1))
  fi
  echo "N=$N: $(python -c "
import json
d=json.loads(open('$O/scale_n$N.json').read())
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'cfg4', round(d['config4']['ms_per_batch'],2), d.get('exchange',{}).get('ms_per_step_nccl_gather'), d.get('exchange',{}).get('verified_equal_to_nccl_gather'), d['clocks']['sm_mhz'])
" 2>&1)"
  [ -f $O/sweep_n$N.jsonl ] && tail -2 $O/sweep_n$N.jsonl | cut -c1-260
done

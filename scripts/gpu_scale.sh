#!/bin/bash
# Multi-GPU evidence on one box: bench at N = 1,2,4,8 (whatever the box has) and config 4.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
P=29610
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
    timeout 300 python scripts/config4.py > gpurun_out/config4_n1.json 2> gpurun_out/config4_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    P=$((P+1))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P scripts/config4.py > gpurun_out/config4_n$N.json 2> gpurun_out/config4_n$N.err
    P=$((P+1))
  fi
  echo "N=$N: $(cut -c1-160 gpurun_out/scale_n$N.json)"; cat gpurun_out/config4_n$N.json
done

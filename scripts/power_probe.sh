#!/bin/bash
# What the MLP kernel's L2 -> SM weight stream costs in power (tools/l2_stream_power.cu): board power while every SM streams the
# 1.2 MB weight set from L2 at a given rate, with and without cluster multicast.
O=gpurun_out; mkdir -p $O
run() {  # name, args...
  name=$1; shift
  nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader,nounits -lms 200 > $O/power_$name.csv &
  SMI=$!
  sleep 0.6
  ./tools/l2_stream_power "$@" > $O/power_$name.json
  kill $SMI; wait $SMI 2>/dev/null
  python - <<PY
import json, statistics
rows=[l.strip().split(',') for l in open('$O/power_$name.csv') if l.strip()]
p=[float(r[0]) for r in rows][3:-1]; c=[float(r[1]) for r in rows][3:-1]
d=json.loads(open('$O/power_$name.json').read())
d.update(name='$name', power_w_median=statistics.median(p) if p else None, power_w_max=max(p) if p else None, sm_mhz_median=statistics.median(c) if c else None, samples=len(p))
print(json.dumps(d))
PY
}
run idle_spin -rate 0 -sec 4
run paced50 -rate 50 -sec 4
run paced50_mc -rate 50 -sec 4 -mc 1
run paced25 -rate 25 -sec 4
run paced100 -rate 100 -sec 4
run paced100_mc -rate 100 -sec 4 -mc 1
run unthrottled -rate -1 -sec 3
run unthrottled_mc -rate -1 -sec 3 -mc 1

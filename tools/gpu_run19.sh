#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2', round(j['value']), round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in r['kernels'].items()}, round(r['frac'],4), 'e2e', round(j['e2e']['ms_per_step'],2))"; }
for cfg in "default:" "bwd16:CCB_BATCH_BWD=16"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench19_$name.json 2> gpurun_out/bench19_$name.err || tail -3 gpurun_out/bench19_$name.err
  show gpurun_out/bench19_$name.json $name
done

#!/bin/bash
# A/B harness: tools/gpu_ab.sh "name:ENV=1 ENV2=x" ...   (bench.py kernel times per configuration)
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2', round(j['value']), round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in r['kernels'].items()}, round(r['frac'],4), 'e2e', round(j['e2e']['ms_per_step'],2))"; }
for cfg in "$@"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -3 gpurun_out/ab_$name.err
  show gpurun_out/ab_$name.json $name
done

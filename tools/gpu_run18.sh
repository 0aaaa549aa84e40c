#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2', round(j['value']), round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in r['kernels'].items()}, round(r['frac'],4), 'e2e', round(j['e2e']['ms_per_step'],2), 'host', round(j['host_enqueue_ms_per_step'],2), j['clocks'].get('samples'))"; }
for cfg in "default:" "ca:CCB_LIB_NAME=libctc_crf_b200_ca.so" "balsum:CCB_BALANCE_SUM=1" "ca_b8:CCB_LIB_NAME=libctc_crf_b200_ca.so CCB_BATCH_FWD=8"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench18_$name.json 2> gpurun_out/bench18_$name.err || tail -3 gpurun_out/bench18_$name.err
  show gpurun_out/bench18_$name.json $name
done
echo "== host probe"; timeout 300 python tools/host_probe.py 2>&1 | tail -12
echo "== pytest CA variant"; CCB_LIB_NAME=libctc_crf_b200_ca.so timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest18ca.log 2>&1; tail -3 gpurun_out/pytest18ca.log

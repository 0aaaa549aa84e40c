#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2', round(j['value']), round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in r['kernels'].items()}, round(r['frac'],4), 'e2e', round(j['e2e']['ms_per_step'],2), 'host', round(j['host_enqueue_ms_per_step'],2), j['clocks'].get('samples'))"; }
for cfg in "default:" "nosampler:CCB_BENCH_NO_SAMPLER=1" "nopadskip:CCB_DEBUG=4" "balsum:CCB_BALANCE_SUM=1" "default2:" "nosampler2:CCB_BENCH_NO_SAMPLER=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench17_$name.json 2> gpurun_out/bench17_$name.err || tail -3 gpurun_out/bench17_$name.err
  show gpurun_out/bench17_$name.json $name
done
echo "== timeline"; timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd|bwd|slowest|ratio"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest17.log 2>&1; tail -3 gpurun_out/pytest17.log

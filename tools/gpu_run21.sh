#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2', round(j['value']), round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in r['kernels'].items()}, round(r['frac'],4), 'e2e', round(j['e2e']['ms_per_step'],2), j.get('raw_logit_entry'))"; }
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench21.json 2> gpurun_out/bench21.err || tail -3 gpurun_out/bench21.err
show gpurun_out/bench21.json default
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-cuda --N 256 --T 3000 --varlen > gpurun_out/bench21_varlen.json 2> gpurun_out/bench21_varlen.err || tail -3 gpurun_out/bench21_varlen.err
show gpurun_out/bench21_varlen.json varlen256

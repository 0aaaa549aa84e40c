#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest20.log 2>&1; tail -5 gpurun_out/pytest20.log
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2', round(j['value']), round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in r['kernels'].items()}, round(r['frac'],4), 'e2e', round(j['e2e']['ms_per_step'],2), j.get('raw_logit_entry'))"; }
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench20.json 2> gpurun_out/bench20.err || tail -3 gpurun_out/bench20.err
show gpurun_out/bench20.json default
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-cuda --N 256 --T 3000 --varlen > gpurun_out/bench20_varlen.json 2> gpurun_out/bench20_varlen.err || tail -3 gpurun_out/bench20_varlen.err
show gpurun_out/bench20_varlen.json varlen256
echo "== sanitizer (ctc cluster kernel, smoke)"; timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san20_mem.log 2>&1; tail -3 gpurun_out/san20_mem.log
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san20_race.log 2>&1; tail -3 gpurun_out/san20_race.log

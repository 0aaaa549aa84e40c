#!/bin/bash
# SURVEY 8d configurations through bench.py (one line each) -> gpurun_out/configs.txt
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2 |', j['config']['workload'], '|', j['dtype'], '|', round(j['value']), 'frames/s |', round(j['ms_per_step'],2), 'ms/step | den fwd/bwd', {k:round(v['ms'],2) for k,v in r['kernels'].items()}, '| e2e', round(j['e2e']['value']))"; }
run() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline --no-ref-cuda --steps 5 --warmup 3 "$@" > gpurun_out/cfg_$name.json 2> gpurun_out/cfg_$name.err || tail -2 gpurun_out/cfg_$name.err; show gpurun_out/cfg_$name.json $name; }
{
run cfg2_N32_T800 --N 32 --T 800
run cfg3_bf16_V72 --dtype bf16 --V 72
run cfg3_bf16_V218 --dtype bf16
run cfg4_N16_T2000_5M --N 16 --T 2000 --H 100000
run cfg5_N256_varlen --N 256 --T 3000 --varlen
run headline
} 2>&1 | tee gpurun_out/configs.txt

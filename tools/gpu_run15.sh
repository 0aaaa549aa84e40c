#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest15.log 2>&1; tail -4 gpurun_out/pytest15.log
echo "== cfg3 bf16 V=72"; timeout 600 python bench.py --steps 3 --warmup 3 --V 72 --dtype bf16 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b15.err | python -c "
import json,sys;j=json.loads(sys.stdin.read());print(j['value'],j['ms_per_step'],j['roofline']['frac'],j['config']['den_graph'])"; tail -2 gpurun_out/b15.err
echo "== cfg4 per-GPU shape: N=16,T=2000, 5M-arc graph"; timeout 900 python bench.py --steps 2 --warmup 3 --N 16 --T 2000 --H 100000 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b15.err | python -c "
import json,sys;j=json.loads(sys.stdin.read());print(j['value'],j['ms_per_step'],j['roofline']['frac'],j['config']['den_graph'])"; tail -2 gpurun_out/b15.err
echo "== N=128"; timeout 900 python bench.py --steps 2 --warmup 3 --N 128 --T 1000 --no-cpu-baseline --no-ref-cuda 2> gpurun_out/b15.err | python -c "
import json,sys;j=json.loads(sys.stdin.read());print(j['value'],j['ms_per_step'],j['roofline']['frac'],j['config']['den_graph'])"; tail -2 gpurun_out/b15.err

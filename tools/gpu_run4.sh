#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu4.log 2>&1; tail -8 gpurun_out/pytest_gpu4.log; grep "max |grad" gpurun_out/pytest_gpu4.log
echo "== timeline"; timeout 300 python tools/timeline.py --T 200 2>&1 | tail -8
echo "== bench warps=16"; CCB_DEN_WARPS=16 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -2 gpurun_out/bench4.err; python -c "
import json;j=json.load(open('gpurun_out/bench4.json'));print(j['value'],j['ms_per_step'],j['roofline']['kernels'],j['roofline']['frac'],j['e2e']['value'])"
